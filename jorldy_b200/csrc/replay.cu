// HBM-resident replay ring: row scatter (store) and row gather (sample) for any transition field.
//
// Replaces the python-dict ring of jorldy/core/buffer/replay_buffer.py:16-23 (`self.buffer[self.buffer_index] = transition`)
// and the B-way `np.stack` of `stack_transition` (base.py:42-56) behind `sample` (:25-31) / PERBuffer.sample
// (per_buffer.py:95-96).  A field is a [capacity, row_bytes] byte matrix (uint8 frame stacks 28 224 B, f32 vectors 16..44 B,
// n-step reward/done rows 4 n B); rows are moved with 16-byte vectors when the row size and the base pointers allow it.
// Algorithmic traffic: 2 x row_bytes per transition (read + write), e.g. 2 x 56 448 B per sampled Atari transition
// (state + next_state) — SURVEY.md 8(d) "PER sample: gather of the transition dominates".
#include "common.cuh"

namespace {

template <typename V>
__global__ void replay_move_rows_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                        const int64_t* __restrict__ dst_rows, const int64_t* __restrict__ src_rows,
                                        long long vec_per_row) {
  // grid.y = row of the batch; grid.x covers the row's vectors.  dst_rows / src_rows: ring positions (NULL = batch order)
  const long long r = blockIdx.y;
  const long long dr = dst_rows ? dst_rows[r] : r, sr = src_rows ? src_rows[r] : r;
  const V* s = reinterpret_cast<const V*>(src) + sr * vec_per_row;
  V* d = reinterpret_cast<V*>(dst) + dr * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_row; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}

int move_rows(uint8_t* dst, const uint8_t* src, const int64_t* dst_rows, const int64_t* src_rows, int n, long long row_bytes,
              cudaStream_t s) {
  if (!dst || !src || n <= 0 || row_bytes <= 0 || n > 65535) return JB_ERR_INVALID;
  const bool v16 = row_bytes % 16 == 0 && (((uintptr_t)dst | (uintptr_t)src) & 15) == 0;
  const bool v4 = row_bytes % 4 == 0 && (((uintptr_t)dst | (uintptr_t)src) & 3) == 0;
  const long long vec = v16 ? row_bytes / 16 : (v4 ? row_bytes / 4 : row_bytes);
  const int threads = vec >= 256 ? 256 : (int)(((vec + 31) / 32) * 32);
  int gx = (int)((vec + threads - 1) / threads);
  if (gx > 64) gx = 64;
  dim3 grid(gx, n);
  if (v16) replay_move_rows_kernel<uint4><<<grid, threads, 0, s>>>(dst, src, dst_rows, src_rows, vec);
  else if (v4) replay_move_rows_kernel<uint32_t><<<grid, threads, 0, s>>>(dst, src, dst_rows, src_rows, vec);
  else replay_move_rows_kernel<uint8_t><<<grid, threads, 0, s>>>(dst, src, dst_rows, src_rows, vec);
  return jb_check_launch();
}

}  // namespace

// ring[pos[i]] = batch[i] for i < n (ring write; pos are ring positions, duplicates: the caller passes distinct positions)
JB_API int jb_replay_store(void* ring, const void* batch, const int64_t* pos, int n, long long row_bytes, void* stream) {
  if (!pos) return JB_ERR_INVALID;
  return move_rows((uint8_t*)ring, (const uint8_t*)batch, pos, nullptr, n, row_bytes, (cudaStream_t)stream);
}

// batch[i] = ring[idx[i]] for i < n (minibatch gather; indices may repeat)
JB_API int jb_replay_gather(const void* ring, const int64_t* idx, int n, long long row_bytes, void* batch, void* stream) {
  if (!idx) return JB_ERR_INVALID;
  return move_rows((uint8_t*)batch, (const uint8_t*)ring, nullptr, idx, n, row_bytes, (cudaStream_t)stream);
}
