// Synthetic Atari-shaped environment: 84x84 uint8 frames, 4-frame stack, sign-clipped reward,
// episodic done — the observation CONTRACT of jorldy/core/env/atari.py (state (1,4,84,84) uint8
// C-contiguous, newest frame last: atari.py:56-61,145-160; first state = first frame tiled x4:
// atari.py:112; reward in {-1,0,1}: atari.py:151-152) without ALE emulation, as the north star
// prescribes ("Atari paths fed by a synthetic 84x84x4 uint8 frame generator of identical
// dtype/layout").
//
// Generator (documented choice): pixels i.i.d. U{0..255} from
// Philox(seed, env id, frame counter); reward = +1 / -1 with probability 0.05 each, else 0;
// done ~ Bernoulli(1/1000); reset frame = a fresh random frame tiled over the 4 stack slots.
// All draws are pure functions of (seed, env, frame counter) -> the CPU oracle regenerates them
// bit-exactly (oracle/frames.py).
//
// One CTA per env.  HBM traffic per env-step: read 3 frames + write 4 (stack shift, 49 392 B) + the
// terminal `next_obs` copy (28 224 B) when the caller asks for it; HBM-write-bound by design.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int FRAME = 84 * 84;          // 7056 bytes
constexpr int STACK = 4;

// fills dst[0..7056) with the random frame number `fidx` of env `stream`
__device__ __forceinline__ void gen_frame(uint8_t* __restrict__ dst, uint64_t seed, uint64_t stream, uint64_t fidx) {
  // 7056 bytes = 441 x 16-byte Philox outputs
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int q = threadIdx.x; q < FRAME / 16; q += blockDim.x) {
    jb_philox4 r = jb_philox(seed, stream, fidx * 512 + (uint64_t)q);
    d4[q] = make_uint4(r.x, r.y, r.z, r.w);
  }
}

// event draws of frame fidx: u0 -> reward, u1 -> done
__device__ __forceinline__ void frame_events(uint64_t seed, uint64_t stream, uint64_t fidx, float& reward, bool& done) {
  jb_philox4 r = jb_philox(seed, stream, fidx * 512 + 511);
  const float u0 = jb_u01_float(r.x), u1 = jb_u01_float(r.y);
  reward = u0 < 0.05f ? 1.f : (u0 < 0.10f ? -1.f : 0.f);
  done = u1 < 0.001f;
}

__global__ void frames_reset_kernel(uint8_t* __restrict__ obs, int64_t* __restrict__ fcount, float* __restrict__ score,
                                    uint64_t seed, uint64_t stream_base, int n) {
  const int e = blockIdx.x;
  if (e >= n) return;
  uint8_t* o = obs + (size_t)e * STACK * FRAME;
  const uint64_t f = (uint64_t)fcount[e];
  gen_frame(o, seed, stream_base + e, f);
  __syncthreads();
  const uint4* s4 = reinterpret_cast<const uint4*>(o);
  for (int k = 1; k < STACK; ++k) {
    uint4* d4 = reinterpret_cast<uint4*>(o + (size_t)k * FRAME);
    for (int q = threadIdx.x; q < FRAME / 16; q += blockDim.x) d4[q] = s4[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) { fcount[e] = (int64_t)f + 1; score[e] = 0.f; }
}

__global__ void frames_step_kernel(uint8_t* __restrict__ obs, int64_t* __restrict__ fcount, float* __restrict__ score,
                                   uint8_t* __restrict__ next_obs, float* __restrict__ reward, float* __restrict__ done,
                                   float* __restrict__ stats, int auto_reset, uint64_t seed, uint64_t stream_base, int n) {
  const int e = blockIdx.x;
  if (e >= n) return;
  uint8_t* o = obs + (size_t)e * STACK * FRAME;
  const uint64_t f = (uint64_t)fcount[e];
  // shift the stack: slot k <- slot k+1 (each thread moves the same 16-byte lane through all slots)
  for (int q = threadIdx.x; q < FRAME / 16; q += blockDim.x) {
    uint4 a = reinterpret_cast<const uint4*>(o + 1 * (size_t)FRAME)[q];
    uint4 b = reinterpret_cast<const uint4*>(o + 2 * (size_t)FRAME)[q];
    uint4 c = reinterpret_cast<const uint4*>(o + 3 * (size_t)FRAME)[q];
    reinterpret_cast<uint4*>(o)[q] = a;
    reinterpret_cast<uint4*>(o + 1 * (size_t)FRAME)[q] = b;
    reinterpret_cast<uint4*>(o + 2 * (size_t)FRAME)[q] = c;
  }
  gen_frame(o + 3 * (size_t)FRAME, seed, stream_base + e, f);
  float r; bool d;
  frame_events(seed, stream_base + e, f, r, d);
  __syncthreads();
  if (next_obs) {
    uint4* dn = reinterpret_cast<uint4*>(next_obs + (size_t)e * STACK * FRAME);
    const uint4* s4 = reinterpret_cast<const uint4*>(o);
    for (int q = threadIdx.x; q < STACK * FRAME / 16; q += blockDim.x) dn[q] = s4[q];
  }
  float sc = score[e] + r;
  uint64_t fnext = f + 1;
  if (d && auto_reset) {
    __syncthreads();
    gen_frame(o, seed, stream_base + e, fnext);
    __syncthreads();
    const uint4* s4 = reinterpret_cast<const uint4*>(o);
    for (int k = 1; k < STACK; ++k) {
      uint4* d4 = reinterpret_cast<uint4*>(o + (size_t)k * FRAME);
      for (int q = threadIdx.x; q < FRAME / 16; q += blockDim.x) d4[q] = s4[q];
    }
    if (threadIdx.x == 0 && stats) { atomicAdd(&stats[0], 1.0f); atomicAdd(&stats[1], sc); }
    sc = 0.f;
    fnext += 1;
  }
  if (threadIdx.x == 0) {
    reward[e] = r; done[e] = d ? 1.f : 0.f;
    fcount[e] = (int64_t)fnext; score[e] = sc;
  }
}

}  // namespace

// obs: [n,4,84,84] uint8 (current stacked observation, updated in place); fcount: [n] int64 frame
// counters; score: [n] f32.
JB_API int jb_env_frames_reset(uint8_t* obs, int64_t* fcount, float* score, uint64_t seed, uint64_t stream_base, int n,
                               void* stream) {
  if (!obs || !fcount || !score || n <= 0) return JB_ERR_INVALID;
  frames_reset_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(obs, fcount, score, seed, stream_base, n);
  return jb_check_launch();
}

// next_obs (may be NULL): receives the post-step stack BEFORE any auto-reset (the transition's
// next_state, run_mode.py:91 semantics); obs then holds the observation to act on next.
JB_API int jb_env_frames_step(uint8_t* obs, int64_t* fcount, float* score, uint8_t* next_obs, float* reward, float* done,
                              float* stats, int auto_reset, uint64_t seed, uint64_t stream_base, int n, void* stream) {
  if (!obs || !fcount || !score || !reward || !done || n <= 0) return JB_ERR_INVALID;
  frames_step_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(obs, fcount, score, next_obs, reward, done, stats, auto_reset,
                                                         seed, stream_base, n);
  return jb_check_launch();
}
