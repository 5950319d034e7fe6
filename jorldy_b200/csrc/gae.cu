// Generalised Advantage Estimation: TD residual + backward scan + returns + per-row
// standardisation in one launch.
//
// Replaces the no-grad block of PPO.learn (jorldy/core/agent/ppo.py:95-110):
//   delta = reward + (1 - done) * gamma * next_value - value
//   adv[:, t] += (1 - done[:, t]) * gamma * lambda * adv[:, t + 1]    for t = T-2 .. 0
//   ret = adv + value
//   adv = (adv - mean_row) / (std_row + 1e-7)                         (unbiased std)
// which the reference runs as T-1 strided python-loop launches.
//
// Layout: every array is [N, T] row-major (actor-major, ppo.py:97 `view(-1, n_step)`).
// Arithmetic: the recurrence is evaluated in the reference's operation order with
// round-to-nearest f32 intrinsics (no FMA contraction), so delta/adv/ret are bit-exact
// against torch CPU; the row mean/std are accumulated in f64 (torch's CPU std kernel
// also accumulates in double) and rounded once.
//
// Mapping: one CTA owns ROWS (8) rows. Time is walked backwards in chunks of 64 columns; for
// each chunk the 8 warps load one row each, [8 x 64] tiles with lanes along t (coalesced 128-B
// segments), compute delta and the decay coefficient elementwise into shared memory,
// then warp 0 runs the 8 independent sequential scans out of shared memory (padded
// rows: conflict-free), and all warps stream adv/ret back out coalesced.
// Algorithmic HBM bytes: 16 B read + 8 B written per transition (+8 B re-read/written when
// standardising; that second pass is L2-resident for one CTA's rows).
#include "common.cuh"

namespace {

constexpr int ROWS = 8;     // rows per CTA: 8 x more CTAs than 32-row tiles keep ~7 CTAs resident per SM, which is what hides
                            // the load -> scan -> store latency chain of a chunk (ncu r02: 32-row CTAs reached 21 % of the HBM peak)
constexpr int CHUNK = 64;
constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS)
gae_kernel(const float* __restrict__ reward, const float* __restrict__ done,
           const float* __restrict__ value, const float* __restrict__ next_value,
           const float* __restrict__ last_value, int N, int T, float gamma, float lambda,
           int standardize, float* __restrict__ adv, float* __restrict__ ret) {
  __shared__ float s_delta[ROWS][CHUNK + 1];
  __shared__ float s_coef[ROWS][CHUNK + 1];
  __shared__ float s_carry[ROWS];
  __shared__ double s_sum[ROWS], s_sumsq[ROWS];

  const int row0 = blockIdx.x * ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < ROWS) { s_carry[threadIdx.x] = 0.f; s_sum[threadIdx.x] = 0.0; s_sumsq[threadIdx.x] = 0.0; }
  __syncthreads();

  const int n_chunks = (T + CHUNK - 1) / CHUNK;
  for (int c = n_chunks - 1; c >= 0; --c) {
    const int t0 = c * CHUNK;
    // ---- load + elementwise: warp w handles rows w, w+8, w+16, w+24; lanes along t ----
#pragma unroll
    for (int rr = 0; rr < ROWS / 8; ++rr) {
      const int r = warp + rr * 8;
      const int row = row0 + r;
#pragma unroll
      for (int h = 0; h < CHUNK / 32; ++h) {
        const int tl = lane + h * 32;
        const int t = t0 + tl;
        float dl = 0.f, cf = 0.f;
        if (row < N && t < T) {
          const size_t idx = (size_t)row * T + t;
          const float d = done[idx];
          const float one_minus = __fadd_rn(1.0f, -d);
          float nv;
          if (next_value) nv = next_value[idx];
          else nv = (t + 1 < T) ? value[idx + 1] : last_value[row];
          // reward + ((1-done)*gamma)*next_value - value      (ppo.py:95)
          const float tmp = __fmul_rn(__fmul_rn(one_minus, gamma), nv);
          dl = __fadd_rn(__fadd_rn(reward[idx], tmp), -value[idx]);
          // ((1-done)*gamma)*lambda                           (ppo.py:100)
          cf = __fmul_rn(__fmul_rn(one_minus, gamma), lambda);
        }
        s_delta[r][tl] = dl;
        s_coef[r][tl] = cf;
      }
    }
    __syncthreads();
    // ---- sequential backward scan: warp 0, lane = row ----
    if (warp == 0 && lane < ROWS) {
      const int row = row0 + lane;
      float carry = s_carry[lane];
      double sm = s_sum[lane], sq = s_sumsq[lane];
      const int tmax = min(CHUNK, T - t0);
      for (int tl = tmax - 1; tl >= 0; --tl) {
        const int t = t0 + tl;
        float a = s_delta[lane][tl];
        if (t != T - 1) a = __fadd_rn(a, __fmul_rn(s_coef[lane][tl], carry));
        carry = a;
        s_delta[lane][tl] = a;
        sm += (double)a; sq += (double)a * (double)a;
      }
      if (row < N) { s_carry[lane] = carry; s_sum[lane] = sm; s_sumsq[lane] = sq; }
    }
    __syncthreads();
    // ---- write adv (raw) and ret = adv + value ----
#pragma unroll
    for (int rr = 0; rr < ROWS / 8; ++rr) {
      const int r = warp + rr * 8;
      const int row = row0 + r;
#pragma unroll
      for (int h = 0; h < CHUNK / 32; ++h) {
        const int tl = lane + h * 32;
        const int t = t0 + tl;
        if (row < N && t < T) {
          const size_t idx = (size_t)row * T + t;
          const float a = s_delta[r][tl];
          adv[idx] = a;
          ret[idx] = __fadd_rn(a, value[idx]);
        }
      }
    }
    __syncthreads();
  }

  if (standardize) {
    // (adv - mean) / (std + 1e-7), unbiased std                 (ppo.py:105-108)
#pragma unroll
    for (int rr = 0; rr < ROWS / 8; ++rr) {
      const int r = warp + rr * 8;
      const int row = row0 + r;
      if (row >= N) continue;
      const double mean_d = s_sum[r] / (double)T;
      double var_d = (s_sumsq[r] - s_sum[r] * mean_d) / (double)(T > 1 ? T - 1 : 1);
      if (var_d < 0) var_d = 0;
      const float mean = (float)mean_d;
      const float denom = __fadd_rn((float)sqrt(var_d), 1e-7f);
      for (int t = lane; t < T; t += 32) {
        const size_t idx = (size_t)row * T + t;
        adv[idx] = __fdiv_rn(__fadd_rn(adv[idx], -mean), denom);
      }
    }
  }
}

}  // namespace

// reward/done/value (and next_value if non-null) are [N,T] f32 device arrays. When next_value is
// null, V(s') is taken as value[:, t+1] for t < T-1 and last_value[N] for the final column (valid
// because (1-done) masks the only rows where next_state != state[t+1]).
JB_API int jb_gae(const float* reward, const float* done, const float* value, const float* next_value,
                  const float* last_value, int N, int T, float gamma, float lambda, int standardize,
                  float* adv, float* ret, void* stream) {
  if (N <= 0 || T <= 0 || !reward || !done || !value || !adv || !ret) return JB_ERR_INVALID;
  if (!next_value && !last_value) return JB_ERR_INVALID;
  int blocks = jb_div_up(N, ROWS);
  gae_kernel<<<blocks, THREADS, 0, (cudaStream_t)stream>>>(reward, done, value, next_value, last_value, N, T,
                                                         gamma, lambda, standardize, adv, ret);
  return jb_check_launch();
}
