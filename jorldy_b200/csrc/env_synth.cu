// Synthetic continuous-control env with the DIMENSIONS of a MuJoCo task (BASELINE configs[4]: obs 11 / act 3 = Hopper-v3),
// one thread per env instance.
//
// The reference steps MuJoCo through gym (jorldy/core/env/mujoco.py:25-58: rescale the action from [-1,1] to the
// action-space bounds, env.step, expand dims).  MuJoCo is a third-party simulator that is neither vendored nor
// installable here, and the north star replaces such emulators by synthetic generators of identical dtype/layout
// (SURVEY.md §8d row 5).  This generator:
//     s' = tanh(W_s s + W_a a) + 0.01 xi,  xi ~ N(0, I)      (fixed seeded W_s [D,D], W_a [D,A])
//     reward = -||s'||^2 / D,   done ~ Bernoulli(p_done) or elapsed == max_steps (gym TimeLimit, Hopper-v3: 1000)
//     reset: s ~ U(-0.05, 0.05)^D
// keeps JORLDY's wrapper semantics: float32 observation (1, D) per env, reward / done per env, and
// `state = next_state if not done else env.reset()` (run_mode.py:91) as an in-kernel auto-reset.
// Arithmetic is plain fp32 with explicit round-to-nearest mul/add (no FMA contraction) in ascending index order so
// that the numpy restatement (oracle/classic_control.py::SyntheticControlBatch) computes the same bits up to
// tanhf / the Box-Muller transcendental functions.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int MAX_D = 32, MAX_A = 8;

__device__ __forceinline__ void synth_reset_draw(uint64_t seed, uint64_t stream, uint64_t episode, int D, float* s) {
  // counters 2^40 + 16 * episode + k: disjoint from the per-step counters below
  for (int k = 0; k < (D + 3) / 4; ++k) {
    jb_philox4 r = jb_philox(seed, stream, (1ull << 40) + 16ull * episode + (uint64_t)k);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int q = 0; q < 4 && 4 * k + q < D; ++q) s[4 * k + q] = __fadd_rn(-0.05f, __fmul_rn(0.1f, jb_u01_float(w[q])));
  }
}

__global__ void synth_reset_kernel(float* __restrict__ obs, int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                   float* __restrict__ score, uint64_t seed, uint64_t stream_base, int n, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s[MAX_D];
  const int64_t ep = episode[i];
  synth_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, D, s);
  episode[i] = ep + 1;
  for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = s[k];
  elapsed[i] = 0;
  score[i] = 0.f;
}

__global__ void synth_step_kernel(float* __restrict__ obs, int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                  int64_t* __restrict__ tcount, float* __restrict__ score, const float* __restrict__ action,
                                  const float* __restrict__ Ws, const float* __restrict__ Wa, float* __restrict__ next_obs,
                                  float* __restrict__ reward, float* __restrict__ done, float* __restrict__ stats,
                                  int auto_reset, int max_steps, float p_done, uint64_t seed, uint64_t stream_base, int n,
                                  int D, int A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s[MAX_D], a[MAX_A], sn[MAX_D];
  for (int k = 0; k < D; ++k) s[k] = obs[(size_t)i * D + k];
  for (int k = 0; k < A; ++k) a[k] = action[(size_t)i * A + k];
  const uint64_t stream = stream_base + (uint64_t)i;
  const uint64_t t = (uint64_t)tcount[i];
  tcount[i] = (int64_t)(t + 1);
  float sq = 0.f;
  for (int k0 = 0; k0 < D; k0 += 4) {
    // 4 normals per Philox call: Box-Muller on (x, y) and (z, w); counter = 16 * t + k0 / 4
    jb_philox4 r = jb_philox(seed, stream, 16ull * t + (uint64_t)(k0 >> 2));
    float nrm[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t a0 = h ? r.z : r.x, b0 = h ? r.w : r.y;
      const float u1 = (float)((a0 >> 8) + 1u) * (1.0f / 16777216.0f);   // (0, 1]
      const float u2 = jb_u01_float(b0);
      const float rad = sqrtf(__fmul_rn(-2.0f, logf(u1)));
      nrm[2 * h] = __fmul_rn(rad, cospif(__fmul_rn(2.0f, u2)));
      nrm[2 * h + 1] = __fmul_rn(rad, sinpif(__fmul_rn(2.0f, u2)));
    }
    for (int q = 0; q < 4 && k0 + q < D; ++q) {
      const int k = k0 + q;
      float z = 0.f;
      for (int j = 0; j < D; ++j) z = __fadd_rn(z, __fmul_rn(Ws[k * D + j], s[j]));
      for (int j = 0; j < A; ++j) z = __fadd_rn(z, __fmul_rn(Wa[k * A + j], a[j]));
      const float v = __fadd_rn(tanhf(z), __fmul_rn(0.01f, nrm[q]));
      sn[k] = v;
      sq = __fadd_rn(sq, __fmul_rn(v, v));
    }
  }
  const float rew = -sq / (float)D;
  jb_philox4 rd = jb_philox(seed, stream, 16ull * t + 15ull);
  int el = elapsed[i] + 1;
  const bool d = (jb_u01_float(rd.x) < p_done) || (el >= max_steps);
  float sc = score[i] + rew;                                 // env.score accumulates the env reward (mujoco.py:50)
  for (int k = 0; k < D; ++k) next_obs[(size_t)i * D + k] = sn[k];
  reward[i] = rew;
  done[i] = d ? 1.f : 0.f;
  if (d && auto_reset) {
    if (stats) { atomicAdd(&stats[0], 1.0f); atomicAdd(&stats[1], sc); }
    const int64_t ep = episode[i];
    synth_reset_draw(seed, stream, (uint64_t)ep, D, sn);
    episode[i] = ep + 1;
    el = 0; sc = 0.f;
  }
  for (int k = 0; k < D; ++k) obs[(size_t)i * D + k] = sn[k];
  elapsed[i] = el;
  score[i] = sc;
}

}  // namespace

JB_API int jb_env_synth_reset(float* obs, int32_t* elapsed, int64_t* episode, float* score, uint64_t seed,
                              uint64_t stream_base, int n, int D, void* stream) {
  if (n <= 0 || !obs || !elapsed || !episode || !score || D <= 0 || D > MAX_D) return JB_ERR_INVALID;
  synth_reset_kernel<<<jb_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(obs, elapsed, episode, score, seed, stream_base, n, D);
  return jb_check_launch();
}

JB_API int jb_env_synth_step(float* obs, int32_t* elapsed, int64_t* episode, int64_t* tcount, float* score,
                             const float* action, const float* Ws, const float* Wa, float* next_obs, float* reward,
                             float* done, float* stats, int auto_reset, int max_steps, float p_done, uint64_t seed,
                             uint64_t stream_base, int n, int D, int A, void* stream) {
  if (n <= 0 || !obs || !elapsed || !episode || !tcount || !score || !action || !Ws || !Wa || !next_obs || !reward || !done)
    return JB_ERR_INVALID;
  if (D <= 0 || D > MAX_D || A <= 0 || A > MAX_A) return JB_ERR_INVALID;
  synth_step_kernel<<<jb_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(obs, elapsed, episode, tcount, score, action, Ws, Wa,
                                                                         next_obs, reward, done, stats, auto_reset, max_steps,
                                                                         p_done, seed, stream_base, n, D, A);
  return jb_check_launch();
}
