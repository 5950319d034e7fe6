// Per-row PPO loss forward + backward on the pre-activation head outputs, shared by the
// stand-alone loss kernel (ppo.cu) and the persistent minibatch-loop kernel (ppo_fused.cu) so both
// produce bit-identical values.
//
// Reference: jorldy/core/agent/ppo.py:127-162 and torch.distributions Categorical / Normal (see the
// header of ppo.cu for the restated definitions).
#pragma once
#include "common.cuh"

namespace jbppo {

constexpr int MAX_A = 8;
constexpr float F32_EPS = 1.1920928955078125e-07f;   // torch.finfo(float32).eps

struct HP { float eps_clip, vf_coef, ent_coef; };

struct RowOut {
  float dpol[2 * MAX_A];   // d loss / d policy head outputs: discrete [A] logits; continuous [A] mu then [A] log_std
  float dv1, dv2;          // value-head gradient if critic_loss1 / critic_loss2 is the max (already x vf_coef / B)
  float sq1, sq2;          // (v - ret)^2, (v_clip - ret)^2
  float surr_min, ent;     // min(surr1, surr2), entropy (continuous: summed over dims)
  float ratio, pmin;       // ratio; exp(log_prob) (continuous: min over dims)
};

// All per-action loops run over the compile-time bound MAX_A with an `a < A` guard: arrays stay in
// registers (a run-time trip count would index them dynamically and put them in local memory) and the
// arithmetic order is the plain ascending-a order of the restated definitions.
template <int NA = MAX_A>
__device__ __forceinline__ void log_softmax_row(const float (&lg)[MAX_A], int A, float (&lsm)[MAX_A]) {
  float mx = lg[0];
#pragma unroll
  for (int a = 1; a < NA; ++a) if (a < A) mx = fmaxf(mx, lg[a]);
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < NA; ++a) if (a < A) s += expf(lg[a] - mx);
  const float ls = logf(s);
#pragma unroll
  for (int a = NA; a < MAX_A; ++a) lsm[a] = 0.f;
#pragma unroll
  for (int a = 0; a < NA; ++a) lsm[a] = a < A ? (lg[a] - mx) - ls : 0.f;
}

__device__ __forceinline__ float atanh_clamped(float a) {
  const float hi = (float)(1.0 - 1e-7), lo = (float)(-1.0 + 1e-7);
  return atanhf(fminf(fmaxf(a, lo), hi));
}

__device__ __forceinline__ void surrogate(float ratio, float adv, float eps, float& smin, float& g) {
  const float s1 = ratio * adv;
  const float rc = fminf(fmaxf(ratio, 1.f - eps), 1.f + eps);
  const float s2 = rc * adv;
  const float inr = (ratio >= 1.f - eps && ratio <= 1.f + eps) ? 1.f : 0.f;
  smin = fminf(s1, s2);
  if (s1 < s2) g = adv;
  else if (s1 > s2) g = adv * inr;
  else g = 0.5f * adv + 0.5f * adv * inr;        // torch.minimum splits ties
}

// o: the row's head outputs [nout] (at least 2*MAX_A... entries readable up to index nout-1); a_disc / a_cont:
// the stored action; lpo: log_prob_old (1 or A values)
// NA: compile-time bound on A (the loops run to NA, guarded by a < A): row<., 2> costs a quarter of row<., 8>
template <bool CONT, int NA = MAX_A>
__device__ __forceinline__ void row(const float* o, int A, int a_disc, const float* a_cont, float adv, float ret,
                                    float vold, const float* lpo, HP hp, float invB, RowOut& r) {
  float v = 0.f;                                   // o[CONT ? 2A : A] without a run-time register index
#pragma unroll
  for (int q = 0; q < 2 * NA + 1; ++q) if (q == (CONT ? 2 * A : A)) v = o[q];
  const float dv_raw = v - vold;
  const float vclip = vold + fminf(fmaxf(dv_raw, -hp.eps_clip), hp.eps_clip);
  const float in_clip = (dv_raw >= -hp.eps_clip && dv_raw <= hp.eps_clip) ? 1.f : 0.f;
  const float d1 = v - ret, d2 = vclip - ret;
  r.sq1 = d1 * d1; r.sq2 = d2 * d2;
  r.dv1 = hp.vf_coef * invB * 2.f * d1;
  r.dv2 = hp.vf_coef * invB * 2.f * d2 * in_clip;
#pragma unroll
  for (int q = 0; q < 2 * MAX_A; ++q) r.dpol[q] = 0.f;
  if (!CONT) {
    float lg[MAX_A], lsm[MAX_A], pi[MAX_A], p[MAX_A], lc[MAX_A], inr[MAX_A];
#pragma unroll
    for (int a = 0; a < MAX_A; ++a) lg[a] = (a < NA && a < A) ? o[a] : 0.f;
    log_softmax_row<NA>(lg, A, lsm);
    float S = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) { pi[a] = 0.f; if (a < A) { pi[a] = expf(lsm[a]); S += pi[a]; } }
    float ent = 0.f, logp = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      p[a] = 1.f; lc[a] = 0.f; inr[a] = 0.f;
      if (a < A) {
        p[a] = pi[a] / S;
        const float pc = fminf(fmaxf(p[a], F32_EPS), 1.f - F32_EPS);
        inr[a] = (p[a] >= F32_EPS && p[a] <= 1.f - F32_EPS) ? 1.f : 0.f;
        lc[a] = logf(pc);
        ent -= lc[a] * p[a];
        if (a == a_disc) logp = lc[a];
      }
    }
    const float ratio = expf(logp - lpo[0]);
    float smin, gr;
    surrogate(ratio, adv, hp.eps_clip, smin, gr);
    r.surr_min = smin; r.ent = ent; r.ratio = ratio; r.pmin = expf(logp);
    const float dlogp = -gr * ratio * invB;            // d(actor_loss)/d log_prob
    const float dent = -hp.ent_coef * invB;            // d(ent_coef * entropy_loss)/d entropy_b
    float dp[MAX_A], dot = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      dp[a] = 0.f;
      if (a < A) {
        float t = dent * (-(lc[a] + inr[a]));
        if (a == a_disc) t += dlogp * inr[a] / p[a];
        dp[a] = t;
        dot += t * pi[a];
      }
    }
    float dlsm[MAX_A], sum_dlsm = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      dlsm[a] = 0.f;
      if (a < A) {
        const float dpi = dp[a] / S - dot / (S * S);
        dlsm[a] = dpi * pi[a];
        sum_dlsm += dlsm[a];
      }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) if (a < A) r.dpol[a] = dlsm[a] - expf(lsm[a]) * sum_dlsm;
  } else {
    const float log_sqrt_2pi = 0.9189385332046727f;
    float dsum = 0.f, ent = 0.f;
    float omu[MAX_A], ols[MAX_A], mu[MAX_A], sd[MAX_A], ls[MAX_A], z[MAX_A], dmu_[MAX_A], dls_[MAX_A];
#pragma unroll
    for (int a = 0; a < NA; ++a) {                 // o[a] and o[A + a] without run-time register indices
      omu[a] = a < A ? o[a] : 0.f;
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 2 * NA; ++q) if (q == A + a && a < A) t = o[q];
      ols[a] = t;
    }
    float pmin = INFINITY;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      mu[a] = sd[a] = ls[a] = z[a] = 0.f;
      if (a < A) {
        mu[a] = fminf(fmaxf(omu[a], -5.f), 5.f);
        ls[a] = tanhf(ols[a]);
        sd[a] = expf(ls[a]);
        z[a] = atanh_clamped(a_cont[a]);
        const float d = z[a] - mu[a];
        const float logp = -(d * d) / (2.f * (sd[a] * sd[a])) - logf(sd[a]) - log_sqrt_2pi;
        dsum += logp - lpo[a];
        ent += 0.5f + 0.5f * 1.8378770664093453f + logf(sd[a]);
        pmin = fminf(pmin, expf(logp));
      }
    }
    const float ratio = expf(dsum);
    float smin, gr;
    surrogate(ratio, adv, hp.eps_clip, smin, gr);
    r.surr_min = smin; r.ent = ent; r.ratio = ratio; r.pmin = pmin;
    const float dlogp = -gr * ratio * invB;
    const float dent = -hp.ent_coef * invB / (float)A;   // entropy_loss = -mean over B*A elements
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      dmu_[a] = dls_[a] = 0.f;
      if (a < A) {
        const float d = z[a] - mu[a];
        const float var = sd[a] * sd[a];
        const float dmu = dlogp * d / var;
        const float dsd = dlogp * (d * d / (var * sd[a]) - 1.f / sd[a]) + dent / sd[a];
        const float in_mu = (omu[a] >= -5.f && omu[a] <= 5.f) ? 1.f : 0.f;
        dmu_[a] = dmu * in_mu;
        dls_[a] = dsd * sd[a] * (1.f - ls[a] * ls[a]);
      }
    }
#pragma unroll
    for (int q = 0; q < 2 * NA; ++q) {
      float t = 0.f;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        if (a < A && q == a) t = dmu_[a];
        if (a < A && q == A + a) t = dls_[a];
      }
      r.dpol[q] = t;
    }
  }
}

// weights of the two critic means in max(c1, c2) (torch.maximum splits ties)
__device__ __forceinline__ void critic_weights(float c1, float c2, float& w1, float& w2) {
  w1 = (c1 > c2) ? 1.f : ((c1 == c2) ? 0.5f : 0.f);
  w2 = 1.f - w1;
}

}  // namespace jbppo
