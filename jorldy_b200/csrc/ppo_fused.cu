// Persistent cooperative kernel: a whole run of PPO minibatch steps (ppo.py:118-175 — forward, clipped
// surrogate / clipped value / entropy loss, backward, global-norm clip, Adam) in ONE launch.
//
// Why: at the reference's minibatch size (256 x [4-512-512-(A+1)]) one step is ~0.4 GFLOP and < 4 MB of
// traffic, i.e. microseconds of work, and one learn() is n_epoch * N*T/B = 6144 strictly sequential
// steps.  As 13 separate launches per step (round-1 "graph" path) each step cost 71 us, almost all of
// it launch gaps and cold per-kernel load latency (profiles/r01_launches_ppo_summary.md).  Here one
// CTA per SM stays resident for the whole epoch, phases are separated by a hand-rolled grid barrier
// (one atomic + spin, ~1 us), and producer/consumer fusion removes intermediate tensors:
//
//   P1  layer-2 forward tiles (32x32, fp32 FFMA, k split 4 ways in-CTA).  The A panel h1 =
//       relu(x W1^T + b1) is GENERATED in shared memory from the gathered state rows (K = D <= 16), never
//       read from HBM; the W2 panel arrives by cp.async.cg.  Column-tile-0 CTAs also publish h1 / xg.
//   P2  one warp per row: narrow heads + the whole loss row math (shared with ppo.cu through
//       ppo_rowmath.cuh): d loss/d logits, the two candidate value gradients, per-row statistics.
//   P3  backward tiles.  dh2 = (dout Wh) * relu'(h2) is GENERATED on the fly into the A panel of
//       both products that consume it — dW2 = dh2^T h1 (+ db2 as the panel row-sum) and
//       dh1 = (dh2 W2) * relu'(h1) — plus the head weight gradients.  Every job folds its tile's sum of
//       squares into a per-CTA accumulator (global-norm clipping needs ||g|| before any update).
//   P4  dW1 / db1 column jobs, learn()-level statistics (CTA 0), publish the per-CTA norm partials.
//   P5  Adam on a static 1/gridDim slice of the flat parameter buffer, clip coefficient from the
//       fixed-order fold of the partials.
//
// Determinism: static job -> CTA maps, fixed-order reductions, no float atomics: bit-reproducible run
// to run.  Coherence: every buffer written inside the kernel is read with ld.global.cg / cp.async.cg
// (L2), and the barrier's gpu-scope fences order the phases.
// Constraints (else the host uses the multi-launch path): B % 32 == 0, H % 32 == 0, H <= 512, D <= 16,
// nout <= 8, single GPU (no gradient all-reduce between backward and Adam).
#include <cstdlib>
#include "common.cuh"
#include "ppo_rowmath.cuh"
#include "../../include/jorldy_b200_fused.h"

namespace {

constexpr int NT = 256;
constexpr int PK = 512;
constexpr int MAXD = 16;
constexpr int MAXO = 8;
constexpr int ROWBUF = 24;          // floats per row: dpol[16] dv1 dv2 sq1 sq2 surr_min ent ratio pmin
constexpr int SMALL_FLOATS = 2048;  // xs[32*16], ds[32*8], reduction scratch
constexpr int PANEL_FLOATS = PK * 36;   // >= 32 * (PK + 4)

typedef jb_ppo_fused_args Args;

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }

__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_commit_wait() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
}

// ---- grid barrier (all CTAs co-resident: cooperative launch) ------------------------------------------
__device__ __forceinline__ void grid_bar(unsigned int* ctr, unsigned int& epoch, unsigned int nctas) {
  __syncthreads();                       // CTA scope: every thread's phase writes happen-before thread 0's release
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned int target = epoch * nctas;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// ---- panel staging ---------------------------------------------------------------------------------
// KC: sm[r*(kp+4) + k] = G[(r0+r)*ld + k0 + k], r < 32, k < kp
__device__ __forceinline__ void stage_kc(float* sm, const float* G, int ld, int r0, int k0, int kp) {
  const int c16s = kp >> 2, total = 32 * c16s, stride = kp + 4;
  for (int e = threadIdx.x; e < total; e += NT) {
    const int row = e / c16s, c = e - row * c16s;
    cp16(&sm[row * stride + c * 4], &G[(size_t)(r0 + row) * ld + k0 + c * 4]);
  }
}
// nonKC: sm[k*36 + c] = G[(k0+k)*ld + r0 + c], k < kp, c < 32
__device__ __forceinline__ void stage_nonkc(float* sm, const float* G, int ld, int r0, int k0, int kp) {
  const int total = kp * 8;
  for (int e = threadIdx.x; e < total; e += NT) {
    const int k = e >> 3, c = e & 7;
    cp16(&sm[k * 36 + c * 4], &G[(size_t)(k0 + k) * ld + r0 + c * 4]);
  }
}

// ---- 32x32 tile product over a staged panel ----------------------------------------------------------
// 256 threads = 16 k-groups x (4 x 4) threads, each thread an 8x8 register micro-tile: 16 LDS.128 per
// 256 FFMA.  (The first version used 4 k-groups x 4x4 micro-tiles = 8 LDS.128 per 64 FFMA; every LDS.128
// costs 4 shared-memory phases whatever the broadcast pattern, so that shape was shared-memory-bandwidth
// bound at <= 50 % of the FFMA rate.)  Rows/cols are interleaved (r = ty + 4 i) for k-contiguous panels so
// a quarter-warp's 16-byte reads fall in distinct banks, contiguous (r = 8 ty + i) for [k][32] panels.
constexpr int KG = 16;
template <bool A_KC, bool B_KC>
__device__ __forceinline__ void tile_mma(const float* As, const float* Bs, int kp, float (&acc)[8][8], float* rs) {
  const int tid = threadIdx.x, grp = tid >> 4, t = tid & 15, tx = t & 3, ty = t >> 2;
  const int a_stride = kp + 4, b_stride = kp + 4;
  for (int k = grp * 4; k < kp; k += 4 * KG) {
    float a[8][4], b[8][4];
    if (A_KC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(&As[(ty + 4 * i) * a_stride + k]);
        a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v0 = *reinterpret_cast<const float4*>(&As[(k + q) * 36 + ty * 8]);
        const float4 v1 = *reinterpret_cast<const float4*>(&As[(k + q) * 36 + ty * 8 + 4]);
        a[0][q] = v0.x; a[1][q] = v0.y; a[2][q] = v0.z; a[3][q] = v0.w;
        a[4][q] = v1.x; a[5][q] = v1.y; a[6][q] = v1.z; a[7][q] = v1.w;
      }
    }
    if (B_KC) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[(tx + 4 * j) * b_stride + k]);
        b[j][0] = v.x; b[j][1] = v.y; b[j][2] = v.z; b[j][3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v0 = *reinterpret_cast<const float4*>(&Bs[(k + q) * 36 + tx * 8]);
        const float4 v1 = *reinterpret_cast<const float4*>(&Bs[(k + q) * 36 + tx * 8 + 4]);
        b[0][q] = v0.x; b[1][q] = v0.y; b[2][q] = v0.z; b[3][q] = v0.w;
        b[4][q] = v1.x; b[5][q] = v1.y; b[6][q] = v1.z; b[7][q] = v1.w;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i][q], b[j][q], acc[i][j]);
    if (rs) {
#pragma unroll
      for (int i = 0; i < 8; ++i) rs[i] += (a[i][0] + a[i][1]) + (a[i][2] + a[i][3]);
    }
  }
}

template <bool A_KC>
__device__ __forceinline__ int tile_row(int ty, int i) { return A_KC ? (ty + 4 * i) : (ty * 8 + i); }

// fold the 16 k-groups in a fixed order; thread gets outputs e = tid + r*256 -> (m = e>>5, n = e&31).
// `red` (16 x (32*36 + 16) floats = 75 KB) aliases the dead A panel and the head of the dead B panel.
// Layout: row stride 36, odd groups shifted by 16 banks, columns rotated by 4*(m>>3) for [k][32] A panels:
// the 64 partial-tile stores of a warp are conflict-free (k-contiguous panels) or at most 2-way.
template <bool A_KC>
__device__ __forceinline__ int red_index(int grp, int m, int n) {
  const int rot = A_KC ? 0 : 4 * (m >> 3);
  return grp * (32 * 36 + 16) + m * 36 + ((n + rot) & 31) + (grp & 1) * 16;   // group stride keeps the shifted rows apart
}
template <bool A_KC, bool B_KC>
__device__ __forceinline__ void tile_reduce(const float (&acc)[8][8], float* red, float (&outv)[4]) {
  const int tid = threadIdx.x, grp = tid >> 4, t = tid & 15, tx = t & 3, ty = t >> 2;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      red[red_index<A_KC>(grp, tile_row<A_KC>(ty, i), tile_row<B_KC>(tx, j))] = acc[i][j];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = tid + r * NT, m = e >> 5, n = e & 31;
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < KG; ++g) v += red[red_index<A_KC>(g, m, n)];
    outv[r] = v;
  }
  __syncthreads();
}

// fixed-order block sum of one float (all threads call; result valid in thread 0 only)
__device__ __forceinline__ float block_sum0(float v, float* scratch /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) for (int w = 0; w < NT / 32; ++w) t += scratch[w];
  return t;
}

struct HeadTab { const float* w[MAXO]; const float* b[MAXO]; float* gw[MAXO]; float* gb[MAXO]; };

__global__ void __launch_bounds__(NT, 1) ppo_epoch_kernel(Args a, int dsm_floats, int skip /* debug: phase bitmask to skip (timing only) */) {
  extern __shared__ __align__(16) float smem[];
  float* s_small = smem;                       // SMALL_FLOATS
  float* dsm = smem + SMALL_FLOATS;            // resolved d loss/d head-outputs [B][MAXO] (P3) / reduction scratch (P4)
  float* As = dsm + dsm_floats;                // PANEL_FLOATS
  float* Bs = As + PANEL_FLOATS;               // PANEL_FLOATS
  float* xs = s_small;                         // [32][MAXD]
  float* scr = s_small + 768;                  // reduction scratch [128]
  int* sidx = reinterpret_cast<int*>(s_small + 1024);   // [32] gathered rollout row ids of the tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int nctas = gridDim.x;
  const int cta = blockIdx.x;
  const int B = a.B, D = a.D, H = a.H, A = a.A, nout = a.nout;
  const int npol = a.continuous ? 2 * A : A;
  const float invB = 1.0f / (float)B;
  const jbppo::HP hp{a.eps_clip, a.vf_coef, a.ent_coef};
  unsigned int epoch = 0;
  float cta_norm = 0.f;                        // thread 0 only
  const long long step0 = *a.step, cursor0 = *a.cursor;
  const float lr = *a.lr;

  HeadTab ht;
  {
    int o = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g)
      for (int q = 0; q < a.nh[g]; ++q, ++o) {
#pragma unroll
        for (int u = 0; u < MAXO; ++u)
          if (u == o) { ht.w[u] = a.Wh[g] + (size_t)q * H; ht.b[u] = a.bh[g] + q; ht.gw[u] = a.gWh[g] + (size_t)q * H; ht.gb[u] = a.gbh[g] + q; }
      }
#pragma unroll
    for (int u = 0; u < MAXO; ++u)
      if (u >= o) { ht.w[u] = a.Wh[0]; ht.b[u] = a.bh[0]; ht.gw[u] = a.gWh[0]; ht.gb[u] = a.gbh[0]; }
  }

  const int MT = B / 32, NTL = H / 32;

  for (int s = 0; s < a.n_steps; ++s) {
    // =========================== P1: h2 = relu(relu(x W1^T + b1) W2^T + b2) ===========================
    for (int job = cta; job < ((skip & 1) ? 0 : MT * NTL); job += (int)nctas) {
      const int mt = job / NTL, nt = job - mt * NTL;
      const int m0 = mt * 32, n0 = nt * 32;
      __syncthreads();
      if (tid < 32) {
        const int r = a.perm[(cursor0 + s) * (long long)B + m0 + tid];
        sidx[tid] = r;
        if (nt == 0) a.cur_idx[m0 + tid] = r;
      }
      stage_kc(Bs, a.W2, H, n0, 0, H);
      __syncthreads();
      for (int e = tid; e < 32 * D; e += NT) {
        const int r = e / D, i = e - r * D;
        const float v = a.state[(size_t)sidx[r] * D + i];
        xs[i * 32 + r] = v;                                   // transposed: [i][r]
        if (nt == 0) a.xg[(size_t)(m0 + r) * D + i] = v;
      }
      __syncthreads();
      for (int k = tid; k < H; k += NT) {
        float w[MAXD];
#pragma unroll
        for (int i = 0; i < MAXD; ++i) w[i] = i < D ? ldcg(a.W1 + (size_t)k * D + i) : 0.f;
        const float bb = ldcg(a.b1 + k);
        float hacc[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) hacc[r] = 0.f;
#pragma unroll
        for (int i = 0; i < MAXD; ++i) {
          if (i < D) {                                         // uniform branch: no wasted issue slots for D < 16
            const float wi = w[i];
#pragma unroll
            for (int r4 = 0; r4 < 8; ++r4) {
              const float4 xv = *reinterpret_cast<const float4*>(&xs[i * 32 + r4 * 4]);
              hacc[r4 * 4 + 0] = fmaf(xv.x, wi, hacc[r4 * 4 + 0]); hacc[r4 * 4 + 1] = fmaf(xv.y, wi, hacc[r4 * 4 + 1]);
              hacc[r4 * 4 + 2] = fmaf(xv.z, wi, hacc[r4 * 4 + 2]); hacc[r4 * 4 + 3] = fmaf(xv.w, wi, hacc[r4 * 4 + 3]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float hv = fmaxf(hacc[r] + bb, 0.f);
          As[r * (H + 4) + k] = hv;
          if (nt == 0) a.h1[(size_t)(m0 + r) * H + k] = hv;
        }
      }
      cp_commit_wait();
      __syncthreads();
      float acc[8][8] = {};
      tile_mma<true, true>(As, Bs, H, acc, nullptr);
      float outv[4];
      tile_reduce<true, true>(acc, As, outv);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = tid + r * NT, m = e >> 5, n = e & 31;
        const float v = outv[r] + ldcg(a.b2 + n0 + n);
        a.h2[(size_t)(m0 + m) * H + n0 + n] = fmaxf(v, 0.f);
      }
    }
    grid_bar(a.barrier, epoch, nctas);

    // =========================== P2: heads + loss rows (warp per row) ==================================
    for (int b = cta * (NT / 32) + warp; b < ((skip & 2) ? 0 : B); b += (int)nctas * (NT / 32)) {
      const float* hrow = a.h2 + (size_t)b * H;
      float hv[PK / 32];
#pragma unroll
      for (int j = 0; j < PK / 32; ++j) { const int k = lane + 32 * j; hv[j] = k < H ? ldcg(hrow + k) : 0.f; }
      float ov[MAXO];
#pragma unroll
      for (int o = 0; o < MAXO; ++o) {
        float acc = 0.f;
        if (o < nout) {
          float wv[PK / 32];
#pragma unroll
          for (int j = 0; j < PK / 32; ++j) { const int k = lane + 32 * j; wv[j] = k < H ? ldcg(ht.w[o] + k) : 0.f; }
#pragma unroll
          for (int j = 0; j < PK / 32; ++j) acc = fmaf(hv[j], wv[j], acc);
          acc = jb_warp_sum(acc) + ldcg(ht.b[o]);
        }
        ov[o] = acc;
      }
      if (lane == 0) {
        const int r = __ldcg(a.cur_idx + b);
        jbppo::RowOut ro;
        if (a.continuous)
          jbppo::row<true>(ov, A, 0, (const float*)a.action + (size_t)r * A, a.adv[r], a.ret[r], a.vold[r],
                           a.logp_old + (size_t)r * A, hp, invB, ro);
        else
          jbppo::row<false>(ov, A, ((const int32_t*)a.action)[r], nullptr, a.adv[r], a.ret[r], a.vold[r],
                            a.logp_old + r, hp, invB, ro);
        float* rb = a.rowbuf + (size_t)b * ROWBUF;
#pragma unroll
        for (int q = 0; q < 16; ++q) rb[q] = ro.dpol[q];
        rb[16] = ro.dv1; rb[17] = ro.dv2; rb[18] = ro.sq1; rb[19] = ro.sq2;
        rb[20] = ro.surr_min; rb[21] = ro.ent; rb[22] = ro.ratio; rb[23] = ro.pmin;
      }
    }
    grid_bar(a.barrier, epoch, nctas);

    // =========================== P3: backward tiles ====================================================
    // critic means (every CTA, fixed order) -> which branch of max(c1, c2) carries the gradient; then the
    // resolved d loss / d head outputs of the whole minibatch into shared memory
    float c1, c2;
    {
      float p1 = 0.f, p2 = 0.f;
      for (int b = tid; b < B; b += NT) { p1 += ldcg(a.rowbuf + (size_t)b * ROWBUF + 18); p2 += ldcg(a.rowbuf + (size_t)b * ROWBUF + 19); }
      const float t1 = block_sum0(p1, scr), t2 = block_sum0(p2, scr + 8);
      if (tid == 0) { scr[16] = t1 * invB; scr[17] = t2 * invB; }
      __syncthreads();
      c1 = scr[16]; c2 = scr[17];
      float w1, w2;
      jbppo::critic_weights(c1, c2, w1, w2);
      for (int e = tid; e < B * MAXO; e += NT) {
        const int m = e >> 3, o = e & 7;
        const float* rb = a.rowbuf + (size_t)m * ROWBUF;
        float v = 0.f;
        if (o < npol) v = ldcg(rb + o);
        else if (o == npol) v = w1 * ldcg(rb + 16) + w2 * ldcg(rb + 17);
        dsm[e] = v;
      }
      __syncthreads();
    }
    const int nJB = MT * NTL;          // dh1 tiles
    const int nJA = NTL * NTL;         // dW2 tiles
    const int nJC = NTL;               // head weight-gradient column jobs
    for (int job = cta; job < ((skip & 4) ? 0 : nJB + nJA + nJC); job += (int)nctas) {
      __syncthreads();
      if (job < nJB) {
        // ---- dh1 tile = ((dout Wh) * relu'(h2)) W2, masked by relu'(h1) -------------------------------
        const int mt = job / NTL, kt = job - mt * NTL;
        const int m0 = mt * 32, k0 = kt * 32;
        stage_kc(As, a.h2, H, m0, 0, H);
        stage_nonkc(Bs, a.W2, H, k0, 0, H);
        float wr[2][MAXO];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int n = tid + u * NT;
#pragma unroll
          for (int o = 0; o < MAXO; ++o) wr[u][o] = (o < nout && n < H) ? ldcg(ht.w[o] + n) : 0.f;
        }
        cp_commit_wait();
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int n = tid + u * NT;
          if (n < H) {
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const float4 d0 = *reinterpret_cast<const float4*>(&dsm[(m0 + r) * MAXO]);
              const float4 d1 = *reinterpret_cast<const float4*>(&dsm[(m0 + r) * MAXO + 4]);
              float dot = d0.x * wr[u][0];
              dot = fmaf(d0.y, wr[u][1], dot); dot = fmaf(d0.z, wr[u][2], dot); dot = fmaf(d0.w, wr[u][3], dot);
              dot = fmaf(d1.x, wr[u][4], dot); dot = fmaf(d1.y, wr[u][5], dot); dot = fmaf(d1.z, wr[u][6], dot); dot = fmaf(d1.w, wr[u][7], dot);
              const float hvv = As[r * (H + 4) + n];
              As[r * (H + 4) + n] = hvv > 0.f ? dot : 0.f;
            }
          }
        }
        __syncthreads();
        float acc[8][8] = {};
        tile_mma<true, false>(As, Bs, H, acc, nullptr);
        float outv[4];
        tile_reduce<true, false>(acc, As, outv);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = tid + r * NT, m = e >> 5, n = e & 31;
          const size_t off = (size_t)(m0 + m) * H + k0 + n;
          a.dh1[off] = ldcg(a.h1 + off) > 0.f ? outv[r] : 0.f;
        }
      } else if (job < nJB + nJA) {
        // ---- dW2 tile [n0.., k0..] = sum_m dh2[m, n] h1[m, k]; db2 = row sums (k tile 0) ----------------
        const int j = job - nJB;
        const int nt = j / NTL, kt = j - nt * NTL;
        const int n0 = nt * 32, k0 = kt * 32;
        float acc[8][8] = {};
        float rs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int c = tid & 31, ml = tid >> 5;
        float wr[MAXO];
#pragma unroll
        for (int o = 0; o < MAXO; ++o) wr[o] = o < nout ? ldcg(ht.w[o] + n0 + c) : 0.f;
        for (int mp = 0; mp < B; mp += PK) {
          const int kp = min(PK, B - mp);
          if (mp > 0) __syncthreads();
          stage_nonkc(As, a.h2, H, n0, mp, kp);
          stage_nonkc(Bs, a.h1, H, k0, mp, kp);
          cp_commit_wait();
          __syncthreads();
#pragma unroll 4
          for (int m = ml; m < kp; m += NT / 32) {
            const float4 d0 = *reinterpret_cast<const float4*>(&dsm[(mp + m) * MAXO]);
            const float4 d1 = *reinterpret_cast<const float4*>(&dsm[(mp + m) * MAXO + 4]);
            float dot = d0.x * wr[0];
            dot = fmaf(d0.y, wr[1], dot); dot = fmaf(d0.z, wr[2], dot); dot = fmaf(d0.w, wr[3], dot);
            dot = fmaf(d1.x, wr[4], dot); dot = fmaf(d1.y, wr[5], dot); dot = fmaf(d1.z, wr[6], dot); dot = fmaf(d1.w, wr[7], dot);
            const float hvv = As[m * 36 + c];
            As[m * 36 + c] = hvv > 0.f ? dot : 0.f;
          }
          __syncthreads();
          tile_mma<false, false>(As, Bs, kp, acc, kt == 0 ? rs : nullptr);
        }
        float outv[4];
        tile_reduce<false, false>(acc, As, outv);
        float sq = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = tid + r * NT, m = e >> 5, n = e & 31;
          a.gW2[(size_t)(n0 + m) * H + k0 + n] = outv[r];
          sq = fmaf(outv[r], outv[r], sq);
        }
        if (kt == 0) {
          const int grp = tid >> 4, t = tid & 15, tx = t & 3, ty = t >> 2;
          if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) As[grp * 32 + ty * 8 + i] = rs[i];
          }
          __syncthreads();
          if (tid < 32) {
            float v = 0.f;
#pragma unroll
            for (int g = 0; g < KG; ++g) v += As[g * 32 + tid];
            a.gb2[n0 + tid] = v;
            sq = fmaf(v, v, sq);
          }
        }
        const float tot = block_sum0(sq, scr);
        if (tid == 0) cta_norm += tot;
      } else {
        // ---- head weight gradients, 32 columns per job; job 0 also the head bias gradients -------------
        const int jt = job - nJB - nJA;
        const int j0 = jt * 32;
        const int c = tid & 31, ml = tid >> 5;
        float hacc[MAXO];
#pragma unroll
        for (int o = 0; o < MAXO; ++o) hacc[o] = 0.f;
        for (int mp = 0; mp < B; mp += PK) {
          const int kp = min(PK, B - mp);
          if (mp > 0) __syncthreads();
          stage_nonkc(Bs, a.h2, H, j0, mp, kp);
          cp_commit_wait();
          __syncthreads();
#pragma unroll 4
          for (int m = ml; m < kp; m += NT / 32) {
            const float hvv = Bs[m * 36 + c];
            const float4 d0 = *reinterpret_cast<const float4*>(&dsm[(mp + m) * MAXO]);
            const float4 d1 = *reinterpret_cast<const float4*>(&dsm[(mp + m) * MAXO + 4]);
            hacc[0] = fmaf(d0.x, hvv, hacc[0]); hacc[1] = fmaf(d0.y, hvv, hacc[1]); hacc[2] = fmaf(d0.z, hvv, hacc[2]);
            hacc[3] = fmaf(d0.w, hvv, hacc[3]); hacc[4] = fmaf(d1.x, hvv, hacc[4]); hacc[5] = fmaf(d1.y, hvv, hacc[5]);
            hacc[6] = fmaf(d1.z, hvv, hacc[6]); hacc[7] = fmaf(d1.w, hvv, hacc[7]);
          }
        }
        float* red = As;               // [8][MAXO][32]
#pragma unroll
        for (int o = 0; o < MAXO; ++o) red[(ml * MAXO + o) * 32 + c] = hacc[o];
        __syncthreads();
        float sq = 0.f;
        if (ml == 0) {
#pragma unroll
          for (int o = 0; o < MAXO; ++o) {
            if (o < nout) {
              float t = red[(0 * MAXO + o) * 32 + c];
              for (int r = 1; r < NT / 32; ++r) t += red[(r * MAXO + o) * 32 + c];
              ht.gw[o][j0 + c] = t;
              sq = fmaf(t, t, sq);
            }
          }
        }
        if (jt == 0) {
          // bias gradients db[o] = sum_m dout[m][o]: 32 partial sums per output, folded in a fixed order
          __syncthreads();
          const int o = tid & 7, part = tid >> 3;
          float t = 0.f;
          for (int m = part; m < B; m += NT / 8) t += dsm[m * MAXO + o];
          red[part * MAXO + o] = t;
          __syncthreads();
          if (tid < MAXO) {
            float tt = 0.f;
            for (int q = 0; q < NT / 8; ++q) tt += red[q * MAXO + tid];
#pragma unroll
            for (int u = 0; u < MAXO; ++u) if (u == tid && u < nout) { *ht.gb[u] = tt; sq = fmaf(tt, tt, sq); }
          }
        }
        const float tot = block_sum0(sq, scr);
        if (tid == 0) cta_norm += tot;
      }
    }
    grid_bar(a.barrier, epoch, nctas);

    // =========================== P4: dW1 / db1, statistics, norm partials ==============================
    for (int job = cta; job < ((skip & 8) ? 0 : NTL); job += (int)nctas) {
      __syncthreads();
      const int j0 = job * 32, c = tid & 31, ml = tid >> 5;
      float wacc[MAXD + 1];
#pragma unroll
      for (int i = 0; i <= MAXD; ++i) wacc[i] = 0.f;
      for (int mp = 0; mp < B; mp += PK) {
        const int kp = min(PK, B - mp);
        if (mp > 0) __syncthreads();
        stage_nonkc(As, a.dh1, H, j0, mp, kp);
        for (int e = tid; e < kp * D; e += NT) { const int m = e / D, i = e - m * D; Bs[m * MAXD + i] = ldcg(a.xg + (size_t)mp * D + e); }
        cp_commit_wait();
        __syncthreads();
        for (int m = ml; m < kp; m += NT / 32) {
          const float dv = As[m * 36 + c];
#pragma unroll
          for (int i = 0; i < MAXD; ++i) if (i < D) wacc[i] = fmaf(dv, Bs[m * MAXD + i], wacc[i]);
          wacc[MAXD] += dv;
        }
      }
      float* red = dsm;                // [8][MAXD+1][32] = 4352 floats (dsm is dead after P3)
#pragma unroll
      for (int i = 0; i <= MAXD; ++i) red[(ml * (MAXD + 1) + i) * 32 + c] = wacc[i];
      __syncthreads();
      float sq = 0.f;
      if (ml == 0) {
#pragma unroll
        for (int i = 0; i <= MAXD; ++i) {
          if (i < D || i == MAXD) {
            float t = red[(0 * (MAXD + 1) + i) * 32 + c];
            for (int r = 1; r < NT / 32; ++r) t += red[(r * (MAXD + 1) + i) * 32 + c];
            if (i < D) a.gW1[(size_t)(j0 + c) * D + i] = t; else a.gb1[j0 + c] = t;
            sq = fmaf(t, t, sq);
          }
        }
      }
      const float tot = block_sum0(sq, scr);
      if (tid == 0) cta_norm += tot;
    }
    if (cta == (int)nctas - 1) {
      // learn()-level statistics of this minibatch (ppo.py:171-175), accumulated on the device
      float ssum = 0.f, esum = 0.f, mr = -INFINITY, mp = INFINITY;
      for (int b = tid; b < B; b += NT) {
        const float* rb = a.rowbuf + (size_t)b * ROWBUF;
        ssum += ldcg(rb + 20); esum += ldcg(rb + 21);
        mr = fmaxf(mr, ldcg(rb + 22)); mp = fminf(mp, ldcg(rb + 23));
      }
      const float t1 = block_sum0(ssum, scr), t2 = block_sum0(esum, scr + 8);
      mr = jb_warp_max(mr); mp = jb_warp_min(mp);
      __syncthreads();
      if (lane == 0) { scr[32 + warp] = mr; scr[48 + warp] = mp; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w) { mr = fmaxf(mr, scr[32 + w]); mp = fminf(mp, scr[48 + w]); }
        a.acc[0] += -t1 * invB;
        a.acc[1] += fmaxf(c1, c2);
        a.acc[2] += -t2 * invB / (a.continuous ? (float)A : 1.f);
        a.acc[3] = fmaxf(a.acc[3], mr);
        a.acc[4] = fminf(a.acc[4], mp);
        a.acc[5] += 1.f;
      }
    }
    if (tid == 0) { a.partials[cta] = cta_norm; cta_norm = 0.f; }
    grid_bar(a.barrier, epoch, nctas);

    // =========================== P5: clip + Adam on this CTA's slice ===================================
    if (!(skip & 16)) {
      double t = 0.0;
      for (int k = tid; k < (int)nctas; k += NT) t += (double)ldcg(a.partials + k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      double* dscr = reinterpret_cast<double*>(scr + 64);
      __syncthreads();
      if (lane == 0) dscr[warp] = t;
      __syncthreads();
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tot += dscr[w];
      const float total_norm = (float)sqrt(tot);
      float coef = 1.f;
      if (a.max_norm > 0.f) { coef = a.max_norm / (total_norm + 1e-6f); coef = coef < 1.f ? coef : 1.f; }
      if (tid == 0) {                                  // bias corrections once per CTA (double pow is ~200 instructions)
        const double tt = (double)(step0 + s + 1);
        const double bc1 = 1.0 - pow((double)a.beta1, tt), bc2 = 1.0 - pow((double)a.beta2, tt);
        scr[96] = (float)((double)lr / bc1);
        scr[97] = (float)sqrt(bc2);
      }
      __syncthreads();
      const float step_size = scr[96], bc2_sqrt = scr[97];
      const float one_m_b1 = 1.f - a.beta1, one_m_b2 = 1.f - a.beta2;
      auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= coef;
        mm = fmaf(gg - mm, one_m_b1, mm);
        vv = fmaf(one_m_b2 * gg, gg, a.beta2 * vv);
        const float denom = sqrtf(vv) / bc2_sqrt + a.adam_eps;
        pp = fmaf(-step_size, mm / denom, pp);
      };
      const long long per = (a.P4 + nctas - 1) / nctas;
      const long long lo = (long long)cta * per, hi = min(a.P4, lo + per);
      float4* p4 = reinterpret_cast<float4*>(a.flat);
      const float4* g4 = reinterpret_cast<const float4*>(a.grad);
      float4* m4 = reinterpret_cast<float4*>(a.am);
      float4* v4 = reinterpret_cast<float4*>(a.av);
      for (long long i = lo + tid; i < hi; i += NT) {
        float4 pp = __ldcg(p4 + i), mm = __ldcg(m4 + i), vv = __ldcg(v4 + i);
        const float4 gg = __ldcg(g4 + i);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
      }
    }
    grid_bar(a.barrier, epoch, nctas);
  }
  if (cta == 0 && tid == 0) { *a.step = step0 + a.n_steps; *a.cursor = cursor0 + a.n_steps; }
}

static int dsm_floats_for(int B) { const int need = B * MAXO; return need > 4352 ? need : 4352; }
static size_t fused_smem(int B) { return sizeof(float) * (size_t)(SMALL_FLOATS + dsm_floats_for(B) + 2 * PANEL_FLOATS); }

// Largest grid the cooperative launch can keep co-resident (one CTA per SM on B200) for minibatch size B.
static int fused_max_ctas(int B) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = fused_smem(B);
  if (smem > 227 * 1024) return 0;
  cudaFuncSetAttribute(ppo_epoch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppo_epoch_kernel, NT, smem);
  return sms * (per_sm > 0 ? 1 : 0);
}

}  // namespace

JB_API int jb_ppo_fused_args_size(void) { return (int)sizeof(jb_ppo_fused_args); }

JB_API int jb_ppo_fused_max_ctas(void) { return fused_max_ctas(256); }

// Runs args->n_steps minibatch steps starting at the device-side cursor.  `args` is a HOST pointer to
// a jb_ppo_fused_args (include/jorldy_b200_fused.h); it is copied at launch.
JB_API int jb_ppo_fused_run(const void* host_args, void* stream) {
  if (!host_args) return JB_ERR_INVALID;
  Args a = *reinterpret_cast<const Args*>(host_args);
  if (a.B <= 0 || a.B % 32 || a.H <= 0 || a.H % 32 || a.H > PK || a.D <= 0 || a.D > MAXD || a.nout <= 0 || a.nout > MAXO ||
      a.A <= 0 || a.A > jbppo::MAX_A || a.n_steps <= 0 || a.P4 <= 0)
    return JB_ERR_INVALID;
  int ctas = fused_max_ctas(a.B);
  if (ctas <= 0) return JB_ERR_INVALID;
  if (ctas > NT) ctas = NT;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(a.barrier, 0, sizeof(unsigned int), s) != cudaSuccess) return JB_ERR_CUDA;
  const size_t smem = fused_smem(a.B);
  int dsm_floats = dsm_floats_for(a.B);
  int skip = 0;
  if (const char* e = getenv("JB_FUSED_SKIP")) skip = atoi(e);     // timing experiments only
  void* kargs[] = {&a, &dsm_floats, &skip};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)ppo_epoch_kernel, dim3(ctas), dim3(NT), kargs, smem, s);
  if (e != cudaSuccess) { cudaGetLastError(); return JB_ERR_CUDA; }
  return JB_OK;
}
