// LMI family, latency-optimised path for samples that carry a kappa of the other families ("prior"): an exact
// definiteness FILTER in front of a one-warp-per-matrix eigen-solver.
//
// Why.  Behind the pruning bound of the linear/quadratic/SOC kernel the LMI work list is short (cfg5: 7 % of the batch)
// and almost all of it is false positives of that bound (cfg5: the LMI binds for 1 sample in 32768).  For such a sample
// the question is not "what is lambda_max(S~(u))" but "is lambda_max(S~(u)) < kappa_prior" -- and that is answered by
// an LDL' factorisation of  tau I - S~(u)  (all pivots positive <=> positive definite), r^3/3 flops and one broadcast
// per step, against the 4r^3/3 flops and two reductions + two broadcasts per step of a Householder tridiagonalisation
// followed by Sturm multisection (lmi.cuh).  Only the samples that FAIL the test (the LMI may bind) are solved.
//
// Layout (padded size 32 x 32; smaller LMIs are zero padded, which changes neither test nor solve: kappa = relu(.)):
// lane j of a warp owns COLUMN j of the matrix, all 32 rows in registers.  By symmetry column j is row j, so the plan
// section LMIW stores F~z_a row-major with a row stride of 36 words: lane j reads its column as eight 16-byte loads,
// and the 8 lanes of a quarter-warp hit 8 different bank groups (36 j mod 32 = 4 j): conflict-free LDS.128.
//   * Filter: a warp takes MT = 4 samples at once, register-tiled: every F~z word that comes out of shared memory
//     feeds 4 FMAs (4 x 32 accumulators per lane).  The 8-lanes-per-matrix layout of lmi.cuh re-reads F~z once per
//     matrix (n r^2 words each): its contraction was the shared-memory-bandwidth bound of the short-list launch.
//     The 4 LDL' factorisations are independent instruction streams in the same warp (ILP instead of occupancy).
//   * Solver: one warp per failing matrix: Householder with lane j = column j (row k of the trailing matrix is spread
//     over the lanes: no pivot-row bookkeeping), v / w broadcast through shared memory, Sturm multisection with 32
//     probes per round (33-section, 5 rounds), top eigenvector by twisted factorisation + back-transform through the
//     reflectors kept in shared memory, d kappa/du_a = q' F~z_a q with a transpose-reduce.  About half the latency of
//     the 8-lane layout per matrix; used where latency is what counts (few failing samples per warp).
// Results are those of the unfiltered path: a sample that passes keeps the prior (kappa, tag) -- exactly what the merge
// after a full solve would have kept -- and a sample that fails gets the full solve.
//
// This header is free of PTX up to the marker below so that tests/test_lmi_warp_emulated.py can compile the arithmetic
// for the host under the SIMT emulator (tests/emu) and check it against numpy before any GPU time is spent.
#pragma once
#ifndef RAYEN_EMU
#include "common.cuh"
#endif

namespace rayen {

#if defined(RAYEN_LW_TRACE) && !defined(RAYEN_EMU)
// development build only (scripts/lw_trace.py): phase time stamps (clock64) of the first 128 warps' latest chunk, of the
// latest full solve, and the start / end of every CTA (globaltimer)
__device__ long long g_lw_trace[8192];
#define LW_STAMP(id)                                                          \
  do {                                                                        \
    const int w_ = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);       \
    if ((threadIdx.x & 31) == 0 && w_ < 128) g_lw_trace[w_ * 16 + (id)] = clock64(); \
  } while (0)
#define LW_STAMP_SOLVE(id)                                                    \
  do {                                                                        \
    if ((threadIdx.x & 31) == 0) g_lw_trace[4096 + (id)] = clock64();         \
  } while (0)
#define LW_GT(id)                                                             \
  do {                                                                        \
    if (threadIdx.x == 0 && blockIdx.x < 256) {                               \
      long long t_;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                  \
      g_lw_trace[6144 + blockIdx.x * 2 + (id)] = t_;                          \
    }                                                                         \
  } while (0)
#else
#define LW_STAMP(id) do { } while (0)
#define LW_STAMP_SOLVE(id) do { } while (0)
#define LW_GT(id) do { } while (0)
#endif

constexpr int kLwR = 32;                            // padded matrix size of this path
constexpr int kLwRowStride = 36;                    // words per row of an LMIW matrix (see above)
constexpr int kLwMatWords = kLwR * kLwRowStride;    // 1152 words per F~z_a
constexpr int kLwMT = 4;                            // samples per warp in the filter
// per-warp scratch (floats): us[32][4] | cb[2][4][32] | vb[2][32] | wb[2][32] | sd se stau sz sdp sdm rr [32 each] |
//                            rb[2][32] (pivot rows of the tridiagonalisation) | red[2][32] (all-reduce buffers) |
//                            refl[32][32] (reflectors, WITH_GRAD)
constexpr int kLwScrUs = 0, kLwScrCb = 128, kLwScrVb = 384, kLwScrWb = 448, kLwScrD = 512, kLwScrE = 544,
              kLwScrTau = 576, kLwScrZ = 608, kLwScrDp = 640, kLwScrDm = 672, kLwScrRr = 704, kLwScrRb = 736,
              kLwScrRed = 800, kLwScrRefl = 864;
constexpr int kLwScratch = 864 + kLwR * kLwR;
constexpr float kLwFilterMargin = 1e-5f;  // tau = kappa_prior - margin (kappa_prior + |S~|_F): covers the float32 LDL'

__device__ __forceinline__ float lw_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
__device__ __forceinline__ float lw_max(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
  return x;
}
__device__ __forceinline__ float lw_min(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, off));
  return x;
}
__device__ __forceinline__ float4 lw_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Reciprocal / square root as ONE special-function instruction (MUFU.RCP / MUFU.SQRT, ~1 ulp).  The IEEE versions
// (1.0f / x, __frcp_rn, sqrtf) expand to a MUFU plus a Newton fix-up behind a branch to a slow path, and that branch
// serialises what would otherwise be independent instruction streams of one warp (measured: the four reciprocals of an
// LDL' step of four matrices cost ~400 of the step's ~700 cycles).  One ulp in a pivot reciprocal or a Householder tau
// is the size of the rounding these recurrences carry anyway.
#ifdef RAYEN_EMU
__device__ __forceinline__ float lw_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float lw_sqrt(float x) { return sqrtf(x); }
#else
__device__ __forceinline__ float lw_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lw_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#endif

// Sum over the warp through shared memory: one store, one barrier, eight broadcast 16-byte loads and an add tree --
// a shorter dependent chain than five shuffle + add rounds where the sum sits on the critical path of a serial
// recurrence (every step of the tridiagonalisation, every reflector of the back-transform).  `buf`: 32 floats that no
// lane still reads from an earlier call.
#ifndef LW_SMEM_REDUCE
#define LW_SMEM_REDUCE 1
#endif
// the read half of lw_sum_chain: the 32 addends are already in `buf` (stored before the last __syncwarp)
__device__ __forceinline__ float lw_sum_read(const float* __restrict__ buf) {
  float4 a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = lw_ld4(buf + 4 * c);
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c].x = (a[c].x + a[c].y) + (a[c].z + a[c].w);
  return ((a[0].x + a[1].x) + (a[2].x + a[3].x)) + ((a[4].x + a[5].x) + (a[6].x + a[7].x));
}
__device__ __forceinline__ float lw_sum_chain(float x, float* __restrict__ buf, int lane) {
#if LW_SMEM_REDUCE
  buf[lane] = x;
  __syncwarp();
  float4 a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = lw_ld4(buf + 4 * c);
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c].x = (a[c].x + a[c].y) + (a[c].z + a[c].w);
  return ((a[0].x + a[1].x) + (a[2].x + a[3].x)) + ((a[4].x + a[5].x) + (a[6].x + a[7].x));
#else
  return lw_sum(x);
#endif
}

// ----------------------------------------------------------------------------- filter: 4 samples per warp
template <int MT>
struct LwFilter {
  float A[MT][kLwR];  // A[m][i] = entry (i, lane) of sample m's matrix
  // 2-sample chunks keep a copy of the contracted matrices: the factorisation destroys A, and a sample that fails the
  // test is solved from the copy instead of being contracted again (3.8 k cycles on the kernel's critical path: the
  // short-list launch is one filter chunk + one solve long).  4-sample chunks have no registers to spare for that.
  float C[(MT == 2) ? 2 : 1][(MT == 2) ? kLwR : 1];
  __device__ __forceinline__ void keep_copy() {
    if constexpr (MT == 2) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int i = 0; i < kLwR; ++i) C[m][i] = A[m][i];
    }
  }

  // S~_m = sum_a u_m[a] F~z_a for the 4 samples of the chunk; us[a * 4 + m] = u_m[a]
  __device__ __forceinline__ void contract(const float* __restrict__ FW, int n, const float* __restrict__ us, int lane) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < kLwR; ++i) A[m][i] = 0.f;
    const float* Fj = FW + lane * kLwRowStride;
#pragma unroll 2
    for (int a = 0; a < n; ++a) {
      const float4 u4 = lw_ld4(us + 4 * a);
      const float um4[4] = {u4.x, u4.y, u4.z, u4.w};
      float um[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) um[m] = um4[m];
      const float* Fa = Fj + a * kLwMatWords;
#pragma unroll
      for (int c = 0; c < kLwR / 4; ++c) {
        const float4 f = lw_ld4(Fa + 4 * c);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          A[m][4 * c + 0] = fmaf(um[m], f.x, A[m][4 * c + 0]);
          A[m][4 * c + 1] = fmaf(um[m], f.y, A[m][4 * c + 1]);
          A[m][4 * c + 2] = fmaf(um[m], f.z, A[m][4 * c + 2]);
          A[m][4 * c + 3] = fmaf(um[m], f.w, A[m][4 * c + 3]);
        }
      }
    }
  }

  // Steps k = 8 S .. 8 S + 7 of the right-looking LDL' of the 4 matrices.  The step index is a RUNTIME loop variable
  // (a fully unrolled factorisation is ~3300 straight-line instructions that every warp executes once: measured at 7.5
  // cycles per instruction, instruction-fetch bound, against 1.7 in a loop that stays in the instruction cache), so
  // nothing may index the register file with it: the pivot row k sits in shared memory, where the previous step put
  // it with predicated stores out of the unrolled row loop, and rows >= 8 S are updated whether or not they are still
  // live (rows <= k only collect rounding noise that nothing reads).
  template <int S>
  __device__ __forceinline__ void ldlt_stage(float* __restrict__ cb, int lane, bool (&ok)[MT]) {
    constexpr int R0 = 8 * S;
#pragma unroll 1
    for (int k = R0; k < R0 + 8; ++k) {
      const float* buf = cb + (k & 1) * (4 * kLwR);        // row k of the trailing matrices, spread over the lanes
      float* nbuf = cb + ((k + 1) & 1) * (4 * kLwR);
      __syncwarp();
      // all loads of the step first, all stores last: a store to nbuf inside the per-matrix loop would order the next
      // matrix's loads behind it (same base pointer) and serialise the four independent factorisations
      float my[MT], nxt[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float d = buf[m * kLwR + k];
        ok[m] = ok[m] && (d > 0.f);  // NaN fails
        my[m] = buf[m * kLwR + lane] * lw_rcp(d);  // M[k][lane] / d
        nxt[m] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < (kLwR - R0) / 4; ++c) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 x = lw_ld4(buf + m * kLwR + R0 + 4 * c);
          const float cr[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = R0 + 4 * c + ii;
            const float nv = fmaf(-cr[ii], my[m], A[m][i]);
            A[m][i] = nv;
            // row k + 1 becomes the next pivot row (picked up as a value, straight out of the update: a separate pass
            // over A[m][k + 1] makes the compiler index the array dynamically and puts all of A in local memory)
            if (i > R0 && i <= R0 + 8 && i == k + 1) nxt[m] = nv;
          }
        }
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) nbuf[m * kLwR + lane] = nxt[m];
    }
  }

  // Is lambda_max(S~_m) < kprior[m] beyond doubt?  LDL' of  M = tau_m I - S~_m,  tau_m = kprior_m - margin: every
  // pivot positive <=> M positive definite <=> lambda_max(S~_m) < tau_m.  The float32 factorisation is the exact one of
  // M + E with |E| <= c r eps |M| (c r eps ~ 2e-6 at r = 32, far less in practice) and |M|_2 <= tau + |S~|_F, so with
  // margin = 1e-5 (kprior + |S~|_F) a pass proves lambda_max(S~_m) < kprior_m; a failure proves nothing and costs a
  // full solve.  Returns the 4-bit pass mask (identical in every lane).  Destroys A.
  __device__ __forceinline__ unsigned passes(const float (&kprior)[MT], float* __restrict__ cb, int lane) {
    bool ok[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float f2 = 0.f;
#pragma unroll
      for (int i = 0; i < kLwR; ++i) f2 = fmaf(A[m][i], A[m][i], f2);
      const float fro = lw_sqrt(lw_sum(f2));
      const float tau = kprior[m] - kLwFilterMargin * (kprior[m] + fro);
#pragma unroll
      for (int i = 0; i < kLwR; ++i) A[m][i] = ((i == lane) ? tau : 0.f) - A[m][i];
      ok[m] = true;
      cb[m * kLwR + lane] = A[m][0];
    }
    ldlt_stage<0>(cb, lane, ok);
    ldlt_stage<1>(cb, lane, ok);
    ldlt_stage<2>(cb, lane, ok);
    ldlt_stage<3>(cb, lane, ok);
    unsigned mask = 0u;
#pragma unroll
    for (int m = 0; m < MT; ++m) mask |= ok[m] ? (1u << m) : 0u;
    return mask;
  }
};

// ----------------------------------------------------------------------------- solver: one warp per matrix
template <bool WITH_GRAD>
struct LwSolver {
  float W[kLwR];  // W[i] = entry (i, lane)
  float* scr;     // the warp's scratch
  int lane;

  // same arithmetic, operand order included, as LwFilter::contract for sample m: bit-identical S~
  __device__ __forceinline__ void contract_one(const float* __restrict__ FW, int n, const float* __restrict__ us, int m) {
#pragma unroll
    for (int i = 0; i < kLwR; ++i) W[i] = 0.f;
    const float* Fj = FW + lane * kLwRowStride;
#pragma unroll 4
    for (int a = 0; a < n; ++a) {
      const float ua = us[4 * a + m];
      const float* Fa = Fj + a * kLwMatWords;
#pragma unroll
      for (int c = 0; c < kLwR / 4; ++c) {
        const float4 f = lw_ld4(Fa + 4 * c);
        W[4 * c + 0] = fmaf(ua, f.x, W[4 * c + 0]);
        W[4 * c + 1] = fmaf(ua, f.y, W[4 * c + 1]);
        W[4 * c + 2] = fmaf(ua, f.z, W[4 * c + 2]);
        W[4 * c + 3] = fmaf(ua, f.w, W[4 * c + 3]);
      }
    }
  }

  // Steps k = 8 S .. of the Householder tridiagonalisation as a runtime loop (see LwFilter::ldlt_stage for why): row k
  // -- by symmetry column k of the trailing matrix, spread over the lanes -- comes from shared memory, where the
  // previous step left it; reflector k goes to shared memory (WITH_GRAD); rows >= 8 S are updated, the dead ones among
  // them with v_i = w_i = 0, i.e. not at all.
  template <int S>
  __device__ __forceinline__ void householder_stage() {
    constexpr int R0 = 8 * S;
    constexpr int K_END = (R0 + 8 < kLwR - 2) ? R0 + 8 : kLwR - 2;
    float* sd = scr + kLwScrD;
    float* se = scr + kLwScrE;
#pragma unroll 1
    for (int k = R0; k < K_END; ++k) {
      const float* rb = scr + kLwScrRb + (k & 1) * kLwR;
      float* nrb = scr + kLwScrRb + ((k + 1) & 1) * kLwR;
      float* vb = scr + kLwScrVb + (k & 1) * kLwR;
      float* wb = scr + kLwScrWb + (k & 1) * kLwR;
      __syncwarp();
      const float xk = rb[lane];  // entry (k, lane) = (lane, k)
      const float xk1 = rb[k + 1];
      // the squares of this row's tail were stored by whoever wrote the row (the previous step's update, or
      // tridiagonalize() for row 0): one store + barrier less on the step's dependent chain
      const float tail2 = lw_sum_read(scr + kLwScrRed);
      const float sigma = fmaf(xk1, xk1, tail2);
      const float rt = lw_sqrt(sigma);
      const float alpha = (xk1 >= 0.f) ? -rt : rt;
      const bool skip = !(tail2 > 0.f);  // column already tridiagonal (also covers zero padding)
      const float tau = skip ? 0.f : lw_rcp(fmaf(fabsf(xk1), rt, sigma));  // 2 / v'v
      if (lane == k) sd[k] = xk;
      if (lane == k + 1) {
        se[k] = skip ? xk1 : alpha;
        if constexpr (WITH_GRAD) scr[kLwScrTau + k] = tau;
      }
      float vj = (lane > k + 1) ? xk : ((lane == k + 1) ? xk1 - alpha : 0.f);
      if (skip) vj = 0.f;
      if constexpr (WITH_GRAD) scr[kLwScrRefl + k * kLwR + lane] = vj;
      vb[lane] = vj;
      __syncwarp();
      float vr[kLwR - R0];
#pragma unroll
      for (int c = 0; c < (kLwR - R0) / 4; ++c) {
        const float4 x = lw_ld4(vb + R0 + 4 * c);
        vr[4 * c + 0] = x.x;
        vr[4 * c + 1] = x.y;
        vr[4 * c + 2] = x.z;
        vr[4 * c + 3] = x.w;
      }
      // p = tau S v (entries of v at or above row k are zero; four partial sums shorten the dependent chain)
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = R0; i < kLwR; ++i) ps[i & 3] = fmaf(W[i], vr[i - R0], ps[i & 3]);
      const float pj = tau * ((ps[0] + ps[1]) + (ps[2] + ps[3]));
      const float Kc = 0.5f * tau * lw_sum_chain(vj * pj, scr + kLwScrRed + kLwR, lane);
      const float wj = (lane > k) ? fmaf(-Kc, vj, pj) : 0.f;
      wb[lane] = wj;
      __syncwarp();
      // S <- S - v w' - w v'
#pragma unroll
      for (int c = 0; c < (kLwR - R0) / 4; ++c) {
        const float4 x = lw_ld4(wb + R0 + 4 * c);
        const float wr[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = R0 + 4 * c + ii;
          const float nv = fmaf(-vr[4 * c + ii], wj, fmaf(-wr[ii], vj, W[i]));
          W[i] = nv;
          if (i > R0 && i <= R0 + 8 && i == k + 1) {  // row k + 1 becomes the next pivot row
            nrb[lane] = nv;
            scr[kLwScrRed + lane] = (lane > k + 2) ? nv * nv : 0.f;  // ... and its tail's squares go with it
          }
        }
      }
    }
  }

  // Householder tridiagonalisation: diagonal -> sd, sub-diagonal -> se (se[31] = 0), tau_k = 2 / |v_k|^2 -> stau
  __device__ __forceinline__ void tridiagonalize() {
    scr[kLwScrRb + lane] = W[0];
    scr[kLwScrRed + lane] = (lane > 1) ? W[0] * W[0] : 0.f;
    householder_stage<0>();
    householder_stage<1>();
    householder_stage<2>();
    householder_stage<3>();
    float* sd = scr + kLwScrD;
    float* se = scr + kLwScrE;
    if (lane == kLwR - 2) {
      sd[kLwR - 2] = W[kLwR - 2];
      se[kLwR - 2] = W[kLwR - 1];
    }
    if (lane == kLwR - 1) {
      sd[kLwR - 1] = W[kLwR - 1];
      se[kLwR - 1] = 0.f;
    }
    __syncwarp();
  }

  // Largest eigenvalue of the tridiagonal matrix, clipped at 0 from below (kappa = relu(lambda_max)): 32 probes per
  // round, each deciding "x above the whole spectrum?" with the positive-product Sturm sequence of lmi.cuh; 5 rounds
  // shrink the Gershgorin interval by 32 * 33^4 = 3.8e7.
  __device__ __forceinline__ float lambda_max_relu() {
    const float* sd = scr + kLwScrD;
    const float* se = scr + kLwScrE;
    float d[kLwR], e2[kLwR];
    {
      const float di = sd[lane];
      const float rad = fabsf(lane > 0 ? se[lane - 1] : 0.f) + fabsf(se[lane]);
      const float dmax = lw_max(di);
      float hi = lw_max(di + rad);
      const float lo_g = lw_min(di - rad);
      const float scale = fmaxf(fmaxf(fabsf(hi), fabsf(lo_g)), 1e-30f);
      const float inv_scale = 1.0f / scale;
      float eprev = 0.f;
#pragma unroll
      for (int c = 0; c < kLwR / 4; ++c) {
        const float4 dv = lw_ld4(sd + 4 * c), ev = lw_ld4(se + 4 * c);
        const float da[4] = {dv.x, dv.y, dv.z, dv.w}, ea[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          d[4 * c + ii] = da[ii] * inv_scale;
          e2[4 * c + ii] = (eprev * inv_scale) * (eprev * inv_scale);  // e2[i] couples rows i-1 and i
          eprev = ea[ii];
        }
      }
      float lo = fmaxf(dmax, 0.f) * inv_scale;
      hi = fmaf(1e-6f, scale, hi) * inv_scale;
      hi = fmaxf(hi, lo);  // whole spectrum <= 0: degenerate interval, the rounds return lo = 0
      for (int round = 0; round < 5; ++round) {
        // the first round also probes x = lo itself (32 sections): lo = 0 above the whole spectrum means
        // lambda_max < 0, and the answer is then exactly 0 (relu), not the midpoint of a tiny interval above it
        const int shift = round ? 1 : 0;
        const float h = (hi - lo) / static_cast<float>(32 + shift);
        const float x = fmaf(h, static_cast<float>(lane + shift), lo);
        float s0 = 1.f, s1 = x - d[0];
        float mn = s1;
#pragma unroll
        for (int i = 1; i < kLwR; ++i) {
          const float sn = fmaf(x - d[i], s1, -e2[i] * s0);
          mn = fminf(mn, sn);
          s0 = s1;
          s1 = sn;
          // every 4 steps: (x - d_i) can be ~1e-7 for several rows in a row (zero padding, clustered eigenvalues),
          // and 8 such factors underflow
          if ((i & 3) == 3 && i + 1 < kLwR) {
            const int ex = (__float_as_int(fmaxf(fabsf(s0), fabsf(s1))) >> 23) & 0xff;
            const float sc = __int_as_float((254 - max(min(ex, 253), 1)) << 23);
            s0 *= sc;
            s1 *= sc;
          }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, mn > 0.f);
        const int first = mask ? (__ffs(mask) - 1) : 32;  // first probe above the spectrum
        const float new_lo = (first + shift == 0) ? lo : fmaf(h, static_cast<float>(first + shift - 1), lo);
        const float new_hi = (first == 32) ? hi : fmaf(h, static_cast<float>(first + shift), lo);
        lo = new_lo;
        hi = new_hi;
      }
      // below 1e-7 of the matrix scale lambda_max is 0 to the resolution of a float32 matrix (a negative semi-definite
      // S~, e.g. a zero-padded negative definite one): exactly 0 then, the same answer pruning gives
      const float mid = 0.5f * (lo + hi);
      return (mid > 1e-7f) ? mid * scale : 0.f;
    }
  }

  // Unit eigenvector of lambda (twisted factorisation of T - lambda I, then q = H_0 ... H_29 z through the reflectors
  // in the scratch); returns this lane's entry and leaves q in scr[kLwScrZ ..].
  __device__ __forceinline__ float eigenvector(float lam) {
    static_assert(WITH_GRAD, "eigenvector needs the reflectors");
    const float* d = scr + kLwScrD;
    const float* e = scr + kLwScrE;
    float* dp = scr + kLwScrDp;
    float* dm = scr + kLwScrDm;
    float* z = scr + kLwScrZ;
    float* rr = scr + kLwScrRr;
    const float di = d[lane], ei = e[lane];
    const float scale = lw_max(fabsf(di) + fabsf(ei));
    const float tiny = fmaxf(1e-12f * scale, 1e-30f);
    // the two pivot recurrences run on lanes 0 and 1 at the same time (same instruction stream, direction per lane)
    if (lane <= 1) {
      const bool fw = lane == 0;
      float* out = fw ? dp : dm;
      const int step = fw ? 1 : -1;
      int i = fw ? 0 : kLwR - 1;
      float piv = d[i] - lam;
#pragma unroll 8
      for (int t = 0; t < kLwR; ++t) {
        if (t > 0) {
          const float ee = fw ? e[i - 1] : e[i];
          piv = fmaf(-ee * ee, lw_rcp(piv), d[i] - lam);
        }
        if (fabsf(piv) < tiny) piv = -tiny;
        out[i] = piv;
        i += step;
      }
    }
    __syncwarp();
    // twist index: argmin_i |dp_i + dm_i - (d_i - lam)|, lowest index on ties
    float best = fabsf(dp[lane] + dm[lane] - (di - lam));
    int kt = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ok = __shfl_xor_sync(0xffffffffu, kt, off);
      if (ob < best || (ob == best && ok < kt)) {
        best = ob;
        kt = ok;
      }
    }
    // z_{i-1} = -e_{i-1}/dp_{i-1} z_i below the twist, z_{i+1} = -e_i/dm_{i+1} z_i above it
    {
      float r = 0.f;
      if (lane < kt) r = -ei * lw_rcp(dp[lane]);
      else if (lane > kt) r = -e[lane - 1] * lw_rcp(dm[lane]);
      rr[lane] = r;
    }
    __syncwarp();
    if (lane <= 1) {
      const int step = (lane == 0) ? -1 : 1;
      float zi = 1.f;
      int i = kt + step;
      for (int t = 1; t < kLwR; ++t) {
        if (i >= 0 && i < kLwR) {
          zi *= rr[i];
          z[i] = zi;
        }
        i += step;
      }
    }
    if (lane == 2) z[kt] = 1.f;
    __syncwarp();
    const float zi = z[lane];
    float qj = zi * rsqrtf(lw_sum(zi * zi));
    // q = H_0 ... H_29 z; reflector k was left in shared memory by the tridiagonalisation
#pragma unroll 1
    for (int k = kLwR - 3; k >= 0; --k) {
      const float r = scr[kLwScrRefl + k * kLwR + lane];
      // tau_k = 2 / v_k'v_k (0 for a skipped step)
      const float c = scr[kLwScrTau + k] * lw_sum_chain(r * qj, scr + kLwScrRed + (k & 1) * kLwR, lane);
      qj = fmaf(-c, r, qj);
    }
    __syncwarp();  // every lane has read z
    z[lane] = qj;
    __syncwarp();
    return qj;
  }

  // d kappa/du_a = q' F~z_a q for a < n; the value of a = lane is returned in every lane < n (0 beyond)
  __device__ __forceinline__ float eig_gradient(const float* __restrict__ FW, int n, float qj) {
    const float* z = scr + kLwScrZ;
    float qr[kLwR];
#pragma unroll
    for (int c = 0; c < kLwR / 4; ++c) {
      const float4 x = lw_ld4(z + 4 * c);
      qr[4 * c + 0] = x.x;
      qr[4 * c + 1] = x.y;
      qr[4 * c + 2] = x.z;
      qr[4 * c + 3] = x.w;
    }
    const float* Fj = FW + lane * kLwRowStride;
    float mine = 0.f;
    // four matrices per trip: their four warp sums overlap
#pragma unroll 1
    for (int a0 = 0; a0 < n; a0 += 4) {
      float part[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int a = (a0 + t < n) ? a0 + t : n - 1;
        const float* Fa = Fj + a * kLwMatWords;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < kLwR / 4; ++c) {
          const float4 f = lw_ld4(Fa + 4 * c);
          acc[0] = fmaf(f.x, qr[4 * c + 0], acc[0]);
          acc[1] = fmaf(f.y, qr[4 * c + 1], acc[1]);
          acc[2] = fmaf(f.z, qr[4 * c + 2], acc[2]);
          acc[3] = fmaf(f.w, qr[4 * c + 3], acc[3]);
        }
        part[t] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * qj;  // column `lane` of q' F~z_a q
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int t = 0; t < 4; ++t) part[t] += __shfl_xor_sync(0xffffffffu, part[t], off);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane == a0 + t) mine = part[t];
    }
    return mine;
  }
};

// value of x[m] for a runtime m < 4 without dynamic register indexing
template <class T, int MT>
__device__ __forceinline__ T lw_pick(const T (&x)[MT], int m) {
  T r = x[0];
#pragma unroll
  for (int i = 1; i < MT; ++i)
    if (m == i) r = x[i];
  return r;
}

// What a chunk needs besides the sample data (plain pointers: also filled in by the emulator harness)
struct LwCtx {
  const float* FW;       // staged LMIW matrices (shared memory on the GPU)
  float* scr;            // this warp's scratch, kLwScratch floats
  const float* y0;       // [k]
  const float* nmat;     // N rows, stride nstride (unused when n_is_identity)
  int n, k, nstride, n_is_identity, mode;
};

// y = y0 + alpha N u for one sample (u of sample m in the scratch as us[a * 4 + m]); lanes over the ambient coordinates
__device__ __forceinline__ void lw_write_y(const LwCtx& C, float* __restrict__ yrow, float alpha, int m, int lane) {
  const float* us = C.scr + kLwScrUs;
  for (int i = lane; i < C.k; i += 32) {
    float rho;
    if (C.n_is_identity) {
      rho = us[4 * i + m];
    } else {
      rho = 0.f;
      const float* nrow = C.nmat + i * C.nstride;
      for (int a = 0; a < C.n; ++a) rho = fmaf(__ldg(nrow + a), us[4 * a + m], rho);
    }
    yrow[i] = fmaf(alpha, rho, C.y0[i]);
  }
}

// One chunk of up to 4 samples: sample ids b[m] (valid[m]), whose kappa_io / active_io hold the prior of the other
// families.  Writes y for all of them, kappa / active for those whose LMI binds, d kappa/du (WITH_GRAD) for those that
// need it in backward.
// `solve_budget` (in/out): how many failing samples this warp may still solve itself; the others are appended to
// fail_list (fail_count: running counter in global memory) for the launch behind this one.
template <bool WITH_GRAD, int MT = kLwMT>
__device__ __forceinline__ void lw_process_chunk(const LwCtx& C, const long long (&b)[MT], const bool (&valid)[MT],
                                                 const float* __restrict__ v, long long ldv, float* __restrict__ y,
                                                 float* __restrict__ kappa_io, int* __restrict__ active_io,
                                                 float* __restrict__ dkappa, int lane, bool use_filter,
                                                 int& solve_budget, int* __restrict__ fail_list,
                                                 int* __restrict__ fail_count) {
  float* us = C.scr + kLwScrUs;
  const int n = C.n;
  float kprior[MT], s[MT], beta[MT];
  int tprior[MT];
  LW_STAMP(0);
  {
    float x[MT], ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      x[m] = (valid[m] && lane < n) ? __ldg(v + b[m] * ldv + lane) : 0.f;
      kprior[m] = valid[m] ? kappa_io[b[m]] : 0.f;
      tprior[m] = valid[m] ? active_io[b[m]] : 0;
      beta[m] = (valid[m] && C.mode == RAYEN_MODE_RAYEN_OLD) ? __ldg(v + b[m] * ldv + n) : 0.f;
      ss[m] = x[m] * x[m];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int m = 0; m < MT; ++m) ss[m] += __shfl_xor_sync(0xffffffffu, ss[m], off);
    float4 u4;
    float um[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      s[m] = sqrtf(ss[m]);
      um[m] = x[m] * (1.0f / fmaxf(s[m], 1e-12f));
    }
    u4.x = um[0]; u4.y = um[1]; u4.z = um[2]; u4.w = um[3];
    __syncwarp();  // the previous chunk's readers of the scratch are done
    *reinterpret_cast<float4*>(us + 4 * lane) = u4;
    __syncwarp();
  }

  LW_STAMP(1);
  unsigned pass = 0u;
  LwFilter<MT> F;
  if (use_filter) {
    F.contract(C.FW, n, us, lane);
    F.keep_copy();
    LW_STAMP(2);
    pass = F.passes(kprior, C.scr + kLwScrCb, lane);
  }
  LW_STAMP(3);
  // ---- samples the filter settled: the prior (kappa, tag) stands; scale step (reference :472-474 / :464-465, :512-514)
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    if (valid[m] && ((pass >> m) & 1u)) {
      const float alpha = (C.mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(beta[m]) + kprior[m])
                                                          : fminf(1.0f / kprior[m], s[m]);
      lw_write_y(C, y + b[m] * C.k, alpha, m, lane);
    }
  }
  LW_STAMP(4);
  // ---- the others: this warp solves as many as its budget allows (the whole warp on each, one after the other: the
  // low-latency path for the usual case of a rare failure); what is left goes to the fail list
  unsigned todo = 0u;
#pragma unroll
  for (int m = 0; m < MT; ++m) todo |= (valid[m] && !((pass >> m) & 1u)) ? (1u << m) : 0u;
  {
    unsigned hand_over = 0u;
    int keep = solve_budget;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if ((todo >> m) & 1u) {
        if (keep > 0) --keep;
        else hand_over |= 1u << m;
      }
    }
    solve_budget = keep;
    if (hand_over != 0u) {
      int slot = 0;
      if (lane == 0) slot = atomicAdd(fail_count, __popc(hand_over));
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (lane < MT && ((hand_over >> lane) & 1u))
        fail_list[slot + __popc(hand_over & ((1u << lane) - 1u))] = static_cast<int>(lw_pick(b, lane));
      todo &= ~hand_over;
    }
  }
#pragma unroll 1
  for (int m = 0; m < MT; ++m) {
    if (!((todo >> m) & 1u)) continue;  // warp-uniform
    const long long bm = lw_pick(b, m);
    const float k0 = lw_pick(kprior, m), sm = lw_pick(s, m), bt = lw_pick(beta, m);
    const int t0 = lw_pick(tprior, m);
    LwSolver<WITH_GRAD> S;
    S.scr = C.scr;
    S.lane = lane;
    LW_STAMP_SOLVE(0);
    if constexpr (MT == 2) {
      if (use_filter) {
        // the filter's own contraction of this sample (same arithmetic as contract_one: bit-identical)
#pragma unroll
        for (int i = 0; i < kLwR; ++i) S.W[i] = (m == 0) ? F.C[0][i] : F.C[1][i];
      } else {
        S.contract_one(C.FW, n, us, m);
      }
    } else {
      S.contract_one(C.FW, n, us, m);
    }
    LW_STAMP_SOLVE(1);
    S.tridiagonalize();
    LW_STAMP_SOLVE(2);
    const float lam = S.lambda_max_relu();
    LW_STAMP_SOLVE(3);
    float kap = fmaxf(lam, 0.f);
    int tag = kap > 0.f ? make_tag(RAYEN_FAM_LMI, 0) : make_tag(RAYEN_FAM_NONE, 0);
    if (!(kap > k0)) {
      kap = k0;
      tag = t0;
    }
    if (lane == 0) {
      kappa_io[bm] = kap;
      active_io[bm] = tag;
    }
    const float alpha = (C.mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(bt) + kap) : fminf(1.0f / kap, sm);
    lw_write_y(C, y + bm * C.k, alpha, m, lane);
    if constexpr (WITH_GRAD) {
      bool need = tag_family(tag) == RAYEN_FAM_LMI && kap > 0.f;
      if (need && C.mode == RAYEN_MODE_RAYEN) need = (1.0f / kap < sm);
      LW_STAMP_SOLVE(4);
      if (need) {
        const float qj = S.eigenvector(lam);
        LW_STAMP_SOLVE(5);
        const float g = S.eig_gradient(C.FW, n, qj);
        if (lane < n) dkappa[bm * n + lane] = g;
        LW_STAMP_SOLVE(6);
      }
    }
    __syncwarp();
  }
  LW_STAMP(5);
}

#ifndef RAYEN_EMU
// ============================================================================= PTX below: the kernel around the chunks
constexpr int kLwThreads = 256;
template <int V>
struct LwTag {
  static constexpr int value = V;
};

__host__ __device__ constexpr size_t lmi_warp_smem_bytes(int n, int threads) {
  return 192 + static_cast<size_t>(n) * kLwMatWords * 4 + static_cast<size_t>(threads / 32) * kLwScratch * 4;
}

// Samples come from the work list of the linear/quadratic/SOC kernel (list mode) or are the whole batch (dense mode,
// pruning off); in both cases kappa_io / active_io hold the prior.  Chunk c of 4 samples goes to CTA c % gridDim,
// warp c / gridDim, so that a short list spreads over all SMs and their four schedulers.  A warp solves at most
// `solves_per_warp` failing samples itself (latency: the usual case is a handful of failures in the whole batch); the
// rest is left in fail_list for a second launch of this kernel (filter off, unlimited budget) right behind this one, so
// that every solve is the same arithmetic whoever runs it.
template <bool WITH_GRAD>
__global__ void __launch_bounds__(kLwThreads, 1)
    lmi_forward_warp_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                            float* __restrict__ kappa_io, int* __restrict__ active_io, long long B, int mode,
                            const int* work_list, const int* work_count,  // (written by the kernel before: not __restrict__)
                            float* __restrict__ dkappa, int use_filter, int solves_per_warp,
                            int* __restrict__ fail_list, int* __restrict__ fail_count, int* __restrict__ next_chunk) {
  pdl_launch_dependents();  // the fail list's consumer may be scheduled; it waits before it reads
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* y0s = reinterpret_cast<float*>(smem_raw + 64);   // y0 (k <= 32 words): read by every sample's scale step
  float* fw = reinterpret_cast<float*>(smem_raw + 192);
  const int fw_words = P.n * kLwMatWords;
  LW_GT(0);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) y0s[threadIdx.x] = (static_cast<int>(threadIdx.x) < P.k) ? P.blob[P.off_y0 + threadIdx.x] : 0.f;
  __syncthreads();
  pdl_wait();  // launched behind the linear/quadratic/SOC kernel: its kappa / active / work list must be complete
  const long long total = work_list ? static_cast<long long>(ld_after_wait(work_count)) : B;
  // with the filter a warp takes 4 samples at a time (register tiling of the contraction: every F~z word feeds 4 FMAs)
  // -- or 2 when the list is short enough for every warp of the grid to get at most one chunk of 2: a chunk is one
  // warp's dependent chain, and with one busy warp per scheduler nothing hides its latencies; twice as many warps on
  // half the work each finish sooner (cfg5: 2115 samples on 1184 warps).  Without the filter the samples are solved one
  // by one anyway, so one sample per warp spreads a short list over more warps
  const long long n_warps_total = static_cast<long long>(gridDim.x) * (kLwThreads / 32);
  const int per = use_filter ? ((total <= 2 * n_warps_total) ? 2 : kLwMT) : 1;
  const long long n_chunks = (total + per - 1) / per;
  const bool cta_has_work = static_cast<long long>(blockIdx.x) < n_chunks;
  if (threadIdx.x == 0 && cta_has_work) stage_bulk(fw, P.blob + P.off_lmiw, fw_words, &bars[0]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LW_STAMP(8);
  LwCtx C;
  C.FW = fw;
  C.scr = fw + fw_words + warp * kLwScratch;
  C.y0 = (P.k <= 32) ? y0s : P.blob + P.off_y0;  // (more ambient coordinates than that: from L1/L2)
  C.nmat = P.blob + P.off_nmat;
  C.n = P.n; C.k = P.k; C.nstride = P.np + 4; C.n_is_identity = P.n_is_identity; C.mode = mode;
  bool staged = false;
  int solve_budget = solves_per_warp;
  // every warp starts with the chunk of its own number; further chunks are handed out by a counter (next_chunk, zeroed
  // before the launch), so that a warp that had eigen-solves to do takes fewer chunks: with a static stride the kernel
  // waits for the unluckiest warp (cfg5 "loose", 13 % of the samples LMI-bound: 0.48 instead of 0.30 ms)
  long long c = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  auto run = [&](auto mt_tag) {
    constexpr int MT = decltype(mt_tag)::value;
    while (c < n_chunks) {
      long long b[MT];
      bool valid[MT];
      {
        const long long idx = c * per + (lane & 3);
        const bool ok = (lane & 3) < per && idx < total;
        const int mine = ok ? (work_list ? work_list[idx] : static_cast<int>(idx)) : 0;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          b[m] = __shfl_sync(0xffffffffu, mine, m);
          valid[m] = m < per && c * per + m < total;
        }
      }
      LW_STAMP(6);
      if (!staged) {
        mbar_wait(&bars[0], 0);
        staged = true;
      }
      LW_STAMP(7);
      lw_process_chunk<WITH_GRAD, MT>(C, b, valid, v, ldv, y, kappa_io, active_io, dkappa, lane, use_filter != 0, solve_budget,
                                      fail_list, fail_count);
      if (next_chunk) {
        int t = 0;
        if (lane == 0) t = atomicAdd(next_chunk, 1);
        c = n_warps_total + __shfl_sync(0xffffffffu, t, 0);
      } else {
        c += n_warps_total;
      }
    }
  };
  if (per == 2) run(LwTag<2>{});
  else run(LwTag<kLwMT>{});
  if (!staged && cta_has_work) mbar_wait(&bars[0], 0);  // never exit with a bulk copy in flight
#if defined(RAYEN_LW_TRACE)
  __syncthreads();
  LW_GT(1);
#endif
}
#endif  // RAYEN_EMU

}  // namespace rayen
