// LMI family: kappa = relu(lambda_max(sum_a u_a F~z_a)), fused with the shift-and-scale step.
//
// The reference materialises S = einsum(F, rho), L'(-S)L and calls eigvalsh on [B,r,r]
// (constraint_module.py:401-449).  Here the congruence is folded into the constants on the host and
// one kernel does, per sample and without touching HBM in between:
//   1. the contraction S~ = sum_a u_a F~z_a straight into registers,
//   2. a Householder tridiagonalisation of S~ in registers,
//   3. lambda_max of the tridiagonal matrix by parallel multisection on Sturm counts,
//   4. (backward only) the eigenvector: twisted factorisation + back-transform through the
//      reflectors, then d kappa/du_a = q' F~z_a q,
//   5. the merge with the kappa of the other families and the scale step / the closed-form g_v.
//
// Layout: a matrix of padded size RP (4, 8, 16, 32) is owned by LPM = RP/4 consecutive lanes; lane q
// keeps columns q, q+LPM, q+2LPM, q+3LPM (all RP rows) in registers: 4*RP floats.  Owning 4 columns
// instead of 1 cuts the broadcast traffic of the two rank-1 vectors per Householder step (the LSU
// bound of a column-per-lane layout) by 4x and gives every lane 4 independent FMA chains; the
// interleaved ownership lets whole column slots drop out at compile time as the reduction proceeds.
// A warp carries 32/LPM matrices.  Vectors that every lane of a matrix needs go through a small
// per-matrix shared-memory scratch; sums go through width-LPM shuffles.
#pragma once
#include "common.cuh"

namespace rayen {

#ifdef RAYEN_LMI_TRACE
// development build only (scripts/lmi_trace.py): phase time stamps of warp 0 / warp 7 of the first CTAs
__device__ long long g_lmi_trace[4096];
#define LMI_STAMP(slot)                                                                      \
  do {                                                                                       \
    if (blockIdx.x < 16 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == 7)) \
      g_lmi_trace[blockIdx.x * 64 + ((threadIdx.x >> 5) ? 32 : 0) + (slot)] = clock64();     \
  } while (0)
#define LMI_GT(slot)                                                                         \
  do {                                                                                       \
    if (threadIdx.x == 0 && blockIdx.x < 256) {                                              \
      long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
      g_lmi_trace[2048 + blockIdx.x * 2 + (slot)] = t_;                                      \
    }                                                                                        \
  } while (0)
#else
#define LMI_STAMP(slot) do { } while (0)
#define LMI_GT(slot) do { } while (0)
#endif

constexpr int kLmiThreads = 256;     // backward kernel (keeps the reflectors: needs the registers)
constexpr int kLmiFwdThreads = 256;  // forward kernel default (384 = 12 warps/SM with spills: RAYEN_LMI_THREADS=384)
constexpr int kLmiMaxN = 32;

template <int RP>
struct LmiCfg {
  static constexpr int LPM = RP / 4;             // lanes per matrix
  static constexpr int MPW = 32 / LPM;           // matrices per warp
  static constexpr int PPL = (LPM >= 8) ? 1 : 8 / LPM;  // Sturm points per lane: 8 points per round
  static constexpr int ROUNDS = 8;               // 9^8 = 4.3e7 > 2^25
  // scratch floats per matrix; the pad makes consecutive matrices start LPM (>= 4) banks apart so that
  // neither the float4 broadcasts nor the scalar stores of the matrices of one warp collide
  static constexpr int SCR = kLmiMaxN + 7 * RP + (LPM >= 4 ? LPM : 4);
  static constexpr int NPL = kLmiMaxN / LPM;     // entries of an n-vector per lane
};

template <int LPM>
__device__ __forceinline__ int group_or(int x) {
#pragma unroll
  for (int off = LPM >> 1; off > 0; off >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

// scratch layout (floats): [0,32) u | v | w | d | e | aux0 | aux1 | tau   (each RP long after u)
template <int RP, bool WANT_GRAD, bool F_SMEM>
struct LmiSolver {
  using C = LmiCfg<RP>;
  static constexpr int LPM = C::LPM;

  float A[RP][4];
  float* scr;
  int q;         // lane within the matrix group
  int grp_base;  // first lane of the group

  __device__ __forceinline__ float* su() { return scr; }
  __device__ __forceinline__ float* sv() { return scr + kLmiMaxN; }
  __device__ __forceinline__ float* sw() { return scr + kLmiMaxN + RP; }
  __device__ __forceinline__ float* sd() { return scr + kLmiMaxN + 2 * RP; }
  __device__ __forceinline__ float* se() { return scr + kLmiMaxN + 3 * RP; }
  __device__ __forceinline__ float* sx0() { return scr + kLmiMaxN + 4 * RP; }
  __device__ __forceinline__ float* sx1() { return scr + kLmiMaxN + 5 * RP; }
  __device__ __forceinline__ float* stau() { return scr + kLmiMaxN + 6 * RP; }  // 2 / |v_k|^2 of reflector k

  // ---- 0. u = v / max(|v|, eps) into the scratch; returns |v|
  __device__ __forceinline__ float load_direction(const float* __restrict__ vrow, int n, bool valid) {
    float ss = 0.f;
    float* u = su();
    for (int a = q; a < kLmiMaxN; a += LPM) {
      const float x = (valid && a < n) ? __ldg(vrow + a) : 0.f;
      u[a] = x;
      ss = fmaf(x, x, ss);
    }
    ss = group_sum<LPM>(ss);
    const float s = sqrtf(ss);
    const float inv = 1.0f / fmaxf(s, kNormEps);
    for (int a = q; a < kLmiMaxN; a += LPM) u[a] *= inv;
    __syncwarp();
    return s;
  }

  // ---- 1. S~ = sum_a u_a F~z_a; F is [a][row][lane q][slot t], so a lane reads one float4 per row
  __device__ __forceinline__ void contract(const float* __restrict__ F, int n) {
#pragma unroll
    for (int i = 0; i < RP; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) A[i][t] = 0.f;
    const float* u = su();
    const float* Fq = F + 4 * q;
    for (int a = 0; a < n; ++a) {
      const float ua = u[a];
      const float* Fa = Fq + a * (RP * RP);
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        float4 f;
        if constexpr (F_SMEM)
          f = ld4(Fa + i * RP);
        else
          f = __ldg(reinterpret_cast<const float4*>(Fa + i * RP));
        A[i][0] = fmaf(ua, f.x, A[i][0]);
        A[i][1] = fmaf(ua, f.y, A[i][1]);
        A[i][2] = fmaf(ua, f.z, A[i][2]);
        A[i][3] = fmaf(ua, f.w, A[i][3]);
      }
    }
  }

  // same contraction with the coefficients read from global memory and a trailing coefficient of 1 (the
  // constant matrix of the LMI): used by the violation checker, S = -F(y)
  __device__ __forceinline__ void contract_affine(const float* __restrict__ F, const float* __restrict__ coef, int k,
                                                  bool valid) {
#pragma unroll
    for (int i = 0; i < RP; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) A[i][t] = 0.f;
    const float* Fq = F + 4 * q;
    for (int a = 0; a <= k; ++a) {
      const float ua = valid ? (a < k ? __ldg(coef + a) : 1.0f) : 0.f;
      const float* Fa = Fq + a * (RP * RP);
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        float4 f;
        if constexpr (F_SMEM)
          f = ld4(Fa + i * RP);
        else
          f = __ldg(reinterpret_cast<const float4*>(Fa + i * RP));
        A[i][0] = fmaf(ua, f.x, A[i][0]);
        A[i][1] = fmaf(ua, f.y, A[i][1]);
        A[i][2] = fmaf(ua, f.z, A[i][2]);
        A[i][3] = fmaf(ua, f.w, A[i][3]);
      }
    }
  }

  // ---- 2. Householder tridiagonalisation.  Afterwards sd()/se() hold the diagonal / sub-diagonal and,
  //         if WANT_GRAD, row k of A holds this lane's part of reflector k (zeros in dead columns).
  // The RP-2 reduction steps run as 4 runtime loops ("stages") instead of RP-2 unrolled bodies: stage S
  // covers k in [LPM*S, LPM*(S+1)), during which column slots < S are dead, slot S is partly live and
  // rows < LPM*S are dead -- all compile-time facts, so the register file is still indexed statically
  // while the code stays small enough for the instruction cache.  The pivot row k (= column k, by
  // symmetry) is carried from step to step in xo[]: it is picked out of the register file with a few
  // predicated moves while the previous step updates the LPM rows that can be next.
  template <int S>
  __device__ __forceinline__ void householder_stage(float (&xo)[4]) {
    constexpr int R0 = LPM * S;
    constexpr int K_END = (LPM * (S + 1) < RP - 2) ? LPM * (S + 1) : RP - 2;
    constexpr int I4 = R0 / 4;                                       // first float4 of rows that can be live
    constexpr int EDGE = (R0 + LPM < RP - 1) ? R0 + LPM : RP - 1;    // last row that can become the pivot row
    if constexpr (R0 < RP - 2) {
      for (int k = R0; k < K_END; ++k) {
        const int kk = k - R0;
        if (q == kk) sd()[k] = xo[S];  // d_k: column k lives in lane kk, slot S
        // x_{k+1} = A[k][k+1] lives in lane (kk+1) % LPM, slot S or (when the lane index wraps) S+1
        const bool wrap = (kk + 1 == LPM);
        float cand = xo[S];
        if constexpr (S < 3) {
          if (wrap) cand = xo[S + 1];
        }
        const float xk1 = __shfl_sync(0xffffffffu, cand, grp_base + (wrap ? 0 : kk + 1));
        float loc = 0.f;
#pragma unroll
        for (int t = S; t < 4; ++t) {
          const int j = q + LPM * t;
          if (j > k + 1) loc = fmaf(xo[t], xo[t], loc);
        }
        const float tail2 = group_sum<LPM>(loc);
        const float sigma = fmaf(xk1, xk1, tail2);
        const float rt = sqrtf(sigma);
        const float alpha = (xk1 >= 0.f) ? -rt : rt;
        const bool skip = !(tail2 > 0.f);  // column already tridiagonal (also covers zero padding)
        const float tau = skip ? 0.f : 1.0f / fmaf(fabsf(xk1), rt, sigma);
        if (q == kk) {
          se()[k] = skip ? xk1 : alpha;
          if constexpr (WANT_GRAD) stau()[k] = tau;  // = 2 / v'v: the back-transform needs no second reduction
        }
        float vo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) vo[t] = 0.f;
#pragma unroll
        for (int t = S; t < 4; ++t) {
          const int j = q + LPM * t;
          float val = (j > k + 1) ? xo[t] : ((j == k + 1) ? (xk1 - alpha) : 0.f);
          if (skip) val = 0.f;
          vo[t] = val;
          sv()[j] = val;
        }
        if constexpr (WANT_GRAD) {
#pragma unroll
          for (int i = R0; i < K_END; ++i)
            if (i == k) {
#pragma unroll
              for (int t = S; t < 4; ++t) A[i][t] = vo[t];
            }
        }
        __syncwarp();
        float vr[RP];
#pragma unroll
        for (int i4 = I4; i4 < RP / 4; ++i4) {
          const float4 x = ld4(sv() + 4 * i4);
          vr[4 * i4 + 0] = x.x;
          vr[4 * i4 + 1] = x.y;
          vr[4 * i4 + 2] = x.z;
          vr[4 * i4 + 3] = x.w;
        }
        // p = tau * A v (dead rows carry v_i = 0)
        float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 4 * I4; i < RP; ++i)
#pragma unroll
          for (int t = S; t < 4; ++t) p[t] = fmaf(A[i][t], vr[i], p[t]);
        loc = 0.f;
#pragma unroll
        for (int t = S; t < 4; ++t) {
          p[t] *= tau;
          loc = fmaf(vo[t], p[t], loc);
        }
        const float Kc = 0.5f * tau * group_sum<LPM>(loc);
        float wo[4];
#pragma unroll
        for (int t = S; t < 4; ++t) {
          const int j = q + LPM * t;
          wo[t] = (j > k) ? fmaf(-Kc, vo[t], p[t]) : 0.f;
          sw()[j] = wo[t];
        }
        __syncwarp();
        // A <- A - v w' - w v'; the row that becomes the next pivot row is copied out on the way
#pragma unroll
        for (int i4 = I4; i4 < RP / 4; ++i4) {
          const float4 x = ld4(sw() + 4 * i4);
          const float wr[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = 4 * i4 + ii;
#pragma unroll
            for (int t = S; t < 4; ++t) A[i][t] = fmaf(-vr[i], wo[t], fmaf(-wr[ii], vo[t], A[i][t]));
            if (i > R0 && i <= EDGE) {
              if (i == k + 1) {
#pragma unroll
                for (int t = S; t < 4; ++t) xo[t] = A[i][t];
              }
            }
          }
        }
        __syncwarp();  // scratch v/w are rewritten by the next step
      }
    }
  }

  __device__ __forceinline__ void tridiagonalize() {
    // sw must start clean: entries of dead column slots are read (times v_i = 0) but never rewritten
    float xo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      sw()[q + LPM * t] = 0.f;
      sv()[q + LPM * t] = 0.f;
      xo[t] = A[0][t];
    }
    householder_stage<0>(xo);
    householder_stage<1>(xo);
    householder_stage<2>(xo);
    householder_stage<3>(xo);
    // trailing 2x2 block
    constexpr int K2 = RP - 2, K1 = RP - 1;
    if (q == K2 % LPM) {
      sd()[K2] = A[K2][K2 / LPM];
      se()[K2] = A[K1][K2 / LPM];
    }
    if (q == K1 % LPM) {
      sd()[K1] = A[K1][K1 / LPM];
      se()[K1] = 0.f;
    }
    __syncwarp();
  }

  // ---- 3. largest eigenvalue of the tridiagonal matrix, clipped at 0 (kappa = relu(lambda_max)).
  // Parallel multisection: 8 probes per matrix and round (one per lane), each probe decides "x above the
  // whole spectrum?" with a Sturm sequence; 8 rounds shrink the Gershgorin interval by 9^8.
  __device__ __forceinline__ float lambda_max_relu() {
    float d[RP], e2[RP];
    float dmax = -3.0e38f, hi = -3.0e38f, lo_g = 3.0e38f;
    {
      float eprev = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < RP / 4; ++i4) {
        const float4 dv = ld4(sd() + 4 * i4), ev = ld4(se() + 4 * i4);
        const float da[4] = {dv.x, dv.y, dv.z, dv.w}, ea[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float rad = fabsf(eprev) + fabsf(ea[ii]);
          dmax = fmaxf(dmax, da[ii]);
          hi = fmaxf(hi, da[ii] + rad);
          lo_g = fminf(lo_g, da[ii] - rad);
          d[4 * i4 + ii] = da[ii];
          e2[4 * i4 + ii] = eprev * eprev;  // e2[i] couples rows i-1 and i
          eprev = ea[ii];
        }
      }
    }
    // Work on T / scale so that |d_i - x| <= 2 and e_i^2 <= 1, in the PRODUCT form of the Sturm sequence with the
    // alternating sign folded in:  s_0 = 1, s_1 = x - d_0, s_{i+1} = (x - d_i) s_i - e_i^2 s_{i-1}
    // (s_i = (-1)^i det(T_i - x I); two dependent FMAs per step instead of a reciprocal): x is above the whole
    // spectrum iff every s_i is strictly positive, i.e. iff the running minimum is.  Only signs matter, so
    // (s_{i-1}, s_i) is rescaled by a positive power of two every 4 steps to stay clear of overflow / underflow.
    const float scale = fmaxf(fmaxf(fabsf(hi), fabsf(lo_g)), 1e-30f);
    const float inv_scale = 1.0f / scale;
#pragma unroll
    for (int i = 0; i < RP; ++i) {
      d[i] *= inv_scale;
      e2[i] *= inv_scale * inv_scale;
    }
    float lo = fmaxf(dmax, 0.f) * inv_scale;
    hi = fmaf(1e-6f, scale, hi) * inv_scale;
    hi = fmaxf(hi, lo);  // whole spectrum <= 0: degenerate interval, the rounds below return lo = 0 (no early
                         // exit: the shuffles below need every lane of the warp)
    for (int round = 0; round < C::ROUNDS; ++round) {
      // the first round also probes x = lo itself (8 sections): lo = 0 above the whole spectrum means lambda_max < 0,
      // and the answer is then exactly 0 (relu), not the midpoint of a tiny interval above it
      const int shift = round ? 1 : 0;
      const float h = (hi - lo) / static_cast<float>(8 + shift);
      int bits = 0;
#pragma unroll
      for (int pp = 0; pp < C::PPL; ++pp) {
        const int pt = q * C::PPL + pp;
        const float x = fmaf(h, static_cast<float>(pt + shift), lo);
        float s0 = 1.f, s1 = x - d[0];
        float mn = s1;
#pragma unroll
        for (int i = 1; i < RP; ++i) {
          const float sn = fmaf(x - d[i], s1, -e2[i] * s0);
          mn = fminf(mn, sn);
          s0 = s1;
          s1 = sn;
          if ((i & 3) == 3 && i + 1 < RP) {
            const int ex = (__float_as_int(fmaxf(fabsf(s0), fabsf(s1))) >> 23) & 0xff;
            const float sc = __int_as_float((254 - max(min(ex, 253), 1)) << 23);
            s0 *= sc;
            s1 *= sc;
          }
        }
        bits |= (mn > 0.f) ? (1 << pt) : 0;
      }
      const int mask = group_or<LPM>(bits) & 0xff;
      const int first = mask ? (__ffs(mask) - 1) : 8;  // first probe above the spectrum
      const float new_lo = (first + shift == 0) ? lo : fmaf(h, static_cast<float>(first + shift - 1), lo);
      const float new_hi = (first == 8) ? hi : fmaf(h, static_cast<float>(first + shift), lo);
      lo = new_lo;
      hi = new_hi;
    }
    // below 1e-7 of the matrix scale lambda_max is 0 to the resolution of a float32 matrix (a negative semi-definite
    // S~, e.g. a zero-padded negative definite one): exactly 0 then, the same answer pruning gives
    const float mid = 0.5f * (lo + hi);
    return (mid > 1e-7f) ? mid * scale : 0.f;
  }

  // ---- 4. unit eigenvector of lambda (backward only): twisted factorisation of T - lambda I on one
  // lane per matrix, then q = H_0 ... H_{RP-3} z with the reflectors kept in the dead rows of A.
  __device__ __forceinline__ void eigenvector(float lam, float (&qo)[4]) {
    static_assert(WANT_GRAD, "eigenvector needs the reflectors");
    float* dp = sx0();
    float* dm = sx1();
    float* z = sv();
    float* rr = sw();  // recurrence ratios of z (the w scratch is free after the tridiagonalisation)
    const float* d = sd();
    const float* e = se();
    // Gershgorin-type scale, every lane for itself (RP independent loads, no cross-lane traffic)
    float scale = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < RP / 4; ++i4) {
      const float4 dv = ld4(d + 4 * i4), ev = ld4(e + 4 * i4);
      scale = fmaxf(scale, fmaxf(fmaxf(fabsf(dv.x) + fabsf(ev.x), fabsf(dv.y) + fabsf(ev.y)),
                                 fmaxf(fabsf(dv.z) + fabsf(ev.z), fabsf(dv.w) + fabsf(ev.w))));
    }
    const float tiny = fmaxf(1e-12f * scale, 1e-30f);
    // the two pivot recurrences of the twisted factorisation run on two different lanes at the same time
    constexpr int QB = (LPM > 1) ? 1 : 0;  // lane of the backward recurrence
    // (same instruction stream, direction chosen per lane: divergent branches would serialise them)
    auto pivots = [&](bool fw) {
      float* out = fw ? dp : dm;
      const int di = fw ? 1 : -1;
      int i = fw ? 0 : RP - 1;
      float piv = d[i] - lam;
#pragma unroll 8
      for (int t = 0; t < RP; ++t) {
        if (t > 0) {
          const float ee = fw ? e[i - 1] : e[i];
          piv = fmaf(-ee * ee, __frcp_rn(piv), d[i] - lam);
        }
        if (fabsf(piv) < tiny) piv = -tiny;
        out[i] = piv;
        i += di;
      }
    };
    if constexpr (QB == 0) {
      if (q == 0) {
        pivots(true);
        pivots(false);
      }
    } else {
      if (q <= QB) pivots(q == 0);
    }
    __syncwarp();
    // twist index: argmin_i |dp_i + dm_i - (d_i - lam)|, lowest index on ties; every lane looks at its 4 entries
    float best = 3.0e38f;
    int kt = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = q + LPM * t;
      const float gam = fabsf(dp[i] + dm[i] - (d[i] - lam));
      if (gam < best) {
        best = gam;
        kt = i;
      }
    }
#pragma unroll
    for (int off = LPM >> 1; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ok = __shfl_xor_sync(0xffffffffu, kt, off);
      if (ob < best || (ob == best && ok < kt)) {
        best = ob;
        kt = ok;
      }
    }
    // z_{i-1} = -e_{i-1}/dp_{i-1} z_i below the twist, z_{i+1} = -e_i/dm_{i+1} z_i above it: the ratios are
    // independent of each other, so all lanes compute them first and the serial part is one multiply per entry
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = q + LPM * t;
      float r = 0.f;
      if (i < kt) r = -e[i] * __frcp_rn(dp[i]);                 // multiplies z_{i+1} to give z_i
      else if (i > kt) r = -e[i - 1] * __frcp_rn(dm[i]);        // multiplies z_{i-1} to give z_i
      rr[i] = r;
    }
    __syncwarp();
    float nrm = 0.f;
    auto fill = [&](bool down) {
      const int di = down ? -1 : 1;
      float zi = 1.f;
      int i = kt + di;
      for (int t = 1; t < RP; ++t) {
        if (i >= 0 && i < RP) {
          zi *= rr[i];
          z[i] = zi;
          nrm = fmaf(zi, zi, nrm);
        }
        i += di;
      }
    };
    if constexpr (QB == 0) {
      if (q == 0) {
        fill(true);
        fill(false);
      }
    } else {
      if (q <= QB) fill(q == 0);
    }
    if (q == 0) {
      z[kt] = 1.f;
      nrm += 1.f;
    }
    nrm = group_sum<LPM>(nrm);  // lanes other than 0 / QB contribute 0
    const float inv = rsqrtf(nrm);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) qo[t] = z[q + LPM * t] * inv;
    // q = H_0 ... H_{RP-3} z; reflector k sits in row k of A (this lane's columns)
    back_stage<3>(qo);
    back_stage<2>(qo);
    back_stage<1>(qo);
    back_stage<0>(qo);
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) z[q + LPM * t] = qo[t];
    __syncwarp();
  }

  template <int S>
  __device__ __forceinline__ void back_stage(float (&qo)[4]) {
    constexpr int R0 = LPM * S;
    constexpr int K_END = (LPM * (S + 1) < RP - 2) ? LPM * (S + 1) : RP - 2;
    if constexpr (R0 < RP - 2) {
      for (int k = K_END - 1; k >= R0; --k) {
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = R0; i < K_END; ++i)
          if (i == k) {
#pragma unroll
            for (int t = S; t < 4; ++t) r[t] = A[i][t];
          }
        float dot = 0.f;
#pragma unroll
        for (int t = S; t < 4; ++t) dot = fmaf(r[t], qo[t], dot);
        const float c = stau()[k] * group_sum<LPM>(dot);  // tau_k = 2 / v_k'v_k (0 for a skipped step)
#pragma unroll
        for (int t = S; t < 4; ++t) qo[t] = fmaf(-c, r[t], qo[t]);
      }
    }
  }

  // d kappa/du_a = q' F~z_a q for the entries a = q + LPM*slot owned by this lane
  __device__ __forceinline__ void eig_gradient(const float* __restrict__ F, int n, const float (&qo)[4],
                                               float (&dk)[C::NPL]) {
    const float* qs = sv();  // left there by eigenvector()
    float qa[RP];
#pragma unroll
    for (int i4 = 0; i4 < RP / 4; ++i4) {
      const float4 x = ld4(qs + 4 * i4);
      qa[4 * i4 + 0] = x.x;
      qa[4 * i4 + 1] = x.y;
      qa[4 * i4 + 2] = x.z;
      qa[4 * i4 + 3] = x.w;
    }
#pragma unroll
    for (int sl = 0; sl < C::NPL; ++sl) dk[sl] = 0.f;
    const float* Fq = F + 4 * q;
    for (int a = 0; a < n; ++a) {
      const float* Fa = Fq + a * (RP * RP);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        float4 f;
        if constexpr (F_SMEM)
          f = ld4(Fa + i * RP);
        else
          f = __ldg(reinterpret_cast<const float4*>(Fa + i * RP));
        acc[0] = fmaf(qa[i], f.x, acc[0]);
        acc[1] = fmaf(qa[i], f.y, acc[1]);
        acc[2] = fmaf(qa[i], f.z, acc[2]);
        acc[3] = fmaf(qa[i], f.w, acc[3]);
      }
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) part = fmaf(acc[t], qo[t], part);
      part = group_sum<LPM>(part);
#pragma unroll
      for (int sl = 0; sl < C::NPL; ++sl)
        if (a == q + LPM * sl) dk[sl] = part;
    }
    __syncwarp();
  }

  // The same quantity with the whole warp working on the matrix of lane group G (its unit eigenvector was left in
  // that group's scratch by eigenvector()): lane group h takes a = h, h + MPW, ... and writes d kappa/du_a straight
  // to out_row[a].  When one sample of a warp needs the gradient -- the usual case behind a pruned work list --
  // this is MPW times shorter than every group walking all n matrices for its own sample.
  // Writes out[row_off + a].
  __device__ __forceinline__ void eig_gradient_coop(const float* __restrict__ F, int n, int G, int my_grp,
                                                    float* __restrict__ out, long long row_off) {
    const float* qs = scr + (G - my_grp) * C::SCR + kLmiMaxN;  // sv() of group G
    float qa[RP], qo[4];
#pragma unroll
    for (int i4 = 0; i4 < RP / 4; ++i4) {
      const float4 x = ld4(qs + 4 * i4);
      qa[4 * i4 + 0] = x.x;
      qa[4 * i4 + 1] = x.y;
      qa[4 * i4 + 2] = x.z;
      qa[4 * i4 + 3] = x.w;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) qo[t] = qs[q + LPM * t];
    const float* Fq = F + 4 * q;
    // the trip count is the same for every lane (the group sums below are full-warp shuffles): groups whose a
    // falls beyond n redo the last matrix and write nothing
    for (int a0 = 0; a0 < n; a0 += C::MPW) {
      const int a = a0 + my_grp;
      const float* Fa = Fq + (a < n ? a : n - 1) * (RP * RP);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        float4 f;
        if constexpr (F_SMEM)
          f = ld4(Fa + i * RP);
        else
          f = __ldg(reinterpret_cast<const float4*>(Fa + i * RP));
        acc[0] = fmaf(qa[i], f.x, acc[0]);
        acc[1] = fmaf(qa[i], f.y, acc[1]);
        acc[2] = fmaf(qa[i], f.z, acc[2]);
        acc[3] = fmaf(qa[i], f.w, acc[3]);
      }
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) part = fmaf(acc[t], qo[t], part);
      part = group_sum<LPM>(part);
      if (q == 0 && a < n) out[row_off + a] = part;
    }
  }
};

// dynamic smem: 64 B barriers | F (if F_SMEM) | per-warp scratch
template <int RP, bool F_SMEM>
__host__ __device__ constexpr size_t lmi_smem_bytes(int n, int threads) {
  return 64 + (F_SMEM ? static_cast<size_t>(n) * RP * RP * 4 : 0) +
         static_cast<size_t>(threads / 32) * LmiCfg<RP>::MPW * LmiCfg<RP>::SCR * 4;
}

template <int RP, bool F_SMEM>
__device__ __forceinline__ const float* lmi_stage(const PlanDev& P, unsigned char* smem_raw, uint64_t* bars,
                                                  float** scratch_base, bool has_work) {
  const float* F;
  if constexpr (F_SMEM) {
    float* fs = reinterpret_cast<float*>(smem_raw + 64);
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && has_work) stage_bulk(fs, P.blob + P.off_lmi, P.lmi_words, &bars[0]);
    F = fs;
    *scratch_base = fs + P.lmi_words;
  } else {
    F = P.blob + P.off_lmi;
    *scratch_base = reinterpret_cast<float*>(smem_raw + 64);
  }
  return F;
}

// ----------------------------------------------------------------------------- forward
// prior kappa/tag (from lqs_forward_kernel) are merged when has_prior != 0; y, kappa, active are written.
// WITH_GRAD: additionally, for the samples whose binding constraint turns out to be the LMI (and whose gradient
// needs it), finish the job while the tridiagonal form and the reflectors are still in registers: top eigenvector,
// d kappa/du_a = q' F~z_a q, stored to dkappa[b, :].  Backward then needs no LMI kernel at all.
template <int RP, bool F_SMEM, int THREADS, bool WITH_GRAD>
__global__ void __launch_bounds__(THREADS, 1)
    lmi_forward_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                       float* __restrict__ kappa_io, int* __restrict__ active_io, long long B, int mode,
                       int has_prior, const int* __restrict__ work_list, const int* __restrict__ work_count,
                       float* __restrict__ dkappa) {
  using C = LmiCfg<RP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* scratch_base;
  LMI_GT(0);
  pdl_wait();  // launched behind the linear/quadratic/SOC kernel: its kappa / active / work list must be complete
  // dense mode: every sample; list mode: only the samples the LQS kernel could not prune
  const long long total = work_list ? static_cast<long long>(ld_after_wait(work_count)) : B;
  // chunk c (MPW samples) belongs to CTA c % gridDim: a CTA without a chunk does not stage F~z at all
  const bool cta_has_work = static_cast<long long>(blockIdx.x) * C::MPW < total;
  LMI_STAMP(0);
  const float* F = lmi_stage<RP, F_SMEM>(P, smem_raw, bars, &scratch_base, cta_has_work);
  LMI_STAMP(1);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LmiSolver<RP, WITH_GRAD, F_SMEM> S;
  S.q = lane % C::LPM;
  S.grp_base = lane - S.q;
  const int grp = lane / C::LPM;
  S.scr = scratch_base + (warp * C::MPW + grp) * C::SCR;
  const int n = P.n, k = P.k;
  const float* y0 = P.blob + P.off_y0;
  const float* nmat = P.blob + P.off_nmat;
  // chunk c of MPW samples goes to CTA c % gridDim, warp c / gridDim: a short work list spreads over all SMs
  // instead of filling the first CTAs (the per-SM shared-memory pipe is what the contraction saturates)
  const long long warp_id = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  bool staged = !F_SMEM;

  for (long long base = warp_id * C::MPW; base < total; base += n_warps * C::MPW) {
    const long long idx = base + grp;
    const bool valid = idx < total;
    const long long b = valid ? (work_list ? static_cast<long long>(work_list[idx]) : idx) : 0;
    LMI_STAMP(2);
    const float s = S.load_direction(v + b * ldv, n, valid);
    LMI_STAMP(3);
    if (!staged) {
      mbar_wait(&bars[0], 0);
      staged = true;
    }
    LMI_STAMP(4);
    S.contract(F, n);
    LMI_STAMP(5);
    S.tridiagonalize();
    LMI_STAMP(6);
    const float lam = S.lambda_max_relu();
    LMI_STAMP(7);
    float kap = fmaxf(lam, 0.f);
    int tag = kap > 0.f ? make_tag(RAYEN_FAM_LMI, 0) : make_tag(RAYEN_FAM_NONE, 0);
    if (has_prior && valid) {
      const float k0 = kappa_io[b];
      const int t0 = active_io[b];
      if (!(kap > k0)) {
        kap = k0;
        tag = t0;
      }
    }
    __syncwarp();  // every lane of the matrix has read the prior before lane 0 overwrites it
    if (valid) {
      if (S.q == 0) {
        kappa_io[b] = kap;
        active_io[b] = tag;
      }
      float alpha;
      if (mode == RAYEN_MODE_RAYEN_OLD)
        alpha = 1.0f / (expf(__ldg(v + b * ldv + n)) + kap);
      else
        alpha = fminf(1.0f / kap, s);
      const float* u = S.su();
      float* yrow = y + b * k;
      if (P.n_is_identity) {
        for (int a = S.q; a < k; a += C::LPM) yrow[a] = fmaf(alpha, u[a], __ldg(y0 + a));
      } else {
        for (int i = S.q; i < k; i += C::LPM) {
          const float* nrow = nmat + i * (P.np + 4);
          float acc = 0.f;
          for (int a = 0; a < n; ++a) acc = fmaf(__ldg(nrow + a), u[a], acc);
          yrow[i] = fmaf(alpha, acc, __ldg(y0 + i));
        }
      }
    }
    LMI_STAMP(8);
    if constexpr (WITH_GRAD) {
      bool need = valid && tag_family(tag) == RAYEN_FAM_LMI && kap > 0.f;
      if (need && mode == RAYEN_MODE_RAYEN) need = (1.0f / kap < s);
      if (__ballot_sync(0xffffffffu, need) != 0u) {  // warp-uniform: the other matrices of the warp just ride along
        __syncwarp();
        float qo[4];
        S.eigenvector(lam, qo);
        LMI_STAMP(9);
        // one needing sample after the other, the whole warp on each
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
#pragma unroll 1
        for (int G = 0; G < C::MPW; ++G) {
          if (!((need_mask >> (G * C::LPM)) & 1u)) continue;
          // (sample indices fit 32 bits: the work lists are int32)
          const int bG = __shfl_sync(0xffffffffu, static_cast<int>(b), G * C::LPM);
          S.eig_gradient_coop(F, n, G, grp, dkappa, static_cast<long long>(bG) * n);
        }
        LMI_STAMP(10);
      }
    }
    __syncwarp();  // the scratch (u) is rewritten by the next sample
    LMI_STAMP(11);
  }
  if constexpr (F_SMEM) {
    if (!staged && cta_has_work) mbar_wait(&bars[0], 0);  // never exit with a bulk copy in flight
  }
#ifdef RAYEN_LMI_TRACE
  __syncthreads();
  LMI_GT(1);
#endif
}

// ----------------------------------------------------------------------------- backward
// Only the samples whose binding constraint is the LMI and whose gradient needs d kappa/du are
// processed; everything else was written by lqs_backward_kernel, which also queued those samples in
// the work list (dense mode without a list: every group checks its own sample).
template <int RP, bool F_SMEM>
__global__ void __launch_bounds__(kLmiThreads, 1)
    lmi_backward_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, const float* __restrict__ gy,
                        const float* __restrict__ kappa, const int* __restrict__ active, float* __restrict__ gv,
                        long long ldgv, long long B, int mode, const int* __restrict__ work_list,
                        const int* __restrict__ work_count) {
  using C = LmiCfg<RP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* scratch_base;
  const long long total = work_list ? static_cast<long long>(ld_after_wait(work_count)) : B;
  // only the CTAs that own a chunk of the (usually short) work list stage F~z
  const bool cta_has_work = static_cast<long long>(blockIdx.x) * C::MPW < total;
  const float* F = lmi_stage<RP, F_SMEM>(P, smem_raw, bars, &scratch_base, cta_has_work);
  bool staged = !F_SMEM;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LmiSolver<RP, true, F_SMEM> S;
  S.q = lane % C::LPM;
  S.grp_base = lane - S.q;
  const int grp = lane / C::LPM;
  S.scr = scratch_base + (warp * C::MPW + grp) * C::SCR;
  const int n = P.n, k = P.k;
  const float* nmat = P.blob + P.off_nmat;
  // chunk c of MPW samples goes to CTA c % gridDim, warp c / gridDim: a short work list spreads over all SMs
  // instead of filling the first CTAs (the per-SM shared-memory pipe is what the contraction saturates)
  const long long warp_id = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);

  for (long long base = warp_id * C::MPW; base < total; base += n_warps * C::MPW) {
    const long long idx = base + grp;
    const long long b = (idx < total) ? (work_list ? static_cast<long long>(work_list[idx]) : idx) : 0;
    bool mine = false;
    float kap = 0.f;
    if (idx < total) {
      kap = __ldg(kappa + b);
      mine = tag_family(__ldg(active + b)) == RAYEN_FAM_LMI && kap > 0.f;
    }
    const float s = S.load_direction(v + b * ldv, n, mine);
    if (mine && mode == RAYEN_MODE_RAYEN) mine = (1.0f / kap < s);
    if (__ballot_sync(0xffffffffu, mine) == 0u) continue;  // warp-uniform
    if (!staged) {
      mbar_wait(&bars[0], 0);
      staged = true;
    }
    // groups that are not `mine` run on u = 0 (a zero matrix) and write nothing
    if (!mine) {
      for (int a = S.q; a < kLmiMaxN; a += C::LPM) S.su()[a] = 0.f;
    }
    __syncwarp();
    S.contract(F, n);
    S.tridiagonalize();
    const float lam = S.lambda_max_relu();
    float qo[4];
    S.eigenvector(lam, qo);
    float dk[C::NPL];
    S.eig_gradient(F, n, qo, dk);

    // closed-form tail with the n-vector spread over the lanes of the matrix (a = q + LPM*slot)
    const float* u = S.su();
    float uo[C::NPL], gz[C::NPL];
    float gzu = 0.f;
#pragma unroll
    for (int sl = 0; sl < C::NPL; ++sl) {
      const int a = S.q + C::LPM * sl;
      uo[sl] = u[a];
      float g = 0.f;
      if (mine && a < n) {
        if (P.n_is_identity) {
          g = __ldg(gy + b * k + a);
        } else {
          for (int i = 0; i < k; ++i) g = fmaf(__ldg(gy + b * k + i), __ldg(nmat + i * (P.np + 4) + a), g);
        }
      }
      gz[sl] = g;
      gzu = fmaf(g, uo[sl], gzu);
    }
    gzu = group_sum<C::LPM>(gzu);
    float gu[C::NPL], guu = 0.f, gbeta = 0.f;
    if (mode == RAYEN_MODE_RAYEN_OLD) {
      const float eb = mine ? expf(__ldg(v + b * ldv + n)) : 1.f;
      const float alpha = 1.0f / (eb + kap);
      const float c = gzu * alpha * alpha;
#pragma unroll
      for (int sl = 0; sl < C::NPL; ++sl) gu[sl] = fmaf(alpha, gz[sl], -c * dk[sl]);
      gbeta = -c * eb;
    } else {
      const float ik = mine ? 1.0f / kap : 0.f;
      const float c = gzu * ik * ik;
#pragma unroll
      for (int sl = 0; sl < C::NPL; ++sl) gu[sl] = fmaf(ik, gz[sl], -c * dk[sl]);
    }
#pragma unroll
    for (int sl = 0; sl < C::NPL; ++sl) guu = fmaf(gu[sl], uo[sl], guu);
    guu = group_sum<C::LPM>(guu);
    if (s < kNormEps) guu = 0.f;
    const float inv_s = 1.0f / fmaxf(s, kNormEps);
    if (mine) {
#pragma unroll
      for (int sl = 0; sl < C::NPL; ++sl) {
        const int a = S.q + C::LPM * sl;
        if (a < n) gv[b * ldgv + a] = (gu[sl] - guu * uo[sl]) * inv_s;
      }
      if (mode == RAYEN_MODE_RAYEN_OLD && S.q == 0) gv[b * ldgv + n] = gbeta;
    }
    __syncwarp();
  }
  if constexpr (F_SMEM) {
    if (!staged && cta_has_work) mbar_wait(&bars[0], 0);
  }
}

}  // namespace rayen
