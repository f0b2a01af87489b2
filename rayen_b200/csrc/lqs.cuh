// Linear / quadratic / SOC families: fused kappa + shift-and-scale forward, closed-form backward.
//
// Thread mapping (forward): a "tile" is TM consecutive samples whose unit directions u live in
// registers (TM x NP floats); L consecutive lanes (L = 1..32, power of two, chosen at launch from the
// batch size) share one tile and split the constraints between them -- 4-row chunks of D, whole
// quadratics, whole cones -- then merge their (kappa, tag) with log2(L) shuffles.  The constants
// are staged once per CTA into shared memory with 1-D TMA bulk copies and read as LDS.128; every
// LDS.128 feeds 4*TM FMAs, which is what keeps the kernel on the FP32 pipe instead of the LSU.
#pragma once
#include "common.cuh"

namespace rayen {

__host__ __device__ constexpr int lqs_max_threads(int np, int tm) {
  return (np * tm >= 128) ? 384 : ((np * tm >= 64) ? 512 : 768);
}

// ----------------------------------------------------------------------------- loading directions
template <int NP>
__device__ __forceinline__ void load_row(const float* __restrict__ row, int n, bool vec_ok, bool valid,
                                         float (&x)[NP]) {
  if (!valid) {
#pragma unroll
    for (int a = 0; a < NP; ++a) x[a] = 0.f;
    return;
  }
  if (vec_ok) {
#pragma unroll
    for (int kk = 0; kk < NP / 4; ++kk) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * kk < n) t = __ldg(reinterpret_cast<const float4*>(row) + kk);
      x[4 * kk + 0] = t.x;
      x[4 * kk + 1] = t.y;
      x[4 * kk + 2] = t.z;
      x[4 * kk + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int a = 0; a < NP; ++a) x[a] = (a < n) ? __ldg(row + a) : 0.f;
  }
}

template <int NP>
__device__ __forceinline__ float normalize_row(float (&x)[NP]) {
  float ss = 0.f;
#pragma unroll
  for (int a = 0; a < NP; ++a) ss = fmaf(x[a], x[a], ss);
  const float s = sqrtf(ss);
  const float inv = 1.0f / fmaxf(s, kNormEps);
#pragma unroll
  for (int a = 0; a < NP; ++a) x[a] *= inv;
  return s;
}

// dot of a 4-float constant group with u[t][4*kk .. 4*kk+3]
#define RAYEN_FMA4(acc, c4, uu, kk)            \
  acc = fmaf((c4).x, (uu)[4 * (kk) + 0], acc); \
  acc = fmaf((c4).y, (uu)[4 * (kk) + 1], acc); \
  acc = fmaf((c4).z, (uu)[4 * (kk) + 2], acc); \
  acc = fmaf((c4).w, (uu)[4 * (kk) + 3], acc);

// ----------------------------------------------------------------------------- kappa of the three families
// ||T u||^2 for a packed upper-triangular T (row i keeps columns 4*floor(i/4)..NP-1).
template <int NP, int TM>
__device__ __forceinline__ void tri_norm2(const float* __restrict__ tri, const float (&u)[TM][NP],
                                          float (&out)[TM]) {
#pragma unroll
  for (int t = 0; t < TM; ++t) out[t] = 0.f;
  int pos = 0;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float r[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) r[t] = 0.f;
#pragma unroll
    for (int kk = i / 4; kk < NP / 4; ++kk) {
      const float4 g = ld4(tri + pos);
      pos += 4;
#pragma unroll
      for (int t = 0; t < TM; ++t) { RAYEN_FMA4(r[t], g, u[t], kk) }
    }
#pragma unroll
    for (int t = 0; t < TM; ++t) out[t] = fmaf(r[t], r[t], out[t]);
  }
}

template <int NP, int TM>
__device__ __forceinline__ void dot_np(const float* __restrict__ c, const float (&u)[TM][NP], float (&out)[TM]) {
#pragma unroll
  for (int t = 0; t < TM; ++t) out[t] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NP / 4; ++kk) {
    const float4 g = ld4(c + 4 * kk);
#pragma unroll
    for (int t = 0; t < TM; ++t) { RAYEN_FMA4(out[t], g, u[t], kk) }
  }
}

// Largest root of -A kappa^2 + 2 hb kappa + cq = 0 (A > 0), i.e. the reference's
// solveSecondOrderEq (constraint_module.py:339-348) with a' = -A, b' = 2 hb, c' = cq, written in the
// cancellation-free form.  A non-positive result means the ray never leaves the cone.
__device__ __forceinline__ float soc_root(float A, float hb, float cq, float* root_out) {
  const float disc = fmaxf(fmaf(hb, hb, A * cq), 0.f);
  const float root = sqrtf(disc);
  if (root_out) *root_out = root;
  if (hb >= 0.f) return (hb + root) / A;
  const float den = root - hb;  // > 0
  return cq / den;
}

// Wolkowicz-Styan upper bound of lambda_max(S~(u)) = mean + radius (mean = tr S~/r, radius = sqrt((r-1)/r) times the
// Frobenius norm of the trace-free part), plus what the float32 evaluation of its two dot products can be off by:
// `margin` (absolute, from the plan: rayen_b200.h BOUND) and 1e-5 of the terms.  A sample is pruned iff this is below
// the kappa of the other families; being conservative only costs time.
__device__ __forceinline__ float lmi_upper_bound(float mean, float radius, float margin) {
  return mean + radius + fmaf(1e-5f, fabsf(mean) + radius, margin);
}

template <int NP, int TM>
__device__ __forceinline__ void kappa_lqs(const PlanDev& P, const float* __restrict__ cst,
                                          const float (&u)[TM][NP], int lane_l, int L, float (&best)[TM],
                                          int (&tag)[TM]) {
  // ---- linear rows: kappa_j = D_j . u                        (reference constraint_module.py:353)
  {
    const float* lin = cst;
    const int nchunks = P.m_pad >> 2;
    for (int c = lane_l; c < nchunks; c += L) {
      const float* p = lin + c * P.lin_stride;
      float acc[4][TM];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < TM; ++t) acc[i][t] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NP / 4; ++kk) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 d = ld4(p + kk * 16 + i * 4);
#pragma unroll
          for (int t = 0; t < TM; ++t) { RAYEN_FMA4(acc[i][t], d, u[t], kk) }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < TM; ++t)
          if (acc[i][t] > best[t]) {
            best[t] = acc[i][t];
            tag[t] = make_tag(RAYEN_FAM_LINEAR, 4 * c + i);
          }
    }
  }
  // ---- quadratics: kappa_i = phi_z . u + ||G u||             (reference constraint_module.py:360-381)
  {
    const float* quad = cst + (P.off_quad - P.off_lin);
    for (int q = lane_l; q < P.n_quad; q += L) {
      const float* base = quad + q * P.quad_stride;
      float lin_part[TM], nrm2[TM];
      dot_np<NP, TM>(base, u, lin_part);
      tri_norm2<NP, TM>(base + NP, u, nrm2);
#pragma unroll
      for (int t = 0; t < TM; ++t) {
        const float kap = lin_part[t] + sqrtf(nrm2[t]);
        if (kap > best[t]) {
          best[t] = kap;
          tag[t] = make_tag(RAYEN_FAM_QUAD, q);
        }
      }
    }
  }
  // ---- second-order cones                                      (reference constraint_module.py:383-399)
  {
    const float* soc = cst + (P.off_soc - P.off_lin);
    constexpr int TRI = (NP / 4) * (NP / 4 + 1) * 8;  // packed triangular words
    for (int j = lane_l; j < P.n_soc; j += L) {
      const float* base = soc + j * P.soc_stride;
      float cu[TM], hb[TM], nrm2[TM];
      dot_np<NP, TM>(base, u, cu);
      dot_np<NP, TM>(base + NP, u, hb);
      tri_norm2<NP, TM>(base + 2 * NP, u, nrm2);
      const float A = base[2 * NP + TRI];
#pragma unroll
      for (int t = 0; t < TM; ++t) {
        const float cq = fmaf(-cu[t], cu[t], nrm2[t]);
        const float kap = soc_root(A, hb[t], cq, nullptr);
        if (kap > best[t]) {
          best[t] = kap;
          tag[t] = make_tag(RAYEN_FAM_SOC, j);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- forward kernel
// grid: persistent CTAs; dynamic smem = 64 B of mbarriers + the LQS constant region (SMEM == true).
template <int NP, int TM, bool SMEM>
__global__ void __launch_bounds__(lqs_max_threads(NP, TM), 1)
    lqs_forward_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                       float* __restrict__ kappa_out, int* __restrict__ active_out, long long B, int mode,
                       int L, int lmi_follows, int prune, int* __restrict__ work_list,
                       int* __restrict__ work_count, const MapArgs M) {
  pdl_launch_dependents();  // see common.cuh
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const float* cst;
  if constexpr (SMEM) {
    float* cs = reinterpret_cast<float*>(smem_raw + 64);
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // linear rows first so that their FMAs start while the rest is still in flight
      const int lin_words = P.off_quad - P.off_lin;
      stage_bulk(cs, P.blob + P.off_lin, lin_words, &bars[0]);
      stage_bulk(cs + lin_words, P.blob + P.off_quad, P.lqs_words - lin_words, &bars[1]);
    }
    cst = cs;
  } else {
    cst = P.blob + P.off_lin;
  }

  const int lane = threadIdx.x & 31;
  const int lane_l = lane & (L - 1);
  const int tiles_per_warp = 32 / L;
  const long long warp_id = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long n_tiles = (B + TM - 1) / TM;
  const int n = P.n, k = P.k;
  const bool vec_in = ((n & 3) == 0) && ((ldv & 3) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  const bool vec_out = ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  bool staged = !SMEM;

  for (long long tile_base = warp_id * tiles_per_warp; tile_base < n_tiles; tile_base += n_warps * tiles_per_warp) {
    const long long tile = tile_base + lane / L;
    float u[TM][NP];
    float s[TM], beta[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      const long long b = tile * TM + t;
      const bool valid = (tile < n_tiles) && (b < B);
      if (M.x)  // fused mapper: every lane of the sample computes the same row (the v_out stores coincide)
        map_row<NP>(M, b, n, valid, u[t]);
      else
        load_row<NP>(v + b * ldv, n, vec_in, valid, u[t]);
      beta[t] = (mode == RAYEN_MODE_RAYEN_OLD && valid) ? __ldg(v + b * ldv + n) : 0.f;
      s[t] = normalize_row<NP>(u[t]);
    }
    if (!staged) {
      mbar_wait(&bars[0], 0);
    }
    float best[TM];
    int tag[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      best[t] = 0.f;
      tag[t] = make_tag(RAYEN_FAM_NONE, 0);
    }
    if (!staged) {
      // the linear chunk loop only touches the first section; the rest must have landed before the
      // quadratic loop starts, so wait for both here (the second wait is almost always free).
      mbar_wait(&bars[1], 0);
      staged = true;
    }
    kappa_lqs<NP, TM>(P, cst, u, lane_l, L, best, tag);
#pragma unroll
    for (int t = 0; t < TM; ++t) group_argmax(best[t], tag[t], L);

    // ---- LMI pruning: an upper bound of lambda_max(S~(u)) that needs no eigen-solve.  If it is below the
    // kappa found so far, the LMI cannot bind and this kernel finishes the sample itself.
    bool pruned[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) pruned[t] = false;
    if (lmi_follows && prune) {
      constexpr int TRI = (NP / 4) * (NP / 4 + 1) * 8;
      const float* bnd = cst + (P.off_bound - P.off_lin);
      float tu[TM], nrm2[TM];
      dot_np<NP, TM>(bnd, u, tu);
      tri_norm2<NP, TM>(bnd + NP, u, nrm2);
      const float r = bnd[NP + TRI];
      const float inv_r = 1.0f / r;
#pragma unroll
      for (int t = 0; t < TM; ++t) {
        // nrm2 = |T_c u|^2, T_c the factor of the CENTRED Gram matrix (a sum of squares: no cancellation)
        const float ub = lmi_upper_bound(tu[t] * inv_r, sqrtf((r - 1.0f) * inv_r * nrm2[t]), P.lmi_bound_margin);
        pruned[t] = ub < best[t];
      }
    }

    const float* y0 = cst + (P.off_y0 - P.off_lin);
    const float* nmat = cst + (P.off_nmat - P.off_lin);
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      const long long b = tile * TM + t;
      const bool valid = (tile < n_tiles) && (b < B);
      if (!valid) continue;
      const float kap = best[t];
      const bool finish = !lmi_follows || pruned[t];
      if (lane_l == 0) {
        if (kappa_out) kappa_out[b] = kap;
        if (active_out) active_out[b] = tag[t];
        if (!finish && work_list) work_list[atomicAdd(work_count, 1)] = static_cast<int>(b);
      }
      if (!finish) continue;
      // shift-and-scale (reference constraint_module.py:472-474 / :464-465, :512-514)
      float alpha;
      if (mode == RAYEN_MODE_RAYEN_OLD)
        alpha = 1.0f / (expf(beta[t]) + kap);
      else
        alpha = fminf(1.0f / kap, s[t]);
      float* yrow = y + b * k;
      if (P.n_is_identity) {
#pragma unroll
        for (int kk = 0; kk < NP / 4; ++kk) {
          if (4 * kk < k && (kk & (L - 1)) == lane_l) {
            const float4 c = ld4(y0 + 4 * kk);
            float4 o;
            o.x = fmaf(alpha, u[t][4 * kk + 0], c.x);
            o.y = fmaf(alpha, u[t][4 * kk + 1], c.y);
            o.z = fmaf(alpha, u[t][4 * kk + 2], c.z);
            o.w = fmaf(alpha, u[t][4 * kk + 3], c.w);
            if (vec_out) {
              *reinterpret_cast<float4*>(yrow + 4 * kk) = o;
            } else {
              if (4 * kk + 0 < k) yrow[4 * kk + 0] = o.x;
              if (4 * kk + 1 < k) yrow[4 * kk + 1] = o.y;
              if (4 * kk + 2 < k) yrow[4 * kk + 2] = o.z;
              if (4 * kk + 3 < k) yrow[4 * kk + 3] = o.w;
            }
          }
        }
      } else {
        for (int i = lane_l; i < k; i += L) {
          const float* nrow = nmat + i * (NP + 4);
          float acc = 0.f;
#pragma unroll
          for (int kk = 0; kk < NP / 4; ++kk) {
            const float4 g = ld4(nrow + 4 * kk);
            RAYEN_FMA4(acc, g, u[t], kk)
          }
          yrow[i] = fmaf(alpha, acc, y0[i]);
        }
      }
    }
  }
  if constexpr (SMEM) {
    // a CTA that got no tile must still not exit while its bulk copies are in flight
    if (!staged) {
      mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], 0);
    }
  }
}

// ----------------------------------------------------------------------------- backward
// d kappa / d u of the binding constraint, z-space, constants read through L1/L2.
template <int NP>
__device__ __forceinline__ void tri_grad(const float* __restrict__ tri, const float (&u)[NP], float (&g)[NP],
                                         float* norm2) {
  // g = T'(T u), norm2 = |T u|^2
  float ss = 0.f;
#pragma unroll
  for (int a = 0; a < NP; ++a) g[a] = 0.f;
  int pos = 0;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float4 row[NP / 4];
    float r = 0.f;
#pragma unroll
    for (int kk = i / 4; kk < NP / 4; ++kk) {
      row[kk] = __ldg(reinterpret_cast<const float4*>(tri + pos));
      pos += 4;
      RAYEN_FMA4(r, row[kk], u, kk)
    }
    ss = fmaf(r, r, ss);
#pragma unroll
    for (int kk = i / 4; kk < NP / 4; ++kk) {
      g[4 * kk + 0] = fmaf(row[kk].x, r, g[4 * kk + 0]);
      g[4 * kk + 1] = fmaf(row[kk].y, r, g[4 * kk + 1]);
      g[4 * kk + 2] = fmaf(row[kk].z, r, g[4 * kk + 2]);
      g[4 * kk + 3] = fmaf(row[kk].w, r, g[4 * kk + 3]);
    }
  }
  *norm2 = ss;
}

template <int NP>
__device__ __forceinline__ void dkappa_lqs(const PlanDev& P, int tag, float kap, const float (&u)[NP],
                                           float (&dk)[NP]) {
  const int fam = tag_family(tag), idx = tag_index(tag);
  const float* blob = P.blob;
#pragma unroll
  for (int a = 0; a < NP; ++a) dk[a] = 0.f;
  if (fam == RAYEN_FAM_LINEAR) {
    const float* p = blob + P.off_lin + (idx >> 2) * P.lin_stride + (idx & 3) * 4;
#pragma unroll
    for (int kk = 0; kk < NP / 4; ++kk) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(p + kk * 16));
      dk[4 * kk + 0] = d.x;
      dk[4 * kk + 1] = d.y;
      dk[4 * kk + 2] = d.z;
      dk[4 * kk + 3] = d.w;
    }
  } else if (fam == RAYEN_FAM_QUAD) {
    const float* base = blob + P.off_quad + idx * P.quad_stride;
    float g[NP], nrm2;
    tri_grad<NP>(base + NP, u, g, &nrm2);
    const float root = sqrtf(nrm2);
    const float inv = root > 0.f ? 1.0f / root : 0.f;
#pragma unroll
    for (int a = 0; a < NP; ++a) dk[a] = fmaf(g[a], inv, __ldg(base + a));
  } else if (fam == RAYEN_FAM_SOC) {
    constexpr int TRI = (NP / 4) * (NP / 4 + 1) * 8;
    const float* base = blob + P.off_soc + idx * P.soc_stride;
    float g[NP], nrm2, cu = 0.f, hb = 0.f;
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      cu = fmaf(__ldg(base + a), u[a], cu);
      hb = fmaf(__ldg(base + NP + a), u[a], hb);
    }
    tri_grad<NP>(base + 2 * NP, u, g, &nrm2);
    const float A = __ldg(base + 2 * NP + TRI);
    const float cq = fmaf(-cu, cu, nrm2);
    float root;
    (void)soc_root(A, hb, cq, &root);
    // d kappa/du = (kappa h + R'R u - (c.u) c) / sqrt(disc);  the reference's autograd is NaN at
    // disc == 0 (tangent ray, measure zero) -- emit 0 there.
    const float inv = root > 0.f ? 1.0f / root : 0.f;
#pragma unroll
    for (int a = 0; a < NP; ++a)
      dk[a] = (fmaf(kap, __ldg(base + NP + a), g[a]) - cu * __ldg(base + a)) * inv;
  }
}

// Shared tail of both backward kernels: from g_z, u, s, kappa and d kappa/du to g_v.
template <int NP>
__device__ __forceinline__ void backward_tail(int mode, float s, float kap, float beta, const float (&u)[NP],
                                              const float (&gz)[NP], const float (&dk)[NP], bool boundary,
                                              float (&gv)[NP], float* gbeta) {
  float gzu = 0.f;
#pragma unroll
  for (int a = 0; a < NP; ++a) gzu = fmaf(gz[a], u[a], gzu);
  float gu[NP];
  if (mode == RAYEN_MODE_RAYEN_OLD) {
    const float eb = expf(beta);
    const float alpha = 1.0f / (eb + kap);
    const float c = gzu * alpha * alpha;
#pragma unroll
    for (int a = 0; a < NP; ++a) gu[a] = fmaf(alpha, gz[a], -c * dk[a]);
    *gbeta = -c * eb;
  } else {
    if (!boundary) {
      // z = z0 + v: identity Jacobian (and 0 at v == 0, the norm's subgradient)
#pragma unroll
      for (int a = 0; a < NP; ++a) gv[a] = (s > 0.f) ? gz[a] : 0.f;
      return;
    }
    const float ik = 1.0f / kap;
    const float c = gzu * ik * ik;
#pragma unroll
    for (int a = 0; a < NP; ++a) gu[a] = fmaf(ik, gz[a], -c * dk[a]);
  }
  float guu = 0.f;
#pragma unroll
  for (int a = 0; a < NP; ++a) guu = fmaf(gu[a], u[a], guu);
  const float inv_s = 1.0f / fmaxf(s, kNormEps);
  if (s < kNormEps) guu = 0.f;  // normalize() divides by the constant eps there
#pragma unroll
  for (int a = 0; a < NP; ++a) gv[a] = (gu[a] - guu * u[a]) * inv_s;
}

// g_z = N' g_y
template <int NP>
__device__ __forceinline__ void load_gz(const PlanDev& P, const float* __restrict__ gyrow, bool vec_ok,
                                        float (&gz)[NP]) {
  if (P.n_is_identity) {
    load_row<NP>(gyrow, P.n, vec_ok, true, gz);
  } else {
#pragma unroll
    for (int a = 0; a < NP; ++a) gz[a] = 0.f;
    const float* nmat = P.blob + P.off_nmat;
    for (int i = 0; i < P.k; ++i) {
      const float gi = __ldg(gyrow + i);
      const float* nrow = nmat + i * (NP + 4);
#pragma unroll
      for (int kk = 0; kk < NP / 4; ++kk) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(nrow + 4 * kk));
        gz[4 * kk + 0] = fmaf(gi, c.x, gz[4 * kk + 0]);
        gz[4 * kk + 1] = fmaf(gi, c.y, gz[4 * kk + 1]);
        gz[4 * kk + 2] = fmaf(gi, c.z, gz[4 * kk + 2]);
        gz[4 * kk + 3] = fmaf(gi, c.w, gz[4 * kk + 3]);
      }
    }
  }
}

template <int NP>
__device__ __forceinline__ void store_row(float* __restrict__ row, int n, bool vec_ok, const float (&x)[NP]) {
  if (vec_ok) {
#pragma unroll
    for (int kk = 0; kk < NP / 4; ++kk)
      if (4 * kk < n)
        *reinterpret_cast<float4*>(row + 4 * kk) = make_float4(x[4 * kk], x[4 * kk + 1], x[4 * kk + 2], x[4 * kk + 3]);
  } else {
#pragma unroll
    for (int a = 0; a < NP; ++a)
      if (a < n) row[a] = x[a];
  }
}

// ---- warp-tile row I/O: the 32 rows of a warp are one contiguous chunk of a dense [B, NP] tensor, so the warp
// moves it with fully coalesced 16-byte accesses (512 B per instruction) and transposes through a padded
// shared-memory tile; per-thread row accesses touch 32 different lines per instruction and lean on L1 instead.
template <int NP>
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ base, int rows_valid, float* tile, int lane,
                                               float (&x)[NP]) {
  constexpr int C4 = NP / 4, TS = NP + 4;
  const float4* src = reinterpret_cast<const float4*>(base);
#pragma unroll
  for (int it = 0; it < C4; ++it) {
    const int i = it * 32 + lane;
    const int r = i / C4, c4 = i % C4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) t = __ldg(src + i);
    *reinterpret_cast<float4*>(tile + r * TS + 4 * c4) = t;
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < C4; ++kk) {
    const float4 t = *reinterpret_cast<const float4*>(tile + lane * TS + 4 * kk);
    x[4 * kk + 0] = t.x;
    x[4 * kk + 1] = t.y;
    x[4 * kk + 2] = t.z;
    x[4 * kk + 3] = t.w;
  }
  __syncwarp();
}
// Two tensors at once: all global loads are issued before the first shared-memory store, so the two tiles cost one
// memory round trip instead of two.
template <int NP>
__device__ __forceinline__ void warp_load_rows2(const float* __restrict__ base_a, const float* __restrict__ base_b,
                                                int rows_valid, float* tile_a, float* tile_b, int lane,
                                                float (&xa)[NP], float (&xb)[NP]) {
  constexpr int C4 = NP / 4, TS = NP + 4;
  const float4* sa = reinterpret_cast<const float4*>(base_a);
  const float4* sb = reinterpret_cast<const float4*>(base_b);
  float4 ta[C4], tb[C4];
#pragma unroll
  for (int it = 0; it < C4; ++it) {
    const int i = it * 32 + lane;
    const bool ok = i / C4 < rows_valid;
    ta[it] = ok ? __ldg(sa + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    tb[it] = ok ? __ldg(sb + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int it = 0; it < C4; ++it) {
    const int i = it * 32 + lane;
    const int r = i / C4, c4 = i % C4;
    *reinterpret_cast<float4*>(tile_a + r * TS + 4 * c4) = ta[it];
    *reinterpret_cast<float4*>(tile_b + r * TS + 4 * c4) = tb[it];
  }
  __syncwarp();
#pragma unroll
  for (int kk = 0; kk < C4; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(tile_a + lane * TS + 4 * kk);
    const float4 b = *reinterpret_cast<const float4*>(tile_b + lane * TS + 4 * kk);
    xa[4 * kk + 0] = a.x; xa[4 * kk + 1] = a.y; xa[4 * kk + 2] = a.z; xa[4 * kk + 3] = a.w;
    xb[4 * kk + 0] = b.x; xb[4 * kk + 1] = b.y; xb[4 * kk + 2] = b.z; xb[4 * kk + 3] = b.w;
  }
  __syncwarp();
}
template <int NP>
__device__ __forceinline__ void warp_store_rows(float* __restrict__ base, int rows_valid, float* tile, int lane,
                                                const float (&x)[NP]) {
  constexpr int C4 = NP / 4, TS = NP + 4;
#pragma unroll
  for (int kk = 0; kk < C4; ++kk)
    *reinterpret_cast<float4*>(tile + lane * TS + 4 * kk) = make_float4(x[4 * kk], x[4 * kk + 1], x[4 * kk + 2], x[4 * kk + 3]);
  __syncwarp();
  float4* dst = reinterpret_cast<float4*>(base);
#pragma unroll
  for (int it = 0; it < C4; ++it) {
    const int i = it * 32 + lane;
    const int r = i / C4, c4 = i % C4;
    if (r < rows_valid) dst[i] = *reinterpret_cast<const float4*>(tile + r * TS + 4 * c4);
  }
  __syncwarp();
}

// d kappa/du of ONE quadratic- or cone-bound sample with the whole warp on it (lane j owns component j).  All NP rows
// of the packed triangular factor are loaded first (one coalesced line each, one L2 round trip), T u is NP
// independent warp shuffle sums, T'(T u) one FMA per row.  The per-thread dkappa_lqs walks ~270 dependent 16-byte
// loads and ~1000 FMAs on a single lane: with one or two such samples in a warp that was the tail of the whole
// backward kernel (CTAs with such a sample: 6.4-7.3 us, CTAs without: 2.9 us; scripts/bwd_trace.py).
template <int NP>
__device__ __forceinline__ float dkappa_tri_warp(const PlanDev& P, int fam, int idx, float kap, float uj, int lane) {
  constexpr int TRI = (NP / 4) * (NP / 4 + 1) * 8;
  const bool soc = fam == RAYEN_FAM_SOC;
  const float* item = soc ? P.blob + P.off_soc + idx * P.soc_stride : P.blob + P.off_quad + idx * P.quad_stride;
  const float* tri = item + (soc ? 2 * NP : NP);
  const bool in = lane < NP;
  float t[NP], w[NP];
  {
    int pos = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int c0 = 4 * (i / 4);
      t[i] = (in && lane >= c0) ? __ldg(tri + pos + (lane - c0)) : 0.f;
      pos += NP - c0;
    }
  }
  const float cj = in ? __ldg(item + lane) : 0.f;                 // phi_z (quadratic) or c_z (cone)
  const float hj = (in && soc) ? __ldg(item + NP + lane) : 0.f;
  const float A = soc ? __ldg(item + 2 * NP + TRI) : 1.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) w[i] = t[i] * uj;
  float cu = cj * uj, hb = hj * uj;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int i = 0; i < NP; ++i) w[i] += __shfl_xor_sync(0xffffffffu, w[i], off);
    cu += __shfl_xor_sync(0xffffffffu, cu, off);
    hb += __shfl_xor_sync(0xffffffffu, hb, off);
  }
  float g = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    ss = fmaf(w[i], w[i], ss);
    g = fmaf(w[i], t[i], g);
  }
  if (!soc) {
    const float root = sqrtf(ss);
    const float inv = root > 0.f ? 1.0f / root : 0.f;
    return fmaf(g, inv, cj);                                      // phi_z + G'(G u)/|G u|   (reference :360-381)
  }
  float root;
  (void)soc_root(A, hb, fmaf(-cu, cu, ss), &root);
  // (kappa h + R'R u - (c_z.u) c_z)/sqrt(disc) (reference :383-399, :339-348); the reference's autograd is NaN at
  // disc == 0 (tangent ray, measure zero) -- emit 0 there
  const float inv = root > 0.f ? 1.0f / root : 0.f;
  return (fmaf(kap, hj, g) - cu * cj) * inv;
}

#ifdef RAYEN_BWD_TRACE
// development build only (scripts/bwd_trace.py): per-CTA globaltimer start/end and phase stamps of warp 0
__device__ long long g_bwd_trace[8192];
#define BWD_STAMP(slot) do { if ((threadIdx.x & 127) == 0 && blockIdx.x < 256) g_bwd_trace[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
#define BWD_GT(slot) do { if (threadIdx.x == 0 && blockIdx.x < 256) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_bwd_trace[4096 + blockIdx.x * 2 + (slot)] = t_; } } while (0)
#else
#define BWD_STAMP(slot) do { } while (0)
#define BWD_GT(slot) do { } while (0)
#endif

constexpr int kBwdThreads = 128;
template <int NP>
__host__ __device__ constexpr size_t lqs_bwd_smem_bytes() {
  return static_cast<size_t>(kBwdThreads / 32) * 2 * 32 * (NP + 4) * sizeof(float);  // two tiles per warp
}

// One thread per sample, one warp per 32 consecutive samples.  Samples whose binding constraint is the LMI (and
// that need d kappa/du the forward pass did not leave behind) are appended to a work list for lmi_backward_kernel.
template <int NP>
__global__ void __launch_bounds__(kBwdThreads)
    lqs_backward_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, const float* __restrict__ gy,
                        const float* __restrict__ kappa, const int* __restrict__ active, float* __restrict__ gv,
                        long long ldgv, long long B, int mode, int* __restrict__ work_list,
                        int* __restrict__ work_count, const float* __restrict__ dkappa) {
  extern __shared__ __align__(16) float bwd_tiles[];
  BWD_GT(0);
  BWD_STAMP(0);
  const int n = P.n;
  const int lane = threadIdx.x & 31;
  float* tile = bwd_tiles + (threadIdx.x >> 5) * 2 * 32 * (NP + 4);
  float* tile2 = tile + 32 * (NP + 4);
  const bool vec_v = ((n & 3) == 0) && ((ldv & 3) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
  const bool vec_gy = ((P.k & 3) == 0) && ((reinterpret_cast<uintptr_t>(gy) & 15) == 0);
  const bool vec_gv = ((n & 3) == 0) && ((ldgv & 3) == 0) && ((reinterpret_cast<uintptr_t>(gv) & 15) == 0);
  // dense contiguous tensors of full width: the warp-tile path
  const bool tile_v = vec_v && n == NP && ldv == n;
  const bool tile_gy = vec_gy && P.n_is_identity && P.k == NP;
  const bool tile_gv = vec_gv && n == NP && ldgv == n;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x + (threadIdx.x & ~31); base < B; base += stride) {
    const long long b = base + lane;
    const bool valid = b < B;
    const int rows_valid = (B - base < 32) ? static_cast<int>(B - base) : 32;
    const float kap = valid ? __ldg(kappa + b) : 1.f;
    const int tag = valid ? __ldg(active + b) : 0;
    float u[NP], gz[NP], dk[NP], g[NP];
    if (tile_v && tile_gy) {
      warp_load_rows2<NP>(v + base * n, gy + base * P.k, rows_valid, tile, tile2, lane, u, gz);
    } else {
      if (tile_v)
        warp_load_rows<NP>(v + base * n, rows_valid, tile, lane, u);
      else
        load_row<NP>(v + b * ldv, n, vec_v, valid, u);
      if (tile_gy) {
        warp_load_rows<NP>(gy + base * P.k, rows_valid, tile, lane, gz);
      } else if (valid) {
        load_gz<NP>(P, gy + b * P.k, vec_gy, gz);
      } else {
#pragma unroll
        for (int a = 0; a < NP; ++a) gz[a] = 0.f;
      }
    }
    // the row of D of a linear-bound sample is fetched as soon as the tag is known (before |v| says whether the
    // sample sits on the boundary at all): one more load in flight next to the tiles
    const int fam = tag_family(tag);
#pragma unroll
    for (int a = 0; a < NP; ++a) dk[a] = 0.f;
    if (valid && fam == RAYEN_FAM_LINEAR) {
      const float* p = P.blob + P.off_lin + (tag_index(tag) >> 2) * P.lin_stride + (tag_index(tag) & 3) * 4;
#pragma unroll
      for (int kk = 0; kk < NP / 4; ++kk) {
        const float4 d = __ldg(reinterpret_cast<const float4*>(p + kk * 16));
        dk[4 * kk + 0] = d.x;
        dk[4 * kk + 1] = d.y;
        dk[4 * kk + 2] = d.z;
        dk[4 * kk + 3] = d.w;
      }
    }
    BWD_STAMP(1);
    const float s = normalize_row<NP>(u);
    BWD_STAMP(2);
    const float beta = (mode == RAYEN_MODE_RAYEN_OLD && valid) ? __ldg(v + b * ldv + n) : 0.f;
    const bool boundary = valid && ((mode == RAYEN_MODE_RAYEN_OLD) ? (kap > 0.f) : (1.0f / kap < s));
    const bool lmi_bound = boundary && fam == RAYEN_FAM_LMI;
    // needs the eigenvector and the forward pass did not leave d kappa/du behind: queued for lmi_backward_kernel
    const bool queued = lmi_bound && !dkappa;
    if (queued && work_list) work_list[atomicAdd(work_count, 1)] = static_cast<int>(b);
    if (lmi_bound && !queued) {
#pragma unroll
      for (int a = 0; a < NP; ++a) dk[a] = (a < n) ? __ldg(dkappa + b * n + a) : 0.f;
    }
    // quadratic / cone gradients: a few such samples per warp are taken one at a time by the whole warp (the row
    // tiles still hold the raw v); many of them run in parallel, one per thread
    const bool needs_tri = boundary && !queued && (fam == RAYEN_FAM_QUAD || fam == RAYEN_FAM_SOC);
    unsigned todo = __ballot_sync(0xffffffffu, needs_tri);
    if (tile_v && tile_gy && todo != 0u && __popc(todo) <= 6) {
      constexpr int TS = NP + 4;
      const float inv_s = 1.0f / fmaxf(s, kNormEps);
      while (todo) {
        const int r = __ffs(todo) - 1;
        todo &= todo - 1;
        const int fam_r = __shfl_sync(0xffffffffu, fam, r), idx_r = __shfl_sync(0xffffffffu, tag_index(tag), r);
        const float kap_r = __shfl_sync(0xffffffffu, kap, r), inv_r = __shfl_sync(0xffffffffu, inv_s, r);
        const float uj = (lane < NP) ? tile[r * TS + lane] * inv_r : 0.f;
        const float dj = dkappa_tri_warp<NP>(P, fam_r, idx_r, kap_r, uj, lane);
        if (lane < NP) tile2[r * TS + lane] = dj;   // g_y of this row is already in registers
        __syncwarp();
        if (lane == r) {
#pragma unroll
          for (int kk = 0; kk < NP / 4; ++kk) {
            const float4 d = *reinterpret_cast<const float4*>(tile2 + r * TS + 4 * kk);
            dk[4 * kk + 0] = d.x;
            dk[4 * kk + 1] = d.y;
            dk[4 * kk + 2] = d.z;
            dk[4 * kk + 3] = d.w;
          }
        }
      }
      __syncwarp();
    } else if (needs_tri) {
      dkappa_lqs<NP>(P, tag, kap, u, dk);
    }
    BWD_STAMP(3);
    float gbeta = 0.f;
    backward_tail<NP>(mode, s, kap, beta, u, gz, dk, boundary && !queued, g, &gbeta);
    BWD_STAMP(4);
    if (tile_gv && __all_sync(0xffffffffu, !queued)) {
      warp_store_rows<NP>(gv + base * n, rows_valid, tile, lane, g);
    } else if (valid && !queued) {
      store_row<NP>(gv + b * ldgv, n, vec_gv, g);
    }
    if (valid && !queued && mode == RAYEN_MODE_RAYEN_OLD) gv[b * ldgv + n] = gbeta;
    BWD_STAMP(5);
  }
  BWD_GT(1);
}

}  // namespace rayen
