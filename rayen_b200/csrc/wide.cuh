// Wide sets (32 < n <= 4096): linear + quadratic + SOC kappa, the scale step and the closed-form backward when a
// direction no longer fits a thread's registers.  Same arithmetic as lqs.cuh (reference constraint_module.py:351-399,
// :468-474, :512-514), different mapping: the directions of a tile of samples sit in shared memory, every constraint is
// a set of rows of ONE matrix W (plan section WIDE, stored transposed so that 32 lanes read 32 consecutive rows of a
// column with one coalesced load), a lane owns a row and carries its dot products with the TS (8 or 16) samples of
// the tile in registers, and the unit of work of a warp is a task = one group of 64 rows (two per lane); the groups of a quadratic /
// cone leave partial sums in shared-memory slots that a finalize pass adds up in a fixed order (deterministic).
#pragma once
#include "common.cuh"
#include "lqs.cuh"

namespace rayen {

constexpr int kWideThreads = 256;     // forward: 8 warps
constexpr int kWideWarps = kWideThreads / 32;
// backward: one CTA per sample, 128 threads (n < 512: the per-sample chain is latency-bound, small CTAs keep more samples
// in flight per SM) or 256 threads (wider sets: more rows of the binding item in flight)
constexpr int kWideBwdSwitchN = 512;
constexpr int kWideMaxN = 12288;      // widest subspace: 4-sample tiles of the forward and the backward vectors still fit shared memory
constexpr int kWideMagic = 0x57494445;
constexpr int kWideVersion = 3;
constexpr int kWideGroupRows = 64;   // rows per task: a lane owns two consecutive rows (one 8-byte load per column)
constexpr int kWideSlots = 256;       // partial-sum slots per round (plan.py WIDE_SLOTS)
constexpr int kWideRoundItems = 64;   // items per round (plan.py WIDE_ROUND_ITEMS)

struct WideDev {
  const float* blob;
  int n, k;
  int r_pad, n_tasks, off_tasks, off_wt, off_nt, off_nrow, k32, np, off_items, n_quad, n_soc, n_rounds, off_rounds;
  int off_y0, n_is_identity;
};

// forward smem: us[n][TS] | s_norm, s_beta, s_alpha, sbest [TS each] | stag [TS] | wbest, wtag [warps][TS] |
//               part_sq [slots][TS] | head [items][2][TS] | ksplit [warps][64 rows][TS] (column-split partial sums)
__host__ __device__ inline size_t wide_fwd_smem_bytes(int n, int ts) {
  return (static_cast<size_t>(n) * ts + 5 * ts + 2 * kWideWarps * ts + static_cast<size_t>(kWideSlots) * ts +
          static_cast<size_t>(kWideRoundItems) * 2 * ts + static_cast<size_t>(kWideWarps) * kWideGroupRows * ts) * sizeof(float);
}
__host__ __device__ inline size_t wide_bwd_smem_bytes(int n) {
  // u, gz, dk: n each; t: n + 2 (+ pad); 8 words of reduction scratch
  return (static_cast<size_t>(4) * (n + 4) + 8) * sizeof(float);
}

__device__ __forceinline__ float warp_sum32(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

template <int TS>
__device__ __forceinline__ void wide_fma_col(float w, const float4* __restrict__ urow, float (&acc)[TS]) {
#pragma unroll
  for (int q = 0; q < TS / 4; ++q) {
    const float4 a = urow[q];
    acc[4 * q + 0] = fmaf(w, a.x, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(w, a.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(w, a.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(w, a.w, acc[4 * q + 3]);
  }
}

// BLK (sets with n >= kWideBlockedN): columns are accumulated in blocks of kWideAccBlock and the block sums added to a
// running total: the rounding of a float32 dot product of n terms then grows like sqrt(block) + sqrt(n / block) instead
// of sqrt(n) (n = 4000: ~2e-7 of the terms' scale instead of ~2e-6).  The second set of accumulators costs registers
// (measured: -28 % at n = 64, -41 % at n = 256 when applied everywhere), so narrower sets keep the plain running sum,
// whose rounding is below 1e-6 there anyway.
constexpr int kWideAccBlock = 128;
constexpr int kWideBlockedN = 512;
constexpr int kWideCompensatedN = kWideBlockedN;  // the row walk's block sums are added with a Kahan term (wide_dot2, CMP)

// acc[s] = sum_{j >= j0} Wt[j][row] * us[j][s]: the column walk of one row against the tile's directions; eight
// column loads in flight per lane
template <int TS, bool BLK>
__device__ __forceinline__ void wide_dot(const float* __restrict__ wcol, int r_pad, int j0, int n,
                                         const float4* __restrict__ us4, float (&tot)[TS]) {
  if constexpr (!BLK) {
    // narrow enough for a plain running sum (no second set of accumulators: registers)
#pragma unroll
    for (int s = 0; s < TS; ++s) tot[s] = 0.f;
    int j = j0;
    for (; j + 8 <= n; j += 8) {
      float w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = __ldg(wcol + static_cast<size_t>(j + q) * r_pad);
#pragma unroll
      for (int q = 0; q < 8; ++q) wide_fma_col<TS>(w[q], us4 + static_cast<size_t>(j + q) * (TS / 4), tot);
    }
    for (; j < n; ++j) wide_fma_col<TS>(__ldg(wcol + static_cast<size_t>(j) * r_pad), us4 + static_cast<size_t>(j) * (TS / 4), tot);
  } else {
    float acc[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) tot[s] = acc[s] = 0.f;
    int j = j0;
    while (j + 8 <= n) {
      const int stop = (j + kWideAccBlock < n) ? j + kWideAccBlock : n;
      for (; j + 8 <= stop; j += 8) {
        float w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = __ldg(wcol + static_cast<size_t>(j + q) * r_pad);
#pragma unroll
        for (int q = 0; q < 8; ++q) wide_fma_col<TS>(w[q], us4 + static_cast<size_t>(j + q) * (TS / 4), acc);
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        tot[s] += acc[s];
        acc[s] = 0.f;
      }
    }
    for (; j < n; ++j) wide_fma_col<TS>(__ldg(wcol + static_cast<size_t>(j) * r_pad), us4 + static_cast<size_t>(j) * (TS / 4), acc);
#pragma unroll
    for (int s = 0; s < TS; ++s) tot[s] += acc[s];
  }
}

// Two consecutive rows per lane (wcol2 points at the lane's row pair, 8-byte aligned): per column one 8-byte load and
// TS/4 broadcast 16-byte loads of the tile feed 2*TS FMAs
template <int TS>
__device__ __forceinline__ void wide_fma_col2(float2 w, const float4* __restrict__ urow, float (&acc0)[TS], float (&acc1)[TS]) {
#pragma unroll
  for (int q = 0; q < TS / 4; ++q) {
    const float4 a = urow[q];
    acc0[4 * q + 0] = fmaf(w.x, a.x, acc0[4 * q + 0]);
    acc0[4 * q + 1] = fmaf(w.x, a.y, acc0[4 * q + 1]);
    acc0[4 * q + 2] = fmaf(w.x, a.z, acc0[4 * q + 2]);
    acc0[4 * q + 3] = fmaf(w.x, a.w, acc0[4 * q + 3]);
    acc1[4 * q + 0] = fmaf(w.y, a.x, acc1[4 * q + 0]);
    acc1[4 * q + 1] = fmaf(w.y, a.y, acc1[4 * q + 1]);
    acc1[4 * q + 2] = fmaf(w.y, a.z, acc1[4 * q + 2]);
    acc1[4 * q + 3] = fmaf(w.y, a.w, acc1[4 * q + 3]);
  }
}
template <int TS, bool BLK, bool CMP>
__device__ __forceinline__ void wide_dot2(const float* __restrict__ wcol2, int r_pad, int j0, int n,
                                          const float4* __restrict__ us4, float (&tot0)[TS], float (&tot1)[TS]) {
  // eight columns in flight per lane (sixteen in the 8-sample kernel measured 2x SLOWER on B200: 26.6 -> 54.7 us at n = 64,
  // 179 -> 354 us at n = 1000 -- kept at eight); block sums as in wide_dot
  constexpr int U = 8;
  if constexpr (!BLK) {
#pragma unroll
    for (int s = 0; s < TS; ++s) tot0[s] = tot1[s] = 0.f;
    int j = j0;
    for (; j + U <= n; j += U) {
      float2 w[U];
#pragma unroll
      for (int q = 0; q < U; ++q) w[q] = __ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j + q) * r_pad));
#pragma unroll
      for (int q = 0; q < U; ++q) wide_fma_col2<TS>(w[q], us4 + static_cast<size_t>(j + q) * (TS / 4), tot0, tot1);
    }
    for (; j < n; ++j)
      wide_fma_col2<TS>(__ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j) * r_pad)),
                        us4 + static_cast<size_t>(j) * (TS / 4), tot0, tot1);
  } else if constexpr (CMP) {
    // CMP (every blocked build, n >= 512): dot products of 10^3..10^4 terms that cancel to
    // 1 / |v| ~ 1e-2 of their terms' scale.  Plain blocks of 128 leave ~2.5e-5 relative error in kappa for the worst of
    // 2000 samples (measured: the 1 x 10000 point of the reference's sweep ended 1.09e-5 outside its row); blocks of 16
    // added to the running total with a Kahan compensation term bring that to the rounding of the inputs (~3e-6).
    float acc0[TS], acc1[TS], c0[TS], c1[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) tot0[s] = tot1[s] = acc0[s] = acc1[s] = c0[s] = c1[s] = 0.f;
    auto fold = [&]() {
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        const float y0 = acc0[s] - c0[s], t0 = tot0[s] + y0;
        c0[s] = (t0 - tot0[s]) - y0;
        tot0[s] = t0;
        const float y1 = acc1[s] - c1[s], t1 = tot1[s] + y1;
        c1[s] = (t1 - tot1[s]) - y1;
        tot1[s] = t1;
        acc0[s] = acc1[s] = 0.f;
      }
    };
    int j = j0;
    while (j + U <= n) {
      const int stop = (j + 16 < n) ? j + 16 : n;
      for (; j + U <= stop; j += U) {
        float2 w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) w[q] = __ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j + q) * r_pad));
#pragma unroll
        for (int q = 0; q < U; ++q) wide_fma_col2<TS>(w[q], us4 + static_cast<size_t>(j + q) * (TS / 4), acc0, acc1);
      }
      fold();
    }
    for (; j < n; ++j)
      wide_fma_col2<TS>(__ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j) * r_pad)),
                        us4 + static_cast<size_t>(j) * (TS / 4), acc0, acc1);
    fold();
  } else {
    float acc0[TS], acc1[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) tot0[s] = tot1[s] = acc0[s] = acc1[s] = 0.f;
    int j = j0;
    while (j + U <= n) {
      const int stop = (j + kWideAccBlock < n) ? j + kWideAccBlock : n;
      for (; j + U <= stop; j += U) {
        float2 w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) w[q] = __ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j + q) * r_pad));
#pragma unroll
        for (int q = 0; q < U; ++q) wide_fma_col2<TS>(w[q], us4 + static_cast<size_t>(j + q) * (TS / 4), acc0, acc1);
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        tot0[s] += acc0[s];
        tot1[s] += acc1[s];
        acc0[s] = acc1[s] = 0.f;
      }
    }
    for (; j < n; ++j)
      wide_fma_col2<TS>(__ldg(reinterpret_cast<const float2*>(wcol2 + static_cast<size_t>(j) * r_pad)),
                        us4 + static_cast<size_t>(j) * (TS / 4), acc0, acc1);
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      tot0[s] += acc0[s];
      tot1[s] += acc1[s];
    }
  }
}

// the same walk as a function call (the column-split branch of the forward kernel)
template <int TS, bool BLK, bool CMP>
__device__ __noinline__ void wide_dot2_call(const float* __restrict__ wcol2, int r_pad, int j0, int n,
                                            const float4* __restrict__ us4, float (&tot0)[TS], float (&tot1)[TS]) {
  wide_dot2<TS, BLK, CMP>(wcol2, r_pad, j0, n, us4, tot0, tot1);
}

// value of x[lane] for lane < TS without dynamic register indexing
template <int TS>
__device__ __forceinline__ float pick_lane(const float (&x)[TS], int lane) {
  float r = x[0];
#pragma unroll
  for (int s = 1; s < TS; ++s)
    if (lane == s) r = x[s];
  return r;
}

// (c, ct) replaces (best, tag) when larger; equal positive values go to the lower tag (the reference's order)
__device__ __forceinline__ void wide_take(float c, int ct, float& best, int& tag) {
  if (c > best || (c == best && c > 0.f && ct < tag)) {
    best = c;
    tag = ct;
  }
}

// ----------------------------------------------------------------------------- forward
// grid: one CTA per tile of TS samples (grid-stride); dynamic smem = wide_fwd_smem_bytes(n, TS).
template <int TS, bool BLK, bool CMP = BLK>
__global__ void __launch_bounds__(kWideThreads)
    wide_forward_kernel(const WideDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                        float* __restrict__ kappa_out, int* __restrict__ active_out, long long B, int mode) {
  extern __shared__ __align__(16) float wide_smem[];
  const int n = P.n, k = P.k;
  float* us = wide_smem;                             // [n][TS]
  float* s_norm = us + static_cast<size_t>(n) * TS;  // [TS]
  float* s_beta = s_norm + TS;
  float* s_alpha = s_beta + TS;
  float* sbest = s_alpha + TS;
  int* stag = reinterpret_cast<int*>(sbest + TS);
  float* wbest = sbest + 2 * TS;                     // [kWideWarps][TS]
  int* wtag = reinterpret_cast<int*>(wbest + kWideWarps * TS);
  float* part_sq = wbest + 2 * kWideWarps * TS;      // [kWideSlots][TS]
  float* head = part_sq + kWideSlots * TS;           // [kWideRoundItems][2][TS]
  float* ksplit = head + kWideRoundItems * 2 * TS;   // [kWideWarps][64][TS]
  const float4* us4 = reinterpret_cast<const float4*>(us);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* blob = P.blob;
  const int* tasks = reinterpret_cast<const int*>(blob + P.off_tasks);
  const int* rounds = reinterpret_cast<const int*>(blob + P.off_rounds);
  const int* items = reinterpret_cast<const int*>(blob + P.off_items);
  const float* wt = blob + P.off_wt;
  const float* y0 = blob + P.off_y0;
  const long long n_tiles = (B + TS - 1) / TS;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();  // the previous tile's readers of us / s_alpha are done
    // ---- stage: a warp normalises one sample of the tile at a time (reference constraint_module.py:470)
    for (int sl = warp; sl < TS; sl += kWideWarps) {
      const long long b = tile * TS + sl;
      const bool valid = b < B;
      const float* vrow = v + (valid ? b : 0) * ldv;
      float ss = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float x = valid ? __ldg(vrow + j) : 0.f;
        us[j * TS + sl] = x;
        ss = fmaf(x, x, ss);
      }
      ss = warp_sum32(ss);
      const float s = sqrtf(ss);
      const float inv = 1.0f / fmaxf(s, kNormEps);
      __syncwarp();
      for (int j = lane; j < n; j += 32) us[j * TS + sl] *= inv;
      if (lane == 0) {
        s_norm[sl] = s;
        s_beta[sl] = (valid && mode == RAYEN_MODE_RAYEN_OLD) ? __ldg(vrow + n) : 0.f;
        sbest[sl] = 0.f;
        stag[sl] = 0;
      }
    }
    __syncthreads();

    // ---- linear rows: every lane keeps the running (max, row) of the rows it has seen        (:353)
    float mx[TS];
    int mr[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      mx[s] = 0.f;
      mr[s] = 0;
    }
    for (int rd = 0; rd < P.n_rounds; ++rd) {
      const int t0 = __ldg(rounds + 4 * rd), t1 = __ldg(rounds + 4 * rd + 1), i0 = __ldg(rounds + 4 * rd + 2),
                i1 = __ldg(rounds + 4 * rd + 3);
      // ---- tasks of the round: one group of 64 rows each (two per lane), dealt to the warps in order (heaviest first).
      // A round with fewer tasks than half the warps (few rows in a wide space: 1 ... 100 rows at n = 10^4) splits every
      // task's COLUMN range over `split` warps -- otherwise one warp walks all n columns with eight loads in flight while
      // the others wait at the barrier (measured at 100 rows x 10000: barrier stalls 30 per issue, 4.6 ms per 2000
      // samples) -- and adds the partial sums in a fixed order.  The split depends only on the plan: a sample's arithmetic
      // does not depend on its tile.
      const int nt = t1 - t0;
      const int split = (nt > 0 && 2 * nt <= kWideWarps) ? kWideWarps / nt : 1;
      // what a warp does with the finished dot products of a task (acc0 / acc1: rows 2 lane, 2 lane + 1 of the group)
      auto consume = [&](int kind, int idx, int slot, float (&acc0)[TS], float (&acc1)[TS]) {
        if (kind == 1) {
          // a lane's rows arrive in ascending order: strict > keeps the lowest
#pragma unroll
          for (int s = 0; s < TS; ++s) {
            if (acc0[s] > mx[s]) {
              mx[s] = acc0[s];
              mr[s] = idx + 2 * lane;
            }
            if (acc1[s] > mx[s]) {
              mx[s] = acc1[s];
              mr[s] = idx + 2 * lane + 1;
            }
          }
        } else if (kind == 2) {
          // 64 rows of the triangular factor of a quadratic / cone: a partial |T u|^2                      (:360-399)
          float part[TS];
#pragma unroll
          for (int s = 0; s < TS; ++s) part[s] = warp_sum32(fmaf(acc0[s], acc0[s], acc1[s] * acc1[s]));
          if (lane < TS) part_sq[slot * TS + lane] = pick_lane<TS>(part, lane);
        } else {
          // header rows of 32 items of the round: {phi_z . u, 0} of a quadratic, {c_z . u, h . u} of a cone
          const int il = idx + lane;
          if (il < i1 - i0) {
#pragma unroll
            for (int s = 0; s < TS; ++s) {
              head[(il * 2 + 0) * TS + s] = acc0[s];
              head[(il * 2 + 1) * TS + s] = acc1[s];
            }
          }
        }
      };
      if (split > 1) {
        // (the column walk of this branch is a CALL, wide_dot2_call: a second inlined copy of it cost registers in the
        // common path -- 64 -> 127 at TS = 8 -- and a shared call site de-optimised that path's loop: 3.9 -> 7.3 ms at
        // 500 x 10000)
        const int ti = warp / split, q = warp - ti * split;
        const bool takes = ti < nt;          // (nt * split may be less than the number of warps)
        int kind = 0, idx = 0, slot = 0;
        float acc0[TS], acc1[TS];
        if (takes) {
          const int* tk = tasks + (t0 + ti) * 8;
          kind = __ldg(tk);
          idx = __ldg(tk + 3);
          slot = __ldg(tk + 5);
          const int row = __ldg(tk + 1), j0 = __ldg(tk + 2);
          // slice q of [j0, n), whole blocks of 16 columns (the unit of the compensated sums)
          const int len = ((n - j0 + split - 1) / split + 15) & ~15;
          const int jlo = (j0 + q * len < n) ? j0 + q * len : n, jhi = (jlo + len < n) ? jlo + len : n;
          wide_dot2_call<TS, BLK, CMP>(wt + row + 2 * lane, P.r_pad, jlo, jhi, us4, acc0, acc1);
          if (q != 0) {
            float* mine = ksplit + (warp * kWideGroupRows + 2 * lane) * TS;
#pragma unroll
            for (int s = 0; s < TS; ++s) {
              mine[s] = acc0[s];
              mine[TS + s] = acc1[s];
            }
          }
        }
        __syncthreads();
        if (takes && q == 0) {
          for (int qq = 1; qq < split; ++qq) {
            const float* other = ksplit + ((warp + qq) * kWideGroupRows + 2 * lane) * TS;
#pragma unroll
            for (int s = 0; s < TS; ++s) {
              acc0[s] += other[s];
              acc1[s] += other[TS + s];
            }
          }
          consume(kind, idx, slot, acc0, acc1);
        }
      } else {
        for (int t = t0 + warp; t < t1; t += kWideWarps) {
          const int* tk = tasks + t * 8;
          const int kind = __ldg(tk), row = __ldg(tk + 1), j0 = __ldg(tk + 2), idx = __ldg(tk + 3), slot = __ldg(tk + 5);
          float acc0[TS], acc1[TS];
          wide_dot2<TS, BLK, CMP>(wt + row + 2 * lane, P.r_pad, j0, n, us4, acc0, acc1);
          consume(kind, idx, slot, acc0, acc1);
        }
      }
      __syncthreads();
      // ---- finalize the items of the round: a warp per sample, lanes over the items, parts added in slot order
      if (i1 > i0) {
        for (int sl = warp; sl < TS; sl += kWideWarps) {
          float best = 0.f;
          int tag = 0;
          for (int il = lane; il < i1 - i0; il += 32) {
            const int* it = items + (i0 + il) * 8;
            const int kind = __ldg(it + 1), fidx = __ldg(it + 2), slot0 = __ldg(it + 3), nparts = __ldg(it + 4);
            float nrm2 = 0.f;
            for (int q = 0; q < nparts; ++q) nrm2 += part_sq[(slot0 + q) * TS + sl];
            const float a0 = head[(il * 2 + 0) * TS + sl];
            float kap;
            if (kind == 2) {
              kap = a0 + sqrtf(nrm2);
            } else {
              const float A = __int_as_float(__ldg(it + 5));
              kap = soc_root(A, head[(il * 2 + 1) * TS + sl], fmaf(-a0, a0, nrm2), nullptr);
            }
            wide_take(kap, make_tag(kind == 2 ? RAYEN_FAM_QUAD : RAYEN_FAM_SOC, fidx), best, tag);
          }
          group_argmax(best, tag, 32);
          if (lane == 0) {
            float b0 = sbest[sl];
            int t0s = stag[sl];
            wide_take(best, tag, b0, t0s);
            sbest[sl] = b0;
            stag[sl] = t0s;
          }
        }
        __syncthreads();  // the next round reuses the slots
      }
    }
    // ---- linear rows: lanes -> warp (ties -> lowest row), warps -> CTA through shared memory
#pragma unroll
    for (int s = 0; s < TS; ++s) group_argmax(mx[s], mr[s], 32);
    if (lane < TS) {
      int row = mr[0];
#pragma unroll
      for (int s = 1; s < TS; ++s)
        if (lane == s) row = mr[s];
      wbest[warp * TS + lane] = pick_lane<TS>(mx, lane);
      wtag[warp * TS + lane] = make_tag(RAYEN_FAM_LINEAR, row);
    }
    __syncthreads();
    // ---- combine (ties -> lowest tag = the reference's evaluation order), alpha (:472-474 / :464-465)
    if (tid < TS) {
      float best = sbest[tid];
      int tag = stag[tid];
      for (int w = 0; w < kWideWarps; ++w) wide_take(wbest[w * TS + tid], wtag[w * TS + tid], best, tag);
      const long long b = tile * TS + tid;
      if (b < B) {
        if (kappa_out) kappa_out[b] = best;
        if (active_out) active_out[b] = tag;
      }
      s_alpha[tid] = (mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(s_beta[tid]) + best) : fminf(1.0f / best, s_norm[tid]);
    }
    __syncthreads();
    // ---- y = y0 + alpha N u                                   (reference constraint_module.py:512-514)
    if (P.n_is_identity) {
      for (int e = tid; e < TS * k; e += kWideThreads) {
        const int s = e / k, j = e - s * k;
        const long long b = tile * TS + s;
        if (b < B) y[b * k + j] = fmaf(s_alpha[s], us[j * TS + s], __ldg(y0 + j));
      }
    } else {
      const float* nt = blob + P.off_nt;
      for (int i = tid; i < k; i += kWideThreads) {
        float acc[TS];
        wide_dot<TS, BLK>(nt + i, P.k32, 0, n, us4, acc);
        const float c = __ldg(y0 + i);
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          const long long b = tile * TS + s;
          if (b < B) y[b * k + i] = fmaf(s_alpha[s], acc[s], c);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- backward
// Sum over the CTA (THREADS threads), same value in every thread, fixed order (deterministic).
template <int THREADS>
__device__ __forceinline__ float wide_block_sum(float x, float* red) {
  x = warp_sum32(x);
  __syncthreads();  // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) total += red[w];
  return total;
}

// One CTA per sample (grid-stride): the closed form of SURVEY 3.3 with the vectors in shared memory.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    wide_backward_kernel(const WideDev P, const float* __restrict__ v, long long ldv, const float* __restrict__ gy,
                         const float* __restrict__ kappa, const int* __restrict__ active, float* __restrict__ gv,
                         long long ldgv, long long B, int mode, const float* __restrict__ dkappa_lmi) {
  extern __shared__ __align__(16) float wide_smem[];
  const int n = P.n, k = P.k, np4 = n + 4;
  float* u = wide_smem;
  float* gz = u + np4;
  float* dk = gz + np4;
  float* tt = dk + np4;   // n + 2 entries used
  float* red = tt + np4;  // 8 words
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* blob = P.blob;
  const float* wt = blob + P.off_wt;
  const int* items = reinterpret_cast<const int*>(blob + P.off_items);

  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* vrow = v + b * ldv;
    const float* gyrow = gy + b * static_cast<long long>(k);
    float part = 0.f;
    for (int j = tid; j < n; j += THREADS) {
      const float x = __ldg(vrow + j);
      u[j] = x;
      part = fmaf(x, x, part);
      dk[j] = 0.f;
    }
    const float ss = wide_block_sum<THREADS>(part, red);
    const float s = sqrtf(ss);
    const float inv_norm = 1.0f / fmaxf(s, kNormEps);
    for (int j = tid; j < n; j += THREADS) u[j] *= inv_norm;
    const float beta = (mode == RAYEN_MODE_RAYEN_OLD) ? __ldg(vrow + n) : 0.f;
    // g_z = N' g_y
    if (P.n_is_identity) {
      for (int j = tid; j < n; j += THREADS) gz[j] = __ldg(gyrow + j);
    } else {
      const float* nrow = blob + P.off_nrow;
      for (int a = tid; a < n; a += THREADS) {
        float acc = 0.f;
#pragma unroll 8
        for (int i = 0; i < k; ++i) acc = fmaf(__ldg(nrow + static_cast<size_t>(i) * P.np + a), __ldg(gyrow + i), acc);
        gz[a] = acc;
      }
    }
    const float kap = __ldg(kappa + b);
    const int tag = __ldg(active + b);
    const int fam = tag_family(tag), idx = tag_index(tag);
    const bool boundary = (mode == RAYEN_MODE_RAYEN_OLD) ? (kap > 0.f) : (1.0f / kap < s);
    __syncthreads();  // u, gz, dk complete

    // ---- d kappa / du of the binding constraint (uniform per CTA)
    if (boundary && fam == RAYEN_FAM_LINEAR) {
      for (int j = tid; j < n; j += THREADS) dk[j] = __ldg(wt + static_cast<size_t>(j) * P.r_pad + idx);
    } else if (boundary && fam == RAYEN_FAM_LMI) {
      // left in the workspace by the eigen-solve of lmi_big.cuh: d kappa/du_a = q' F~z_a q
      if (dkappa_lmi)
        for (int j = tid; j < n; j += THREADS) dk[j] = dkappa_lmi[b * n + j];
    } else if (boundary && (fam == RAYEN_FAM_QUAD || fam == RAYEN_FAM_SOC)) {
      const int hdr = 2;  // tt[0], tt[1]: the header dot products (phi_z.u | c_z.u, h.u); tt[2 + i]: row i of the factor
      const int* it = items + (fam == RAYEN_FAM_QUAD ? idx : P.n_quad + idx) * 8;
      const int rb = __ldg(it), hrow = __ldg(it + 6);
      const float* wi = wt + rb;       // the n rows of the triangular factor
      const float* wh = wt + hrow;     // the two header rows
      // t = W_item u: a thread per row, coalesced over the rows; row i of the triangular factor starts at column i;
      // sixteen column loads in flight per thread (the walk is L2-latency-bound)
      for (int r = tid; r < hdr + n; r += THREADS) {
        const float* wp = (r < hdr) ? wh + r : wi + (r - hdr);
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
        int j = (r < hdr) ? 0 : r - hdr;
        for (; j + 16 <= n; j += 16) {
          float w[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) w[q] = __ldg(wp + static_cast<size_t>(j + q) * P.r_pad);
#pragma unroll
          for (int q = 0; q < 16; ++q) a4[q & 3] = fmaf(w[q], u[j + q], a4[q & 3]);
        }
        for (; j < n; ++j) a4[0] = fmaf(__ldg(wp + static_cast<size_t>(j) * P.r_pad), u[j], a4[0]);
        tt[r] = (a4[0] + a4[1]) + (a4[2] + a4[3]);
      }
      __syncthreads();
      float p2 = 0.f;
      for (int r = hdr + tid; r < hdr + n; r += THREADS) p2 = fmaf(tt[r], tt[r], p2);
      const float nrm2 = wide_block_sum<THREADS>(p2, red);
      float scale_g, scale_h = 0.f, scale_c = 0.f;  // dk = scale_g * T'(T u) + scale_h * row1 + scale_c * row0
      if (fam == RAYEN_FAM_QUAD) {
        const float root = sqrtf(nrm2);
        scale_g = root > 0.f ? 1.0f / root : 0.f;
        scale_c = 1.f;  // + phi_z
      } else {
        const float A = __int_as_float(__ldg(it + 5));
        const float cu = tt[0], hb = tt[1];
        const float cq = fmaf(-cu, cu, nrm2);
        float root;
        (void)soc_root(A, hb, cq, &root);
        // (kappa h + R'R u - (c.u) c) / sqrt(disc); 0 at a tangent ray, where the reference's autograd is NaN
        const float inv = root > 0.f ? 1.0f / root : 0.f;
        scale_g = inv;
        scale_h = kap * inv;
        scale_c = -cu * inv;
      }
      // (T'(T u))_a = sum_{r <= a} T[r][a] t_r: a warp per four components, lanes over the rows (coalesced), shuffle sums
      for (int a4 = warp * 4; a4 < n; a4 += (THREADS / 32) * 4) {
        const int amax = (a4 + 3 < n) ? a4 + 3 : n - 1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int r = lane; r <= amax; r += 32) {
          const float t = tt[hdr + r];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (a4 + q < n && r <= a4 + q)
              acc[q] = fmaf(__ldg(wi + static_cast<size_t>(a4 + q) * P.r_pad + r), t, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float g = warp_sum32(acc[q]);
          if (lane == 0 && a4 + q < n) {
            const float* col = wh + static_cast<size_t>(a4 + q) * P.r_pad;
            dk[a4 + q] = fmaf(scale_h, __ldg(col + 1), fmaf(scale_c, __ldg(col), scale_g * g));
          }
        }
      }
    }
    __syncthreads();

    // ---- tail: g_u, projection onto the tangent of the unit sphere, 1/|v|  (same as lqs.cuh backward_tail)
    float pz = 0.f;
    for (int j = tid; j < n; j += THREADS) pz = fmaf(gz[j], u[j], pz);
    const float gzu = wide_block_sum<THREADS>(pz, red);
    float* grow = gv + b * ldgv;
    float c1, c2;  // g_u = c1 g_z - c2 dk
    if (mode == RAYEN_MODE_RAYEN_OLD) {
      const float eb = expf(beta);
      const float alpha = 1.0f / (eb + kap);
      c1 = alpha;
      c2 = gzu * alpha * alpha;
      if (tid == 0) grow[n] = -c2 * eb;
    } else {
      if (!boundary) {
        for (int j = tid; j < n; j += THREADS) grow[j] = (s > 0.f) ? gz[j] : 0.f;
        continue;  // uniform per CTA
      }
      const float ik = 1.0f / kap;
      c1 = ik;
      c2 = gzu * ik * ik;
    }
    float pu = 0.f;
    for (int j = tid; j < n; j += THREADS) {
      const float g = fmaf(c1, gz[j], -c2 * dk[j]);
      dk[j] = g;  // own entries only
      pu = fmaf(g, u[j], pu);
    }
    float guu = wide_block_sum<THREADS>(pu, red);
    if (s < kNormEps) guu = 0.f;
    for (int j = tid; j < n; j += THREADS) grow[j] = (dk[j] - guu * u[j]) * inv_norm;
  }
}

}  // namespace rayen
