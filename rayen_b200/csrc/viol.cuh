// rayen_violation_f32: max constraint residual of y against the ORIGINAL constraints (ambient space), float32.
// Not on the training hot path: it is the on-device version of the violation metric (BASELINE.json's second
// metric; the reference computes it per sample with a cvxpy projection, constraints.py:549-559).
#pragma once
#include "common.cuh"
#include "lmi.cuh"

namespace rayen {

constexpr int kViolThreads = 256;

// init + row . y over k4 (multiple of 4) entries.  Blocks of 64 terms are summed in float32 and the block sums added
// with a compensated (Neumaier) sum: the metric of a 4000-dimensional row then carries ~1e-7 of the terms' scale
// instead of the ~1e-5 of a plain running sum, i.e. it can certify the 1e-5 feasibility bar in wide sets too.
__device__ __forceinline__ float viol_dot(const float* __restrict__ row, const float* __restrict__ yb, int k4, float init) {
  float tot = init, comp = 0.f;
  for (int i0 = 0; i0 < k4; i0 += 64) {
    const int i1 = (i0 + 64 < k4) ? i0 + 64 : k4;
    float acc = 0.f;
    for (int i = i0; i < i1; i += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(row + i));
      const float4 yy = ld4(yb + i);
      acc = fmaf(a.x, yy.x, fmaf(a.y, yy.y, fmaf(a.z, yy.z, fmaf(a.w, yy.w, acc))));
    }
    const float sum = tot + acc;
    const float bp = sum - tot;
    comp += (tot - (sum - bp)) + (acc - bp);
    tot = sum;
  }
  return tot + comp;
}

// One warp per sample; lanes split the rows of every constraint; y sits in shared memory.
__global__ void __launch_bounds__(kViolThreads)
    viol_lqs_kernel(const PlanDev P, const float* __restrict__ y, long long ldy, float* __restrict__ viol, long long B) {
  extern __shared__ __align__(16) float ybuf_all[];
  const int k = P.k, k4 = (k + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* yb = ybuf_all + warp * k4;
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int row_w = k4 + 4;
  for (long long b = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp; b < B; b += n_warps) {
    __syncwarp();
    for (int i = lane; i < k4; i += 32) yb[i] = (i < k) ? __ldg(y + b * ldy + i) : 0.f;
    __syncwarp();
    float worst = -3.0e38f;
    const float* p = P.blob + P.off_viol;
    // A1 y <= b1
    for (int j = lane; j < P.viol_in; j += 32) {
      const float* row = p + static_cast<size_t>(j) * row_w;
      const float acc = viol_dot(row, yb, k4, 0.f);
      worst = fmaxf(worst, acc - __ldg(row + k4));
    }
    p += static_cast<size_t>(P.viol_in) * row_w;
    // A2 y = b2
    for (int j = lane; j < P.viol_eq; j += 32) {
      const float* row = p + static_cast<size_t>(j) * row_w;
      const float acc = viol_dot(row, yb, k4, 0.f);
      worst = fmaxf(worst, fabsf(acc - __ldg(row + k4)));
    }
    p += static_cast<size_t>(P.viol_eq) * row_w;
    // (1/2) y'P y + q'y + r <= 0
    for (int qi = 0; qi < P.n_quad; ++qi) {
      float part = 0.f;
      for (int r = lane; r < k; r += 32) {
        const float* row = p + static_cast<size_t>(r) * k4;
        const float acc = viol_dot(row, yb, k4, 0.f);
        part = fmaf(yb[r], fmaf(0.5f, acc, __ldg(p + static_cast<size_t>(k4) * k4 + r)), part);
      }
      part = group_sum<32>(part);
      worst = fmaxf(worst, part + __ldg(p + static_cast<size_t>(k4) * k4 + k4));
      p += static_cast<size_t>(k4) * k4 + k4 + 4;
    }
    // ||M y + s|| - c'y - d <= 0
    for (int si = 0; si < P.n_soc; ++si) {
      const int rm = static_cast<int>(__ldg(p));
      const float d = __ldg(p + 1);
      const float* c = p + 4;
      const float* rows = c + k4;
      float cy = 0.f, ss = 0.f;
      for (int i = lane; i < k; i += 32) cy = fmaf(__ldg(c + i), yb[i], cy);
      for (int r = lane; r < rm; r += 32) {
        const float* row = rows + static_cast<size_t>(r) * row_w;
        const float acc = viol_dot(row, yb, k4, __ldg(row + k4));
        ss = fmaf(acc, acc, ss);
      }
      cy = group_sum<32>(cy);
      ss = group_sum<32>(ss);
      worst = fmaxf(worst, sqrtf(ss) - cy - d);
      p = rows + static_cast<size_t>(rm) * row_w;
    }
    worst = group_max<32>(worst);
    if (lane == 0) viol[b] = worst;
  }
}

// relu(-lambda_min(F(y))) with the LMI solver on S = -F(y); merged into viol[b] (runs after viol_lqs_kernel).
template <int RP, bool F_SMEM>
__global__ void __launch_bounds__(kLmiThreads, 1)
    viol_lmi_kernel(const PlanDev P, const float* __restrict__ y, long long ldy, float* __restrict__ viol, long long B) {
  using C = LmiCfg<RP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  const int k = P.k;
  const float* F;
  float* scratch_base;
  const int words = (k + 1) * RP * RP;
  if constexpr (F_SMEM) {
    float* fs = reinterpret_cast<float*>(smem_raw + 64);
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) stage_bulk(fs, P.blob + P.off_lmineg, words, &bars[0]);
    F = fs;
    scratch_base = fs + words;
  } else {
    F = P.blob + P.off_lmineg;
    scratch_base = reinterpret_cast<float*>(smem_raw + 64);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LmiSolver<RP, false, F_SMEM> S;
  S.q = lane % C::LPM;
  S.grp_base = lane - S.q;
  const int grp = lane / C::LPM;
  S.scr = scratch_base + (warp * C::MPW + grp) * C::SCR;
  const long long warp_id = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  const long long n_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  bool staged = !F_SMEM;
  for (long long base = warp_id * C::MPW; base < B; base += n_warps * C::MPW) {
    const long long b = base + grp;
    const bool valid = b < B;
    if (!staged) {
      mbar_wait(&bars[0], 0);
      staged = true;
    }
    S.contract_affine(F, y + (valid ? b : 0) * ldy, k, valid);
    S.tridiagonalize();
    const float lam = S.lambda_max_relu();
    if (valid && S.q == 0) viol[b] = fmaxf(viol[b], lam);
    __syncwarp();
  }
  if constexpr (F_SMEM) {
    if (!staged) mbar_wait(&bars[0], 0);
  }
}

}  // namespace rayen
