// Wide sets (32 < n <= 4096): linear + quadratic + SOC kappa, the scale step and the closed-form backward when a
// direction no longer fits a thread's registers.  Same arithmetic as lqs.cuh (reference constraint_module.py:351-399,
// :468-474, :512-514), different mapping: the directions of a tile of samples sit in shared memory, every constraint is
// a set of rows of ONE matrix W (plan section WIDE, stored transposed so that 32 lanes read 32 consecutive rows of a
// column with one coalesced load), a lane owns a row and carries its dot products with the kWideTS samples of the
// tile in registers, and a warp owns a task (up to 128 linear rows, or one quadratic / cone).
#pragma once
#include "common.cuh"
#include "lqs.cuh"

namespace rayen {

constexpr int kWideTS = 8;            // samples per CTA tile (forward)
constexpr int kWideThreads = 256;     // forward: 8 warps = 8 samples staged, 8 tasks in flight
constexpr int kWideWarps = kWideThreads / 32;
constexpr int kWideBwdThreads = 128;  // backward: one CTA per sample
constexpr int kWideMagic = 0x57494445;

struct WideDev {
  const float* blob;
  int n, k;
  int r_pad, n_tasks, off_tasks, off_wt, off_nt, off_nrow, k32, np, off_items, n_quad, n_soc, off_soc_a;
  int off_y0, n_is_identity;
};

__host__ __device__ inline size_t wide_fwd_smem_bytes(int n) {
  return (static_cast<size_t>(n) * kWideTS + 4 * kWideTS + 2 * kWideWarps * kWideTS) * sizeof(float);
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

// acc[s] = sum_j Wt[j][row] * us[j][s]: the column walk of one row against the tile's directions
__device__ __forceinline__ void wide_dot(const float* __restrict__ wcol, int r_pad, int n, const float4* __restrict__ us4,
                                         float (&acc)[kWideTS]) {
#pragma unroll
  for (int s = 0; s < kWideTS; ++s) acc[s] = 0.f;
  int j = 0;
  for (; j + 4 <= n; j += 4) {
    float w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = __ldg(wcol + static_cast<size_t>(j + q) * r_pad);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = us4[2 * (j + q)], b = us4[2 * (j + q) + 1];
      acc[0] = fmaf(w[q], a.x, acc[0]);
      acc[1] = fmaf(w[q], a.y, acc[1]);
      acc[2] = fmaf(w[q], a.z, acc[2]);
      acc[3] = fmaf(w[q], a.w, acc[3]);
      acc[4] = fmaf(w[q], b.x, acc[4]);
      acc[5] = fmaf(w[q], b.y, acc[5]);
      acc[6] = fmaf(w[q], b.z, acc[6]);
      acc[7] = fmaf(w[q], b.w, acc[7]);
    }
  }
  for (; j < n; ++j) {
    const float w = __ldg(wcol + static_cast<size_t>(j) * r_pad);
    const float4 a = us4[2 * j], b = us4[2 * j + 1];
    acc[0] = fmaf(w, a.x, acc[0]);
    acc[1] = fmaf(w, a.y, acc[1]);
    acc[2] = fmaf(w, a.z, acc[2]);
    acc[3] = fmaf(w, a.w, acc[3]);
    acc[4] = fmaf(w, b.x, acc[4]);
    acc[5] = fmaf(w, b.y, acc[5]);
    acc[6] = fmaf(w, b.z, acc[6]);
    acc[7] = fmaf(w, b.w, acc[7]);
  }
}

// value of x[lane] for lane < kWideTS without dynamic register indexing
__device__ __forceinline__ float pick_lane(const float (&x)[kWideTS], int lane) {
  float r = x[0];
#pragma unroll
  for (int s = 1; s < kWideTS; ++s)
    if (lane == s) r = x[s];
  return r;
}

// ----------------------------------------------------------------------------- forward
// grid: one CTA per tile of kWideTS samples (grid-stride); dynamic smem = wide_fwd_smem_bytes(n).
__global__ void __launch_bounds__(kWideThreads)
    wide_forward_kernel(const WideDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                        float* __restrict__ kappa_out, int* __restrict__ active_out, long long B, int mode) {
  extern __shared__ __align__(16) float wide_smem[];
  const int n = P.n, k = P.k;
  float* us = wide_smem;                                  // [n][kWideTS]
  float* s_norm = us + static_cast<size_t>(n) * kWideTS;  // [kWideTS]
  float* s_beta = s_norm + kWideTS;
  float* s_alpha = s_beta + kWideTS;
  float* s_pad = s_alpha + kWideTS;
  float* wbest = s_pad + kWideTS;                         // [kWideWarps][kWideTS]
  int* wtag = reinterpret_cast<int*>(wbest + kWideWarps * kWideTS);
  const float4* us4 = reinterpret_cast<const float4*>(us);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* blob = P.blob;
  const int* tasks = reinterpret_cast<const int*>(blob + P.off_tasks);
  const float* wt = blob + P.off_wt;
  const float* y0 = blob + P.off_y0;
  const long long n_tiles = (B + kWideTS - 1) / kWideTS;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();  // the previous tile's readers of us / s_alpha are done
    // ---- stage: warp w normalises sample w of the tile (reference constraint_module.py:470)
    {
      const long long b = tile * kWideTS + warp;
      const bool valid = b < B;
      const float* vrow = v + (valid ? b : 0) * ldv;
      float ss = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float x = valid ? __ldg(vrow + j) : 0.f;
        us[j * kWideTS + warp] = x;
        ss = fmaf(x, x, ss);
      }
      ss = warp_sum32(ss);
      const float s = sqrtf(ss);
      const float inv = 1.0f / fmaxf(s, kNormEps);
      __syncwarp();
      for (int j = lane; j < n; j += 32) us[j * kWideTS + warp] *= inv;
      if (lane == 0) {
        s_norm[warp] = s;
        s_beta[warp] = (valid && mode == RAYEN_MODE_RAYEN_OLD) ? __ldg(vrow + n) : 0.f;
      }
    }
    __syncthreads();

    // ---- tasks: lanes 0..kWideTS-1 of a warp keep the warp's running (kappa, tag) of sample `lane`
    float wb = 0.f;
    int wtg = 0;
    for (int t = warp; t < P.n_tasks; t += kWideWarps) {
      const int kind = __ldg(tasks + t * 8 + 0), rb = __ldg(tasks + t * 8 + 1), ng = __ldg(tasks + t * 8 + 2),
                idx = __ldg(tasks + t * 8 + 3);
      const float A = __int_as_float(__ldg(tasks + t * 8 + 4));
      float cand;
      int ctag;
      if (kind == 1) {
        // linear rows: kappa_j = D_j . u                       (reference constraint_module.py:353)
        float mx[kWideTS];
        int mr[kWideTS];
#pragma unroll
        for (int s = 0; s < kWideTS; ++s) {
          mx[s] = 0.f;
          mr[s] = 0;
        }
        for (int g = 0; g < ng; ++g) {
          float acc[kWideTS];
          wide_dot(wt + rb + g * 32 + lane, P.r_pad, n, us4, acc);
#pragma unroll
          for (int s = 0; s < kWideTS; ++s)
            if (acc[s] > mx[s]) {
              mx[s] = acc[s];
              mr[s] = idx + g * 32 + lane;
            }
        }
#pragma unroll
        for (int s = 0; s < kWideTS; ++s) group_argmax(mx[s], mr[s], 32);  // ties -> lowest row
        cand = pick_lane(mx, lane);
        int row = mr[0];
#pragma unroll
        for (int s = 1; s < kWideTS; ++s)
          if (lane == s) row = mr[s];
        ctag = make_tag(RAYEN_FAM_LINEAR, row);
      } else {
        // quadratic {phi_z | G}: kappa = phi_z.u + |G u|;  cone {c_z, h | R}: root of the quadratic (:360-399)
        const int hdr = (kind == 2) ? 1 : 2;
        float sq[kWideTS], a0[kWideTS], a1[kWideTS];
#pragma unroll
        for (int s = 0; s < kWideTS; ++s) sq[s] = a0[s] = a1[s] = 0.f;
        for (int g = 0; g < ng; ++g) {
          float acc[kWideTS];
          wide_dot(wt + rb + g * 32 + lane, P.r_pad, n, us4, acc);
          const int rl = g * 32 + lane;
#pragma unroll
          for (int s = 0; s < kWideTS; ++s) {
            if (rl == 0) a0[s] = acc[s];
            else if (rl < hdr) a1[s] = acc[s];
            else sq[s] = fmaf(acc[s], acc[s], sq[s]);
          }
        }
        float kap[kWideTS];
#pragma unroll
        for (int s = 0; s < kWideTS; ++s) {
          const float nrm2 = warp_sum32(sq[s]);
          const float h0 = __shfl_sync(0xffffffffu, a0[s], 0);
          const float h1 = __shfl_sync(0xffffffffu, a1[s], 1);
          if (kind == 2) {
            kap[s] = h0 + sqrtf(nrm2);
          } else {
            const float cq = fmaf(-h0, h0, nrm2);
            kap[s] = soc_root(A, h1, cq, nullptr);
          }
        }
        cand = pick_lane(kap, lane);
        ctag = make_tag(kind == 2 ? RAYEN_FAM_QUAD : RAYEN_FAM_SOC, idx);
      }
      if (cand > wb) {  // tasks arrive in tag order within a warp: strict > keeps the lowest tag
        wb = cand;
        wtg = ctag;
      }
    }
    if (lane < kWideTS) {
      wbest[warp * kWideTS + lane] = wb;
      wtag[warp * kWideTS + lane] = wtg;
    }
    __syncthreads();
    // ---- combine the warps (ties -> lowest tag = the reference's evaluation order), alpha (:472-474 / :464-465)
    if (tid < kWideTS) {
      float best = 0.f;
      int tag = 0;
      for (int w = 0; w < kWideWarps; ++w) {
        const float c = wbest[w * kWideTS + tid];
        const int ct = wtag[w * kWideTS + tid];
        if (c > best || (c == best && c > 0.f && ct < tag)) {
          best = c;
          tag = ct;
        }
      }
      const long long b = tile * kWideTS + tid;
      if (b < B) {
        if (kappa_out) kappa_out[b] = best;
        if (active_out) active_out[b] = tag;
      }
      s_alpha[tid] = (mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(s_beta[tid]) + best) : fminf(1.0f / best, s_norm[tid]);
    }
    __syncthreads();
    // ---- y = y0 + alpha N u                                   (reference constraint_module.py:512-514)
    if (P.n_is_identity) {
      for (int e = tid; e < kWideTS * k; e += kWideThreads) {
        const int s = e / k, j = e - s * k;
        const long long b = tile * kWideTS + s;
        if (b < B) y[b * k + j] = fmaf(s_alpha[s], us[j * kWideTS + s], __ldg(y0 + j));
      }
    } else {
      const float* nt = blob + P.off_nt;
      for (int i = tid; i < k; i += kWideThreads) {
        float acc[kWideTS];
        wide_dot(nt + i, P.k32, n, us4, acc);
        const float c = __ldg(y0 + i);
#pragma unroll
        for (int s = 0; s < kWideTS; ++s) {
          const long long b = tile * kWideTS + s;
          if (b < B) y[b * k + i] = fmaf(s_alpha[s], acc[s], c);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- backward
// Sum over the CTA (kWideBwdThreads threads), same value in every thread, fixed order (deterministic).
__device__ __forceinline__ float wide_block_sum(float x, float* red) {
  x = warp_sum32(x);
  __syncthreads();  // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < kWideBwdThreads / 32; ++w) total += red[w];
  return total;
}

// One CTA per sample (grid-stride): the closed form of SURVEY 3.3 with the vectors in shared memory.
__global__ void __launch_bounds__(kWideBwdThreads)
    wide_backward_kernel(const WideDev P, const float* __restrict__ v, long long ldv, const float* __restrict__ gy,
                         const float* __restrict__ kappa, const int* __restrict__ active, float* __restrict__ gv,
                         long long ldgv, long long B, int mode) {
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
    for (int j = tid; j < n; j += kWideBwdThreads) {
      const float x = __ldg(vrow + j);
      u[j] = x;
      part = fmaf(x, x, part);
      dk[j] = 0.f;
    }
    const float ss = wide_block_sum(part, red);
    const float s = sqrtf(ss);
    const float inv_norm = 1.0f / fmaxf(s, kNormEps);
    for (int j = tid; j < n; j += kWideBwdThreads) u[j] *= inv_norm;
    const float beta = (mode == RAYEN_MODE_RAYEN_OLD) ? __ldg(vrow + n) : 0.f;
    // g_z = N' g_y
    if (P.n_is_identity) {
      for (int j = tid; j < n; j += kWideBwdThreads) gz[j] = __ldg(gyrow + j);
    } else {
      const float* nrow = blob + P.off_nrow;
      for (int a = tid; a < n; a += kWideBwdThreads) {
        float acc = 0.f;
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
      for (int j = tid; j < n; j += kWideBwdThreads) dk[j] = __ldg(wt + static_cast<size_t>(j) * P.r_pad + idx);
    } else if (boundary && (fam == RAYEN_FAM_QUAD || fam == RAYEN_FAM_SOC)) {
      const int hdr = (fam == RAYEN_FAM_QUAD) ? 1 : 2;
      const int rb = __ldg(items + (fam == RAYEN_FAM_QUAD ? idx : P.n_quad + idx));
      const float* wi = wt + rb;
      // t = W_item u: a thread per row, coalesced over the rows
      for (int r = tid; r < hdr + n; r += kWideBwdThreads) {
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(__ldg(wi + static_cast<size_t>(j) * P.r_pad + r), u[j], acc);
        tt[r] = acc;
      }
      __syncthreads();
      float p2 = 0.f;
      for (int r = hdr + tid; r < hdr + n; r += kWideBwdThreads) p2 = fmaf(tt[r], tt[r], p2);
      const float nrm2 = wide_block_sum(p2, red);
      float scale_g, scale_h = 0.f, scale_c = 0.f;  // dk = scale_g * T'(T u) + scale_h * row1 + scale_c * row0
      if (fam == RAYEN_FAM_QUAD) {
        const float root = sqrtf(nrm2);
        scale_g = root > 0.f ? 1.0f / root : 0.f;
        scale_c = 1.f;  // + phi_z
      } else {
        const float A = __ldg(blob + P.off_soc_a + idx);
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
      // (T'(T u))_a: a warp per component, lanes over the rows (coalesced), shuffle sum
      for (int a = warp; a < n; a += kWideBwdThreads / 32) {
        const float* col = wi + static_cast<size_t>(a) * P.r_pad;
        float acc = 0.f;
        for (int r = hdr + lane; r < hdr + n; r += 32) acc = fmaf(__ldg(col + r), tt[r], acc);
        acc = warp_sum32(acc);
        if (lane == 0) {
          float d = scale_g * acc + scale_c * __ldg(col);
          if (hdr == 2) d = fmaf(scale_h, __ldg(col + 1), d);
          dk[a] = d;
        }
      }
    }
    __syncthreads();

    // ---- tail: g_u, projection onto the tangent of the unit sphere, 1/|v|  (same as lqs.cuh backward_tail)
    float pz = 0.f;
    for (int j = tid; j < n; j += kWideBwdThreads) pz = fmaf(gz[j], u[j], pz);
    const float gzu = wide_block_sum(pz, red);
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
        for (int j = tid; j < n; j += kWideBwdThreads) grow[j] = (s > 0.f) ? gz[j] : 0.f;
        continue;  // uniform per CTA
      }
      const float ik = 1.0f / kap;
      c1 = ik;
      c2 = gzu * ik * ik;
    }
    float pu = 0.f;
    for (int j = tid; j < n; j += kWideBwdThreads) {
      const float g = fmaf(c1, gz[j], -c2 * dk[j]);
      dk[j] = g;  // own entries only
      pu = fmaf(g, u[j], pu);
    }
    float guu = wide_block_sum(pu, red);
    if (s < kNormEps) guu = 0.f;
    for (int j = tid; j < n; j += kWideBwdThreads) grow[j] = (dk[j] - guu * u[j]) * inv_norm;
  }
}

}  // namespace rayen
