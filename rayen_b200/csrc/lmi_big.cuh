// LMI family beyond the register-resident kernels: LMI sizes 33 <= r <= 320 with any subspace dimension, and LMIs of any
// size together with a wide subspace (n > 32).  Reference: constraint_module.py:401-449 (S = einsum(all_F, rho),
// L'(-S)L, eigvalsh, relu(lambda_max)), which takes any size; its own sweep (examples/scripts/time_analysis.py:159-175)
// runs r_F = 10 ... 300 with k = 100 ... 10000.
//
// Two kernels per batch chunk, behind the kernel of the other families (lqs*.cuh or wide.cuh), which has already left
// (kappa, tag, y) of the linear / quadratic / SOC constraints ("the prior"):
//   1. lmib_contract_kernel: S~(v) = sum_a v_a F~z_a for a chunk of samples as ONE FP32 GEMM
//        C[B_c x P] = V[B_c x n] . F[n x P],   P = r (r + 1) / 2 padded to 4  (plan section LMIB: the lower triangles of
//      the congruence-transformed, subspace-folded matrices F~z_a of plan.py, row-major packed),
//      128 x 128 x 8 tiles, 8 x 8 outputs per thread, operands staged through shared memory with a register prefetch.
//      The result goes to the workspace (L2 / HBM): B_c x P floats; forward_impl cuts the batch into chunks that fit.
//      (S~ is linear in v: the eigen kernel divides by |v| -- lambda_max(S~(u)) = lambda_max(S~(v)) / |v|.)
//   2. lmib_solve_kernel<THREADS>: one CTA per sample.
//        a. Wolkowicz-Styan bound straight from the packed entries, in the cancellation-free centred form
//           (mean + sqrt((r-1)/r (sum_i (S_ii - mean)^2 + 2 sum_{i>j} S_ij^2))): below the prior kappa => the LMI cannot
//           bind, the sample is finished (nothing is written).
//        a'. exact definiteness filter: LDL' of (kprior - margin) I - S~(u) on the packed triangle; all pivots positive
//           => the LMI cannot bind, the sample is finished after r^3/6 FMAs (a failure falls through to the solve);
//        b. the packed lower triangle (row i at word i (i + 1) / 2) is copied to shared memory as it is, scaled by
//           1 / |v|: r (r + 1) / 2 words, 205 KB at r = 320.  Thread i owns row i.  Entry (i, j), j <= i, of 32
//           consecutive rows falls into 32 different banks (triangular numbers are a complete residue system modulo a
//           power of two), entry (j, i), j > i, of 32 consecutive i is 32 consecutive words: both walks are conflict-free;
//        c. Householder tridiagonalisation: p = tau A v as a walk along row i up to the diagonal and down column i below
//           it, rank-2 update of the lower triangle of the trailing block, the reflector stays in column k
//           (back-transform);
//        d. lambda_max by multisection: every thread probes one shift with the pivot recurrence of T - xI (all pivots
//           negative <=> x above the spectrum), THREADS sections per round;
//        e. merge with the prior (ties keep the earlier family, like torch.max), and when the LMI binds: kappa, tag, and
//           y <- y0 + (y - y0) alpha_new / alpha_prior (the other families' kernel wrote y with alpha_prior; both alphas
//           are recomputed here from kappa and |v| exactly as that kernel did);
//        f. with gradients: top eigenvector by twisted factorisation of T - lambda I, back-transformed through the
//           reflectors, d kappa/du_a = q' F~z_a q = <F_a packed, w>, w_e = (2 - [i = j]) q_i q_j, one warp per a.
// Deterministic: fixed-order block reductions, no atomics.
//
// Free of PTX so that tests/test_lmi_big_emulated.py can compile both kernels for the host under the SIMT emulator.
#pragma once
#ifndef RAYEN_EMU
#include "common.cuh"
#define LB_STATIC_SHARED __shared__
#else
#define LB_STATIC_SHARED static   // the emulator runs one block at a time: a function-local static is the block's shared memory
#endif

namespace rayen {

constexpr int kLbMaxR = 320;          // largest LMI of this path (thread per row, <= 320 threads; 215 KB of shared memory)
constexpr int kLbTileM = 128, kLbTileN = 128, kLbTileK = 8, kLbGemmThreads = 256;

struct LmiBigDev {
  const float* blob;
  int n, k, r;
  int p4;          // words per packed matrix: r (r + 1) / 2 rounded up to 4
  int off_lmib;    // [n][p4]
  int off_y0;
};

__host__ __device__ inline int lmib_tri(int i) { return i * (i + 1) / 2; }  // first word of packed row i
// vectors in front of the matrix: vv, ww, d, e, tau, dp, dm, q  (r4 each) + 40 words of reduction scratch
__host__ __device__ inline size_t lmib_vec_words(int r) { return static_cast<size_t>(8) * ((r + 3) / 4 * 4) + 40; }
__host__ __device__ inline size_t lmib_smem_bytes(int r) {
  const size_t p4 = (static_cast<size_t>(r) * (r + 1) / 2 + 3) / 4 * 4;
  return (lmib_vec_words(r) + p4) * sizeof(float);
}

// ----------------------------------------------------------------------------- 1. contraction GEMM
// C[b][e] = C0[e] + sum_a V[b][a] F[a][e]   (b < Bc, e < p4, a < n); V rows ldv apart, F rows p4 apart (16-byte
// aligned); C0 (nullable): a constant row, the F_k of the violation metric's affine pencil F(y) = sum_i y_i F_i + F_k.
__global__ void __launch_bounds__(kLbGemmThreads)
    lmib_contract_kernel(const float* __restrict__ V, long long ldv, const float* __restrict__ F, int n, int p4,
                         float* __restrict__ C, long long Bc, const float* __restrict__ C0) {
  LB_STATIC_SHARED __align__(16) float As[2][kLbTileK][kLbTileM];  // [k][sample]
  LB_STATIC_SHARED __align__(16) float Bs[2][kLbTileK][kLbTileN];  // [k][entry]
  const int tid = threadIdx.x;
  const int tiles_n = (p4 + kLbTileN - 1) / kLbTileN;  // grid: tiles_n x ceil(Bc / 128) blocks, flattened
  const long long m0 = static_cast<long long>(blockIdx.x / tiles_n) * kLbTileM;
  const int n0 = static_cast<int>(blockIdx.x % tiles_n) * kLbTileN;
  // loaders: A tile 128 samples x 8 k: thread t loads sample t / 2, k = 4 (t % 2) .. + 3 (four scalar loads: rows of v
  // need not be 16-byte aligned); B tile 8 k x 128 entries: thread t loads k = t / 32, entries 4 (t % 32) .. + 3
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;
  const int b_k = tid >> 5, b_col = (tid & 31) * 4;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 8 x 8 outputs each: rows ty*4.. and 64+ty*4.., cols likewise
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[4];
  float4 rb;
  const bool a_ok = m0 + a_row < Bc;
  const float* vrow = V + (m0 + a_row) * ldv;
  const bool b_ok = n0 + b_col < p4;  // p4 is a multiple of 4: a 4-group is in or out as a whole

  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) ra[q] = (a_ok && k0 + a_k + q < n) ? __ldg(vrow + k0 + a_k + q) : 0.f;
    if (b_ok && k0 + b_k < n)
      rb = __ldg(reinterpret_cast<const float4*>(F + static_cast<size_t>(k0 + b_k) * p4 + n0 + b_col));
    else
      rb = float4{0.f, 0.f, 0.f, 0.f};
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) As[buf][a_k + q][a_row] = ra[q];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = rb;
  };

  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < n; k0 += kLbTileK) {
    const bool more = k0 + kLbTileK < n;
    if (more) fetch(k0 + kLbTileK);
#pragma unroll
    for (int kk = 0; kk < kLbTileK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= Bc) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + h * 64 + tx * 4;
      if (col < p4) {
        float4 c0 = float4{0.f, 0.f, 0.f, 0.f};
        if (C0) c0 = __ldg(reinterpret_cast<const float4*>(C0 + col));
        *reinterpret_cast<float4*>(C + static_cast<size_t>(row) * p4 + col) =
            float4{acc[i][4 * h + 0] + c0.x, acc[i][4 * h + 1] + c0.y, acc[i][4 * h + 2] + c0.z, acc[i][4 * h + 3] + c0.w};
      }
    }
  }
}

// ----------------------------------------------------------------------------- 2. eigen-solve, one CTA per sample
__device__ __forceinline__ float lb_warp_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
// Sum over the CTA, same value in every thread, fixed order.  `red`: 16 words nobody else uses.
template <int THREADS>
__device__ __forceinline__ float lb_block_sum(float x, float* red) {
  x = lb_warp_sum(x);
  __syncthreads();  // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) total += red[w];
  return total;
}
template <int THREADS>
__device__ __forceinline__ float lb_block_max(float x, float* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float total = red[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) total = fmaxf(total, red[w]);
  return total;
}
template <int THREADS>
__device__ __forceinline__ int lb_block_min_int(int x, int* red) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  int total = red[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) total = min(total, red[w]);
  return total;
}

constexpr int kLbFlagGrad = 1;       // leave d kappa/du of the LMI-bound boundary samples in dkappa
constexpr int kLbFlagGradOnly = 2;   // backward without a forward-computed gradient: kappa / active / y are inputs only
constexpr int kLbFlagLambdaOut = 4;  // violation metric: out = max(out, relu(lambda_max)); no prior, no y

// Scale step of the sample as the other families' kernels compute it (lqs.cuh / wide.cuh forward; reference :472-474,
// :464-465): RAYEN alpha = min(1 / kappa, |v|), RAYEN_old alpha = 1 / (e^beta + kappa).
__device__ __forceinline__ float lb_alpha(float kap, float s, float beta, int mode) {
  return (mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(beta) + kap) : fminf(1.0f / kap, s);
}

// (minimum CTAs per SM: keeps the register count where the shared-memory footprint, not the register file, limits the
// number of matrices in flight -- 128 registers at 128 threads cost one of five CTAs per SM at r = 100: 1.20 -> 1.41 ms)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 64) ? 10 : ((THREADS <= 128) ? 5 : ((THREADS <= 256) ? 2 : 1)))
    lmib_solve_kernel(const LmiBigDev P, const float* S, const float* __restrict__ v, long long ldv,
                      float* __restrict__ y, float* __restrict__ kappa_io, int* __restrict__ active_io,
                      float* __restrict__ dkappa, long long Bc, int mode, int flags, float* Sw,  // (Sw may be S itself)
                      int* __restrict__ grad_list, int* __restrict__ grad_count) {
  extern __shared__ __align__(16) float lmib_smem[];
  const int r = P.r, r4 = (r + 3) / 4 * 4, n = P.n, k = P.k, p4 = P.p4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* vv = lmib_smem;
  float* ww = vv + r4;
  float* sd = ww + r4;
  float* se = sd + r4;
  float* stau = se + r4;
  float* dp = stau + r4;
  float* dm = dp + r4;
  float* sq = dm + r4;
  float* red = sq + r4;           // 16 words
  int* redi = reinterpret_cast<int*>(red + 16);  // 16 words
  float* A = red + 40;            // packed lower triangle, p4 words (16-byte aligned)
  const bool want_grad = (flags & (kLbFlagGrad | kLbFlagGradOnly)) != 0;
  const bool grad_only = (flags & kLbFlagGradOnly) != 0;
  const bool lambda_out = (flags & kLbFlagLambdaOut) != 0;
  const int ti = lmib_tri(tid);   // row tid of the packed matrix

  for (long long b = blockIdx.x; b < Bc; b += gridDim.x) {
    __syncthreads();  // the previous sample's readers of the shared vectors are done
    const float* srow = S + static_cast<size_t>(b) * p4;
    const float* vrow = v + b * ldv;
    // ---- |v|
    float part = 0.f;
    for (int j = tid; j < n; j += THREADS) {
      const float x = __ldg(vrow + j);
      part = fmaf(x, x, part);
    }
    const float s = sqrtf(lb_block_sum<THREADS>(part, red));
    const float inv = lambda_out ? 1.0f : 1.0f / fmaxf(s, kNormEps);
    const float beta = (mode == RAYEN_MODE_RAYEN_OLD && !lambda_out) ? __ldg(vrow + n) : 0.f;
    const float kprior = lambda_out ? 0.f : kappa_io[b];
    const int tprior = lambda_out ? 0 : active_io[b];
    if (grad_only) {
      // only the samples whose LMI binds on the boundary need anything
      bool need = tag_family(tprior) == RAYEN_FAM_LMI && kprior > 0.f;
      if (need && mode == RAYEN_MODE_RAYEN) need = (1.0f / kprior < s);
      if (!need) continue;  // uniform per CTA
    }

    // ---- b. the packed matrix into shared memory, scaled to the unit direction (16-byte loads: rows of S are 16-byte
    //         aligned, p4 is a multiple of 4)
    for (int e = tid; e < p4 / 4; e += THREADS) {
      float4 x = *reinterpret_cast<const float4*>(srow + 4 * e);
      x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
      *reinterpret_cast<float4*>(A + 4 * e) = x;
    }
    __syncthreads();
    // ---- a. pruning bound (diagonal entry of row i at tri(i) + i): lambda_max <= mean + sqrt((r-1)/r dev2), dev2 the
    //         squared Frobenius norm of the trace-free part as a sum of squares
    float fro;
    {
      const float tr = (tid < r) ? A[ti + tid] : 0.f;
      const float mean = lb_block_sum<THREADS>(tr, red) / static_cast<float>(r);
      float dev = 0.f;
      if (tid < r) {
        for (int j = 0; j < tid; ++j) {
          const float x = A[ti + j];
          dev = fmaf(2.f * x, x, dev);
        }
        const float x = A[ti + tid] - mean;
        dev = fmaf(x, x, dev);
      }
      const float dev2 = lb_block_sum<THREADS>(dev, red);
      fro = sqrtf(fmaf(static_cast<float>(r) * mean, mean, dev2));
      if (!grad_only) {
        // allowance for the float32 sums: 2e-6 of the scale involved
        const float rad = sqrtf(dev2 * (static_cast<float>(r - 1) / static_cast<float>(r)));
        const float ub = (mean + rad) + 2e-6f * (fabsf(mean) + rad);
        if (ub <= kprior) continue;  // uniform per CTA: the LMI cannot bind (kprior = 0: lambda_max <= 0, kappa_LMI = 0)
      }
    }

    // ---- a'. exact definiteness filter (the idea of lmi_warp.cuh at this size): LDL' of M = tau I - S~(u) on the packed
    //          triangle, tau = kprior - margin (kprior + |S~|_F).  All pivots positive <=> M positive definite <=>
    //          lambda_max(S~) < tau: the LMI cannot bind and the sample is finished -- r^3/6 FMAs and two barriers per step
    //          against ~r^3/2 FMAs, two block reductions and four barriers per Householder step, no multisection.  The
    //          float32 factorisation is the exact one of M + E, |E| <= c r eps |M|: margin = 3.2e-7 r (1e-5 at r = 32 as
    //          in lmi_warp.cuh, 1e-4 at r = 320) keeps ~5x headroom over that, so a pass is a proof; a failure proves
    //          nothing: the matrix is loaded again and solved.
    if (!grad_only && !lambda_out && kprior > 0.f) {
      const float tau = kprior - 3.2e-7f * static_cast<float>(r) * (kprior + fro);
      for (int e = tid; e < p4; e += THREADS) A[e] = -A[e];
      __syncthreads();
      if (tid < r) A[ti + tid] += tau;
      __syncthreads();
      bool pass = true;
      for (int kk = 0; kk < r; ++kk) {
        const float d = A[lmib_tri(kk) + kk];   // the same word for every thread
        if (!(d > 0.f)) {                       // uniform (NaN fails)
          pass = false;
          break;
        }
        if (kk == r - 1) break;
        const bool mine = tid > kk && tid < r;
        if (mine) vv[tid] = A[ti + kk];
        __syncthreads();
        if (mine) {
          const float l = vv[tid] * (1.0f / d);
          float* ar = A + ti;
          int j = kk + 1;
          for (; (j & 3) && j <= tid; ++j) ar[j] = fmaf(-l, vv[j], ar[j]);
          for (; j + 4 <= tid + 1; j += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vv + j);
            ar[j] = fmaf(-l, v4.x, ar[j]);
            ar[j + 1] = fmaf(-l, v4.y, ar[j + 1]);
            ar[j + 2] = fmaf(-l, v4.z, ar[j + 2]);
            ar[j + 3] = fmaf(-l, v4.w, ar[j + 3]);
          }
          for (; j <= tid; ++j) ar[j] = fmaf(-l, vv[j], ar[j]);
        }
        __syncthreads();
      }
      if (pass) continue;  // uniform per CTA
      // the factorisation destroyed the matrix: load it again
      __syncthreads();
      for (int e = tid; e < p4 / 4; e += THREADS) {
        float4 x = *reinterpret_cast<const float4*>(srow + 4 * e);
        x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
        *reinterpret_cast<float4*>(A + 4 * e) = x;
      }
      __syncthreads();
    }

    // ---- c. Householder tridiagonalisation (lower form): thread i owns row i
    for (int kk = 0; kk + 2 < r; ++kk) {
      const int i = tid;
      const float xi = (i > kk && i < r) ? A[ti + kk] : 0.f;
      const float tail2 = lb_block_sum<THREADS>((i > kk + 1) ? xi * xi : 0.f, red);
      const float x1 = A[lmib_tri(kk + 1) + kk];
      const bool skip = !(tail2 > 0.f);  // column already tridiagonal
      const float sigma = fmaf(x1, x1, tail2);
      const float rt = sqrtf(sigma);
      const float alpha = (x1 >= 0.f) ? -rt : rt;
      const float tau = skip ? 0.f : 1.0f / fmaf(fabsf(x1), rt, sigma);  // 2 / v'v with v_1 = x1 - alpha
      const float vi = skip ? 0.f : ((i == kk + 1) ? x1 - alpha : ((i > kk + 1 && i < r) ? xi : 0.f));
      if (i < r) vv[i] = vi;
      if (i == kk) {
        sd[kk] = A[ti + kk];
        se[kk] = skip ? x1 : alpha;
        stau[kk] = tau;
      }
      __syncthreads();
      if (skip) continue;  // uniform
      if (i == kk + 1) A[ti + kk] = vi;  // the reflector stays in column kk (its other entries are already there)
      // p = tau A v over the trailing block: along row i up to the diagonal, then down column i.  v comes in 16-byte
      // broadcast loads (one per four entries), the column walk keeps a running offset (row j + 1 starts j + 1 words
      // behind row j): ~2.3 instructions per entry instead of ~6 -- the walk is issue-bound, not bandwidth-bound
      float pi = 0.f;
      if (i > kk && i < r) {
        const float* ar = A + ti;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
        int j = kk + 1;
        for (; (j & 3) && j <= i; ++j) p0 = fmaf(ar[j], vv[j], p0);
        for (; j + 4 <= i + 1; j += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vv + j);
          p0 = fmaf(ar[j], v4.x, p0);
          p1 = fmaf(ar[j + 1], v4.y, p1);
          p2 = fmaf(ar[j + 2], v4.z, p2);
          p3 = fmaf(ar[j + 3], v4.w, p3);
        }
        for (; j <= i; ++j) p0 = fmaf(ar[j], vv[j], p0);
        j = i + 1;
        int off = lmib_tri(j) + i;
        for (; (j & 3) && j < r; ++j) {
          p1 = fmaf(A[off], vv[j], p1);
          off += j + 1;
        }
        for (; j + 4 <= r; j += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vv + j);
          const int o1 = off + j + 1, o2 = o1 + j + 2, o3 = o2 + j + 3;
          p0 = fmaf(A[off], v4.x, p0);
          p1 = fmaf(A[o1], v4.y, p1);
          p2 = fmaf(A[o2], v4.z, p2);
          p3 = fmaf(A[o3], v4.w, p3);
          off = o3 + j + 4;
        }
        for (; j < r; ++j) {
          p2 = fmaf(A[off], vv[j], p2);
          off += j + 1;
        }
        pi = tau * ((p0 + p1) + (p2 + p3));
      }
      const float Kc = 0.5f * tau * lb_block_sum<THREADS>(vi * pi, red);
      const float wi = (i > kk && i < r) ? fmaf(-Kc, vi, pi) : 0.f;
      if (i < r) ww[i] = wi;
      __syncthreads();
      // A <- A - v w' - w v' over the lower triangle of the trailing block
      if (i > kk && i < r) {
        float* ar = A + ti;
        int j = kk + 1;
        for (; (j & 3) && j <= i; ++j) ar[j] = fmaf(-vi, ww[j], fmaf(-wi, vv[j], ar[j]));
        for (; j + 4 <= i + 1; j += 4) {
          const float4 v4 = *reinterpret_cast<const float4*>(vv + j);
          const float4 w4 = *reinterpret_cast<const float4*>(ww + j);
          const float a0 = ar[j], a1 = ar[j + 1], a2 = ar[j + 2], a3 = ar[j + 3];
          ar[j] = fmaf(-vi, w4.x, fmaf(-wi, v4.x, a0));
          ar[j + 1] = fmaf(-vi, w4.y, fmaf(-wi, v4.y, a1));
          ar[j + 2] = fmaf(-vi, w4.z, fmaf(-wi, v4.z, a2));
          ar[j + 3] = fmaf(-vi, w4.w, fmaf(-wi, v4.w, a3));
        }
        for (; j <= i; ++j) ar[j] = fmaf(-vi, ww[j], fmaf(-wi, vv[j], ar[j]));
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (r >= 2) {
        sd[r - 2] = A[lmib_tri(r - 2) + (r - 2)];
        se[r - 2] = A[lmib_tri(r - 1) + (r - 2)];
      }
      sd[r - 1] = A[lmib_tri(r - 1) + (r - 1)];
      se[r - 1] = 0.f;
    }
    __syncthreads();

    // ---- d. lambda_max of the tridiagonal matrix by multisection
    float lam;
    {
      float gmax = -3.0e38f, gabs = 0.f, dmax = -3.0e38f;
      for (int i = tid; i < r; i += THREADS) {
        const float rad = fabsf(i > 0 ? se[i - 1] : 0.f) + fabsf(se[i]);
        gmax = fmaxf(gmax, sd[i] + rad);
        gabs = fmaxf(gabs, fabsf(sd[i]) + rad);
        dmax = fmaxf(dmax, sd[i]);
      }
      gmax = lb_block_max<THREADS>(gmax, red);
      gabs = lb_block_max<THREADS>(gabs, red);
      dmax = lb_block_max<THREADS>(dmax, red);
      const float scale = fmaxf(gabs, 1e-30f);
      const float isc = 1.0f / scale;
      // lambda_max >= the largest diagonal entry; kappa = relu(lambda_max): search [max(dmax, 0), Gershgorin]
      float lo = fmaxf(dmax, 0.f) * isc;
      float hi = fmaxf(fmaf(1e-6f, scale, gmax) * isc, lo);
      bool below_zero = false;
      const int rounds = (THREADS >= 128) ? 4 : 5;
      for (int round = 0; round < rounds; ++round) {
        const float h = (hi - lo) / static_cast<float>(THREADS);
        const float x = fmaf(h, static_cast<float>(tid), lo);
        // pivots of x I - T: all positive <=> x above the whole spectrum
        float q = x - sd[0] * isc;
        bool above = q > 0.f;
        for (int i = 1; i < r && above; ++i) {
          const float e = se[i - 1] * isc;
          q = (x - sd[i] * isc) - e * e / q;
          above = q > 0.f;
        }
        const int first = lb_block_min_int<THREADS>(above ? tid : THREADS, redi);
        if (first == 0) {
          // even lo is above the spectrum: only possible for lo = 0 (lambda_max >= dmax): lambda_max < 0, kappa = 0
          below_zero = true;
          break;
        }
        const float nlo = fmaf(h, static_cast<float>(first - 1), lo);
        const float nhi = (first == THREADS) ? hi : fmaf(h, static_cast<float>(first), lo);
        lo = nlo;
        hi = nhi;
      }
      const float mid = 0.5f * (lo + hi);
      lam = (below_zero || !(mid > 1e-7f)) ? 0.f : mid * scale;
    }

    if (lambda_out) {
      if (tid == 0) kappa_io[b] = fmaxf(kappa_io[b], lam);
      continue;
    }
    // ---- e. merge with the prior, scale step
    float kap = lam;
    int tag = make_tag(RAYEN_FAM_LMI, 0);
    const bool binds = kap > kprior;
    if (!binds) {
      kap = kprior;
      tag = tprior;
    }
    if (!grad_only && binds) {
      if (tid == 0) {
        kappa_io[b] = kap;
        active_io[b] = tag;
      }
      const float a_old = lb_alpha(kprior, s, beta, mode);
      const float a_new = lb_alpha(kap, s, beta, mode);
      if (a_old > 0.f && a_new != a_old) {
        const float ratio = a_new / a_old;
        float* yrow = y + b * static_cast<long long>(k);
        const float* y0 = P.blob + P.off_y0;
        for (int i = tid; i < k; i += THREADS) {
          const float c = __ldg(y0 + i);
          yrow[i] = fmaf(yrow[i] - c, ratio, c);
        }
      }
    }
    if (!want_grad) continue;
    {
      bool need = tag_family(tag) == RAYEN_FAM_LMI && kap > 0.f;
      if (need && mode == RAYEN_MODE_RAYEN) need = (1.0f / kap < s);
      if (!need) continue;  // uniform
    }

    // ---- f. eigenvector of lam: twisted factorisation of T - lam I ...
    {
      float sc = 0.f;
      for (int i = tid; i < r; i += THREADS) sc = fmaxf(sc, fabsf(sd[i]) + fabsf(se[i]));
      sc = lb_block_max<THREADS>(sc, red);
      const float tiny = fmaxf(1e-12f * sc, 1e-30f);
      if (tid == 0 || tid == 32) {
        const bool fw = tid == 0;
        float* out = fw ? dp : dm;
        const int step = fw ? 1 : -1;
        int i = fw ? 0 : r - 1;
        float piv = sd[i] - lam;
        for (int t = 0; t < r; ++t) {
          if (t > 0) {
            const float ee = fw ? se[i - 1] : se[i];
            piv = fmaf(-ee * ee, 1.0f / piv, sd[i] - lam);
          }
          if (fabsf(piv) < tiny) piv = -tiny;
          out[i] = piv;
          i += step;
        }
      }
      __syncthreads();
      // twist index: argmin |dp_i + dm_i - (d_i - lam)|, lowest index on ties
      // two-step argmin (exact): the minimum value, then the lowest index that attains it
      float gmin = 3.0e38f;
      for (int i = tid; i < r; i += THREADS) gmin = fminf(gmin, fabsf(dp[i] + dm[i] - (sd[i] - lam)));
      gmin = -lb_block_max<THREADS>(-gmin, red);
      int kt = r;
      for (int i = tid; i < r; i += THREADS)
        if (fabsf(dp[i] + dm[i] - (sd[i] - lam)) == gmin) kt = min(kt, i);
      kt = lb_block_min_int<THREADS>(kt, redi);
      // z_{i-1} = -e_{i-1} / dp_{i-1} z_i below the twist, z_{i+1} = -e_i / dm_{i+1} z_i above it
      for (int i = tid; i < r; i += THREADS) {
        float rr = 0.f;
        if (i < kt) rr = -se[i] / dp[i];
        else if (i > kt) rr = -se[i - 1] / dm[i];
        ww[i] = rr;
      }
      __syncthreads();
      if (tid == 0 || tid == 32) {
        const int step = (tid == 0) ? -1 : 1;
        float zi = 1.f;
        for (int i = kt + step; i >= 0 && i < r; i += step) {
          zi *= ww[i];
          sq[i] = zi;
        }
      }
      if (tid == 64 % THREADS) sq[kt] = 1.f;
      __syncthreads();
      float qi = (tid < r) ? sq[tid] : 0.f;
      qi *= 1.0f / sqrtf(lb_block_sum<THREADS>(qi * qi, red));
      // ... back-transformed: q = H_0 ... H_{r-3} z, reflector kk in column kk of A (rows kk + 1 ..)
      for (int kk = r - 3; kk >= 0; --kk) {
        const float tau = stau[kk];
        if (tau == 0.f) continue;  // uniform (shared value)
        const float vi = (tid > kk && tid < r) ? A[ti + kk] : 0.f;
        const float c = tau * lb_block_sum<THREADS>(vi * qi, red);
        qi = fmaf(-c, vi, qi);
      }
      __syncthreads();
      if (tid < r) sq[tid] = qi;
      __syncthreads();
      // gradient weights over the packed order: w_e = (2 - [i = j]) q_i q_j (overwrites the matrix: no longer needed)
      float* wts = A;
      if (tid < r) {
        const float qrow = qi;
        float* wp = wts + ti;
        for (int j = 0; j < tid; ++j) wp[j] = 2.f * qrow * sq[j];
        wp[tid] = qrow * qrow;
      }
      for (int e = r * (r + 1) / 2 + tid; e < p4; e += THREADS) wts[e] = 0.f;
      __syncthreads();
      if (grad_list != nullptr) {
        // d kappa / du_a = <F_a, w> for all a is a row of the GEMM  W [samples x P] . F' [P x n]: leave w in this sample's
        // (now dead) row of the contraction buffer and the sample's number on a list; lmib_grad_gemm_kernel does the rest
        // for all LMI-bound samples of the chunk at once (streaming F once per 64 samples instead of once per sample)
        float4* dst = reinterpret_cast<float4*>(Sw + static_cast<size_t>(b) * p4);
        const float4* w4 = reinterpret_cast<const float4*>(wts);
        for (int e = tid; e < p4 / 4; e += THREADS) dst[e] = w4[e];
        if (tid == 0) grad_list[atomicAdd(grad_count, 1)] = static_cast<int>(b);
      } else {
        // d kappa / du_a = <F_a, w>: one warp per a, coalesced 16-byte loads of the packed row
        const float* F = P.blob + P.off_lmib;
        for (int a = warp; a < n; a += THREADS / 32) {
          const float4* fr = reinterpret_cast<const float4*>(F + static_cast<size_t>(a) * p4);
          const float4* w4 = reinterpret_cast<const float4*>(wts);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int e = lane; e < p4 / 4; e += 32) {
            const float4 f = __ldg(fr + e);
            const float4 w = w4[e];
            a0 = fmaf(f.x, w.x, a0);
            a1 = fmaf(f.y, w.y, a1);
            a2 = fmaf(f.z, w.z, a2);
            a3 = fmaf(f.w, w.w, a3);
          }
          const float g = lb_warp_sum((a0 + a1) + (a2 + a3));
          if (lane == 0) dkappa[b * n + a] = g;
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------- 3. gradient GEMM (NT)
// dkappa[list[m]][a] = sum_e Sw[list[m]][e] F[a][e]   (m < *count, a < n, e < p4): the rows the solve kernel left behind
// (w_e = (2 - [i = j]) q_i q_j) against the packed matrices -- both operands K-contiguous.  64 x 64 x 16 tiles, 4 x 4
// outputs per thread; the grid covers the whole chunk, CTAs beyond the list length leave at once.
constexpr int kLbGradTile = 64, kLbGradK = 16, kLbGradThreads = 256;
__global__ void __launch_bounds__(kLbGradThreads)
    lmib_grad_gemm_kernel(const float* __restrict__ Sw, const int* __restrict__ list, const int* __restrict__ count,
                          const float* __restrict__ F, int n, int p4, float* __restrict__ dkappa) {
  LB_STATIC_SHARED __align__(16) float As[kLbGradK][kLbGradTile + 4];
  LB_STATIC_SHARED __align__(16) float Bs[kLbGradK][kLbGradTile + 4];
  LB_STATIC_SHARED int rows[kLbGradTile];
  const int tid = threadIdx.x;
  const int tiles_n = (n + kLbGradTile - 1) / kLbGradTile;
  const int m0 = static_cast<int>(blockIdx.x / tiles_n) * kLbGradTile;
  const int n0 = static_cast<int>(blockIdx.x % tiles_n) * kLbGradTile;
  const int total = *count;
  if (m0 >= total) return;  // uniform
  if (tid < kLbGradTile) rows[tid] = (m0 + tid < total) ? list[m0 + tid] : -1;
  __syncthreads();
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row lr of the tile, k = lk .. lk + 3
  const int arow = rows[lr];
  const float* ap = (arow >= 0) ? Sw + static_cast<size_t>(arow) * p4 : nullptr;
  const float* bp = (n0 + lr < n) ? F + static_cast<size_t>(n0 + lr) * p4 : nullptr;
  const int ty = tid >> 4, tx = tid & 15;        // outputs: rows ty * 4 .., columns tx * 4 ..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p4; k0 += kLbGradK) {
    float4 a4 = float4{0.f, 0.f, 0.f, 0.f}, b4 = float4{0.f, 0.f, 0.f, 0.f};
    if (k0 + lk < p4) {   // p4 is a multiple of 4: a 4-group is in or out as a whole
      if (ap) a4 = *reinterpret_cast<const float4*>(ap + k0 + lk);
      if (bp) b4 = __ldg(reinterpret_cast<const float4*>(bp + k0 + lk));
    }
    __syncthreads();  // the previous tile's readers are done
    As[lk + 0][lr] = a4.x; As[lk + 1][lr] = a4.y; As[lk + 2][lr] = a4.z; As[lk + 3][lr] = a4.w;
    Bs[lk + 0][lr] = b4.x; Bs[lk + 1][lr] = b4.y; Bs[lk + 2][lr] = b4.z; Bs[lk + 3][lr] = b4.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kLbGradK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r_ = rows[ty * 4 + i];
    if (r_ < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = n0 + tx * 4 + j;
      if (a < n) dkappa[static_cast<size_t>(r_) * n + a] = acc[i][j];
    }
  }
}

}  // namespace rayen
