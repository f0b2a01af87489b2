// LMI forward with the contraction S~(u) = sum_a u_a F~z_a on the 5th-generation tensor cores.
//
// The contraction of a batch is a genuine dense GEMM (reference constraint_module.py:412-421 forms it with
// einsum + two batched matmuls):   S[sample, (i,j)] = sum_a U[sample, a] * Fz[(i,j), a]
// i.e. D = U W' with U [samples x K] (K = n padded to 8/16/32) and W [RP*RP x K].  lmi_forward_kernel (lmi.cuh)
// evaluates it on the FP32 pipe out of shared memory, n*RP*RP words of LDS traffic per matrix -- the LSU-bound
// 30 % of a dense n = 32 pass.  Here the same CTA that solves the eigenproblems runs the GEMM with tcgen05.mma
// (kind::tf32, M = 128, N = 128, K = 8, error-compensated 3xTF32 exactly like lqs_tc.cuh) into tensor memory,
// drains the accumulator through a small shared-memory staging buffer into the solver's register layout
// (RP*RP words per matrix instead of n*RP*RP), and goes on with the Householder / Sturm solver of lmi.cuh
// unchanged.  Nothing leaves the SM.
//
// One pass of a CTA = SPP samples (8 warps x MPW matrices: 32 for RP = 32, 64 for RP = 16):
//   1. every lane group loads and normalises its direction (scratch) and writes it, split into TF32 hi + lo,
//      as ONE ROW of the K-major U operand tile; sample slot s sits in row 32*(s / SPQ) + s % SPQ, so each of
//      the four TMEM lane quadrants holds SPQ = SPP/4 samples and all eight warps can drain (a warp may only
//      read the 32 TMEM lanes of its quadrant).  The other rows of the 128-row tile stay zero: the tensor pipe
//      is 25-50 % used, which is irrelevant (a pass costs ~6 k MMA cycles against ~60 k solver cycles).
//   2. W (128-row panels, hi + lo) is staged by 1-D TMA bulk copies: once per CTA when all panels fit in shared
//      memory (K = 8), else through a ring that the next pass's first panels already refill during the solver
//      phase; one elected thread issues copies and MMAs, up to four panels ahead of the drain (four TMEM
//      accumulators of 128 columns = all 512 columns).  Panels are ordered like the solver's registers: column e = i*RP + 4q + t is entry
//      (row i, column q + LPM*t) of the matrix, so a lane's four columns of one row are one float4.
//   3. drain: warp w reads TMEM lanes of quadrant w%4, columns [64*(w/4), +64) of the panel and the SPQ lanes
//      that carry samples store them to staging[sample][entry]; after half a matrix (4 panels = 16 rows at
//      RP = 32) the solver lanes pull their rows into registers and the staging buffer is reused.
//   4. tridiagonalise, Sturm multisection, merge, scale step (and eigenvector + d kappa/du): lmi.cuh.
#pragma once
#include "lmi.cuh"
#include "lqs_tc.cuh"

namespace rayen {

constexpr int kLmiTcThreads = 256;
constexpr int kLmiTcPanel = 128;   // entries (MMA N) per panel
constexpr int kLmiTcBufs = 4;       // TMEM accumulators (4 x 128 columns = all 512)
// Panels handed to the tensor core ahead of the one being drained.  The issuing thread also drains, and
// tcgen05.mma back-pressures its issuer once a few instructions are queued (measured: ~100 cycles per MMA when
// the queue is full), so a deep look-ahead stalls the drain it belongs to: one panel ahead is the optimum.
constexpr int kLmiTcAhead = 1;
constexpr int kLmiTcMaxStages = 8;  // W ring stages; == panels of the matrix means W stays resident

// Predicated issue: the whole warp walks the issue path in lock step and only the elected lane's instruction takes
// effect.  (An `if (lane == 0)` region whose other 31 lanes sit in __syncwarp ran 5-10x slower per instruction.)
__device__ __forceinline__ void umma_tf32_if(uint32_t elected, uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t elected, uint64_t* bar) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)),
      "r"(elected)
      : "memory");
}

template <int RP>
struct LmiTcCfg {
  using C = LmiCfg<RP>;
  static constexpr int SPP = (kLmiTcThreads / 32) * C::MPW;       // samples per pass
  static constexpr int SPQ = SPP / 4;                             // samples per TMEM lane quadrant
  static constexpr int NPAN = RP * RP / kLmiTcPanel;              // panels per matrix
  static constexpr int HP = NPAN < 4 ? NPAN : 4;                  // panels per staging fill
  static constexpr int NHALF = NPAN / HP;
  static constexpr int ROWS_H = HP * kLmiTcPanel / RP;            // matrix rows per staging fill
  static constexpr int STG = HP * kLmiTcPanel + 4;                // staging floats per sample (+4: bank spread)
  static_assert(RP == 16 || RP == 32, "the tensor-core contraction covers RP = 16 and 32");
  static_assert(SPQ <= 32 && NPAN % 2 == 0, "layout assumptions");
};

// dynamic smem: 128 B barriers | U tile (hi, lo) | W ring | staging | F (if F_SMEM) | per-warp scratch
template <int RP>
__host__ __device__ constexpr size_t lmi_tc_smem_bytes(int n, int kp, int stages, bool f_smem) {
  using T = LmiTcCfg<RP>;
  return 128 + static_cast<size_t>(2 * kp * 128) * 4 + static_cast<size_t>(stages) * 2 * kp * kLmiTcPanel * 4 +
         static_cast<size_t>(T::SPP) * T::STG * 4 + (f_smem ? static_cast<size_t>(n) * RP * RP * 4 : 0) +
         static_cast<size_t>(kLmiTcThreads / 32) * LmiCfg<RP>::MPW * LmiCfg<RP>::SCR * 4;
}

template <int RP, bool F_SMEM, bool WITH_GRAD>
__global__ void __launch_bounds__(kLmiTcThreads, 1)
    lmi_forward_tc_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                          float* __restrict__ kappa_io, int* __restrict__ active_io, long long B, int mode,
                          int has_prior, const int* __restrict__ work_list, const int* __restrict__ work_count,
                          float* __restrict__ dkappa) {
  using C = LmiCfg<RP>;
  using T = LmiTcCfg<RP>;
  constexpr uint32_t SBO = 128, LBO = (128 / 8) * 128;  // both operands are 128-row K-major tiles
  constexpr uint32_t IDESC = umma_idesc_tf32(128, kLmiTcPanel);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* w_full = &bars[0];     // [kLmiTcMaxStages]
  uint64_t* mma_done = &bars[8];   // [kLmiTcBufs]
  uint64_t* f_bar = &bars[12];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[14]);
  const int ns = P.lmitc_stages;             // ring depth chosen by the host from the shared-memory budget
  const bool resident = ns >= T::NPAN;       // every panel has its own stage: loaded once, never refilled
  const int kp = P.tc_kp;
  const int a_tile = kp * 128;               // floats of one operand tile (hi or lo), U and W alike
  float* u_tiles = reinterpret_cast<float*>(smem_raw + 128);       // [hi, lo][a_tile]
  float* w_ring = u_tiles + 2 * a_tile;                             // [stage][hi, lo][a_tile]
  float* staging = w_ring + ns * 2 * a_tile;                        // [SPP][STG]
  float* f_smem = staging + T::SPP * T::STG;
  float* scratch_base = f_smem + (F_SMEM ? P.lmi_words : 0);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LMI_STAMP(0);
  pdl_wait();  // launched behind the linear/quadratic/SOC kernel: its kappa / active / work list must be complete
  const long long total = work_list ? static_cast<long long>(ld_after_wait(work_count)) : B;
  // samples per pass: a short list is spread over all CTAs (whole warps, at least one)
  int per_pass = T::SPP;
  if (total < static_cast<long long>(T::SPP) * gridDim.x) {
    const long long want = (total + gridDim.x - 1) / gridDim.x;
    per_pass = static_cast<int>((want + C::MPW - 1) / C::MPW) * C::MPW;
    if (per_pass < C::MPW) per_pass = C::MPW;
  }
  const long long n_pass = (total + per_pass - 1) / per_pass;
  const long long my_passes = n_pass > blockIdx.x ? (n_pass - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t total_panels = static_cast<uint32_t>(my_passes) * T::NPAN;
  const float* w_src = P.blob + P.off_lmitc;

  if (tid == 0) {
    for (int i = 0; i < kLmiTcMaxStages; ++i) mbar_init(&w_full[i], 1);
    for (int i = 0; i < kLmiTcBufs; ++i) mbar_init(&mma_done[i], 1);
    mbar_init(f_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows of the U tile that carry no sample stay zero for the whole kernel
  for (int i = tid; i < 2 * a_tile / 4; i += kLmiTcThreads)
    reinterpret_cast<float4*>(u_tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  LMI_STAMP(1);

  auto issue_w = [&](uint32_t g) {  // thread 0: panel g % NPAN -> ring stage g % ns
    const uint32_t stage = g % ns;
    const uint32_t bytes = static_cast<uint32_t>(2 * a_tile) * 4u;
    mbar_expect_tx(&w_full[stage], bytes);
    const char* src = reinterpret_cast<const char*>(w_src) + static_cast<size_t>(g % T::NPAN) * bytes;
    char* dst = reinterpret_cast<char*>(w_ring) + static_cast<size_t>(stage) * bytes;
    for (uint32_t done = 0; done < bytes; done += 8192u)
      bulk_g2s(dst + done, src + done, (bytes - done < 8192u) ? bytes - done : 8192u, &w_full[stage]);
  };
  // operand descriptors are built once; a K step (8 columns = two 16-byte chunks) advances the start-address
  // field (16-byte units, bits 0-13) by 2*LBO/16, a ring stage by the stage size
  const uint64_t desc_u_hi = umma_smem_desc(smem_u32(u_tiles), LBO, SBO);
  const uint64_t desc_u_lo = umma_smem_desc(smem_u32(u_tiles) + a_tile * 4, LBO, SBO);
  const uint64_t desc_w0 = umma_smem_desc(smem_u32(w_ring), LBO, SBO);
  const uint64_t kstep = (2 * LBO) >> 4, lo_off = static_cast<uint64_t>(a_tile * 4) >> 4;
  const uint32_t elected = (lane == 0) ? 1u : 0u;
  auto issue_mma = [&](uint32_t g) {  // warp 0, all lanes: panel g into accumulator g % 4
    const uint32_t stage = g % ns;
    mbar_wait(&w_full[stage], resident ? 0u : ((g / ns) & 1));
    tc_fence_after();
    if (g % T::NPAN == 0) LMI_STAMP(5);
    const uint64_t d_whi0 = desc_w0 + static_cast<uint64_t>(stage) * 2 * lo_off;
    const uint32_t d_tmem = tmem_base + (g % kLmiTcBufs) * kLmiTcPanel;
    const int nk = kp >> 3;
#pragma unroll 1
    for (int ks = 0; ks < nk; ++ks) {
      const uint64_t d_uhi = desc_u_hi + ks * kstep, d_ulo = desc_u_lo + ks * kstep;
      const uint64_t d_whi = d_whi0 + ks * kstep, d_wlo = d_whi + lo_off;
      umma_tf32_if(elected, d_tmem, d_uhi, d_whi, IDESC, ks > 0 ? 1u : 0u);
      umma_tf32_if(elected, d_tmem, d_ulo, d_whi, IDESC, 1u);
      umma_tf32_if(elected, d_tmem, d_uhi, d_wlo, IDESC, 1u);
    }
    umma_commit_if(elected, &mma_done[g % kLmiTcBufs]);
    if (g % T::NPAN == 0) LMI_STAMP(6);
  };
  auto w_ready = [&](uint32_t g) -> bool {  // has panel g landed in its ring stage? (non-blocking)
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(&w_full[g % ns])), "r"(resident ? 0u : ((g / ns) & 1))
        : "memory");
    return ok != 0;
  };

  if (tid == 0 && my_passes > 0) {
    for (int i = 0; i < ns && i < static_cast<int>(total_panels); ++i) issue_w(i);
    if constexpr (F_SMEM && WITH_GRAD) stage_bulk(f_smem, P.blob + P.off_lmi, P.lmi_words, f_bar);
  }
  bool f_ready = !(F_SMEM && WITH_GRAD);
  const float* F = (F_SMEM && WITH_GRAD) ? f_smem : P.blob + P.off_lmi;

  LmiSolver<RP, WITH_GRAD, F_SMEM> S;
  S.q = lane % C::LPM;
  S.grp_base = lane - S.q;
  const int grp = lane / C::LPM;
  const int slot = warp * C::MPW + grp;                          // sample slot of this lane group in the pass
  const int u_row = 32 * (slot / T::SPQ) + slot % T::SPQ;        // its row of the U tile / TMEM lane
  S.scr = scratch_base + slot * C::SCR;
  const int n = P.n, k = P.k;
  const float* y0 = P.blob + P.off_y0;
  const float* nmat = P.blob + P.off_nmat;
  const int quad = warp & 3, chalf = warp >> 2;

  uint32_t g = 0;  // running panel counter of this CTA (same in every thread)
  for (long long pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
    const long long idx = pass * per_pass + slot;
    const bool valid = slot < per_pass && idx < total;
    const long long b = valid ? (work_list ? static_cast<long long>(work_list[idx]) : idx) : 0;
    LMI_STAMP(2);
    const float s = S.load_direction(v + b * ldv, n, valid);
    LMI_STAMP(3);
    // ---- 1. this sample's row of the U operand (hi + lo), 16-byte K chunks kc = q, q + LPM, ...
    {
      const float* u = S.su();
      float* hi = u_tiles + (u_row >> 3) * 32 + (u_row & 7) * 4;
      float* lo = hi + a_tile;
      for (int kc = S.q; kc < kp / 4; kc += C::LPM) {
        const float4 x = ld4(u + 4 * kc);
        float4 h4, l4;
        h4.x = tf32_rna(x.x); l4.x = tf32_rna(x.x - h4.x);
        h4.y = tf32_rna(x.y); l4.y = tf32_rna(x.y - h4.y);
        h4.z = tf32_rna(x.z); l4.z = tf32_rna(x.z - h4.z);
        h4.w = tf32_rna(x.w); l4.w = tf32_rna(x.w - h4.w);
        *reinterpret_cast<float4*>(hi + kc * 16 * 32) = h4;
        *reinterpret_cast<float4*>(lo + kc * 16 * 32) = l4;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    LMI_STAMP(4);
    // ---- 2./3. GEMM panels and drain, half a matrix at a time
    const uint32_t g_end = g + T::NPAN;
    uint32_t next_mma = g;  // warp 0: next panel to hand to the tensor core
#pragma unroll
    for (int h = 0; h < T::NHALF; ++h) {
      for (int p = 0; p < T::HP; ++p, ++g) {
        if (warp == 0) {
          while (next_mma <= g) issue_mma(next_mma++);  // the panel everybody is about to wait for
          // look-ahead: accumulators of panels < g are drained; only panels whose W has already landed are
          // issued (the elected thread must not block the drain)
          while (next_mma < g_end && next_mma <= g + kLmiTcAhead && w_ready(next_mma)) issue_mma(next_mma++);
          if (h == 0 && p == 0) LMI_STAMP(7);
        }
        __syncwarp();
        // one polling lane per warp: 256 threads spinning on the barrier starve the issuing thread of LSU slots
        if (lane == 0) mbar_wait(&mma_done[g % kLmiTcBufs], (g / kLmiTcBufs) & 1);
        __syncwarp();
        tc_fence_after();
        LMI_STAMP(8 + 2 * (h * T::HP + p));
        // MMA(g) is complete: its ring stage is free for the panel ns further on (possibly of the next pass)
        if (tid == 0 && !resident && g + ns < total_panels) issue_w(g + ns);
        float x[64];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (g % kLmiTcBufs) * kLmiTcPanel + chalf * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16(taddr + 16 * c, x + 16 * c);
        tmem_wait_ld();
        if (lane < T::SPQ) {
          float* dst = staging + (quad * T::SPQ + lane) * T::STG + p * kLmiTcPanel + chalf * 64;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        }
        tc_fence_before();
        __syncthreads();
        LMI_STAMP(9 + 2 * (h * T::HP + p));
      }
      // the solver lanes pull rows [h*ROWS_H, +ROWS_H) of their matrix: one float4 per row
      {
        const float* src = staging + slot * T::STG + 4 * S.q;
#pragma unroll
        for (int i = 0; i < T::ROWS_H; ++i) {
          const float4 f = ld4(src + i * RP);
          S.A[h * T::ROWS_H + i][0] = f.x;
          S.A[h * T::ROWS_H + i][1] = f.y;
          S.A[h * T::ROWS_H + i][2] = f.z;
          S.A[h * T::ROWS_H + i][3] = f.w;
        }
      }
      if (h + 1 < T::NHALF) __syncthreads();  // staging is refilled by the next half
    }
    // warps without a sample in this pass took part in the drain only
    if (static_cast<long long>(warp) * C::MPW >= per_pass || pass * per_pass + static_cast<long long>(warp) * C::MPW >= total)
      continue;
    // ---- 4. the solver of lmi.cuh, unchanged
    LMI_STAMP(24);
    S.tridiagonalize();
    LMI_STAMP(25);
    const float lam = S.lambda_max_relu();
    LMI_STAMP(26);
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
    if constexpr (WITH_GRAD) {
      bool need = valid && tag_family(tag) == RAYEN_FAM_LMI && kap > 0.f;
      if (need && mode == RAYEN_MODE_RAYEN) need = (1.0f / kap < s);
      if (__ballot_sync(0xffffffffu, need) != 0u) {  // warp-uniform: the other matrices of the warp just ride along
        __syncwarp();
        float qo[4];
        S.eigenvector(lam, qo);
        if (!f_ready) {
          mbar_wait(f_bar, 0);
          f_ready = true;
        }
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
#pragma unroll 1
        for (int G = 0; G < C::MPW; ++G) {
          if (!((need_mask >> (G * C::LPM)) & 1u)) continue;
          // (sample indices fit 32 bits: the work lists are int32)
          const int bG = __shfl_sync(0xffffffffu, static_cast<int>(b), G * C::LPM);
          S.eig_gradient_coop(F, n, G, grp, dkappa, static_cast<long long>(bG) * n);
        }
      }
    }
    __syncwarp();  // the scratch (u) is rewritten by the next sample
    LMI_STAMP(27);
  }

  // ---- teardown: no bulk copy may be in flight, nobody may still read TMEM when it is released
  if constexpr (F_SMEM && WITH_GRAD) {
    if (!f_ready && my_passes > 0) mbar_wait(f_bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

}  // namespace rayen
