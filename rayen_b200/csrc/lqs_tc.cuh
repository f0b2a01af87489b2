// Linear / quadratic / SOC kappa on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Every constraint of these families is a set of dot products with the unit direction u: a row of D, the
// vectors phi / c_z / h, the rows of the triangular factors G and R (and of the pruning bound's factor).
// Stacked, they are ONE matrix W [rows x K] (K = n padded to 8/16/32), and the work of a tile of 128 samples
// is the GEMM  D = U W'  with U [128 x K].  This kernel runs that GEMM with tcgen05.mma (kind::tf32,
// M = 128, N <= 128, K = 8 per instruction), error-compensated as 3xTF32
//     U W' ~= U_hi W_hi' + U_lo W_hi' + U_hi W_lo'        (drops only the 2^-22 term U_lo W_lo')
// with the accumulator in tensor memory.  The TMEM layout -- lane = sample, column = constraint row -- is
// exactly what the reduction needs: each epilogue thread owns one sample, reads its row of D with
// tcgen05.ld and takes the running max / sums of squares in registers, with no cross-lane traffic at all.
//
// Roles (warp-specialised, one CTA per SM, persistent over super-tiles of 2 x 128 samples):
//   warps 0-7  load v, normalise, write U_hi/U_lo as K-major operand tiles; later read D from TMEM and reduce;
//              finally apply the scale step and write y / kappa / active (and the LMI work list)
//   warp 8     TMA producer: streams the 128-row panels of W (hi + lo) through a 4-stage shared-memory ring
//   warps 9,10 MMA issuers (one elected lane each, one per sample tile); warp 9 owns the TMEM allocation
// Pipelines: W ring full/empty, TMEM accumulator full/empty (two 128-column buffers per sample tile), U ready.
//
// Panels.  128 rows of D are a panel.  The items (quadratics, cones, the LMI pruning bound) come in BATCHES of up to 12:
// their upper-triangular factors are cut into blocks of 8 rows, and block j of all the items of a batch is one panel
// (8 x items rows) whose columns left of 8j are zero -- its GEMM starts at K step j (10 instead of 16 K steps per batch at
// K = 32) and only those K chunks of W are copied into the ring.  Block 0 also carries the two header rows of every item
// (phi_z | c_z | t, then h).  An epilogue thread keeps the running sum of squares of every item of the batch and forms
// the kappas when the last block has arrived.
#pragma once
#include "common.cuh"
#include "lqs.cuh"

namespace rayen {

#ifdef RAYEN_TC_TRACE
__device__ long long g_tc_trace[4096];
#define TC_STAMP(slot) do { if (blockIdx.x == 0) g_tc_trace[(slot)] = clock64(); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#endif

constexpr int kTcPanel = 128;  // rows of W per panel = MMA N (N = 128: 65 cycles per MMA against 57 at N = 96, scripts/mma_probe.cu)
constexpr int kTcStages = 4;
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 3) * 32;  // + TMA warp + two MMA-issuing warps (one per sample tile)
// per panel (plan.py): int32 {kind 0 linear | 1 batch block, first row | block index j, MMA N, first K step, items in the
// batch, last block of the batch, column of the header rows, cone mask | (1 + bound slot) << 16}, 12 x (item type,
// item index) at words 8.., then 12 item
// scalars (float: A of a cone, r of the bound) at words 32.. and their reciprocals at words 44..
constexpr int kTcTableWords = 64;
constexpr int kTcBatchItems = 12;  // items per batch: 8 x 12 = 96 factor rows per block panel + 24 header rows in block 0

// ----------------------------------------------------------------------------- tcgen05 / mbarrier helpers
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes; LBO = byte stride between the two 16-byte K chunks of
  // one MMA, SBO = byte stride between 8-row groups; bits [46,48) = descriptor version 1 (sm_100)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  // c = F32 (bits 4-5), a = b = TF32 (bits 7-9, 10-12), both K-major, n >> 3 at bit 17, m >> 4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tc_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tc_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// FP32-pipe re-evaluation of one cone's kappa from the FP32 constants (SOC section of the blob), for the rare sample
// whose ray is nearly tangent to the cone: there (h.u)^2 + A c' cancels, d kappa/d(h.u) = (1 + h.u/sqrt(disc))/A is
// large, and the 3xTF32 dot products (a few 1e-7 of sum|h_a u_a|) are not good enough.  One thread, ~600 FMAs.
template <int KP>
__device__ __forceinline__ float soc_kappa_fp32(const float* __restrict__ item, const float (&u)[KP]) {
  constexpr int TRI = (KP / 4) * (KP / 4 + 1) * 8;
  float cu = 0.f, hb = 0.f, ss = 0.f;
#pragma unroll
  for (int a = 0; a < KP; ++a) {
    cu = fmaf(__ldg(item + a), u[a], cu);
    hb = fmaf(__ldg(item + KP + a), u[a], hb);
  }
  const float* tri = item + 2 * KP;
  int pos = 0;
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    float r = 0.f;
#pragma unroll
    for (int j = 4 * (i / 4); j < KP; ++j) r = fmaf(__ldg(tri + pos + j - 4 * (i / 4)), u[j], r);
    pos += KP - 4 * (i / 4);
    ss = fmaf(r, r, ss);
  }
  return soc_root(__ldg(item + 2 * KP + TRI), hb, fmaf(-cu, cu, ss), nullptr);
}

// y staging (coalesced stores): per epilogue warp 16 rows of KP floats, row stride KP + 4 words (16-byte aligned rows; the
// 8 lanes of a quarter-warp writing 16 bytes of 8 different rows hit 8 different bank groups)
constexpr int kTcStageRows = 16;
template <int KP>
__host__ __device__ constexpr size_t lqs_tc_smem_bytes(int n_panels, bool y_stage = false) {
  return 256 + static_cast<size_t>((n_panels * kTcTableWords + 3) / 4 * 4) * 4 + 4 * static_cast<size_t>(KP) * 128 * 4 +
         static_cast<size_t>(kTcStages) * 2 * KP * kTcPanel * 4 +
         (y_stage ? static_cast<size_t>(kTcEpiWarps) * kTcStageRows * (KP + 4) * 4 : 0);
}

// ----------------------------------------------------------------------------- the kernel
template <int KP>
__global__ void __launch_bounds__(kTcThreads, 1)
    lqs_tc_forward_kernel(const PlanDev P, const float* __restrict__ v, long long ldv, float* __restrict__ y,
                          float* __restrict__ kappa_out, int* __restrict__ active_out, long long B, int mode,
                          int lmi_follows, int prune, int* __restrict__ work_list, int* __restrict__ work_count,
                          const MapArgs M) {
  constexpr int IB = kTcBatchItems;         // items per batch
  constexpr int KC = KP / 4;                // 16-byte chunks along K
  constexpr int A_TILE = KP * 128;          // floats of one U operand tile (hi or lo)
  constexpr int W_TILE = KP * kTcPanel;     // floats of one W operand tile (hi or lo)
  constexpr uint32_t LBO_A = (128 / 8) * 128, LBO_W = (kTcPanel / 8) * 128, SBO = 128;

  pdl_launch_dependents();  // the LMI kernel behind this one may start scheduling its CTAs (it waits before it reads)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* a_full = &bars[0];
  uint64_t* w_full = &bars[1];                 // [kTcStages]
  uint64_t* w_empty = &bars[1 + kTcStages];    // [kTcStages]
  uint64_t* d_full = &bars[1 + 2 * kTcStages];  // [2]
  uint64_t* d_empty = &bars[3 + 2 * kTcStages]; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[5 + 2 * kTcStages]);
  float* table = reinterpret_cast<float*>(smem_raw + 256);
  const int n_panels = P.tc_panels;
  float* a_tiles = table + (n_panels * kTcTableWords + 3) / 4 * 4;  // [tile 0/1][hi, lo][A_TILE]
  float* w_ring = a_tiles + 4 * A_TILE;                              // [stage][hi, lo][W_TILE]
  float* y_stage = w_ring + kTcStages * 2 * W_TILE;                  // [epilogue warp][16 rows][KP + 4] (P.tc_y_stage)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* tc_base = P.blob + P.off_tc;
  const float* w_src = tc_base + n_panels * kTcTableWords;

  if (tid == 0) {
    mbar_init(a_full, kTcEpiWarps);
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 2);  // one tcgen05.commit per MMA-issuing warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 2);
      mbar_init(&d_empty[i], kTcEpiWarps);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < n_panels * kTcTableWords; i += kTcThreads) table[i] = tc_base[i];
  if (warp == kTcEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_super = (B + 255) / 256;
  const int n = P.n, k = P.k;
  if (tid == 0) TC_STAMP(1000);
#ifdef RAYEN_TC_TRACE
  if (tid == 0) g_tc_trace[2200 + blockIdx.x] = clock64();
  long long gt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0)); if (tid == 0) g_tc_trace[2400 + blockIdx.x] = gt0;
#endif

  if (warp == kTcEpiWarps) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0;  // running panel counter: ring stage g % stages, use g / stages
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
        for (int p = 0; p < n_panels; ++p, ++g) {
          const uint32_t stage = g % kTcStages, use = g / kTcStages;
          mbar_wait(&w_empty[stage], (use & 1) ^ 1);
          TC_STAMP(8 * p + 7);
          // only the K chunks the panel's GEMM reads: from K step ks0 on (the leading columns of a block panel are zero)
          const int ks0 = reinterpret_cast<const int*>(table + p * kTcTableWords)[3];
          const uint32_t skip = static_cast<uint32_t>(2 * ks0) * (LBO_W / 4);   // floats of one operand that are not needed
          const uint32_t part = W_TILE - skip;                                    // floats of W_hi (and of W_lo) to copy
          mbar_expect_tx(&w_full[stage], 2 * part * 4);
          const float* src = w_src + static_cast<size_t>(p) * 2 * W_TILE;
          float* dst = w_ring + stage * 2 * W_TILE;
          // several smaller bulk copies keep more requests in flight than one large copy
          const uint32_t half = (part / 2) & ~3u;
#pragma unroll
          for (int o = 0; o < 2; ++o) {       // W_hi, W_lo
            const uint32_t base = o * W_TILE + skip;
            if (half >= 1024) {
              bulk_g2s(dst + base, src + base, half * 4, &w_full[stage]);
              bulk_g2s(dst + base + half, src + base + half, (part - half) * 4, &w_full[stage]);
            } else {
              bulk_g2s(dst + base, src + base, part * 4, &w_full[stage]);
            }
          }
        }
      }
    }
  } else if (warp > kTcEpiWarps) {
    // ===================================================================== MMA issuers: warp 9 -> sample tile 0,
    // warp 10 -> sample tile 1.  Two issuing threads keep the tensor pipe fed while the other one is busy
    // with its commits / barrier waits (a single issuer left the pipe idle ~25 % of every panel).
    const int t = warp - (kTcEpiWarps + 1);
    if (lane == 0) {
      uint32_t g = 0, s_local = 0;
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x, ++s_local) {
        mbar_wait(a_full, s_local & 1);
        tc_fence_after();
        for (int p = 0; p < n_panels; ++p, ++g) {
          const uint32_t stage = g % kTcStages, use = g / kTcStages;
          const uint32_t buf = g & 1, buf_use = g >> 1;
          TC_STAMP(8 * p + 0);
          mbar_wait(&w_full[stage], use & 1);
          TC_STAMP(8 * p + 1);
          mbar_wait(&d_empty[buf], (buf_use & 1) ^ 1);
          tc_fence_after();
          TC_STAMP(8 * p + 2);
          const uint32_t w_hi = smem_u32(w_ring + stage * 2 * W_TILE), w_lo = w_hi + W_TILE * 4;
          {
            const int* pt = reinterpret_cast<const int*>(table + p * kTcTableWords);
            const uint32_t idesc = umma_idesc_tf32(128, pt[2]);   // N of this panel (a multiple of 16, <= 128)
            const int ks0 = pt[3];                                // first K step: the columns before it are zero
            const uint32_t u_hi = smem_u32(a_tiles + (2 * t) * A_TILE), u_lo = u_hi + A_TILE * 4;
            const uint32_t d_tmem = tmem_base + (2 * t + buf) * kTcPanel;
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks) {
              if (ks >= ks0) {
                const uint64_t d_uhi = umma_smem_desc(u_hi + 2 * ks * LBO_A, LBO_A, SBO);
                const uint64_t d_ulo = umma_smem_desc(u_lo + 2 * ks * LBO_A, LBO_A, SBO);
                const uint64_t d_whi = umma_smem_desc(w_hi + 2 * ks * LBO_W, LBO_W, SBO);
                const uint64_t d_wlo = umma_smem_desc(w_lo + 2 * ks * LBO_W, LBO_W, SBO);
                umma_tf32(d_tmem, d_uhi, d_whi, idesc, ks > ks0 ? 1u : 0u);
                umma_tf32(d_tmem, d_ulo, d_whi, idesc, 1u);
                umma_tf32(d_tmem, d_uhi, d_wlo, idesc, 1u);
              }
            }
          }
          TC_STAMP(8 * p + 3);
          umma_commit(&w_empty[stage]);  // the ring slot is free once these MMAs have read it
          umma_commit(&d_full[buf]);     // ... and the accumulator is complete
        }
      }
    }
  } else {
    // ===================================================================== operand prep + reduction + scale step
    const int t = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool vec_in = ((n & 3) == 0) && ((ldv & 3) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    const bool vec_out = ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const float* y0 = P.blob + P.off_y0;
    const float* nmat = P.blob + P.off_nmat;
    uint32_t g = 0;
    for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
      const long long b = st * 256 + t * 128 + row;
      const bool valid = b < B;
      float u[KP];
      if (M.x)
        map_row<KP>(M, b, n, valid, u);  // fused mapper: v = W x + b, also written to M.v_out
      else
        load_row<KP>(v + b * ldv, n, vec_in, valid, u);
      const float beta = (mode == RAYEN_MODE_RAYEN_OLD && valid) ? __ldg(v + b * ldv + n) : 0.f;
      const float s = normalize_row<KP>(u);
      {
        float* hi = a_tiles + (2 * t) * A_TILE + ((row >> 3) * 32 + (row & 7) * 4);
        float* lo = hi + A_TILE;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          float4 h4, l4;
          h4.x = tf32_rna(u[4 * kc + 0]); l4.x = tf32_rna(u[4 * kc + 0] - h4.x);
          h4.y = tf32_rna(u[4 * kc + 1]); l4.y = tf32_rna(u[4 * kc + 1] - h4.y);
          h4.z = tf32_rna(u[4 * kc + 2]); l4.z = tf32_rna(u[4 * kc + 2] - h4.z);
          h4.w = tf32_rna(u[4 * kc + 3]); l4.w = tf32_rna(u[4 * kc + 3] - h4.w);
          *reinterpret_cast<float4*>(hi + kc * 16 * 32) = h4;
          *reinterpret_cast<float4*>(lo + kc * 16 * 32) = l4;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);

      float best = 0.f;
      int tag = make_tag(RAYEN_FAM_NONE, 0);
      float ub = 3.0e38f;  // pruning bound of the LMI (stays +inf without one)
      int soc_fix = -1;    // a cone whose kappa is ill-conditioned for this sample (see soc_kappa_fp32)
      int win = -1;        // (panel, slot) of the binding item, if an item binds (looked up in the table afterwards)
      float ss[IB], hd[2 * IB];  // running |T u|^2 and the two header dot products of every item of the current batch
#pragma unroll
      for (int i = 0; i < IB; ++i) ss[i] = hd[2 * i] = hd[2 * i + 1] = 0.f;
      for (int p = 0; p < n_panels; ++p, ++g) {
        const uint32_t buf = g & 1, buf_use = g >> 1;
        if (warp == 0 && lane == 0) TC_STAMP(8 * p + 4);
        mbar_wait(&d_full[buf], buf_use & 1);
        tc_fence_after();
        if (warp == 0 && lane == 0) TC_STAMP(8 * p + 5);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (2 * t + buf) * kTcPanel;
        const int* pt = reinterpret_cast<const int*>(table + p * kTcTableWords);
        if (pt[0] == 0) {
          // ---- 128 rows of D, 64 at a time: kappa_j = D_j . u         (reference constraint_module.py:353)
          const int base = pt[1];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float x[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(taddr + 64 * hh + 16 * c, x + 16 * c);
            tmem_wait_ld();
            // max first (3-input max tree), index only when the running best actually moves: the best of a
            // sample changes O(log rows) times, so the per-column compare/select chain is almost never needed
            float m0 = x[0], m1 = x[1], m2 = x[2], m3 = x[3];
#pragma unroll
            for (int j = 4; j < 64; j += 4) {
              m0 = fmaxf(m0, x[j]);
              m1 = fmaxf(m1, x[j + 1]);
              m2 = fmaxf(m2, x[j + 2]);
              m3 = fmaxf(m3, x[j + 3]);
            }
            const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            if (mx > best) {
              best = mx;
              win = -1;
              int arg = 0;
#pragma unroll
              for (int j = 63; j >= 0; --j)
                if (x[j] == mx) arg = j;  // first index on ties, like torch.max
              tag = make_tag(RAYEN_FAM_LINEAR, base + 64 * hh + arg);
            }
          }
        } else {
          // ---- block j of a batch of items: 8 columns per item (rows 8j..8j+7 of its triangular factor times u); block 0
          // also holds the items' header dot products behind them
          const int jb = pt[1], nitems = pt[4];
#pragma unroll
          for (int g = 0; g < IB * 8 / 48; ++g) {      // 48 columns = 6 items at a time
            if (48 * g < 8 * nitems) {                  // (the table is the same for every thread)
              float x[48];
#pragma unroll
              for (int c = 0; c < 3; ++c) tmem_ld16(taddr + 48 * g + 16 * c, x + 16 * c);
              tmem_wait_ld();
#pragma unroll
              for (int q = 0; q < 6; ++q) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) a = fmaf(x[8 * q + c], x[8 * q + c], a);
                ss[6 * g + q] = (jb == 0) ? a : ss[6 * g + q] + a;
              }
            }
          }
          if (jb == 0) {
            const int hoff = pt[6];
#pragma unroll
            for (int c = 0; c < 2 * IB / 8; ++c) {
              if (8 * c < 2 * nitems) tmem_ld8(taddr + hoff + 8 * c, hd + 8 * c);
            }
            tmem_wait_ld();
          }
          if (pt[5]) {
            // ---- last block: the kappas of the batch, in item order (the reference's evaluation order).
            // Branch-free and without loads: item types come as a bit mask, the winner is remembered as a (panel, slot)
            // number that is looked up once after the panel loop, square roots / reciprocals are single MUFU instructions
            // (~1 ulp; the IEEE sqrtf / division expand to a fix-up behind a branch to a slow path).  As a chain of
            // `if (type == ...)` with a table load per item this was the longest phase of the kernel (device-clock
            // trace: 2300-5600 cycles per batch of 11 items, against ~700 for draining a panel).  One ulp is well inside
            // what the 3xTF32 dot products carry (3e-7).
            const int meta = pt[7];
            const int bslot = (meta >> 16) - 1;   // slot of the LMI pruning bound in this batch (-1: none)
            float b_mean = 0.f, b_nrm = 0.f, b_fac = 0.f;
#pragma unroll
            for (int sl = 0; sl < IB; ++sl) {
              const float scal = table[p * kTcTableWords + 32 + sl], inv_scal = table[p * kTcTableWords + 44 + sl];
              const float h0 = hd[2 * sl], h1 = hd[2 * sl + 1], sq = ss[sl];
              const float nrm = tc_sqrt(sq);                          // |G u| (quadratic), |T_c u| (bound)
              const float cq = fmaf(-h0, h0, sq);
              const float disc = fmaf(h1, h1, scal * cq);
              const float root = tc_sqrt(fmaxf(disc, 0.f));
              // kappa = phi_z.u + |G u| (reference :360-381), or the largest root of the cone (:383-399) in the
              // cancellation-free form of soc_root (lqs.cuh): (h1 + root)/A, or c'/(root - h1) when h1 < 0
              const bool soc = (meta >> sl) & 1;
              const float ksoc = (h1 >= 0.f) ? (h1 + root) * inv_scal : cq * tc_rcp(root - h1);
              float kap = soc ? ksoc : h0 + nrm;
              kap = (sl >= nitems || sl == bslot) ? -1.f : kap;       // no item here: never wins (best >= 0)
              // (selects, not branches: a branch per item serialises the items' MUFU latencies)
              b_mean = (sl == bslot) ? h0 * inv_scal : b_mean;
              b_nrm = (sl == bslot) ? nrm : b_nrm;
              b_fac = (sl == bslot) ? (scal - 1.0f) * inv_scal : b_fac;
              // a nearly tangent cone (the discriminant cancels to < 1 % of its terms) that can matter for this sample: the
              // first such cone is left out of the running max and evaluated on the FP32 pipe after the panel loop
              const bool fix = soc && soc_fix < 0 && KP == P.np && disc < 1e-2f * h1 * h1 && kap > 0.9f * best;
              const bool take = !fix && kap > best;
              soc_fix = fix ? p * 16 + sl : soc_fix;
              best = take ? kap : best;
              win = take ? p * 16 + sl : win;
            }
            if (bslot >= 0) {
              // Wolkowicz-Styan bound of lambda_max (LMI pruning); b_nrm = |T_c u| with T_c the factor of the CENTRED Gram
              // matrix: the deviation term is a sum of squares, nothing cancels when S~(u) is close to a multiple of I
              ub = lmi_upper_bound(b_mean, b_nrm * tc_sqrt(b_fac), P.lmi_bound_margin);
            }
          }
        }
        if (warp == 0 && lane == 0) TC_STAMP(8 * p + 6);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[buf]);
      }

      // ---- which item was that (type and index within its family)
      if (win >= 0) {
        const int* wt = reinterpret_cast<const int*>(table + (win >> 4) * kTcTableWords) + 8 + 2 * (win & 15);
        tag = make_tag(wt[0], wt[1]);
      }
      // ---- the rare near-tangent cone that was left out above, in FP32
      if (soc_fix >= 0) {
        const int cone = reinterpret_cast<const int*>(table + (soc_fix >> 4) * kTcTableWords)[9 + 2 * (soc_fix & 15)];
        const float kap = soc_kappa_fp32<KP>(P.blob + P.off_soc + cone * P.soc_stride, u);
        if (kap > best) {
          best = kap;
          tag = make_tag(RAYEN_FAM_SOC, cone);
        }
      }
      // ---- merge / prune / scale step for this thread's sample
      const bool pruned = lmi_follows && prune && (ub < best);  // ub already carries its rounding allowance
      const bool finish = !lmi_follows || pruned;
      {
        // samples the LMI kernel still has to look at: one atomicAdd per warp, not per sample
        const bool enq = valid && !finish && work_list != nullptr;
        const unsigned m = __ballot_sync(0xffffffffu, enq);
        if (m != 0u) {
          const int leader = __ffs(m) - 1;
          int slot0 = 0;
          if (lane == leader) slot0 = atomicAdd(work_count, __popc(m));
          slot0 = __shfl_sync(0xffffffffu, slot0, leader);
          if (enq) work_list[slot0 + __popc(m & ((1u << lane) - 1u))] = static_cast<int>(b);
        }
      }
      if (valid) {
        if (kappa_out) kappa_out[b] = best;
        if (active_out) active_out[b] = tag;
      }
      const bool writes = valid && finish;
      float alpha = 0.f;
      if (writes) alpha = (mode == RAYEN_MODE_RAYEN_OLD) ? 1.0f / (expf(beta) + best) : fminf(1.0f / best, s);
      if (P.n_is_identity && vec_out && P.tc_y_stage) {
        // Coalesced y: the warp's 32 samples are 32 consecutive rows of y.  Half of them at a time go through a
        // shared-memory tile, and every store instruction then writes whole rows -- 512 contiguous bytes per four rows --
        // instead of 16 bytes per lane with a row (128 bytes at k = 32) between lanes.  Same bytes for HBM, but 4x fewer
        // 32-byte sectors per instruction, and when y is a peer / multicast mapping (sharding.forward_gathered: the
        // all-gather fused into this epilogue) full-width NVLink writes instead of one small write per lane.
        const unsigned wmask = __ballot_sync(0xffffffffu, writes);
        if (wmask != 0u) {
          constexpr int ST = KP + 4;
          float* stg = y_stage + warp * (kTcStageRows * ST);
          const int kc = k >> 2;
          float* ybase = y + (b - lane) * k;  // row of lane 0
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __syncwarp();
            if ((lane >> 4) == h && writes) {
              float* dst = stg + (lane & 15) * ST;
#pragma unroll
              for (int kk = 0; kk < KP / 4; ++kk) {
                if (4 * kk < k) {
                  float4 o;
                  o.x = fmaf(alpha, u[4 * kk + 0], __ldg(y0 + 4 * kk + 0));
                  o.y = fmaf(alpha, u[4 * kk + 1], __ldg(y0 + 4 * kk + 1));
                  o.z = fmaf(alpha, u[4 * kk + 2], __ldg(y0 + 4 * kk + 2));
                  o.w = fmaf(alpha, u[4 * kk + 3], __ldg(y0 + 4 * kk + 3));
                  *reinterpret_cast<float4*>(dst + 4 * kk) = o;
                }
              }
            }
            __syncwarp();
            for (int c = lane; c < kTcStageRows * kc; c += 32) {
              const int r = c / kc, cc = c - r * kc;
              if ((wmask >> (16 * h + r)) & 1u)
                *reinterpret_cast<float4*>(ybase + static_cast<long long>(16 * h + r) * k + 4 * cc) =
                    *reinterpret_cast<const float4*>(stg + r * ST + 4 * cc);
            }
          }
        }
      } else if (writes) {
        float* yrow = y + b * k;
        if (P.n_is_identity) {
#pragma unroll
          for (int kk = 0; kk < KP / 4; ++kk) {
            if (4 * kk < k) {
              float4 o;
              o.x = fmaf(alpha, u[4 * kk + 0], __ldg(y0 + 4 * kk + 0));
              o.y = fmaf(alpha, u[4 * kk + 1], __ldg(y0 + 4 * kk + 1));
              o.z = fmaf(alpha, u[4 * kk + 2], __ldg(y0 + 4 * kk + 2));
              o.w = fmaf(alpha, u[4 * kk + 3], __ldg(y0 + 4 * kk + 3));
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
          for (int i = 0; i < k; ++i) {
            const float* nrow = nmat + i * (P.np + 4);
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < KP; ++a)
              if (a < P.np) acc = fmaf(__ldg(nrow + a), u[a], acc);
            yrow[i] = fmaf(alpha, acc, __ldg(y0 + i));
          }
        }
      }
    }
  }

  // ---- teardown: nobody may still be reading TMEM when it is released
  tc_fence_before();
  __syncthreads();
#ifdef RAYEN_TC_TRACE
  if (tid == 0) { g_tc_trace[2000 + blockIdx.x] = clock64() - g_tc_trace[2200 + blockIdx.x]; long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1)); g_tc_trace[2600 + blockIdx.x] = gt1; }
#endif
  if (warp == kTcEpiWarps + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

}  // namespace rayen
