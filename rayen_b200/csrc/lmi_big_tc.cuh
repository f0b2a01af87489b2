// The contraction of the big-LMI path on the tensor cores: S~(v) = V . F as a tcgen05 3xTF32 GEMM with a K loop
// (reference constraint_module.py:412-421: S = einsum('ajk,ial->ijk', all_F, rho) -- for LMIs with a wide subspace this IS
// a dense GEMM, [B x n] . [n x r(r+1)/2], K = n up to thousands; the FP32-pipe version is lmib_contract_kernel).
//
//   C[128 samples x 128 entries] per CTA, K walked in slices of 32:
//     * B operand: plan section LMIBT -- F' cut into [128 entries x 32 k] tiles, TF32-split (hi, lo), stored in the
//       K-major no-swizzle operand layout [k/4][row/8][row%8][k%4]; one (panel, slice) pair of tiles = 32 KB = two 1-D TMA
//       bulk copies into a 4-stage shared-memory ring (SASS UBLKCP), completion on an mbarrier;
//     * A operand: the slice [128 samples x 32 k] of V, read from global by the four operand/epilogue warps (thread =
//       sample row), split hi/lo with cvt.rna.tf32 and written in the same layout into a 2-stage ring;
//     * one elected lane issues, per slice, 4 K-steps x 3 MMAs (hi.hi + lo.hi + hi.lo, error-compensated 3xTF32: the
//       dropped term is 2^-22 relative) of tcgen05.mma.cta_group::1.kind::tf32, M = N = 128, K = 8, into one of TWO
//       128-column TMEM accumulators, each slice starting from zero;
//     * the accumulator of a slice is drained (tcgen05.ld.32x32b.x16, thread = sample = TMEM lane) and ADDED TO 128 FP32
//       REGISTERS per thread while the MMAs of the next slice run into the other accumulator.  Why not accumulate over all
//       of K in TMEM: the tensor core adds with round-toward-zero, a bias of ~0.5 ulp per accumulation that grows
//       linearly -- measured with the whole K loop in TMEM: y off by 1.2e-5 and outputs 3.9e-5 outside the LMI at
//       K = 2000 (750 accumulations), against 3e-7 / 3e-6 with the FP32 GEMM.  Twelve accumulations per slice (the same as
//       lqs_tc.cuh at K = 32) and round-to-nearest adds across slices keep the result at FP32-GEMM accuracy;
//     * epilogue: every thread stores its row of C with 16-byte stores.
// The result replaces lmib_contract_kernel's in layout, not in rounding (3xTF32 vs sequential FP32 FMAs: both ~1e-7 of
// sum |v_a F_a|); the eigen-solve behind it is the same kernel.
#pragma once
#include "common.cuh"
#include "lqs_tc.cuh"

namespace rayen {

constexpr int kLbtThreads = 6 * 32;          // 4 operand/epilogue warps + TMA warp + MMA warp
constexpr int kLbtStagesW = 4, kLbtStagesU = 2;
constexpr int kLbtTile = 128 * 32;           // floats of one operand tile (hi or lo)
constexpr int kLbtPanelsPerCta = 1;          // one 128-entry panel per CTA (two TMEM accumulators alternate over the K slices)

__host__ __device__ constexpr size_t lmib_tc_smem_bytes() {
  return 256 + static_cast<size_t>(kLbtStagesW + kLbtStagesU) * 2 * kLbtTile * 4;
}

// grid: ceil(Bc / 128) x n_panels CTAs, flattened (m tile fastest: CTAs that run together share the B tiles in L2)
__global__ void __launch_bounds__(kLbtThreads, 1)
    lmib_contract_tc_kernel(const float* __restrict__ V, long long ldv, const float* __restrict__ FT, int n, int p4,
                            int n_panels, int k_slices, float* __restrict__ C, long long Bc) {
  constexpr uint32_t LBO = (128 / 8) * 128, SBO = 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* w_full = &bars[0];                               // [4]
  uint64_t* w_empty = &bars[kLbtStagesW];                    // [4]
  uint64_t* u_full = &bars[2 * kLbtStagesW];                 // [2]
  uint64_t* u_empty = &bars[2 * kLbtStagesW + kLbtStagesU];  // [2]
  uint64_t* d_full = &bars[2 * kLbtStagesW + 2 * kLbtStagesU];       // [2]
  uint64_t* d_empty = &bars[2 * kLbtStagesW + 2 * kLbtStagesU + 2];  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[2 * kLbtStagesW + 2 * kLbtStagesU + 4]);
  float* w_ring = reinterpret_cast<float*>(smem_raw + 256);  // [stage][hi, lo][kLbtTile]
  float* u_ring = w_ring + kLbtStagesW * 2 * kLbtTile;       // [stage][hi, lo][kLbtTile]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m_tiles = (Bc + 127) / 128;
  const long long m_tile = blockIdx.x % m_tiles;
  const int group = static_cast<int>(blockIdx.x / m_tiles);
  const int pan0 = group;   // this CTA's panel of 128 entries
  (void)n_panels;

  if (tid == 0) {
    for (int i = 0; i < kLbtStagesW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < kLbtStagesU; ++i) {
      mbar_init(&u_full[i], 4);   // one arrival per operand warp
      mbar_init(&u_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================================================== TMA producer of the B tiles
    if (lane == 0) {
      for (int s = 0; s < k_slices; ++s) {
        const uint32_t stage = s % kLbtStagesW, use = s / kLbtStagesW;
        mbar_wait(&w_empty[stage], (use & 1) ^ 1);
        mbar_expect_tx(&w_full[stage], 2 * kLbtTile * 4);
        const float* src = FT + (static_cast<size_t>(pan0) * k_slices + s) * (2 * kLbtTile);
        float* dst = w_ring + stage * 2 * kLbtTile;
        bulk_g2s(dst, src, kLbtTile * 4, &w_full[stage]);
        bulk_g2s(dst + kLbtTile, src + kLbtTile, kLbtTile * 4, &w_full[stage]);
      }
    }
  } else if (warp == 5) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, 128);
      for (int s = 0; s < k_slices; ++s) {
        const uint32_t ub = s % kLbtStagesU, uuse = s / kLbtStagesU;
        const uint32_t stage = s % kLbtStagesW, use = s / kLbtStagesW;
        const uint32_t buf = s & 1, buse = s >> 1;
        mbar_wait(&u_full[ub], uuse & 1);
        mbar_wait(&w_full[stage], use & 1);
        mbar_wait(&d_empty[buf], (buse & 1) ^ 1);
        tc_fence_after();
        const uint32_t u_hi = smem_u32(u_ring + ub * 2 * kLbtTile), u_lo = u_hi + kLbtTile * 4;
        const uint32_t w_hi = smem_u32(w_ring + stage * 2 * kLbtTile), w_lo = w_hi + kLbtTile * 4;
        const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t d_uhi = umma_smem_desc(u_hi + 2 * ks * LBO, LBO, SBO);
          const uint64_t d_ulo = umma_smem_desc(u_lo + 2 * ks * LBO, LBO, SBO);
          const uint64_t d_whi = umma_smem_desc(w_hi + 2 * ks * LBO, LBO, SBO);
          const uint64_t d_wlo = umma_smem_desc(w_lo + 2 * ks * LBO, LBO, SBO);
          umma_tf32(d_tmem, d_uhi, d_whi, idesc, ks > 0 ? 1u : 0u);   // every slice starts from zero
          umma_tf32(d_tmem, d_ulo, d_whi, idesc, 1u);
          umma_tf32(d_tmem, d_uhi, d_wlo, idesc, 1u);
        }
        umma_commit(&w_empty[stage]);   // the ring slots are free once these MMAs have read them
        umma_commit(&u_empty[ub]);
        umma_commit(&d_full[buf]);      // ... and this slice's accumulator is complete
      }
    }
  } else {
    // ===================================================================== A-operand producers, then the epilogue
    const int row = warp * 32 + lane;     // sample row of the tile = TMEM lane
    const long long b = m_tile * 128 + row;
    const bool valid = b < Bc;
    const float* vrow = V + (valid ? b : 0) * ldv;
    const bool vec = ((ldv & 3) == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);
    float acc[128];
#pragma unroll
    for (int c = 0; c < 128; ++c) acc[c] = 0.f;
    // drain slice `sd`'s accumulator into the registers (round-to-nearest adds), then hand the TMEM buffer back
    auto drain = [&](int sd) {
      const uint32_t buf = sd & 1, buse = sd >> 1;
      mbar_wait(&d_full[buf], buse & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * 128;
#pragma unroll
      for (int c = 0; c < 128; c += 16) {
        float r[16];
        tmem_ld16(taddr + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[c + q] += r[q];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[buf]);
    };
    for (int s = 0; s < k_slices; ++s) {
      const uint32_t ub = s % kLbtStagesU, uuse = s / kLbtStagesU;
      float x[32];
      const int k0 = s * 32;
      if (valid && vec && k0 + 32 <= n) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(vrow + k0) + c);
          x[4 * c] = q.x; x[4 * c + 1] = q.y; x[4 * c + 2] = q.z; x[4 * c + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = (valid && k0 + c < n) ? __ldg(vrow + k0 + c) : 0.f;
      }
      mbar_wait(&u_empty[ub], (uuse & 1) ^ 1);
      float* hi = u_ring + ub * 2 * kLbtTile + ((row >> 3) * 32 + (row & 7) * 4);
      float* lo = hi + kLbtTile;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float4 h4, l4;
        h4.x = tf32_rna(x[4 * kc + 0]); l4.x = tf32_rna(x[4 * kc + 0] - h4.x);
        h4.y = tf32_rna(x[4 * kc + 1]); l4.y = tf32_rna(x[4 * kc + 1] - h4.y);
        h4.z = tf32_rna(x[4 * kc + 2]); l4.z = tf32_rna(x[4 * kc + 2] - h4.z);
        h4.w = tf32_rna(x[4 * kc + 3]); l4.w = tf32_rna(x[4 * kc + 3] - h4.w);
        *reinterpret_cast<float4*>(hi + kc * 16 * 32) = h4;
        *reinterpret_cast<float4*>(lo + kc * 16 * 32) = l4;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&u_full[ub]);
      if (s > 0) drain(s - 1);   // under the MMAs of slice s
    }
    drain(k_slices - 1);
    // ---- epilogue: this thread's row of C
    if (valid) {
      float* crow = C + static_cast<size_t>(b) * p4;
      const int col0 = pan0 * 128;
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        const int col = col0 + c;
        if (col < p4) *reinterpret_cast<float4*>(crow + col) = float4{acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
      }
    }
  }
  // ---- teardown: nobody may still be reading TMEM when it is released
  tc_fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
}

}  // namespace rayen
