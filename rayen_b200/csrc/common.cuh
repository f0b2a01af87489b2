// Shared device-side definitions of the RAYEN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rayen_b200.h"

namespace rayen {

// Device view of a plan: dimensions + word offsets into the constant block (see rayen_b200.h).
struct PlanDev {
  const float* blob;
  int n, k, np, k_pad;
  int m, m_pad, n_quad, n_soc;
  int lmi_r, lmi_rp, n_is_identity, lmi_prune;
  int lin_stride, quad_stride, soc_stride;
  int off_lin, off_quad, off_soc, off_nmat, off_y0, off_bound, off_lmi;
  int lqs_words;   // words [off_lin, off_lin + lqs_words) = LIN | QUAD | SOC | NMAT | Y0 | BOUND, staged to smem
  int lmi_words;   // n * rp * rp
  int off_tc, tc_panels, tc_kp;  // tensor-core section (see rayen_b200.h)
  int off_viol, off_lmineg, viol_in, viol_eq;  // violation checker sections
  int off_lmitc, lmitc_panels, lmitc_stages;   // LMI matrices as a tcgen05 B operand (lmi_tc.cuh); ring depth
  float lmi_bound_margin;                      // float32-rounding allowance of the pruning bound (rayen_b200.h, BOUND)
  int off_lmiw;                                // LMI matrices for lmi_warp.cuh (0: none)
  int tc_y_stage;                              // lqs_tc.cuh: the kernel's shared memory includes the y staging tiles (coalesced stores)
};

// Fused mapper (reference constraint_module.py:261, :525: q = nn.Linear(input_dim, n)(x)): when x != nullptr the
// linear/quadratic/SOC forward kernel computes its own v = W x + b from the layer input, writes it to v_out (the
// backward pass and the LMI kernel need it) and goes on; q never makes the HBM round trip of a separate GEMM launch.
struct MapArgs {
  const float* x;     // [B, in_dim], row stride ldx (multiple of 4, 16-byte aligned rows)
  const float* w;     // [n, in_dim], row stride ldw (multiple of 4, 16-byte aligned rows)
  const float* bias;  // [n] or nullptr
  float* v_out;       // [B, n] dense
  long long ldx, ldw;
  int in_dim;         // multiple of 4
  int pad;
};

constexpr int kFamShift = 24;
__host__ __device__ inline int make_tag(int fam, int idx) { return (fam << kFamShift) | idx; }
__host__ __device__ inline int tag_family(int tag) { return tag >> kFamShift; }
__host__ __device__ inline int tag_index(int tag) { return tag & ((1 << kFamShift) - 1); }

constexpr float kNormEps = 1e-12f;  // torch.nn.functional.normalize eps (reference constraint_module.py:470)

// ----------------------------------------------------------------------------- TMA bulk copy (1-D)
// global -> shared::cta through the async proxy, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One elected thread stages `words` floats (multiple of 4) from global to shared in <= 32 KiB pieces.
__device__ __forceinline__ void stage_bulk(float* dst_smem, const float* src, int words, uint64_t* bar) {
  const uint32_t total = static_cast<uint32_t>(words) * 4u;
  mbar_expect_tx(bar, total);
  uint32_t done = 0;
  while (done < total) {
    uint32_t piece = total - done;
    if (piece > 32768u) piece = 32768u;
    bulk_g2s(reinterpret_cast<char*>(dst_smem) + done, reinterpret_cast<const char*>(src) + done, piece, bar);
    done += piece;
  }
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// The LMI forward kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization right behind the
// linear/quadratic/SOC kernel: its CTAs are scheduled (and run their prologue: barrier init, TMEM allocation, the TMA
// staging of the constant matrices) while the previous kernel drains, and pdl_wait() blocks until that kernel has
// completed and its writes (kappa, active, the work list) are visible.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// A value the PREVIOUS kernel of the stream produced (a work-list counter), read behind pdl_wait().  A plain load through
// a `const T* __restrict__` parameter is an invariant load to the compiler (LDG.CONSTANT) and may be hoisted above the
// wait -- observed in SASS: the counter was read while the producer was still running, and came back as 0.  A volatile
// load is neither hoisted nor served from the non-coherent path.
__device__ __forceinline__ int ld_after_wait(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// ----------------------------------------------------------------------------- small helpers
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// v[a] = bias[a] + sum_k W[a][k] x[b][k] for a < n (0 beyond): the weight rows are warp-uniform 16-byte loads (L1
// broadcast), the input row is this thread's own.
template <int NP>
__device__ __forceinline__ void map_row(const MapArgs& M, long long b, int n, bool valid, float (&u)[NP]) {
#pragma unroll
  for (int a = 0; a < NP; ++a) u[a] = (valid && a < n && M.bias) ? __ldg(M.bias + a) : 0.f;
  if (!valid) return;
  const float* xr = M.x + b * M.ldx;
  for (int k = 0; k < M.in_dim; k += 4) {
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + k));
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      if (a < n) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(M.w + a * M.ldw + k));
        u[a] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, u[a]))));
      }
    }
  }
  float* out = M.v_out + b * n;
#pragma unroll
  for (int a = 0; a < NP; ++a)
    if (a < n) out[a] = u[a];
}

// (value, tag) max over a group of `width` consecutive lanes; ties go to the smaller tag, which is
// the reference's evaluation order (linear rows, then quadratics, cones, LMI; torch.max first index).
__device__ __forceinline__ void group_argmax(float& val, int& tag, int width) {
  for (int off = width >> 1; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, val, off);
    const int ot = __shfl_xor_sync(0xffffffffu, tag, off);
    if (ov > val || (ov == val && ot < tag)) {
      val = ov;
      tag = ot;
    }
  }
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int off = WIDTH >> 1; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
template <int WIDTH>
__device__ __forceinline__ float group_max(float x) {
#pragma unroll
  for (int off = WIDTH >> 1; off > 0; off >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
  return x;
}

}  // namespace rayen
