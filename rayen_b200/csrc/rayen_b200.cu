// librayen_b200.so -- C ABI (include/rayen_b200.h) over the sm_100a kernels in lqs.cuh / lmi.cuh.
// Host side only: plan upload, kernel selection, launch geometry.  No torch types, no exceptions
// across the boundary, no synchronisation except in the *_host_* entry points.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "lmi.cuh"
#include "lmi_tc.cuh"
#include "lmi_warp.cuh"
#include "lqs.cuh"
#include "lqs_tc.cuh"
#include "viol.cuh"
#include "wide.cuh"
#include "lmi_big.cuh"
#include "lmi_big_tc.cuh"

using namespace rayen;

// ----------------------------------------------------------------------------- plan object
struct rayen_plan {
  PlanDev dev;
  int device;
  int sm_count;
  int max_smem_optin;
  int tune_tm, tune_lanes;
  int lmi_fwd_threads;  // 256 or 384 (RAYEN_LMI_THREADS overrides)
  float* d_blob;
  bool lqs_smem;  // LQS constants fit in shared memory
  bool lmi_smem;  // LMI matrices fit in shared memory
  size_t lqs_smem_bytes, lmi_smem_bytes, lmi_bwd_smem_bytes, lmi_grad_smem_bytes, viol_lmi_smem_bytes;
  bool viol_lmi_smem;
  bool wide;      // n > 32: the kernels of wide.cuh on the WIDE section (linear + quadratic + SOC, no LMI)
  WideDev wdev;
  size_t wide_fwd_smem_bytes[3], wide_bwd_smem_bytes;  // forward: tiles of 8 / 16 / 4 samples ([2]: n too wide for 8)
  // LMI beyond the register-resident kernels (lmi_big.cuh): dev.lmi_r stays 0 (the narrow kernels see "no LMI"), the
  // other families' kernel leaves the prior (kappa, tag, y) for every sample and the two kernels of lmi_big.cuh follow
  bool lmi_big;
  LmiBigDev bdev;
  int lmib_threads;        // CTA size of the solve kernel: 64 / 128 / 256 / 320 (>= r)
  size_t lmib_smem_bytes;
  int lmib_ctas_per_sm;
  int64_t lmib_ws_cap;     // bytes of the contracted-matrix buffer a forward call may use (RAYEN_LMIB_WS_MB, default 512)
  int64_t off_lmibt;       // LMIBT section: the contraction's tcgen05 B operand (0: FP32 GEMM only)
  int lmibt_panels, lmibt_slices;
  bool lmib_tc;            // contraction on the tensor cores (lmi_big_tc.cuh; RAYEN_LMIB_TC=0 keeps the FP32 GEMM)
  int off_lminegb;
  bool has_lqs;   // any non-zero linear row / quadratic / cone: otherwise the LQS forward kernel is skipped
  bool prune;     // LMI pruning enabled (needs has_lqs and a BOUND section)
  bool use_tc;    // tensor-core (tcgen05) linear/quadratic/SOC forward kernel
  size_t tc_smem_bytes;
  int lmi_tc_mode;        // LMI contraction as a tcgen05 GEMM inside the LMI forward kernel (lmi_tc.cuh):
                          // 0 never, 1 wherever available, 2 automatic (the measured policy in lmi_use_tc)
  bool lmi_tc_ok;         // ... available for this plan (rp >= 16, section present, fits in shared memory)
  bool lmi_tc_grad_fsmem; // the gradient variant keeps F~z in shared memory next to the GEMM buffers
  size_t lmi_tc_smem_bytes, lmi_tc_grad_smem_bytes;
  // filter + one-warp-per-matrix solver (lmi_warp.cuh) for launches in which the samples carry a prior kappa
  bool lmi_warp_ok;       // LMIW section present and the kernel's shared memory fits
  int lmi_warp_mode;      // 0 never, 1 wherever available, 2 automatic (padded size >= 16)
  bool lmi_warp_filter;   // the definiteness filter in front of the solver (RAYEN_LMI_FILTER=0: solver only)
  int lmi_warp_solves;    // failing samples a warp of the filter kernel solves itself before it hands over (default 1)
  size_t lmi_warp_smem_bytes;
  // host-buffer path only (rayen_forward_backward_host_f32): copy streams and events, created on first use and
  // serialised by host_mu -- the device-pointer entry points never touch them, so the plan stays re-entrant there
  std::mutex* host_mu;
  cudaStream_t host_in, host_out, host_main;  // copy-in, copy-out, kernels of the synchronous step
  cudaEvent_t host_fork, host_join;           // caller's stream -> host_main and back
  cudaEvent_t host_ev[4][2 * 8 + 2];  // per slot: v_c / gy_c landed (reused: forward_c / backward_c done), start, done
  bool host_ready;
  int host_chunks;  // 0 = automatic (RAYEN_HOST_CHUNKS overrides)
  // the synchronous host-buffer step as an instantiated CUDA graph per (buffers, batch): one cudaGraphLaunch instead of
  // ~25 API calls per step (the step moves 2 x 8 MB over PCIe in ~0.2 ms: submission cost shows)
  struct HostGraph {
    const void *v, *gy, *y, *gv, *ws;
    int64_t B;
    cudaStream_t stream;
    cudaGraphExec_t exec;
    unsigned long long used;
  };
  HostGraph host_graphs[8];
  int host_graph_count;
  unsigned long long host_graph_clock;
  bool host_graph_on;
};

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}
#define RAYEN_CUDA(call)                                  \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// An empty kernel, for the launch floor that bench.py reports next to the kernel times.
__global__ void rayen_empty_kernel() {}
extern "C" int rayen_launch_empty(int count, void* stream_) {
  for (int i = 0; i < count; ++i) rayen_empty_kernel<<<148, 128, 0, static_cast<cudaStream_t>(stream_)>>>();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "rayen_launch_empty: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return RAYEN_OK;
}

extern "C" int rayen_abi_version(void) { return RAYEN_ABI_VERSION; }
extern "C" const char* rayen_last_error(void) { return g_err; }
extern "C" int64_t rayen_launch_count(void) { return g_launches.load(); }

// ----------------------------------------------------------------------------- kernel tables
typedef void (*LqsFwdFn)(const PlanDev, const float*, long long, float*, float*, int*, long long, int, int, int, int,
                         int*, int*, const MapArgs);
typedef void (*LqsBwdFn)(const PlanDev, const float*, long long, const float*, const float*, const int*, float*,
                         long long, long long, int, int*, int*, const float*);
typedef void (*LmiFwdFn)(const PlanDev, const float*, long long, float*, float*, int*, long long, int, int,
                         const int*, const int*, float*);
typedef void (*LmiBwdFn)(const PlanDev, const float*, long long, const float*, const float*, const int*, float*,
                         long long, long long, int, const int*, const int*);

static int np_index(int np) { return np == 4 ? 0 : np == 8 ? 1 : np == 16 ? 2 : np == 32 ? 3 : -1; }
static int tm_index(int tm) { return tm == 1 ? 0 : tm == 2 ? 1 : tm == 4 ? 2 : -1; }

template <int NP, int TM>
static LqsFwdFn lqs_fwd_pick(bool smem) {
  return smem ? lqs_forward_kernel<NP, TM, true> : lqs_forward_kernel<NP, TM, false>;
}
template <int NP>
static LqsFwdFn lqs_fwd_pick_tm(int tm, bool smem) {
  switch (tm) {
    case 1: return lqs_fwd_pick<NP, 1>(smem);
    case 2: return lqs_fwd_pick<NP, 2>(smem);
    default: return lqs_fwd_pick<NP, 4>(smem);
  }
}
static LqsFwdFn lqs_fwd_fn(int np, int tm, bool smem) {
  switch (np) {
    case 4: return lqs_fwd_pick_tm<4>(tm, smem);
    case 8: return lqs_fwd_pick_tm<8>(tm, smem);
    case 16: return lqs_fwd_pick_tm<16>(tm, smem);
    default: return lqs_fwd_pick_tm<32>(tm, smem);
  }
}
typedef void (*LqsTcFn)(const PlanDev, const float*, long long, float*, float*, int*, long long, int, int, int, int*,
                        int*, const MapArgs);
static LqsTcFn lqs_tc_fn(int kp) {
  switch (kp) {
    case 8: return lqs_tc_forward_kernel<8>;
    case 16: return lqs_tc_forward_kernel<16>;
    default: return lqs_tc_forward_kernel<32>;
  }
}
static size_t lqs_tc_smem(int kp, int panels, bool y_stage = false) {
  switch (kp) {
    case 8: return lqs_tc_smem_bytes<8>(panels, y_stage);
    case 16: return lqs_tc_smem_bytes<16>(panels, y_stage);
    default: return lqs_tc_smem_bytes<32>(panels, y_stage);
  }
}
static LqsBwdFn lqs_bwd_fn(int np) {
  switch (np) {
    case 4: return lqs_backward_kernel<4>;
    case 8: return lqs_backward_kernel<8>;
    case 16: return lqs_backward_kernel<16>;
    default: return lqs_backward_kernel<32>;
  }
}
template <int THREADS, bool GRAD>
static LmiFwdFn lmi_fwd_fn_t(int rp, bool smem) {
  switch (rp) {
    case 4: return smem ? lmi_forward_kernel<4, true, THREADS, GRAD> : lmi_forward_kernel<4, false, THREADS, GRAD>;
    case 8: return smem ? lmi_forward_kernel<8, true, THREADS, GRAD> : lmi_forward_kernel<8, false, THREADS, GRAD>;
    case 16: return smem ? lmi_forward_kernel<16, true, THREADS, GRAD> : lmi_forward_kernel<16, false, THREADS, GRAD>;
    default: return smem ? lmi_forward_kernel<32, true, THREADS, GRAD> : lmi_forward_kernel<32, false, THREADS, GRAD>;
  }
}
// the gradient-carrying variant keeps the reflectors in registers: 256 threads only
static LmiFwdFn lmi_fwd_fn(int rp, bool smem, int threads, bool grad) {
  if (grad) return lmi_fwd_fn_t<256, true>(rp, smem);
  return threads == 384 ? lmi_fwd_fn_t<384, false>(rp, smem) : lmi_fwd_fn_t<256, false>(rp, smem);
}
static LmiFwdFn lmi_fwd_tc_fn(int rp, bool fsmem, bool grad) {
  if (rp == 16) {
    if (!grad) return lmi_forward_tc_kernel<16, false, false>;
    return fsmem ? lmi_forward_tc_kernel<16, true, true> : lmi_forward_tc_kernel<16, false, true>;
  }
  if (!grad) return lmi_forward_tc_kernel<32, false, false>;
  return fsmem ? lmi_forward_tc_kernel<32, true, true> : lmi_forward_tc_kernel<32, false, true>;
}
static size_t lmi_tc_smem(int rp, int n, int kp, int stages, bool fsmem) {
  return rp == 16 ? lmi_tc_smem_bytes<16>(n, kp, stages, fsmem) : lmi_tc_smem_bytes<32>(n, kp, stages, fsmem);
}
static LmiBwdFn lmi_bwd_fn(int rp, bool smem) {
  switch (rp) {
    case 4: return smem ? lmi_backward_kernel<4, true> : lmi_backward_kernel<4, false>;
    case 8: return smem ? lmi_backward_kernel<8, true> : lmi_backward_kernel<8, false>;
    case 16: return smem ? lmi_backward_kernel<16, true> : lmi_backward_kernel<16, false>;
    default: return smem ? lmi_backward_kernel<32, true> : lmi_backward_kernel<32, false>;
  }
}
typedef void (*ViolLmiFn)(const PlanDev, const float*, long long, float*, long long);
static ViolLmiFn viol_lmi_fn(int rp, bool smem) {
  switch (rp) {
    case 4: return smem ? viol_lmi_kernel<4, true> : viol_lmi_kernel<4, false>;
    case 8: return smem ? viol_lmi_kernel<8, true> : viol_lmi_kernel<8, false>;
    case 16: return smem ? viol_lmi_kernel<16, true> : viol_lmi_kernel<16, false>;
    default: return smem ? viol_lmi_kernel<32, true> : viol_lmi_kernel<32, false>;
  }
}
static size_t lmi_smem(int rp, bool smem, int n, int threads) {
  switch (rp) {
    case 4: return smem ? lmi_smem_bytes<4, true>(n, threads) : lmi_smem_bytes<4, false>(n, threads);
    case 8: return smem ? lmi_smem_bytes<8, true>(n, threads) : lmi_smem_bytes<8, false>(n, threads);
    case 16: return smem ? lmi_smem_bytes<16, true>(n, threads) : lmi_smem_bytes<16, false>(n, threads);
    default: return smem ? lmi_smem_bytes<32, true>(n, threads) : lmi_smem_bytes<32, false>(n, threads);
  }
}

// Launch `fn` behind the previous kernel of the stream with programmatic stream serialization (see common.cuh):
// its CTAs may be scheduled while that kernel drains; the kernel itself waits (pdl_wait) before it reads.
static cudaError_t launch_lmi_forward(LmiFwdFn fn, int blocks, int threads, size_t smem, cudaStream_t stream, bool pdl,
                                      const PlanDev& d, const float* v, long long ldv, float* y, float* kappa,
                                      int* active, long long B, int mode, int has_prior, const int* list,
                                      const int* count, float* dkappa) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fn, d, v, ldv, y, kappa, active, B, mode, has_prior, list, count, dkappa);
}

static int allow_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    // the attribute belongs to the function, not to the plan: always the device maximum, so that plans of different
    // sizes coexist in one process (a later, smaller plan must not lower the limit of an earlier one)
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
        static_cast<size_t>(optin) > bytes)
      bytes = static_cast<size_t>(optin);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  }
  return 0;
}

// ----------------------------------------------------------------------------- plan create / destroy
static int create_wide_plan(const RayenPlanDesc* d, int device, rayen_plan_t** out);

// ----------------------------------------------------------------------------- big LMI (lmi_big.cuh)
typedef void (*LmibSolveFn)(const LmiBigDev, const float*, const float*, long long, float*, float*, int*, float*,
                            long long, int, int, float*, int*, int*);
static LmibSolveFn lmib_solve_fn(int threads) {
  switch (threads) {
    case 64: return lmib_solve_kernel<64>;
    case 128: return lmib_solve_kernel<128>;
    case 256: return lmib_solve_kernel<256>;
    default: return lmib_solve_kernel<320>;
  }
}
static int allow_smem(const void* fn, size_t bytes);
static int lmib_validate(const RayenPlanDesc* d) {
  if (d->lmi_r < 1 || d->lmi_r > kLbMaxR)
    return fail(RAYEN_ERR_UNSUPPORTED, "LMI size %d: the one-CTA-per-matrix path covers 1 <= r <= %d", d->lmi_r, kLbMaxR);
  const int64_t p4 = (static_cast<int64_t>(d->lmi_r) * (d->lmi_r + 1) / 2 + 3) / 4 * 4;
  if (d->lmib_p4 != p4 || d->lmi_rp != 0 || d->off_lmib <= 0 || d->off_lmib % 4 ||
      d->off_lmib + static_cast<int64_t>(d->n) * p4 > d->blob_words || d->off_lminegb <= 0 || d->off_lminegb % 4 ||
      d->off_lminegb + static_cast<int64_t>(d->k + 1) * p4 > d->blob_words)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "LMIB / LMINEGB sections do not fit the block");
  if (d->off_lmibt != 0 &&
      (d->off_lmibt < 0 || d->off_lmibt % 4 || d->lmibt_panels != (p4 + 127) / 128 || d->lmibt_slices != (d->n + 31) / 32 ||
       d->off_lmibt + static_cast<int64_t>(d->lmibt_panels) * d->lmibt_slices * 2 * kLbtTile > d->blob_words))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "LMIBT section does not fit the block");
  return 0;
}
// device must be current; p->sm_count / max_smem_optin / d_blob set
static int lmib_setup(rayen_plan* p, const RayenPlanDesc* d) {
  p->lmi_big = true;
  LmiBigDev& b = p->bdev;
  b.blob = p->d_blob;
  b.n = d->n; b.k = d->k; b.r = d->lmi_r; b.p4 = d->lmib_p4;
  b.off_lmib = static_cast<int>(d->off_lmib);
  b.off_y0 = static_cast<int>(d->off_y0);
  p->off_lminegb = static_cast<int>(d->off_lminegb);
  p->off_lmibt = d->off_lmibt;
  p->lmibt_panels = d->lmibt_panels;
  p->lmibt_slices = d->lmibt_slices;
  p->lmib_tc = d->off_lmibt > 0 && d->lmibt_panels > 0 && d->lmibt_slices > 0 &&
               lmib_tc_smem_bytes() <= static_cast<size_t>(p->max_smem_optin) &&
               !(getenv("RAYEN_LMIB_TC") && atoi(getenv("RAYEN_LMIB_TC")) == 0);
  p->lmib_threads = b.r <= 64 ? 64 : (b.r <= 128 ? 128 : (b.r <= 256 ? 256 : 320));
  p->lmib_smem_bytes = lmib_smem_bytes(b.r);
  if (p->lmib_smem_bytes > static_cast<size_t>(p->max_smem_optin))
    return fail(RAYEN_ERR_UNSUPPORTED, "LMI size %d needs %zu bytes of shared memory", b.r, p->lmib_smem_bytes);
  int per_sm = static_cast<int>(static_cast<size_t>(p->max_smem_optin) / (p->lmib_smem_bytes + 1024));
  const int by_threads = 2048 / p->lmib_threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > 16) per_sm = 16;
  if (per_sm < 1) per_sm = 1;
  p->lmib_ctas_per_sm = per_sm;
  p->lmib_ws_cap = 512ll << 20;
  if (const char* env = getenv("RAYEN_LMIB_WS_MB")) {
    const long long mb = atoll(env);
    if (mb >= 1 && mb <= 65536) p->lmib_ws_cap = mb << 20;
  }
  // the attribute is per function: the device maximum, so that plans of different sizes coexist
  int rc = 0;
  for (int t : {64, 128, 256, 320})
    if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(lmib_solve_fn(t)), p->max_smem_optin);
  if (rc == 0 && p->lmib_tc) rc = allow_smem(reinterpret_cast<const void*>(lmib_contract_tc_kernel), lmib_tc_smem_bytes());
  return rc;
}
// samples per chunk of the contraction buffer for a call of B samples (p4 words each)
static int64_t lmib_chunk_rows(const rayen_plan* p, int64_t B, int p4) {
  int64_t rows = p->lmib_ws_cap / (static_cast<int64_t>(p4) * 4);
  rows = rows / kLbTileM * kLbTileM;
  if (rows < kLbTileM) rows = kLbTileM;
  return rows < B ? rows : B;
}
static int64_t lmib_solve_grid(const rayen_plan* p, int64_t rows) {
  const int64_t cap = static_cast<int64_t>(p->sm_count) * p->lmib_ctas_per_sm;
  return rows < cap ? (rows < 1 ? 1 : rows) : cap;
}
// bytes behind the common workspace prefix: the contraction buffer
static int64_t lmib_ws_extra(const rayen_plan* p, int64_t B, int p4) {
  const int64_t rows = lmib_chunk_rows(p, B, p4);
  return 256 + (rows * p4 * 4 + 255) / 256 * 256;  // 256: the buffer is aligned up inside the caller's workspace
}
// The two kernels of lmi_big.cuh over the batch, chunk by chunk.  F: [nv][p4] packed matrices (LMIB with V = v, or
// LMINEGB with V = y and its constant row as C0); buf: the contraction buffer.
static cudaError_t lmib_run(const rayen_plan* p, const float* F, int nv, const float* V, int64_t ldv, float* y, float* kappa,
                            int32_t* active, float* dkappa, void* buf, int64_t B, int mode, int flags, cudaStream_t stream,
                            const float* C0 = nullptr, int* grad_list = nullptr, int* grad_count = nullptr) {
  LmiBigDev b = p->bdev;
  b.n = nv;
  b.off_lmib = static_cast<int>(F - p->d_blob);
  const int64_t rows = lmib_chunk_rows(p, B, b.p4);
  buf = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf) + 255) / 256 * 256);  // 16-byte stores of the GEMM
  float* S = static_cast<float*>(buf);
  LmibSolveFn sf = lmib_solve_fn(p->lmib_threads);
  const int tiles_n = (b.p4 + kLbTileN - 1) / kLbTileN;
  for (int64_t c0 = 0; c0 < B; c0 += rows) {
    const int64_t bc = (B - c0 < rows) ? B - c0 : rows;
    const int64_t tiles_m = (bc + kLbTileM - 1) / kLbTileM;
    if (p->lmib_tc && C0 == nullptr && F == p->d_blob + p->bdev.off_lmib) {
      // tensor-core contraction: ceil(bc / 128) x ceil(panels / 4) CTAs
      const int64_t groups = (p->lmibt_panels + kLbtPanelsPerCta - 1) / kLbtPanelsPerCta;
      lmib_contract_tc_kernel<<<static_cast<unsigned>(tiles_m * groups), kLbtThreads, lmib_tc_smem_bytes(), stream>>>(
          V + c0 * ldv, ldv, p->d_blob + p->off_lmibt, nv, b.p4, p->lmibt_panels, p->lmibt_slices, S, bc);
    } else {
      lmib_contract_kernel<<<static_cast<unsigned>(tiles_n * tiles_m), kLbGemmThreads, 0, stream>>>(
          V + c0 * ldv, ldv, F, nv, b.p4, S, bc, C0);
    }
    g_launches.fetch_add(1);
    // with gradients: the LMI-bound samples leave their weight rows in S and their numbers on a list; one GEMM over the
    // list turns them into d kappa/du (lmib_grad_gemm_kernel)
    const bool gemm_grad = dkappa != nullptr && grad_list != nullptr && (flags & (kLbFlagGrad | kLbFlagGradOnly));
    if (gemm_grad) {
      cudaError_t me = cudaMemsetAsync(grad_count, 0, sizeof(int), stream);
      if (me != cudaSuccess) return me;
    }
    sf<<<static_cast<unsigned>(lmib_solve_grid(p, bc)), p->lmib_threads, p->lmib_smem_bytes, stream>>>(
        b, S, V + c0 * ldv, ldv, y ? y + c0 * b.k : nullptr, kappa + c0, active ? active + c0 : nullptr,
        dkappa ? dkappa + c0 * nv : nullptr, bc, mode, flags, S, gemm_grad ? grad_list : nullptr,
        gemm_grad ? grad_count : nullptr);
    g_launches.fetch_add(1);
    if (gemm_grad) {
      const int64_t gt_n = (nv + kLbGradTile - 1) / kLbGradTile, gt_m = (bc + kLbGradTile - 1) / kLbGradTile;
      lmib_grad_gemm_kernel<<<static_cast<unsigned>(gt_n * gt_m), kLbGradThreads, 0, stream>>>(
          S, grad_list, grad_count, F, nv, b.p4, dkappa + c0 * nv);
      g_launches.fetch_add(1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}


extern "C" int rayen_plan_create(const RayenPlanDesc* d, int device, rayen_plan_t** out) {
  if (!d || !out) return fail(RAYEN_ERR_BAD_ARGUMENT, "rayen_plan_create: null argument");
  *out = nullptr;
  if (d->abi_version != RAYEN_ABI_VERSION)
    return fail(RAYEN_ERR_ABI, "plan descriptor has ABI %d, library has %d", d->abi_version, RAYEN_ABI_VERSION);
  if (d->n < 1 || d->k < d->n) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad dimensions n=%d k=%d", d->n, d->k);
  if (d->lmi_big && d->lmi_r > 0) {
    const int rcb = lmib_validate(d);
    if (rcb) return rcb;
  }
  if (d->wide) return create_wide_plan(d, device, out);
  const int narrow_r = d->lmi_big ? 0 : d->lmi_r;  // LMI size as the register-resident kernels see it
  if (np_index(d->np) < 0 || d->np < d->n)
    return fail(RAYEN_ERR_UNSUPPORTED, "n=%d (np=%d): the register-resident kernels cover n <= 32 (wide plans: wide = 1)", d->n, d->np);
  if (narrow_r > 0 && (np_index(d->lmi_rp) < 0 || d->lmi_rp < d->lmi_r))
    return fail(RAYEN_ERR_UNSUPPORTED, "LMI size %d (padded %d): the register-resident kernels cover r <= 32 (larger: lmi_big = 1)", d->lmi_r, d->lmi_rp);
  if (d->m_pad % 4 || d->m_pad < d->m || d->m_pad < 4 || d->k_pad % 4 || d->k_pad < d->k)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "bad padding m=%d m_pad=%d k=%d k_pad=%d", d->m, d->m_pad, d->k, d->k_pad);
  if (!d->blob || d->blob_words <= 0 || d->blob_words % 4)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "bad constant block (%lld words)", static_cast<long long>(d->blob_words));
  const int64_t offs[7] = {d->off_lin, d->off_quad, d->off_soc, d->off_nmat, d->off_y0, d->off_bound, d->off_lmi};
  for (int i = 0; i < 7; ++i) {
    if (offs[i] < 0 || offs[i] % 4 || offs[i] >= d->blob_words ||
        (i > 0 && offs[i] < offs[i - 1]))
      return fail(RAYEN_ERR_BAD_ARGUMENT, "section offset %d = %lld is invalid", i, static_cast<long long>(offs[i]));
  }
  const int tri = (d->np / 4) * (d->np / 4 + 1) * 8;
  if (d->lin_chunk_stride < 4 * d->np || d->lin_chunk_stride % 4 ||
      d->off_lin + static_cast<int64_t>(d->m_pad / 4) * d->lin_chunk_stride > d->off_quad)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "linear section does not fit its slot");
  if (d->n_quad < 0 || (d->n_quad > 0 && (d->quad_stride < d->np + tri || d->quad_stride % 4 ||
                                           d->off_quad + static_cast<int64_t>(d->n_quad) * d->quad_stride > d->off_soc)))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "quadratic section does not fit its slot");
  if (d->n_soc < 0 || (d->n_soc > 0 && (d->soc_stride < 2 * d->np + tri + 4 || d->soc_stride % 4 ||
                                         d->off_soc + static_cast<int64_t>(d->n_soc) * d->soc_stride > d->off_nmat)))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "SOC section does not fit its slot");
  if (!d->n_is_identity && d->off_nmat + static_cast<int64_t>(d->k) * (d->np + 4) > d->off_y0)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "N section does not fit its slot");
  if (d->off_y0 + d->k_pad > d->off_bound) return fail(RAYEN_ERR_BAD_ARGUMENT, "y0 section does not fit its slot");
  if (d->lmi_prune && d->off_bound + d->np + tri + 4 > d->off_lmi)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "bound section does not fit its slot");
  if (d->lmi_prune && !(d->lmi_bound_margin >= 0.f))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "lmi_bound_margin must be >= 0");
  if (narrow_r > 0 && d->off_lmi + static_cast<int64_t>(d->n) * d->lmi_rp * d->lmi_rp > d->blob_words)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "LMI section does not fit the block");
  if (d->blob_words > (1ll << 30)) return fail(RAYEN_ERR_UNSUPPORTED, "constant block too large");
  if (d->off_viol < 0 || d->off_viol % 4 || d->off_viol >= d->blob_words || d->off_lmineg < d->off_viol || d->off_lmineg % 4 ||
      d->off_lmineg >= d->blob_words || d->viol_in < 0 || d->viol_eq < 0 ||
      (narrow_r > 0 && d->off_lmineg + static_cast<int64_t>(d->k + 1) * d->lmi_rp * d->lmi_rp > d->blob_words))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "violation sections do not fit the block");
  if (d->tc_panels < 1 || (d->tc_kp != 8 && d->tc_kp != 16 && d->tc_kp != 32) || d->tc_kp < d->np || d->off_tc % 4 ||
      d->off_tc < d->off_lmi ||
      d->off_tc + static_cast<int64_t>(d->tc_panels) * (kTcTableWords + 2 * kTcPanel * d->tc_kp) > d->blob_words)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "tensor-core section does not fit the block");

  if (d->lmitc_panels < 0 || d->off_lmitc < 0 || d->off_lmitc % 4 || d->off_lmitc >= d->blob_words ||
      (d->lmitc_panels > 0 && (d->lmi_rp < 16 || d->lmitc_panels != d->lmi_rp * d->lmi_rp / 128 ||
                               d->off_lmitc + static_cast<int64_t>(d->lmitc_panels) * 2 * 128 * d->tc_kp > d->blob_words)))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "LMI tensor-core section does not fit the block");

  if (d->off_lmiw < 0 || d->off_lmiw % 4 ||
      (narrow_r > 0 && d->off_lmiw > 0 && d->off_lmiw + static_cast<int64_t>(d->n) * kLwMatWords > d->blob_words))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "LMIW section does not fit the block");

  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return fail(RAYEN_ERR_NO_DEVICE, "CUDA device %d is not available (%d devices visible)", device, count);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  RAYEN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RAYEN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    cudaSetDevice(prev);
    return fail(RAYEN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  }

  rayen_plan* p = new (std::nothrow) rayen_plan();
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "out of host memory");
  memset(p, 0, sizeof(*p));
  p->device = device;
  p->host_mu = new (std::nothrow) std::mutex();
  {
    const char* env = getenv("RAYEN_HOST_CHUNKS");
    p->host_chunks = env ? atoi(env) : 0;
    if (p->host_chunks < 0 || p->host_chunks > 8) p->host_chunks = 0;
  }
  p->host_graph_on = !(getenv("RAYEN_HOST_GRAPH") && atoi(getenv("RAYEN_HOST_GRAPH")) == 0);
  p->sm_count = prop.multiProcessorCount;
  p->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  cudaError_t e = cudaMalloc(&p->d_blob, d->blob_words * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(p->d_blob, d->blob, d->blob_words * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->d_blob) cudaFree(p->d_blob);
    delete p->host_mu;
    delete p;
    cudaSetDevice(prev);
    return cuda_fail(e, "uploading the constant block");
  }
  PlanDev& v = p->dev;
  v.blob = p->d_blob;
  v.n = d->n; v.k = d->k; v.np = d->np; v.k_pad = d->k_pad;
  v.m = d->m; v.m_pad = d->m_pad; v.n_quad = d->n_quad; v.n_soc = d->n_soc;
  v.lmi_r = narrow_r; v.lmi_rp = d->lmi_rp; v.n_is_identity = d->n_is_identity;
  v.lin_stride = d->lin_chunk_stride; v.quad_stride = d->quad_stride; v.soc_stride = d->soc_stride;
  v.off_lin = static_cast<int>(d->off_lin); v.off_quad = static_cast<int>(d->off_quad);
  v.off_soc = static_cast<int>(d->off_soc); v.off_nmat = static_cast<int>(d->off_nmat);
  v.off_y0 = static_cast<int>(d->off_y0); v.off_bound = static_cast<int>(d->off_bound);
  v.off_lmi = static_cast<int>(d->off_lmi);
  v.lmi_prune = (d->lmi_prune && narrow_r > 0) ? 1 : 0;
  v.lqs_words = static_cast<int>(d->off_lmi - d->off_lin);
  v.lmi_words = narrow_r > 0 ? d->n * d->lmi_rp * d->lmi_rp : 0;
  v.off_tc = static_cast<int>(d->off_tc); v.tc_panels = d->tc_panels; v.tc_kp = d->tc_kp;
  v.off_viol = static_cast<int>(d->off_viol); v.off_lmineg = static_cast<int>(d->off_lmineg);
  v.viol_in = d->viol_in; v.viol_eq = d->viol_eq;
  v.off_lmitc = static_cast<int>(d->off_lmitc); v.lmitc_panels = d->lmitc_panels;
  v.lmi_bound_margin = d->lmi_bound_margin;
  v.off_lmiw = static_cast<int>(d->off_lmiw);

  p->has_lqs = d->n_quad > 0 || d->n_soc > 0;
  for (int64_t i = d->off_lin; i < d->off_quad && !p->has_lqs; ++i) p->has_lqs = d->blob[i] != 0.0f;
  p->prune = p->has_lqs && v.lmi_prune;
  {
    const char* env = getenv("RAYEN_LMI_PRUNE");
    if (env && atoi(env) == 0) p->prune = false;
  }
  p->lqs_smem_bytes = 64 + static_cast<size_t>(v.lqs_words) * 4;
  p->lqs_smem = p->lqs_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
  if (!p->lqs_smem) p->lqs_smem_bytes = 64;
  p->tc_smem_bytes = lqs_tc_smem(v.tc_kp, v.tc_panels);
  // y staging tiles (coalesced stores, rayen_plan_set_coalesced_output): off by default -- measured neutral to slightly
  // slower on local HBM (cfg5 25.5 vs 25.4 us, cfg3 11.6 vs 11.0 us) -- and what makes y stores through a peer / multicast
  // mapping efficient (sharding.forward_gathered turns it on: fused all-gather epilogue 124 -> 77 us on 2 GPUs)
  if (getenv("RAYEN_TC_YSTAGE") && atoi(getenv("RAYEN_TC_YSTAGE")) == 1 &&
      lqs_tc_smem(v.tc_kp, v.tc_panels, true) <= static_cast<size_t>(p->max_smem_optin)) {
    p->tc_smem_bytes = lqs_tc_smem(v.tc_kp, v.tc_panels, true);
    v.tc_y_stage = 1;
  }
  // measured on B200 (scripts/time_kernels.py): the GEMM formulation wins from K = 16 up; at K = 8 the FP32-pipe
  // kernel is faster (the GEMM is too thin to pay for the TMEM round trip)
  p->use_tc = p->tc_smem_bytes <= static_cast<size_t>(p->max_smem_optin) && v.np >= 16;
  {
    const char* env = getenv("RAYEN_LQS_TC");
    if (env) p->use_tc = atoi(env) != 0 && p->tc_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
  }
  int rc = 0;
  if (p->tc_smem_bytes <= static_cast<size_t>(p->max_smem_optin))
    rc = allow_smem(reinterpret_cast<const void*>(lqs_tc_fn(v.tc_kp)), p->tc_smem_bytes);
  for (int tm = 1; tm <= 4 && rc == 0; tm *= 2)
    rc = allow_smem(reinterpret_cast<const void*>(lqs_fwd_fn(v.np, tm, p->lqs_smem)), p->lqs_smem_bytes);
  if (rc == 0 && v.lmi_r > 0) {
    {
      const char* env = getenv("RAYEN_LMI_THREADS");
      p->lmi_fwd_threads = (env && atoi(env) == 256) ? 256 : ((env && atoi(env) == 384) ? 384 : kLmiFwdThreads);
    }
    p->lmi_smem_bytes = lmi_smem(v.lmi_rp, true, v.n, p->lmi_fwd_threads);
    p->lmi_smem = p->lmi_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
    if (!p->lmi_smem) p->lmi_smem_bytes = lmi_smem(v.lmi_rp, false, v.n, p->lmi_fwd_threads);
    rc = allow_smem(reinterpret_cast<const void*>(lmi_fwd_fn(v.lmi_rp, p->lmi_smem, p->lmi_fwd_threads, false)), p->lmi_smem_bytes);
    p->lmi_grad_smem_bytes = lmi_smem(v.lmi_rp, p->lmi_smem, v.n, 256);
    p->viol_lmi_smem_bytes = lmi_smem(v.lmi_rp, true, v.k + 1, kLmiThreads);
    p->viol_lmi_smem = p->viol_lmi_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
    if (!p->viol_lmi_smem) p->viol_lmi_smem_bytes = lmi_smem(v.lmi_rp, false, v.k + 1, kLmiThreads);
    if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(viol_lmi_fn(v.lmi_rp, p->viol_lmi_smem)), p->viol_lmi_smem_bytes);
    if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(lmi_fwd_fn(v.lmi_rp, p->lmi_smem, 256, true)), p->lmi_grad_smem_bytes);
    p->lmi_bwd_smem_bytes = lmi_smem(v.lmi_rp, p->lmi_smem, v.n, kLmiThreads);
    if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(lmi_bwd_fn(v.lmi_rp, p->lmi_smem)), p->lmi_bwd_smem_bytes);
    if (v.lmitc_panels > 0) {
      // ring depth: all panels resident if that fits, else as many stages (>= 2) as the budget allows
      int stages = v.lmitc_panels < kLmiTcMaxStages ? v.lmitc_panels : kLmiTcMaxStages;
      while (stages > 2 && lmi_tc_smem(v.lmi_rp, v.n, v.tc_kp, stages, false) > static_cast<size_t>(p->max_smem_optin))
        --stages;
      v.lmitc_stages = stages;
      p->lmi_tc_smem_bytes = lmi_tc_smem(v.lmi_rp, v.n, v.tc_kp, stages, false);
      p->lmi_tc_grad_smem_bytes = lmi_tc_smem(v.lmi_rp, v.n, v.tc_kp, stages, true);
      p->lmi_tc_grad_fsmem = p->lmi_tc_grad_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
      if (!p->lmi_tc_grad_fsmem) p->lmi_tc_grad_smem_bytes = p->lmi_tc_smem_bytes;  // F~z through L1/L2 instead
      p->lmi_tc_ok = p->lmi_tc_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
      p->lmi_tc_mode = 2;
      const char* env = getenv("RAYEN_LMI_TC");
      if (env && atoi(env) >= 0 && atoi(env) <= 2) p->lmi_tc_mode = atoi(env);
      if (rc == 0 && p->lmi_tc_ok)
        rc = allow_smem(reinterpret_cast<const void*>(lmi_fwd_tc_fn(v.lmi_rp, false, false)), p->lmi_tc_smem_bytes);
      if (rc == 0 && p->lmi_tc_ok)
        rc = allow_smem(reinterpret_cast<const void*>(lmi_fwd_tc_fn(v.lmi_rp, p->lmi_tc_grad_fsmem, true)),
                        p->lmi_tc_grad_smem_bytes);
    }
  }
  if (rc == 0 && v.lmi_r > 0 && v.off_lmiw > 0) {
    p->lmi_warp_smem_bytes = lmi_warp_smem_bytes(v.n, kLwThreads);
    p->lmi_warp_ok = p->lmi_warp_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
    p->lmi_warp_mode = 2;
    p->lmi_warp_filter = true;
    // Unlimited by default: every warp solves the failing samples of its own chunks, and there is no second launch.
    // Failures are rare where the filter pays (cfg5: 1 in 2100) and spread evenly where they are not (chunks are dealt
    // round-robin); RAYEN_LMI_WARP_SOLVES=k caps the solves per warp and hands the rest to a second launch.
    p->lmi_warp_solves = 0x7fffffff;
    if (const char* sv = getenv("RAYEN_LMI_WARP_SOLVES")) p->lmi_warp_solves = atoi(sv) < 0 ? 0 : atoi(sv);
    const char* env = getenv("RAYEN_LMI_WARP");
    if (env && atoi(env) >= 0 && atoi(env) <= 2) p->lmi_warp_mode = atoi(env);
    env = getenv("RAYEN_LMI_FILTER");
    if (env && atoi(env) == 0) p->lmi_warp_filter = false;
    if (p->lmi_warp_ok) {
      rc = allow_smem(reinterpret_cast<const void*>(lmi_forward_warp_kernel<false>), p->lmi_warp_smem_bytes);
      if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(lmi_forward_warp_kernel<true>), p->lmi_warp_smem_bytes);
    }
  }
  if (rc == 0 && d->lmi_big && d->lmi_r > 0) rc = lmib_setup(p, d);
  // the violation checker keeps one y row per warp in shared memory: large ambient dimensions need the opt-in limit
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(viol_lqs_kernel), p->max_smem_optin);
  cudaSetDevice(prev);
  if (rc != 0) {
    cudaFree(p->d_blob);
    delete p->host_mu;
    delete p;
    return rc;
  }
  *out = p;
  return RAYEN_OK;
}

// Wide plan (32 < n <= 4096; linear + quadratic + SOC): only the WIDE, Y0 and VIOL sections are used on the device.
static int create_wide_plan(const RayenPlanDesc* d, int device, rayen_plan_t** out) {
  if (d->n <= 32 || d->n > kWideMaxN || d->np < d->n || d->np % 4)
    return fail(RAYEN_ERR_UNSUPPORTED, "wide plan: n=%d (np=%d) outside 33..%d", d->n, d->np, kWideMaxN);
  if (d->lmi_r > 0 && !d->lmi_big)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "an LMI together with n=%d > 32 needs lmi_big = 1 (section LMIB)", d->n);
  if (d->k_pad % 4 || d->k_pad < d->k) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad padding k=%d k_pad=%d", d->k, d->k_pad);
  if (!d->blob || d->blob_words <= 0 || d->blob_words % 4 || d->blob_words > (1ll << 30))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "bad constant block (%lld words)", static_cast<long long>(d->blob_words));
  if (d->off_wide < 0 || d->off_wide % 4 || d->off_wide + 16 > d->blob_words || d->off_y0 < 0 || d->off_y0 % 4 ||
      d->off_y0 + d->k_pad > d->blob_words || d->off_viol < 0 || d->off_viol % 4 || d->off_viol >= d->blob_words ||
      d->viol_in < 0 || d->viol_eq < 0 || d->n_quad < 0 || d->n_soc < 0)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: section offsets are invalid");
  const int32_t* h = reinterpret_cast<const int32_t*>(d->blob + d->off_wide);
  WideDev w{};
  w.n = d->n; w.k = d->k;
  w.r_pad = h[1]; w.n_tasks = h[2]; w.off_tasks = h[3]; w.off_wt = h[4]; w.off_nt = h[5]; w.off_nrow = h[6];
  w.k32 = h[7]; w.np = h[8]; w.off_items = h[9]; w.n_quad = h[10]; w.n_soc = h[11]; w.n_rounds = h[12]; w.off_rounds = h[13];
  w.off_y0 = static_cast<int>(d->off_y0); w.n_is_identity = d->n_is_identity;
  const int64_t words = d->blob_words;
  const int n_items = w.n_quad + w.n_soc;
  if (h[0] != kWideMagic || h[14] != kWideVersion || w.r_pad < kWideGroupRows || w.r_pad % kWideGroupRows || w.n_tasks < 1 || w.off_tasks < 0 ||
      w.off_tasks + static_cast<int64_t>(w.n_tasks) * 8 > words || w.off_wt < 0 ||
      w.off_wt + static_cast<int64_t>(w.n) * w.r_pad > words || w.n_quad != d->n_quad || w.n_soc != d->n_soc ||
      w.off_items < 0 || w.off_items + static_cast<int64_t>(n_items) * 8 > words || w.n_rounds < 1 || w.off_rounds < 0 ||
      w.off_rounds + static_cast<int64_t>(w.n_rounds) * 4 > words || w.np != d->np || w.k32 < d->k ||
      (!d->n_is_identity && (w.off_nt <= 0 || w.off_nt + static_cast<int64_t>(w.n) * w.k32 > words || w.off_nrow <= 0 ||
                             w.off_nrow + static_cast<int64_t>(w.k) * w.np > words)))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: the WIDE header does not describe this block");
  const int32_t* tasks = reinterpret_cast<const int32_t*>(d->blob + w.off_tasks);
  for (int t = 0; t < w.n_tasks; ++t) {
    const int32_t* tk = tasks + 8 * t;
    const int kind = tk[0], row = tk[1], j0 = tk[2], idx = tk[3], slot = tk[5];
    if ((kind != 1 && kind != 2 && kind != 4) || row < 0 || row % kWideGroupRows || row + kWideGroupRows > w.r_pad || j0 < 0 ||
        j0 % 4 || j0 > w.n || idx < 0 || (kind == 2 && (slot < 0 || slot >= kWideSlots || idx >= kWideRoundItems)) ||
        (kind == 4 && idx >= kWideRoundItems))
      return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: task %d is invalid", t);
  }
  const int32_t* rounds = reinterpret_cast<const int32_t*>(d->blob + w.off_rounds);
  for (int r = 0; r < w.n_rounds; ++r) {
    const int32_t* rd = rounds + 4 * r;
    if (rd[0] != (r ? rounds[4 * r - 3] : 0) || rd[1] < rd[0] || rd[1] > w.n_tasks || rd[2] != (r ? rounds[4 * r - 1] : 0) ||
        rd[3] < rd[2] || rd[3] - rd[2] > kWideRoundItems || rd[3] > n_items)
      return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: round %d is invalid", r);
  }
  if (rounds[4 * w.n_rounds - 3] != w.n_tasks || rounds[4 * w.n_rounds - 1] != n_items)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: the rounds do not cover the tasks and items");
  const int32_t* items = reinterpret_cast<const int32_t*>(d->blob + w.off_items);
  for (int i = 0; i < n_items; ++i) {
    const int32_t* it = items + 8 * i;
    const int kind = it[1];
    if (it[0] < 0 || it[0] % kWideGroupRows || kind != (i < w.n_quad ? 2 : 3) || it[0] + w.n > w.r_pad ||
        it[2] != (i < w.n_quad ? i : i - w.n_quad) || it[3] < 0 || it[4] < 1 || it[3] + it[4] > kWideSlots || it[6] < 0 ||
        it[6] % 2 || it[6] + 2 > w.r_pad)
      return fail(RAYEN_ERR_BAD_ARGUMENT, "wide plan: item %d is invalid", i);
  }

  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return fail(RAYEN_ERR_NO_DEVICE, "CUDA device %d is not available (%d devices visible)", device, count);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  RAYEN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RAYEN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    cudaSetDevice(prev);
    return fail(RAYEN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  }
  rayen_plan* p = new (std::nothrow) rayen_plan();
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "out of host memory");
  memset(p, 0, sizeof(*p));
  p->device = device;
  p->host_mu = new (std::nothrow) std::mutex();
  p->sm_count = prop.multiProcessorCount;
  p->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  p->wide = true;
  p->has_lqs = true;
  p->host_graph_on = !(getenv("RAYEN_HOST_GRAPH") && atoi(getenv("RAYEN_HOST_GRAPH")) == 0);
  p->wide_fwd_smem_bytes[0] = wide_fwd_smem_bytes(w.n, 8);
  p->wide_fwd_smem_bytes[1] = wide_fwd_smem_bytes(w.n, 16);
  p->wide_fwd_smem_bytes[2] = wide_fwd_smem_bytes(w.n, 4);
  p->wide_bwd_smem_bytes = wide_bwd_smem_bytes(w.n);
  cudaError_t e = cudaMalloc(&p->d_blob, d->blob_words * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, d->blob, d->blob_words * sizeof(float), cudaMemcpyHostToDevice);
  int rc = 0;
  if (e != cudaSuccess) rc = cuda_fail(e, "uploading the constant block");
  // the attribute is per function, not per plan: always the device maximum, so that plans of different n coexist
  if (rc == 0 && (p->wide_fwd_smem_bytes[2] > static_cast<size_t>(p->max_smem_optin) ||
                  p->wide_bwd_smem_bytes > static_cast<size_t>(p->max_smem_optin)))
    rc = fail(RAYEN_ERR_UNSUPPORTED, "wide plan: n=%d needs %zu bytes of shared memory", w.n, p->wide_fwd_smem_bytes[2]);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_forward_kernel<4, true>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_forward_kernel<8, false>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_forward_kernel<16, false>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_forward_kernel<8, true>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_forward_kernel<16, true>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_backward_kernel<128>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(wide_backward_kernel<256>), p->max_smem_optin);
  if (rc == 0) rc = allow_smem(reinterpret_cast<const void*>(viol_lqs_kernel), p->max_smem_optin);
  if (rc == 0 && d->lmi_big && d->lmi_r > 0) rc = lmib_setup(p, d);
  cudaSetDevice(prev);
  if (rc != 0) {
    if (p->d_blob) cudaFree(p->d_blob);
    delete p->host_mu;
    delete p;
    return rc;
  }
  w.blob = p->d_blob;
  p->wdev = w;
  // the violation checker and the host-buffer path read these
  PlanDev& v = p->dev;
  v.blob = p->d_blob;
  v.n = d->n; v.k = d->k; v.np = d->np; v.k_pad = d->k_pad;
  v.m = d->m; v.m_pad = d->m_pad; v.n_quad = d->n_quad; v.n_soc = d->n_soc;
  v.n_is_identity = d->n_is_identity;
  v.off_y0 = static_cast<int>(d->off_y0);
  v.off_viol = static_cast<int>(d->off_viol); v.off_lmineg = static_cast<int>(d->off_lmineg);
  v.viol_in = d->viol_in; v.viol_eq = d->viol_eq;
  *out = p;
  return RAYEN_OK;
}

extern "C" void rayen_plan_destroy(rayen_plan_t* p) {
  if (!p) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaFree(p->d_blob);
  for (int i = 0; i < p->host_graph_count; ++i) cudaGraphExecDestroy(p->host_graphs[i].exec);
  if (p->host_ready) {
    cudaStreamDestroy(p->host_in);
    cudaStreamDestroy(p->host_out);
    cudaStreamDestroy(p->host_main);
    cudaEventDestroy(p->host_fork);
    cudaEventDestroy(p->host_join);
    for (auto& slot_ev : p->host_ev)
      for (cudaEvent_t ev : slot_ev) cudaEventDestroy(ev);
  }
  cudaSetDevice(prev);
  delete p->host_mu;
  delete p;
}

extern "C" int rayen_plan_set_tuning(rayen_plan_t* p, int tm, int lanes) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (tm != 0 && tm_index(tm) < 0) return fail(RAYEN_ERR_BAD_ARGUMENT, "samples_per_thread must be 0, 1, 2 or 4");
  if (lanes != 0 && (lanes < 1 || lanes > 32 || (lanes & (lanes - 1))))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "lanes_per_sample must be 0 or a power of two <= 32");
  p->tune_tm = tm;
  p->tune_lanes = lanes;
  return RAYEN_OK;
}

extern "C" int rayen_plan_set_coalesced_output(rayen_plan_t* p, int enabled) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (p->wide || p->dev.tc_kp == 0) return RAYEN_OK;  // only the tcgen05 kernel stages its output
  const size_t with = lqs_tc_smem(p->dev.tc_kp, p->dev.tc_panels, true), without = lqs_tc_smem(p->dev.tc_kp, p->dev.tc_panels, false);
  const bool on = enabled && with <= static_cast<size_t>(p->max_smem_optin);
  if (on) {
    int prev = 0;
    RAYEN_CUDA(cudaGetDevice(&prev));
    if (prev != p->device) RAYEN_CUDA(cudaSetDevice(p->device));
    const int rc = allow_smem(reinterpret_cast<const void*>(lqs_tc_fn(p->dev.tc_kp)), with);
    if (prev != p->device) cudaSetDevice(prev);
    if (rc) return rc;
  }
  p->dev.tc_y_stage = on ? 1 : 0;
  p->tc_smem_bytes = on ? with : without;
  return RAYEN_OK;
}

extern "C" int rayen_plan_set_tensor_cores(rayen_plan_t* p, int enabled) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  p->use_tc = enabled && p->tc_smem_bytes <= static_cast<size_t>(p->max_smem_optin);
  return RAYEN_OK;
}

#ifdef RAYEN_TC_TRACE
extern "C" int rayen_tc_trace_dump(void) {
  static long long h[4096];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, rayen::g_tc_trace, sizeof(h));
  const long long t0 = h[1000];
  fprintf(stderr, "panel: mma[pre_wfull wfull dempty issued] epi[pre_dfull dfull loaded] prod[issue]  (cycles since kernel start)\n");
  for (int p = 0; p < 20; ++p) {
    fprintf(stderr, "%2d:", p);
    for (int j = 0; j < 8; ++j) fprintf(stderr, " %7lld", h[8 * p + j] - t0);
    fprintf(stderr, "\n");
  }
  long long gmin = h[2400], gmax = h[2600];
  for (int b = 0; b < 128; ++b) { if (h[2400 + b] < gmin) gmin = h[2400 + b]; if (h[2600 + b] > gmax) gmax = h[2600 + b]; }
  fprintf(stderr, "per-CTA: start offset ns / duration cycles\n");
  for (int b = 0; b < 128; b += 8) fprintf(stderr, "cta %3d: start +%6lld ns, end +%6lld ns, cycles %7lld\n", b, h[2400 + b] - gmin, h[2600 + b] - gmin, h[2000 + b]);
  fprintf(stderr, "span %lld ns\n", gmax - gmin);
  return 0;
}
#endif

#ifdef RAYEN_BWD_TRACE
extern "C" int rayen_bwd_trace_read(long long* out) {
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(out, rayen::g_bwd_trace, sizeof(long long) * 8192));
}
#endif

#ifdef RAYEN_LW_TRACE
// development build only (scripts/lw_trace.py): the stamps of lmi_forward_warp_kernel (8192 values)
extern "C" int rayen_lw_trace_read(long long* out) {
  cudaDeviceSynchronize();
  return static_cast<int>(cudaMemcpyFromSymbol(out, rayen::g_lw_trace, sizeof(long long) * 8192));
}
#endif

#ifdef RAYEN_LMI_TRACE
// development build only: copies the phase stamps of lmi_forward_kernel (see LMI_STAMP) to `out` (4096 values)
extern "C" int rayen_lmi_trace_read(long long* out) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, rayen::g_lmi_trace, sizeof(long long) * 4096);
  return static_cast<int>(e);
}
#endif

extern "C" int rayen_plan_set_lmi_tensor_cores(rayen_plan_t* p, int mode) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (mode < 0 || mode > 2) return fail(RAYEN_ERR_BAD_ARGUMENT, "mode must be 0 (never), 1 (always) or 2 (automatic)");
  p->lmi_tc_mode = mode;
  return RAYEN_OK;
}

// Measured on B200 (scripts/lmi_dense_compare.py, scripts/time_kernels.py): the tensor-core
// contraction wins when the GEMM is deep (K = 32: -10 % at r = 32, -23 % at r = 16) and the launch carries no
// gradient work; at K <= 16 the FP32-pipe contraction is cheaper than the GEMM phase's fixed cost, and with
// want_grad the F~z matrices no longer fit in shared memory next to the GEMM buffers.
// A pruned work list is short (one partial pass per CTA): there the GEMM phase's fixed cost is not amortised either.
static bool lmi_use_tc(const rayen_plan* p, bool want_grad, bool list_mode) {
  if (!p->lmi_tc_ok || p->lmi_tc_mode == 0) return false;
  if (p->lmi_tc_mode == 1) return true;
  return p->dev.tc_kp >= 32 && !want_grad && !list_mode;
}

extern "C" int rayen_plan_set_lmi_filter(rayen_plan_t* p, int mode) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (mode < 0 || mode > 2) return fail(RAYEN_ERR_BAD_ARGUMENT, "mode must be 0 (never), 1 (always) or 2 (automatic)");
  p->lmi_warp_mode = mode;
  return RAYEN_OK;
}

// The filter needs a prior kappa to test against (a set with an LMI alone has none: every sample is solved, and there
// the 8-lanes-per-matrix layout of lmi.cuh has the higher throughput); automatic: padded LMI sizes 16 and 32.
static bool lmi_use_warp(const rayen_plan* p, bool has_prior) {
  if (!p->lmi_warp_ok || p->lmi_warp_mode == 0 || !has_prior) return false;
  if (p->lmi_warp_mode == 1) return true;
  return p->dev.lmi_rp >= 16;
}

extern "C" int rayen_plan_set_pruning(rayen_plan_t* p, int enabled) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  p->prune = enabled && p->has_lqs && p->dev.lmi_prune;
  return RAYEN_OK;
}

// workspace: [fwd counter, bwd counter, pad to 256 B][forward work list: B ints][backward work list: B ints]
//            [d kappa/du of the LMI-bound samples: B x n floats, written by forward when want_grad != 0]
static int64_t ws_list_bytes(int64_t B) { return (B * 4 + 255) / 256 * 256; }
static int64_t ws_prefix_bytes(int64_t B, int n) { return 256 + 2 * ws_list_bytes(B) + (B * n * 4 + 255) / 256 * 256; }
extern "C" int64_t rayen_workspace_bytes(const rayen_plan_t* p, int64_t B) {
  if (!p || B < 0) return -1;
  if (p->lmi_big) return ws_prefix_bytes(B, p->dev.n) + lmib_ws_extra(p, B, p->bdev.p4);
  if (p->dev.lmi_r == 0) return 0;
  return ws_prefix_bytes(B, p->dev.n);
}
static float* ws_dkappa(void* workspace, int64_t B) {
  return reinterpret_cast<float*>(static_cast<char*>(workspace) + 256 + 2 * ws_list_bytes(B));
}

extern "C" int rayen_plan_kernel_info(const rayen_plan_t* p, RayenKernelInfo* out) {
  if (!p || !out) return fail(RAYEN_ERR_BAD_ARGUMENT, "null argument");
  memset(out, 0, sizeof(*out));
  cudaFuncAttributes a;
  if (p->wide) {
    RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(wide_forward_kernel<16, false>)));
    out->regs_lqs_fwd = a.numRegs;
    RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(wide_backward_kernel<256>)));
    out->regs_lqs_bwd = a.numRegs;
    out->smem_lqs_bytes = static_cast<int>(p->wide_fwd_smem_bytes[1]);
    out->sm_count = p->sm_count;
    return RAYEN_OK;
  }
  RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(lqs_fwd_fn(p->dev.np, 4, p->lqs_smem))));
  out->regs_lqs_fwd = a.numRegs;
  RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(lqs_bwd_fn(p->dev.np))));
  out->regs_lqs_bwd = a.numRegs;
  if (p->dev.lmi_r > 0) {
    RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(lmi_fwd_fn(p->dev.lmi_rp, p->lmi_smem, p->lmi_fwd_threads, false))));
    out->regs_lmi_fwd = a.numRegs;
    RAYEN_CUDA(cudaFuncGetAttributes(&a, reinterpret_cast<const void*>(lmi_bwd_fn(p->dev.lmi_rp, p->lmi_smem))));
    out->regs_lmi_bwd = a.numRegs;
  }
  out->smem_lqs_bytes = static_cast<int>(p->lqs_smem_bytes);
  out->smem_lmi_bytes = static_cast<int>(p->lmi_smem_bytes);
  out->sm_count = p->sm_count;
  return RAYEN_OK;
}

// ----------------------------------------------------------------------------- launch geometry
static int floor_pow2(long long x) {
  int p = 1;
  while (2ll * p <= x) p *= 2;
  return p;
}
static int ceil_pow2(long long x) {
  int p = 1;
  while (p < x) p *= 2;
  return p;
}

struct LqsGeom {
  int tm, lanes, block, grid;
};

// Pick the samples-per-thread tile TM (register reuse of every LDS.128) and the lanes-per-sample split L
// (parallelism when the batch alone cannot fill 148 SMs).  Rule fitted to a sweep on B200
// (scripts/sweep_lqs.py): aim at ~8 warps per SM; prefer TM = 2 for 32-wide directions (128 registers,
// no spills) and TM = 4 below; never split a sample over more than 8 lanes.
static LqsGeom lqs_geometry(const rayen_plan* p, long long B) {
  const PlanDev& v = p->dev;
  int items = v.m_pad / 4;
  if (v.n_quad > items) items = v.n_quad;
  if (v.n_soc > items) items = v.n_soc;
  int max_lanes = ceil_pow2(items);
  if (max_lanes > 8) max_lanes = 8;
  const long long target = static_cast<long long>(p->sm_count) * 256;
  LqsGeom g{};
  int tm = p->tune_tm;
  int lanes = p->tune_lanes;
  if (tm == 0) {
    for (int cand = (v.np >= 32 ? 2 : 4); cand >= 1; cand /= 2) {
      const long long tiles = (B + cand - 1) / cand;
      long long l = floor_pow2(target / (tiles > 0 ? tiles : 1) > 0 ? target / (tiles > 0 ? tiles : 1) : 1);
      if (l > max_lanes) l = max_lanes;
      tm = cand;
      if (tiles * l * 2 >= target) break;
    }
  }
  const long long tiles = (B + tm - 1) / tm;
  if (lanes == 0) {
    long long l = target / (tiles > 0 ? tiles : 1);
    lanes = floor_pow2(l > 0 ? l : 1);
    if (lanes > max_lanes) lanes = max_lanes;
  }
  const long long cap = static_cast<long long>(p->sm_count) * lqs_max_threads(v.np, tm);
  const long long threads = tiles * lanes;
  int block = lqs_max_threads(v.np, tm);
  if (threads <= cap) {
    long long per_sm = (threads + p->sm_count - 1) / p->sm_count;
    per_sm = (per_sm + 31) / 32 * 32;
    if (per_sm < 64) per_sm = 64;
    if (per_sm < block) block = static_cast<int>(per_sm);
  }
  long long grid = (threads + block - 1) / block;
  if (grid > p->sm_count) grid = p->sm_count;
  if (grid < 1) grid = 1;
  g.tm = tm;
  g.lanes = lanes;
  g.block = block;
  g.grid = static_cast<int>(grid);
  return g;
}

static int check_io(const rayen_plan* p, const void* a, const void* b, long long B, int mode) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (B < 0) return fail(RAYEN_ERR_BAD_ARGUMENT, "negative batch");
  if (B > 0 && (!a || !b)) return fail(RAYEN_ERR_BAD_ARGUMENT, "null tensor");
  if (mode != RAYEN_MODE_RAYEN && mode != RAYEN_MODE_RAYEN_OLD) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad mode %d", mode);
  return 0;
}

// ----------------------------------------------------------------------------- forward / backward
static int forward_impl(const rayen_plan_t* p, const float* v, int64_t ldv, float* y, float* kappa, int32_t* active,
                        int64_t B, int mode, int want_grad, void* workspace, void* stream_, int stage_mask,
                        const MapArgs* map = nullptr);

extern "C" int rayen_forward_f32(const rayen_plan_t* p, const float* v, int64_t ldv, float* y, float* kappa,
                                 int32_t* active, int64_t B, int mode, int want_grad, void* workspace, void* stream_) {
  return forward_impl(p, v, ldv, y, kappa, active, B, mode, want_grad, workspace, stream_, 3);
}

extern "C" int rayen_forward_mapped_f32(const rayen_plan_t* p, const float* x, int64_t ldx, int32_t in_dim,
                                        const float* weight, int64_t ldw, const float* bias, float* v_out, float* y,
                                        float* kappa, int32_t* active, int64_t B, int want_grad, void* workspace,
                                        void* stream_) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (B < 0 || in_dim < 1) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad batch or input dimension");
  if (B > 0 && (!x || !weight || !v_out)) return fail(RAYEN_ERR_BAD_ARGUMENT, "null tensor");
  if (!(p->has_lqs || p->dev.lmi_r == 0))
    return fail(RAYEN_ERR_UNSUPPORTED, "the fused mapper lives in the linear/quadratic/SOC kernel; this plan has only an LMI");
  if ((in_dim & 3) || (ldx & 3) || (ldw & 3) || ldx < in_dim || ldw < in_dim ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15))
    return fail(RAYEN_ERR_UNSUPPORTED, "the fused mapper needs input_dim and the row strides to be multiples of 4 "
                                       "floats and 16-byte aligned tensors");
  MapArgs m{};
  m.x = x; m.w = weight; m.bias = bias; m.v_out = v_out; m.ldx = ldx; m.ldw = ldw; m.in_dim = in_dim;
  return forward_impl(p, v_out, p->dev.n, y, kappa, active, B, RAYEN_MODE_RAYEN, want_grad, workspace, stream_, 3, &m);
}
extern "C" int rayen_forward_stage_f32(const rayen_plan_t* p, const float* v, int64_t ldv, float* y, float* kappa,
                                       int32_t* active, int64_t B, int mode, int want_grad, int stage_mask,
                                       void* workspace, void* stream_) {
  if (stage_mask < 1 || stage_mask > 3) return fail(RAYEN_ERR_BAD_ARGUMENT, "stage_mask must be 1, 2 or 3");
  return forward_impl(p, v, ldv, y, kappa, active, B, mode, want_grad, workspace, stream_, stage_mask);
}
extern "C" int rayen_backward_stage_f32(const rayen_plan_t* p, const float* v, int64_t ldv, const float* gy,
                                        const float* kappa, const int32_t* active, float* gv, int64_t ldgv, int64_t B,
                                        int mode, int have_dkappa, int stage_mask, void* workspace, void* stream_);

static int forward_impl(const rayen_plan_t* p, const float* v, int64_t ldv, float* y, float* kappa, int32_t* active,
                        int64_t B, int mode, int want_grad, void* workspace, void* stream_, int stage_mask,
                        const MapArgs* map) {
  const MapArgs margs = map ? *map : MapArgs{};
  int rc = check_io(p, v, y, B, mode);
  if (rc) return rc;
  if (B == 0) return RAYEN_OK;
  const PlanDev& d = p->dev;
  const int need = d.n + (mode == RAYEN_MODE_RAYEN_OLD ? 1 : 0);
  if (ldv < need) return fail(RAYEN_ERR_BAD_ARGUMENT, "ldv=%lld < %d", static_cast<long long>(ldv), need);
  const bool has_lmi = d.lmi_r > 0;
  if (has_lmi && (!kappa || !active))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "plans with an LMI need the kappa and active outputs");
  if (has_lmi && !workspace) return fail(RAYEN_ERR_BAD_ARGUMENT, "plans with an LMI need the workspace buffer");
  if (p->lmi_big && (!kappa || !active || !workspace))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "plans with an LMI need the kappa and active outputs and the workspace buffer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  if (prev != p->device) RAYEN_CUDA(cudaSetDevice(p->device));
  // big LMI (lmi_big.cuh): behind the other families' kernel, which leaves the prior (kappa, tag, y) of every sample
  auto run_lmi_big = [&]() -> cudaError_t {
    return lmib_run(p, p->d_blob + p->bdev.off_lmib, d.n, v, ldv, y, kappa, active,
                    want_grad ? ws_dkappa(workspace, B) : nullptr, static_cast<char*>(workspace) + ws_prefix_bytes(B, d.n), B,
                    mode, want_grad ? kLbFlagGrad : 0, stream, nullptr,
                    reinterpret_cast<int*>(static_cast<char*>(workspace) + 256), static_cast<int*>(workspace) + 8);
  };

  if (p->wide) {
    if (map) {
      if (prev != p->device) cudaSetDevice(prev);
      return fail(RAYEN_ERR_UNSUPPORTED, "the fused mapper is not available for wide plans (n > 32)");
    }
    cudaError_t we = cudaSuccess;
    if (stage_mask & 1) {
      // tiles of 16 samples halve the L2 traffic of the row matrix; tiles of 8 keep every SM busy on short batches
      static const int force_ts = getenv("RAYEN_WIDE_TS") ? atoi(getenv("RAYEN_WIDE_TS")) : 0;
      bool ts16 = (B + 15) / 16 >= p->sm_count && p->wide_fwd_smem_bytes[1] <= static_cast<size_t>(p->max_smem_optin) / 2;
      if (force_ts == 8) ts16 = false;
      if (force_ts == 16 && p->wide_fwd_smem_bytes[1] <= static_cast<size_t>(p->max_smem_optin)) ts16 = true;
      // sets too wide for a tile of 8 directions in shared memory (n > ~6900) take tiles of 4
      const bool ts4 = p->wide_fwd_smem_bytes[0] > static_cast<size_t>(p->max_smem_optin) || force_ts == 4;
      if (ts4) ts16 = false;
      const int ts = ts4 ? 4 : (ts16 ? 16 : 8);
      long long grid = (B + ts - 1) / ts;
      const long long cap = static_cast<long long>(p->sm_count) * 32;
      if (grid > cap) grid = cap;
      const bool blk = p->wdev.n >= kWideBlockedN;  // blocked accumulation of the long dot products (wide.cuh)
      const size_t wsm = p->wide_fwd_smem_bytes[ts4 ? 2 : (ts16 ? 1 : 0)];
      auto wf = ts4 ? wide_forward_kernel<4, true>
                    : (ts16 ? (blk ? wide_forward_kernel<16, true> : wide_forward_kernel<16, false>)
                            : (blk ? wide_forward_kernel<8, true> : wide_forward_kernel<8, false>));
      if (ts4 && !blk) {  // (only reachable through RAYEN_WIDE_TS=4 on a narrow set: there is no unblocked 4-tile build)
        if (prev != p->device) cudaSetDevice(prev);
        return fail(RAYEN_ERR_UNSUPPORTED, "RAYEN_WIDE_TS=4 needs n >= %d", kWideBlockedN);
      }
      wf<<<static_cast<int>(grid), kWideThreads, wsm, stream>>>(p->wdev, v, ldv, y, kappa, active, B, mode);
      g_launches.fetch_add(1);
      we = cudaGetLastError();
    }
    if (we == cudaSuccess && p->lmi_big && (stage_mask & 2)) we = run_lmi_big();
    if (prev != p->device) cudaSetDevice(prev);
    if (we != cudaSuccess) return cuda_fail(we, "forward launch (wide)");
    return RAYEN_OK;
  }
  int* counters = static_cast<int*>(workspace);
  int* fwd_list = has_lmi ? reinterpret_cast<int*>(static_cast<char*>(workspace) + 256) : nullptr;
  const bool run_lqs = p->has_lqs || !has_lmi;
  const bool use_list = has_lmi && run_lqs && p->prune;
  cudaError_t e = cudaSuccess;
  const bool use_warp = has_lmi && lmi_use_warp(p, run_lqs);
  // counters: [0] forward work list, [1] backward work list, [2] fail list and [3] chunk dispenser of the filter kernel
  if ((stage_mask & 1) && run_lqs) {
    if (use_list || use_warp) e = cudaMemsetAsync(counters, 0, 4 * sizeof(int), stream);
    if (e == cudaSuccess && p->use_tc) {
      long long grid = (B + 255) / 256;
      if (grid > p->sm_count) grid = p->sm_count;
      LqsTcFn f = lqs_tc_fn(d.tc_kp);
      f<<<static_cast<int>(grid), kTcThreads, p->tc_smem_bytes, stream>>>(d, v, ldv, y, kappa, active, B, mode,
                                                                           has_lmi ? 1 : 0, use_list ? 1 : 0,
                                                                           use_list ? fwd_list : nullptr,
                                                                           use_list ? counters : nullptr, margs);
      g_launches.fetch_add(1);
      e = cudaGetLastError();
    } else if (e == cudaSuccess) {
      const LqsGeom g = lqs_geometry(p, B);
      LqsFwdFn f = lqs_fwd_fn(d.np, g.tm, p->lqs_smem);
      f<<<g.grid, g.block, p->lqs_smem_bytes, stream>>>(d, v, ldv, y, kappa, active, B, mode, g.lanes, has_lmi ? 1 : 0,
                                                        use_list ? 1 : 0, use_list ? fwd_list : nullptr,
                                                        use_list ? counters : nullptr, margs);
      g_launches.fetch_add(1);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess && has_lmi && (stage_mask & 2)) {
    const int mpw = 32 / (d.lmi_rp / 4);
    const bool grad = want_grad != 0;
    // the kernel right before this one in the stream is our own linear/quadratic/SOC kernel: launch behind it
    static const bool pdl_on = !(getenv("RAYEN_PDL") && atoi(getenv("RAYEN_PDL")) == 0);
    const bool behind_lqs = pdl_on && (stage_mask & 1) && run_lqs;
    if (use_warp) {
      // The filter kernel settles the samples whose LMI provably cannot bind and solves the others itself (one warp per
      // matrix).  With a cap on the solves per warp (RAYEN_LMI_WARP_SOLVES) the rest goes to a fail list (the region of
      // the backward work list, free until the backward call) and a second launch of the same kernel.
      int* fail_list = reinterpret_cast<int*>(static_cast<char*>(workspace) + 256 + ws_list_bytes(B));
      // counters [2] fail list, [3] chunk dispenser of the filter kernel: zeroed with the others in front of the
      // linear/quadratic/SOC kernel, or here when this stage is launched on its own
      if (!((stage_mask & 1) && run_lqs)) e = cudaMemsetAsync(counters + 2, 0, 2 * sizeof(int), stream);
      // chunks of 4 samples, dealt round-robin to the CTAs (the list length is only known on the device)
      long long blocks = (B + kLwMT - 1) / kLwMT;
      if (blocks > p->sm_count) blocks = p->sm_count;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(blocks));
      cfg.blockDim = dim3(kLwThreads);
      cfg.dynamicSmemBytes = p->lmi_warp_smem_bytes;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = behind_lqs ? 1 : 0;
      const int* list = use_list ? fwd_list : nullptr;
      const int* cnt = use_list ? counters : nullptr;
      float* dk = grad ? ws_dkappa(workspace, B) : nullptr;
      const long long ldv_ = ldv, B_ = B;
      const int filt = p->lmi_warp_filter ? 1 : 0;
      const int budget = p->lmi_warp_solves;
      int* fcnt = counters + 2;
      int* next = (budget == 0x7fffffff) ? counters + 3 : nullptr;   // dynamic chunks only when there is no second launch
      if (e == cudaSuccess)
        e = grad ? cudaLaunchKernelEx(&cfg, lmi_forward_warp_kernel<true>, d, v, ldv_, y, kappa, active, B_, mode, list, cnt, dk,
                                      filt, budget, fail_list, fcnt, next)
                 : cudaLaunchKernelEx(&cfg, lmi_forward_warp_kernel<false>, d, v, ldv_, y, kappa, active, B_, mode, list, cnt,
                                      dk, filt, budget, fail_list, fcnt, next);
      if (e == cudaSuccess && budget != 0x7fffffff) {
        g_launches.fetch_add(1);
        // the same kernel once more on the fail list, filter off, no budget: whoever solves a sample runs the same
        // arithmetic, so results do not depend on how the batch was cut into chunks (bit-identical across pruning
        // on / off, host-buffer chunking, list order)
        cfg.numAttrs = pdl_on ? 1 : 0;
        const int* flist = fail_list;
        const int* fc = fcnt;
        int* none = nullptr;
        const int all = 0x7fffffff, nofilt = 0;
        e = grad ? cudaLaunchKernelEx(&cfg, lmi_forward_warp_kernel<true>, d, v, ldv_, y, kappa, active, B_, mode, flist, fc, dk,
                                      nofilt, all, none, none, none)
                 : cudaLaunchKernelEx(&cfg, lmi_forward_warp_kernel<false>, d, v, ldv_, y, kappa, active, B_, mode, flist, fc,
                                      dk, nofilt, all, none, none, none);
      }
    } else if (lmi_use_tc(p, grad, use_list)) {
      // one CTA per SM, persistent over passes of 8 warps x mpw samples; short batches still start one CTA per
      // warp's worth of samples so that the kernel can spread them (it re-derives the split from the list length)
      long long blocks = (B + mpw - 1) / mpw;
      if (blocks > p->sm_count) blocks = p->sm_count;
      LmiFwdFn lf = lmi_fwd_tc_fn(d.lmi_rp, grad && p->lmi_tc_grad_fsmem, grad);
      e = launch_lmi_forward(lf, static_cast<int>(blocks), kLmiTcThreads,
                             grad ? p->lmi_tc_grad_smem_bytes : p->lmi_tc_smem_bytes, stream, behind_lqs, d, v, ldv, y,
                             kappa, active, B, mode, run_lqs ? 1 : 0, use_list ? fwd_list : nullptr,
                             use_list ? counters : nullptr, grad ? ws_dkappa(workspace, B) : nullptr);
    } else {
      const int threads = grad ? 256 : p->lmi_fwd_threads;
      long long blocks = (B + static_cast<long long>(mpw) * (threads / 32) - 1) / (static_cast<long long>(mpw) * (threads / 32));
      if (blocks > p->sm_count) blocks = p->sm_count;
      LmiFwdFn lf = lmi_fwd_fn(d.lmi_rp, p->lmi_smem, threads, grad);
      e = launch_lmi_forward(lf, static_cast<int>(blocks), threads, grad ? p->lmi_grad_smem_bytes : p->lmi_smem_bytes,
                             stream, behind_lqs, d, v, ldv, y, kappa, active, B, mode, run_lqs ? 1 : 0,
                             use_list ? fwd_list : nullptr, use_list ? counters : nullptr,
                             grad ? ws_dkappa(workspace, B) : nullptr);
    }
    g_launches.fetch_add(1);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e == cudaSuccess && p->lmi_big && (stage_mask & 2)) e = run_lmi_big();
  if (prev != p->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return cuda_fail(e, "forward launch");
  return RAYEN_OK;
}

extern "C" int rayen_backward_f32(const rayen_plan_t* p, const float* v, int64_t ldv, const float* gy,
                                  const float* kappa, const int32_t* active, float* gv, int64_t ldgv, int64_t B,
                                  int mode, int have_dkappa, void* workspace, void* stream_) {
  return rayen_backward_stage_f32(p, v, ldv, gy, kappa, active, gv, ldgv, B, mode, have_dkappa, 3, workspace, stream_);
}

extern "C" int rayen_backward_stage_f32(const rayen_plan_t* p, const float* v, int64_t ldv, const float* gy,
                                        const float* kappa, const int32_t* active, float* gv, int64_t ldgv, int64_t B,
                                        int mode, int have_dkappa, int stage_mask, void* workspace, void* stream_) {
  if (stage_mask < 1 || stage_mask > 3) return fail(RAYEN_ERR_BAD_ARGUMENT, "stage_mask must be 1, 2 or 3");
  int rc = check_io(p, v, gy, B, mode);
  if (rc) return rc;
  if (B == 0) return RAYEN_OK;
  if (!kappa || !active || !gv) return fail(RAYEN_ERR_BAD_ARGUMENT, "backward needs kappa, active and gv");
  const PlanDev& d = p->dev;
  const int need = d.n + (mode == RAYEN_MODE_RAYEN_OLD ? 1 : 0);
  if (ldv < need || ldgv < need)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "ldv=%lld / ldgv=%lld < %d", static_cast<long long>(ldv),
                static_cast<long long>(ldgv), need);
  if ((d.lmi_r > 0 || p->lmi_big) && !workspace)
    return fail(RAYEN_ERR_BAD_ARGUMENT, "plans with an LMI need the workspace buffer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  if (prev != p->device) RAYEN_CUDA(cudaSetDevice(p->device));
  // big LMI: d kappa/du of the LMI-bound samples comes from the workspace -- left there by forward (have_dkappa) or
  // recomputed now by the same two kernels in gradient-only mode (kappa / active / y are inputs only)
  const float* big_dk = p->lmi_big ? ws_dkappa(workspace, B) : nullptr;
  if (p->lmi_big && !have_dkappa && (stage_mask & 1)) {
    cudaError_t ge = lmib_run(p, p->d_blob + p->bdev.off_lmib, d.n, v, ldv, nullptr, const_cast<float*>(kappa),
                              const_cast<int32_t*>(active), ws_dkappa(workspace, B),
                              static_cast<char*>(workspace) + ws_prefix_bytes(B, d.n), B, mode, kLbFlagGradOnly, stream, nullptr,
                              reinterpret_cast<int*>(static_cast<char*>(workspace) + 256), static_cast<int*>(workspace) + 8);
    if (ge != cudaSuccess) {
      if (prev != p->device) cudaSetDevice(prev);
      return cuda_fail(ge, "backward launch (big LMI gradient)");
    }
  }

  if (p->wide) {
    cudaError_t we = cudaSuccess;
    if (stage_mask & 1) {
      long long wgrid = B;
      const long long wcap = static_cast<long long>(p->sm_count) * 64;
      if (wgrid > wcap) wgrid = wcap;
      if (p->wdev.n < kWideBwdSwitchN)
        wide_backward_kernel<128><<<static_cast<int>(wgrid), 128, p->wide_bwd_smem_bytes, stream>>>(
            p->wdev, v, ldv, gy, kappa, active, gv, ldgv, B, mode, big_dk);
      else
        wide_backward_kernel<256><<<static_cast<int>(wgrid), 256, p->wide_bwd_smem_bytes, stream>>>(
            p->wdev, v, ldv, gy, kappa, active, gv, ldgv, B, mode, big_dk);
      g_launches.fetch_add(1);
      we = cudaGetLastError();
    }
    if (prev != p->device) cudaSetDevice(prev);
    if (we != cudaSuccess) return cuda_fail(we, "backward launch (wide)");
    return RAYEN_OK;
  }
  const int block = kBwdThreads;
  long long grid = (B + block - 1) / block;
  const long long cap = static_cast<long long>(p->sm_count) * 12;  // 12 blocks x 18 KB of tile memory per SM at n = 32
  if (grid > cap) grid = cap;
  const bool has_lmi = d.lmi_r > 0;
  int* counters = static_cast<int*>(workspace);
  int* bwd_list = has_lmi ? reinterpret_cast<int*>(static_cast<char*>(workspace) + 256 + ws_list_bytes(B)) : nullptr;
  cudaError_t e = cudaSuccess;
  if (stage_mask & 1) {
    // the backward work list exists only when the LMI gradients were not left behind by the forward call
    if (has_lmi && !have_dkappa) e = cudaMemsetAsync(counters + 1, 0, sizeof(int), stream);
    LqsBwdFn f = lqs_bwd_fn(d.np);
    if (e == cudaSuccess) {
      const size_t tile_bytes = static_cast<size_t>(block / 32) * 2 * 32 * (d.np + 4) * sizeof(float);
      f<<<static_cast<int>(grid), block, tile_bytes, stream>>>(d, v, ldv, gy, kappa, active, gv, ldgv, B, mode, bwd_list,
                                                      has_lmi ? counters + 1 : nullptr,
                                                      p->lmi_big ? big_dk
                                                                 : ((has_lmi && have_dkappa) ? ws_dkappa(workspace, B) : nullptr));
      g_launches.fetch_add(1);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess && has_lmi && !have_dkappa && (stage_mask & 2)) {
    const int mpw = 32 / (d.lmi_rp / 4);
    long long blocks = (B + static_cast<long long>(mpw) * (kLmiThreads / 32) - 1) / (static_cast<long long>(mpw) * (kLmiThreads / 32));
    if (blocks > p->sm_count) blocks = p->sm_count;
    LmiBwdFn lf = lmi_bwd_fn(d.lmi_rp, p->lmi_smem);
    lf<<<static_cast<int>(blocks), kLmiThreads, p->lmi_bwd_smem_bytes, stream>>>(d, v, ldv, gy, kappa, active, gv, ldgv,
                                                                                B, mode, bwd_list, counters + 1);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (prev != p->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return cuda_fail(e, "backward launch");
  return RAYEN_OK;
}

// ----------------------------------------------------------------------------- violation metric
extern "C" int rayen_violation_f32(const rayen_plan_t* p, const float* y, int64_t ldy, float* viol, int64_t B,
                                   void* stream_) {
  if (!p) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (B < 0 || (B > 0 && (!y || !viol))) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad arguments");
  if (B == 0) return RAYEN_OK;
  const PlanDev& d = p->dev;
  if (ldy < d.k) return fail(RAYEN_ERR_BAD_ARGUMENT, "ldy=%lld < k=%d", static_cast<long long>(ldy), d.k);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  if (prev != p->device) RAYEN_CUDA(cudaSetDevice(p->device));
  // one y row per warp in shared memory: as many warps per block (<= 8) as fit
  const size_t row_bytes = static_cast<size_t>((d.k + 3) / 4 * 4) * sizeof(float);
  int warps = kViolThreads / 32;
  while (warps > 1 && warps * row_bytes > static_cast<size_t>(p->max_smem_optin)) --warps;
  long long grid = (B + warps - 1) / warps;
  if (grid > static_cast<long long>(p->sm_count) * 8) grid = static_cast<long long>(p->sm_count) * 8;
  const size_t smem = warps * row_bytes;
  if (smem > static_cast<size_t>(p->max_smem_optin)) {
    if (prev != p->device) cudaSetDevice(prev);
    return fail(RAYEN_ERR_UNSUPPORTED, "violation metric: k=%d needs %zu bytes of shared memory", d.k, smem);
  }
  viol_lqs_kernel<<<static_cast<int>(grid), warps * 32, smem, stream>>>(d, y, ldy, viol, B);
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && d.lmi_r > 0) {
    const int mpw = 32 / (d.lmi_rp / 4);
    long long blocks = (B + static_cast<long long>(mpw) * (kLmiThreads / 32) - 1) / (static_cast<long long>(mpw) * (kLmiThreads / 32));
    if (blocks > p->sm_count) blocks = p->sm_count;
    ViolLmiFn f = viol_lmi_fn(d.lmi_rp, p->viol_lmi_smem);
    f<<<static_cast<int>(blocks), kLmiThreads, p->viol_lmi_smem_bytes, stream>>>(d, y, ldy, viol, B);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && p->lmi_big) {
    // lambda_max(-F(y)) = -lambda_min(F(y)): the two kernels of lmi_big.cuh on the ambient matrices (LMINEGB: rows
    // 0 .. k-1 contract with y, row k is the constant term); the buffer comes from the stream-ordered allocator
    // (a diagnostic call, not the hot path)
    const int64_t bytes = lmib_ws_extra(p, B, p->bdev.p4);
    void* buf = nullptr;
    e = cudaMallocAsync(&buf, static_cast<size_t>(bytes), stream);
    if (e == cudaSuccess) {
      const float* Fn = p->d_blob + p->off_lminegb;
      e = lmib_run(p, Fn, d.k, y, ldy, nullptr, viol, nullptr, nullptr, buf, B, RAYEN_MODE_RAYEN, kLbFlagLambdaOut, stream,
                   Fn + static_cast<size_t>(d.k) * p->bdev.p4);
      cudaError_t fe = cudaFreeAsync(buf, stream);
      if (e == cudaSuccess) e = fe;
    }
  }
  if (prev != p->device) cudaSetDevice(prev);
  if (e != cudaSuccess) return cuda_fail(e, "violation launch");
  return RAYEN_OK;
}

// ----------------------------------------------------------------------------- all-gather of y over peer memory
struct GatherDst {
  float* base[16];
};
// words: floats to move; each destination receives them at `word_offset`.  VEC: 16-byte accesses (everything aligned).
template <bool VEC, bool MULTICAST>
__global__ void __launch_bounds__(256) gather_push_kernel(const float* __restrict__ src, long long words, GatherDst dst,
                                                          int n_dst, long long word_offset) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  if constexpr (VEC) {
    const long long n4 = words >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src) + i);
      if constexpr (MULTICAST) {
        float* p = dst.base[0] + word_offset + 4 * i;
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x.x), "f"(x.y), "f"(x.z),
                     "f"(x.w)
                     : "memory");
      } else {
        for (int r = 0; r < n_dst; ++r) *(reinterpret_cast<float4*>(dst.base[r] + word_offset) + i) = x;
      }
    }
  } else {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < words; i += stride) {
      const float x = __ldg(src + i);
      if constexpr (MULTICAST) {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst.base[0] + word_offset + i), "f"(x) : "memory");
      } else {
        for (int r = 0; r < n_dst; ++r) dst.base[r][word_offset + i] = x;
      }
    }
  }
}

extern "C" int rayen_gather_push_f32(const float* src, int64_t rows, int32_t k, float* const* dst_bases, int32_t n_dst,
                                     float* multicast_base, int64_t row_offset, void* stream_) {
  if (rows < 0 || k < 1 || row_offset < 0) return fail(RAYEN_ERR_BAD_ARGUMENT, "bad gather shape");
  if (rows == 0) return RAYEN_OK;
  if (!src || (!multicast_base && (!dst_bases || n_dst < 1 || n_dst > 16)))
    return fail(RAYEN_ERR_BAD_ARGUMENT, "gather: null source, or neither a multicast mapping nor 1..16 peer buffers");
  GatherDst d{};
  if (multicast_base) {
    d.base[0] = multicast_base;
  } else {
    for (int r = 0; r < n_dst; ++r) {
      if (!dst_bases[r]) return fail(RAYEN_ERR_BAD_ARGUMENT, "gather: peer buffer %d is null", r);
      d.base[r] = dst_bases[r];
    }
  }
  const long long words = rows * static_cast<long long>(k), off = row_offset * static_cast<long long>(k);
  bool vec = (words % 4 == 0) && (off % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
  for (int r = 0; r < (multicast_base ? 1 : n_dst) && vec; ++r) vec = reinterpret_cast<uintptr_t>(d.base[r]) % 16 == 0;
  const long long items = vec ? words / 4 : words;
  long long grid = (items + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (multicast_base) {
    if (vec) gather_push_kernel<true, true><<<static_cast<int>(grid), 256, 0, stream>>>(src, words, d, 1, off);
    else gather_push_kernel<false, true><<<static_cast<int>(grid), 256, 0, stream>>>(src, words, d, 1, off);
  } else {
    if (vec) gather_push_kernel<true, false><<<static_cast<int>(grid), 256, 0, stream>>>(src, words, d, n_dst, off);
    else gather_push_kernel<false, false><<<static_cast<int>(grid), 256, 0, stream>>>(src, words, d, n_dst, off);
  }
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gather launch");
  return RAYEN_OK;
}

// ----------------------------------------------------------------------------- host-buffer path
// The batch is cut into C chunks and the three engines of the GPU work at the same time:
//   copy-in stream   H2D v_0 .. v_{C-1}, then gy_0 .. gy_{C-1}            (event after each)
//   caller's stream  forward_0 .. forward_{C-1}, then backward_0 .. backward_{C-1}   (each waits for its input)
//   copy-out stream  D2H y_c as soon as forward_c is done, then gv_c after backward_c
// PCIe is full duplex, so the step costs about max(bytes in, bytes out) / link rate plus one chunk of compute
// instead of the sum of all three.
constexpr int kHostMaxChunks = 8;
static int64_t round256(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t rayen_host_workspace_bytes(const rayen_plan_t* p, int64_t B) {
  if (!p || B < 0) return -1;
  const int64_t n = p->dev.n, k = p->dev.k;
  // v | gy | y | gv | kappa | active, each rounded up to 256 B, then one kernel workspace per chunk
  return round256(B * n * 4) * 2 + round256(B * k * 4) * 2 + round256(B * 4) * 2 + rayen_workspace_bytes(p, B) +
         kHostMaxChunks * ((p->dev.lmi_r > 0 || p->lmi_big) ? 2048 : 0);
}

static int host_chunk_count(const rayen_plan* p, int64_t B) {
  int c = p->host_chunks;
  if (c == 0) {
    // measured on B200 (scripts/e2e_sweep.py, RAYEN_HOST_TRACE=1): the copies of v / gy / y / gv already overlap
    // each other and the kernels with ONE chunk (gy goes in under the forward kernels, y comes out under the
    // backward kernel).  A second chunk pays only for sets without an LMI once each direction moves >= 8 MB; the
    // eigen-solver's latency-bound kernels take as long on half a batch as on a whole one.
    // Round 2 (the LMI pass of a mixed set now takes ~33 us instead of ~60): two chunks also pay for sets with an LMI
    // from 8 MB per direction on -- the copy-out of y_0 starts while chunk 1 is still in its forward kernels
    // (cfg5 shard, 54 GB/s link: 0.397 -> 0.363 ms per step; four chunks: 0.444, the LMI pass has a fixed latency).
    const int64_t bytes = B * (p->dev.n + p->dev.k) * 4;
    c = (bytes >= ((p->dev.lmi_r == 0) ? (16ll << 20) : (8ll << 20))) ? 2 : 1;
  }
  if (p->lmi_big) c = 1;  // the contraction buffer of lmi_big.cuh is sized for ONE call over the batch
  if (c > kHostMaxChunks) c = kHostMaxChunks;
  if (c > B) c = static_cast<int>(B);
  if (c < 1) c = 1;
  return c;
}

constexpr int kHostSlots = 4;

// Enqueues one forward+backward step on host buffers: copy-in on p->host_in, kernels on `main`, copy-out on
// p->host_out (see the header), and finally makes `main` wait for the last copy-out.  Nothing here synchronises, so
// the whole thing can be stream-captured with `main` as the origin stream.
static int host_enqueue(rayen_plan* p, const float* v_host, const float* gy_host, float* y_host, float* gv_host, int64_t B,
                        void* workspace, cudaStream_t main, int slot, bool order_after_main, bool join, bool trace) {
  const int64_t n = p->dev.n, k = p->dev.k;
  char* w = static_cast<char*>(workspace);
  float* v = reinterpret_cast<float*>(w); w += round256(B * n * 4);
  float* gy = reinterpret_cast<float*>(w); w += round256(B * k * 4);
  float* y = reinterpret_cast<float*>(w); w += round256(B * k * 4);
  float* gv = reinterpret_cast<float*>(w); w += round256(B * n * 4);
  float* kappa = reinterpret_cast<float*>(w); w += round256(B * 4);
  int32_t* active = reinterpret_cast<int32_t*>(w); w += round256(B * 4);
  cudaError_t e = cudaSuccess;
  cudaEvent_t tev[40];
  const char* tname[40];
  int ntev = 0;
  auto mark = [&](cudaStream_t st, const char* name) {
    if (!trace || ntev >= 40) return;
    cudaEventCreate(&tev[ntev]);
    cudaEventRecord(tev[ntev], st);
    tname[ntev++] = name;
  };
  const int C = host_chunk_count(p, B);
  const int64_t per = ((B + C - 1) / C + 255) / 256 * 256;  // chunk size, multiple of the TC kernel's super-tile
  cudaEvent_t* ev_v = p->host_ev[slot];                    // [c]     v_c is on the device
  cudaEvent_t* ev_g = p->host_ev[slot] + kHostMaxChunks;   // [c]     gy_c is on the device, later reused: backward_c done
  cudaEvent_t ev_start = p->host_ev[slot][2 * kHostMaxChunks], ev_done = p->host_ev[slot][2 * kHostMaxChunks + 1];
  int rc = 0;
  mark(main, "start");
  if (order_after_main) {
    // everything is ordered after what is queued on `main` (e.g. an earlier use of the workspace)
    e = cudaEventRecord(ev_start, main);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->host_in, ev_start, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->host_out, ev_start, 0);
  }
  // submitted steps (order_after_main == false): the only earlier user of this slot's workspace and host buffers is the
  // previous step of the same slot, which the caller has waited for (the contract of
  // rayen_forward_backward_host_submit_f32).  The copy-in therefore starts at once -- under the kernels of the step
  // before -- instead of behind everything queued on `main`; the kernels themselves are ordered by `main`, the copy-out
  // by the events below.
  void* ws_c[kHostMaxChunks];
  int64_t lo[kHostMaxChunks], cnt[kHostMaxChunks];
  int used = 0;
  for (int c = 0; c < C; ++c) {
    const int64_t b0 = c * per;
    if (b0 >= B) break;
    lo[used] = b0;
    cnt[used] = (B - b0 < per) ? B - b0 : per;
    const int64_t wb = rayen_workspace_bytes(p, cnt[used]);
    ws_c[used] = wb > 0 ? static_cast<void*>(w) : nullptr;
    w += wb;
    ++used;
  }
  for (int c = 0; c < used && e == cudaSuccess; ++c) {
    e = cudaMemcpyAsync(v + lo[c] * n, v_host + lo[c] * n, cnt[c] * n * 4, cudaMemcpyHostToDevice, p->host_in);
    if (e == cudaSuccess) e = cudaEventRecord(ev_v[c], p->host_in);
    mark(p->host_in, "h2d v");
  }
  for (int c = 0; c < used && e == cudaSuccess; ++c) {
    e = cudaMemcpyAsync(gy + lo[c] * k, gy_host + lo[c] * k, cnt[c] * k * 4, cudaMemcpyHostToDevice, p->host_in);
    if (e == cudaSuccess) e = cudaEventRecord(ev_g[c], p->host_in);
    mark(p->host_in, "h2d gy");
  }
  for (int c = 0; c < used && e == cudaSuccess && rc == 0; ++c) {
    e = cudaStreamWaitEvent(main, ev_v[c], 0);
    if (e == cudaSuccess)
      rc = rayen_forward_f32(p, v + lo[c] * n, n, y + lo[c] * k, kappa + lo[c], active + lo[c], cnt[c],
                             RAYEN_MODE_RAYEN, 1, ws_c[c], main);
    mark(main, "forward");
    if (e == cudaSuccess && rc == 0) e = cudaEventRecord(ev_v[c], main);  // reused: forward_c done
    if (e == cudaSuccess && rc == 0) e = cudaStreamWaitEvent(p->host_out, ev_v[c], 0);
    if (e == cudaSuccess && rc == 0)
      e = cudaMemcpyAsync(y_host + lo[c] * k, y + lo[c] * k, cnt[c] * k * 4, cudaMemcpyDeviceToHost, p->host_out);
    mark(p->host_out, "d2h y");
  }
  for (int c = 0; c < used && e == cudaSuccess && rc == 0; ++c) {
    e = cudaStreamWaitEvent(main, ev_g[c], 0);
    if (e == cudaSuccess)
      rc = rayen_backward_f32(p, v + lo[c] * n, n, gy + lo[c] * k, kappa + lo[c], active + lo[c], gv + lo[c] * n, n,
                              cnt[c], RAYEN_MODE_RAYEN, 1, ws_c[c], main);
    mark(main, "backward");
    if (e == cudaSuccess && rc == 0) e = cudaEventRecord(ev_g[c], main);  // reused: backward_c done
    if (e == cudaSuccess && rc == 0) e = cudaStreamWaitEvent(p->host_out, ev_g[c], 0);
    if (e == cudaSuccess && rc == 0)
      e = cudaMemcpyAsync(gv_host + lo[c] * n, gv + lo[c] * n, cnt[c] * n * 4, cudaMemcpyDeviceToHost, p->host_out);
    mark(p->host_out, "d2h gv");
  }
  // completion of this step = its last copy-out
  cudaError_t e2 = cudaEventRecord(ev_done, p->host_out);
  if (e == cudaSuccess) e = e2;
  if (join) {
    e2 = cudaStreamWaitEvent(main, ev_done, 0);
    if (e == cudaSuccess) e = e2;
  }
  if (trace && ntev > 0) {
    cudaStreamSynchronize(p->host_in);
    cudaStreamSynchronize(p->host_out);
    cudaStreamSynchronize(main);
    fprintf(stderr, "rayen host path, %d chunk(s), B = %lld: end of each piece, us after the start\n", used, static_cast<long long>(B));
    for (int i = 1; i < ntev; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tev[0], tev[i]);
      fprintf(stderr, "  %-9s %8.1f\n", tname[i], ms * 1e3f);
    }
    for (int i = 0; i < ntev; ++i) cudaEventDestroy(tev[i]);
  }
  if (e != cudaSuccess) return cuda_fail(e, "host-buffer forward+backward");
  return rc;
}

// One forward+backward step on host buffers (see the header).  join_and_sync: the synchronous entry point -- the step
// runs on the plan's own main stream, ordered after the caller's stream, replayed from an instantiated CUDA graph when
// the same buffers come back (one cudaGraphLaunch instead of ~25 API calls), and the host blocks until it is complete.
// Otherwise everything is queued on the caller's stream and completion is left to
// rayen_forward_backward_host_wait(slot).
static int host_step(rayen_plan* p, const float* v_host, const float* gy_host, float* y_host, float* gv_host, int64_t B,
                     void* workspace, cudaStream_t stream, int slot, bool join_and_sync) {
  std::lock_guard<std::mutex> guard(*p->host_mu);
  int prev = 0;
  RAYEN_CUDA(cudaGetDevice(&prev));
  if (prev != p->device) RAYEN_CUDA(cudaSetDevice(p->device));
  cudaError_t e = cudaSuccess;
  if (!p->host_ready) {
    e = cudaStreamCreateWithFlags(&p->host_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->host_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->host_main, cudaStreamNonBlocking);
    for (auto& slot_ev : p->host_ev)
      for (cudaEvent_t& ev : slot_ev)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->host_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->host_join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      if (prev != p->device) cudaSetDevice(prev);
      return cuda_fail(e, "creating the copy streams of the host-buffer path");
    }
    p->host_ready = true;
  }
  // RAYEN_HOST_TRACE=1: print the device-side timeline of a synchronous call (development aid, adds event records)
  static const bool trace_env = getenv("RAYEN_HOST_TRACE") && atoi(getenv("RAYEN_HOST_TRACE")) != 0;
  int rc = 0;
  if (!join_and_sync) {
    rc = host_enqueue(p, v_host, gy_host, y_host, gv_host, B, workspace, stream, slot, false, false, false);
    if (prev != p->device) cudaSetDevice(prev);
    return rc;
  }
  // ---- synchronous step on the plan's main stream, ordered after the caller's stream
  cudaStream_t main = p->host_main;
  e = cudaEventRecord(p->host_fork, stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(main, p->host_fork, 0);
  const bool graph_ok = p->host_graph_on && !trace_env;
  rayen_plan::HostGraph* hit = nullptr;
  if (e == cudaSuccess && graph_ok) {
    for (int i = 0; i < p->host_graph_count; ++i) {
      rayen_plan::HostGraph& g = p->host_graphs[i];
      if (g.v == v_host && g.gy == gy_host && g.y == y_host && g.gv == gv_host && g.ws == workspace && g.B == B) hit = &g;
    }
  }
  if (e == cudaSuccess && hit) {
    hit->used = ++p->host_graph_clock;
    e = cudaGraphLaunch(hit->exec, main);
  } else if (e == cudaSuccess && graph_ok) {
    // first time with these buffers: capture the step (copy streams forked from and joined back into `main`), keep the
    // instantiated graph, launch it
    cudaGraph_t graph = nullptr;
    e = cudaStreamBeginCapture(main, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = host_enqueue(p, v_host, gy_host, y_host, gv_host, B, workspace, main, slot, true, true, false);
      cudaError_t e2 = cudaStreamEndCapture(main, &graph);
      if (rc == 0 && e2 != cudaSuccess) e = e2;
      cudaGraphExec_t exec = nullptr;
      if (rc == 0 && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (rc == 0 && e == cudaSuccess) {
        int at = p->host_graph_count;
        if (at == 8) {  // evict the least recently used entry
          at = 0;
          for (int i = 1; i < 8; ++i)
            if (p->host_graphs[i].used < p->host_graphs[at].used) at = i;
          cudaGraphExecDestroy(p->host_graphs[at].exec);
        } else {
          ++p->host_graph_count;
        }
        p->host_graphs[at] = {v_host, gy_host, y_host, gv_host, workspace, B, main, exec, ++p->host_graph_clock};
        e = cudaGraphLaunch(exec, main);
      }
    }
    if (rc == 0 && e != cudaSuccess) {
      // the step could not be captured (e.g. pageable host buffers): run it the plain way, now and from now on
      cudaGetLastError();
      p->host_graph_on = false;
      e = cudaSuccess;
      rc = host_enqueue(p, v_host, gy_host, y_host, gv_host, B, workspace, main, slot, true, true, false);
    }
  } else if (e == cudaSuccess) {
    rc = host_enqueue(p, v_host, gy_host, y_host, gv_host, B, workspace, main, slot, true, true, trace_env);
  }
  // the caller's stream is complete only when the step is, then block the host as documented
  cudaError_t e2 = cudaEventRecord(p->host_join, main);
  if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(stream, p->host_join, 0);
  cudaError_t e3 = cudaStreamSynchronize(main);
  cudaError_t e4 = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = e2;
  if (e == cudaSuccess) e = e3;
  if (e == cudaSuccess) e = e4;
  if (prev != p->device) cudaSetDevice(prev);
  if (rc != 0) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "host-buffer forward+backward");
  return RAYEN_OK;
}

static int host_check(const rayen_plan_t* cp, const void* a, const void* b, const void* c, const void* d, int64_t B,
                      const void* workspace) {
  if (!cp || !a || !b || !c || !d || (!workspace && B > 0)) return fail(RAYEN_ERR_BAD_ARGUMENT, "null argument");
  if (B < 0) return fail(RAYEN_ERR_BAD_ARGUMENT, "negative batch");
  if (!cp->host_mu) return fail(RAYEN_ERR_BAD_ARGUMENT, "plan has no host-path state");
  return 0;
}

extern "C" int rayen_forward_backward_host_f32(const rayen_plan_t* cp, const float* v_host, const float* gy_host,
                                               float* y_host, float* gv_host, int64_t B, void* workspace,
                                               void* stream_) {
  const int rc = host_check(cp, v_host, gy_host, y_host, gv_host, B, workspace);
  if (rc) return rc;
  if (B == 0) return RAYEN_OK;
  return host_step(const_cast<rayen_plan*>(cp), v_host, gy_host, y_host, gv_host, B, workspace,
                   static_cast<cudaStream_t>(stream_), 0, true);
}

extern "C" int rayen_forward_backward_host_submit_f32(const rayen_plan_t* cp, const float* v_host, const float* gy_host,
                                                      float* y_host, float* gv_host, int64_t B, void* workspace,
                                                      void* stream_, int slot) {
  const int rc = host_check(cp, v_host, gy_host, y_host, gv_host, B, workspace);
  if (rc) return rc;
  if (slot < 0 || slot >= kHostSlots) return fail(RAYEN_ERR_BAD_ARGUMENT, "slot must be 0..%d", kHostSlots - 1);
  if (B == 0) return fail(RAYEN_ERR_BAD_ARGUMENT, "empty batch");
  return host_step(const_cast<rayen_plan*>(cp), v_host, gy_host, y_host, gv_host, B, workspace,
                   static_cast<cudaStream_t>(stream_), slot, false);
}

extern "C" int rayen_forward_backward_host_wait(const rayen_plan_t* cp, int slot) {
  if (!cp) return fail(RAYEN_ERR_BAD_ARGUMENT, "null plan");
  if (slot < 0 || slot >= kHostSlots) return fail(RAYEN_ERR_BAD_ARGUMENT, "slot must be 0..%d", kHostSlots - 1);
  if (!cp->host_ready) return fail(RAYEN_ERR_BAD_ARGUMENT, "nothing was submitted on this plan");
  cudaError_t e = cudaEventSynchronize(cp->host_ev[slot][2 * kHostMaxChunks + 1]);
  if (e != cudaSuccess) return cuda_fail(e, "waiting for a submitted host-buffer step");
  return RAYEN_OK;
}
