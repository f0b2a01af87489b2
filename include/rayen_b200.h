/*
 * rayen_b200.h -- C ABI of the B200-native RAYEN feasibility layer (librayen_b200.so).
 *
 * The reference (leggedrobotics/rayen @ 2f007f7c) is pure Python/PyTorch and has no FFI layer of
 * its own; the boundary it exposes is the Python class API (rayen/constraint_module.py:17-533).
 * This header is the thin C boundary that the Python drop-in (rayen_b200/constraint_module.py)
 * binds with ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every tensor, the library owns only the plan;
 *   - device pointers unless the name says "host"; dense row-major; float32 compute ("f32");
 *   - every call returns 0 on success, a negative RAYEN_ERR_* code, or a positive cudaError_t;
 *     nothing throws, exits or synchronises (launch-and-return on the given stream), except the
 *     *_host_* entry points, which synchronise the stream before returning;
 *   - a plan is immutable after creation => re-entrant across host threads and streams;
 *   - rayen_last_error() returns a thread-local message for the last failing call.
 */
#ifndef RAYEN_B200_H
#define RAYEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAYEN_ABI_VERSION 15

/* error codes (negative); positive return values are cudaError_t */
#define RAYEN_OK 0
#define RAYEN_ERR_BAD_ARGUMENT (-1)
#define RAYEN_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels cover (n > 12288, LMI size > 320) */
#define RAYEN_ERR_ABI (-3)
#define RAYEN_ERR_NO_DEVICE (-4)

/* which scale step follows kappa: reference constraint_module.py:468-474 (RAYEN), :460-466 (RAYEN_old) */
#define RAYEN_MODE_RAYEN 0     /* alpha = min(1/kappa, ||v||);       v has n columns            */
#define RAYEN_MODE_RAYEN_OLD 1 /* alpha = 1/(exp(beta)+kappa);       v has n+1 columns (beta last) */

/* family tags stored in the upper 8 bits of the `active` output; low 24 bits = constraint index */
#define RAYEN_FAM_NONE 0
#define RAYEN_FAM_LINEAR 1
#define RAYEN_FAM_QUAD 2
#define RAYEN_FAM_SOC 3
#define RAYEN_FAM_LMI 4

/*
 * Packed, z-space constant block of one feasible set ("plan"), built once on the host by
 * rayen_b200/plan.py from the buffers the reference registers in ConstraintModule.__init__
 * (constraint_module.py:38 D, :43-52 H/L, :59-74 buffers, :99-122 phi/delta).  All offsets are in
 * float32 words from `blob` and are multiples of 4 (16-byte aligned sections).
 *
 *   np      n rounded up to 4, 8, 16 or 32: the register row length of a direction u
 *   LIN     m_pad/4 chunks of 4 rows; chunk c at off_lin + c*lin_chunk_stride;
 *           word [kk*16 + i*4 + e] = D[4c+i][4kk+e]           (D = A_p / (b_p - A_p z0), zero padded)
 *   QUAD    item q at off_quad + q*quad_stride: phi_z[np], then the upper-triangular factor G
 *           (G'G = N' Delta N) row by row, row i holding columns 4*floor(i/4) .. np-1
 *   SOC     item j at off_soc + j*soc_stride: c_z[np], h[np] (h = M_z' beta - tau c_z), the
 *           triangular factor R (R'R = M_z' M_z) packed like G, then {A = tau^2 - beta'beta, 0,0,0}
 *   NMAT    k rows of N (= NA_E), row stride np+4 (absent when N is the identity)
 *   Y0      y0 = N z0 + yp, k_pad words
 *   BOUND   (plans with an LMI) t[np] (t_a = tr F~z_a), the packed triangular factor T_c of the CENTRED Gram matrix
 *           [tr(F~z_a F~z_b) - tr F~z_a tr F~z_b / r]_ab, then {r, margin, 0, 0}: the Wolkowicz-Styan bound
 *           lambda_max(S~(u)) <= t.u/r + sqrt((r-1)/r) |T_c u| that prunes the eigen-solve.  |T_c u|^2 is the squared
 *           Frobenius norm of the trace-free part of S~(u) as a sum of squares (no float32 cancellation); a sample is
 *           pruned only if bound + margin + 1e-5 |bound| < kappa of the other families, margin = lmi_bound_margin
 *           covering the rounding of the two dot products
 *   LMI     F~z_a = sum_i N[i][a] * (-L' F_i L), a < n, each rp x rp (rp = r rounded up to 4, 8, 16
 *           or 32, zero padded), stored [a][row i][lane q][slot t] with column j = q + (rp/4)*t
 *   TC      the LIN/QUAD/SOC/BOUND constants again as the B operand of a tcgen05 GEMM (lqs_tc.cuh): a table
 *           of tc_panels x 32 words {kind, first row, 8 x (item type, item index), pad to 24, 8 item scalars}, then
 *           per panel W_hi and W_lo (128 x tc_kp each, TF32 split) in the K-major no-swizzle operand layout
 *           [k/4][row/8][row%8][k%4].  A linear panel holds 128 rows of D; an item panel holds 128/(8+kp)
 *           items of 8 header rows (phi | c_z, h | t) followed by the kp rows of the triangular factor.
 *   VIOL    the ORIGINAL constraints in the ambient space for rayen_violation_f32 (k4 = k rounded up to 4):
 *           viol_in rows {a[k4], b, 0,0,0} of A1 y <= b1, viol_eq rows of A2 y = b2, per quadratic
 *           {P[k4][k4], q[k4], r,0,0,0}, per cone {r_M, d, 0, 0, c[k4], r_M rows {M_i[k4], s_i,0,0,0}}
 *   LMITC   (lmi_rp >= 16) the LMI matrices again as the B operand of the contraction GEMM S = U W' (lmi_tc.cuh):
 *           lmitc_panels = rp*rp/128 panels of 128 entries, entry e = i*rp + 4q + t <-> F~z_a[i][q + (rp/4)*t]
 *           (the LMI section's order), per panel W_hi and W_lo (128 x tc_kp each, TF32 split) in the operand
 *           layout [k/4][row/8][row%8][k%4]
 *   LMIW    (plans with an LMI) F~z_a once more for the filter + one-warp-per-matrix solver of lmi_warp.cuh: n matrices,
 *           zero padded to 32 x 32, row-major with a row stride of 36 words ([a][row][36]); lane j of a warp reads
 *           row j (= column j, the matrices are symmetric) as eight 16-byte loads, conflict-free in shared memory
 *   WIDE    (wide == 1: 32 < n <= 12288, linear + quadratic + SOC (+ an LMI through LMIB); wide.cuh) 16 int32 words {magic 0x57494445, R_pad,
 *           n_tasks, off_tasks, off_wt, off_nt, off_nrow, k32, np, off_items, n_quad, n_soc, n_rounds, off_rounds,
 *           layout version 3, 0} (offsets in words from `blob`), then
 *             Wt      [n][R_pad], Wt[j][row] = W[row][j]; W stacks the rows of D (zero padded to 64), then per round of
 *                     items a block of header rows (row 2i: phi_z of a quadratic / c_z of a cone, row 2i+1: h of a
 *                     cone / zeros; padded to 64) followed by the n rows of each item's upper-triangular factor
 *                     (G: G'G = N'Delta N; R: R'R = (MN)'(MN); padded to 64)
 *             tasks   n_tasks x 8 int32 {kind 1 linear / 2 factor / 4 header, first row (multiple of 64), first
 *                     non-zero column (multiple of 4), index of the first row (linear) | item within the round
 *                     (factor) | first item within the round (header), 0, partial-sum slot (factor), 0, 0}: a task is
 *                     64 rows, two per lane; within a round the tasks are sorted heaviest first
 *             rounds  n_rounds x 4 int32 {first task, end task, first item, end item}: at most 64 items and 256 slots
 *             items   (n_quad + n_soc) x 8 words {first factor row, kind 2 quadratic / 3 cone, index in its family,
 *                     first slot, slots, float A of a cone, first header row, 0}, quadratics first
 *             NT      [n][k32] = N' and Nrow [k][np] = N (both absent when N is the identity)
 *           The LIN/QUAD/SOC/NMAT/BOUND sections are empty placeholders (np = n rounded up to 4) and TC has no
 *           panels (tc_panels = 0): a wide plan is WIDE + Y0 + VIOL.
 *   LMIB    (lmi_big == 1: lmi_r > 32, or an LMI together with n > 32; lmi_big.cuh) F~z_a, a < n, as packed lower
 *           triangles: entry (i, j), j <= i, at word i (i + 1) / 2 + j of a row of lmib_p4 words -- the [n x lmib_p4]
 *           B operand of the contraction GEMM S~(v) = V . F (FP32 pipe); the eigen-solve runs one CTA per sample on the
 *           result.  LMINEGB: -F_0 .. -F_k of the ambient space in the same packing ([k + 1][lmib_p4]).
 *   LMINEG  -F_0 .. -F_k laid out like LMI ([a][row][lane][slot]): lambda_max(sum_a (y,1)_a (-F_a)) = -lambda_min(F(y))
 */
typedef struct RayenPlanDesc {
  int32_t abi_version; /* must be RAYEN_ABI_VERSION */
  int32_t n;           /* dimension of the subspace (columns of v)  */
  int32_t k;           /* dimension of the ambient space (columns of y) */
  int32_t np;
  int32_t k_pad;
  int32_t m;     /* rows of D (before padding) */
  int32_t m_pad; /* multiple of 4 */
  int32_t n_quad;
  int32_t n_soc;
  int32_t lmi_r;  /* 0 = no LMI */
  int32_t lmi_rp; /* 4, 8, 16 or 32 */
  int32_t n_is_identity;
  int32_t lin_chunk_stride;
  int32_t quad_stride;
  int32_t soc_stride;
  int32_t lmi_prune; /* 1: the BOUND section is valid and pruning may be used */
  int32_t tc_panels; /* number of 128-row panels of the tensor-core section */
  int32_t tc_kp;     /* K of the tensor-core GEMM: max(8, np) */
  int32_t viol_in;   /* inequality rows of the VIOL section */
  int32_t viol_eq;   /* equality rows of the VIOL section */
  int32_t lmitc_panels; /* 128-entry panels of the LMITC section (0: none) */
  int32_t wide;         /* 1: n > 32 -- np is n rounded up to 4, the kernels of wide.cuh read the WIDE section */
  int32_t lmi_big;      /* 1: the LMI is served by lmi_big.cuh (lmi_r > 32, or any lmi_r together with n > 32): section LMIB,
                           lmi_rp = 0 and the register-resident LMI sections (LMI, LMIW, LMITC, LMINEG, BOUND) are empty */
  int32_t lmib_p4;      /* words per packed lower triangle: lmi_r (lmi_r + 1) / 2 rounded up to 4 */
  int32_t lmibt_panels; /* LMIBT: panels of 128 packed entries (0: no tensor-core operand, the FP32 GEMM is used) */
  int32_t lmibt_slices; /* LMIBT: K slices of 32 subspace coordinates */
  float lmi_bound_margin; /* absolute float32-rounding allowance added to the pruning bound (see BOUND) */
  int64_t off_lin, off_quad, off_soc, off_nmat, off_y0, off_bound, off_lmi, off_tc, off_viol, off_lmineg, off_lmitc;
  int64_t off_wide;     /* WIDE section (0 when wide == 0) */
  int64_t off_lmiw;     /* LMIW section (0 without an LMI) */
  int64_t off_lmib;     /* LMIB section: F~z_a, a < n, lower triangles row-major packed ((i, j), j <= i, at i (i + 1) / 2 + j),
                           lmib_p4 words each -- the B operand of the contraction GEMM S~(v) = V . F of lmi_big.cuh */
  int64_t off_lminegb;  /* LMINEGB section: -F_0 .. -F_k of the ambient space in the same packing (violation metric) */
  int64_t off_lmibt;    /* LMIBT section (lmi_big_tc.cuh): F~z' as the B operand of the tcgen05 contraction GEMM -- per (panel of
                           128 entries, slice of 32 coordinates) a TF32-split pair (hi, lo) of [128 x 32] tiles in the
                           operand layout [k/4][row/8][row%8][k%4], 8192 words per pair, panel-major */
  int64_t blob_words;
  const float* blob; /* host pointer, blob_words floats */
} RayenPlanDesc;

typedef struct rayen_plan rayen_plan_t;

int rayen_abi_version(void);
const char* rayen_last_error(void);

/* Copies the constant block to `device` (one cudaMalloc + cudaMemcpy) and selects the kernels.
 * Replaces the .to(device) of the reference's registered buffers. */
int rayen_plan_create(const RayenPlanDesc* desc, int device, rayen_plan_t** out);
void rayen_plan_destroy(rayen_plan_t* plan);

/* Optional launch tuning for sweeps: samples per thread (1, 2 or 4; 0 = auto) and lanes per sample
 * (power of two <= 32; 0 = auto) of the linear/quadratic/SOC kernel. */
int rayen_plan_set_tuning(rayen_plan_t* plan, int samples_per_thread, int lanes_per_sample);
/* LMI pruning on (1, default) / off (0): with pruning the eigen-solve only runs for the samples whose
 * Wolkowicz-Styan bound does not already prove kappa_LMI < kappa of the other families.  Results are
 * identical either way: the bound is evaluated in a cancellation-free form with an explicit allowance for its
 * float32 rounding (see BOUND above), so a pruned sample's LMI cannot bind. */
int rayen_plan_set_pruning(rayen_plan_t* plan, int enabled);
/* Linear/quadratic/SOC forward on the tensor cores (tcgen05 3xTF32 GEMM, default) or on the FP32 pipe (0). */
int rayen_plan_set_tensor_cores(rayen_plan_t* plan, int enabled);
/* Output rows of the tcgen05 kernel through shared-memory tiles: every store instruction writes whole rows of y (512
 * contiguous bytes per four rows) instead of 16 bytes per lane a row apart.  Off by default (neutral on local HBM); on
 * when y is a peer / NVSwitch-multicast mapping (the all-gather fused into the kernel's epilogue, sharding.forward_gathered).
 * Call it before the plan is used concurrently. */
int rayen_plan_set_coalesced_output(rayen_plan_t* plan, int enabled);
/* LMI contraction sum_a u_a F~z_a (reference constraint_module.py:412-421) as a tcgen05 3xTF32 GEMM inside the LMI
 * forward kernel (lmi_tc.cuh; needs lmi_rp >= 16) or on the FP32 pipe out of shared memory (lmi.cuh).
 * mode: 0 never, 1 wherever available, 2 automatic (default): the measured policy -- tensor cores when K = 32 and
 * the call carries no gradient work.  RAYEN_LMI_TC=0/1/2 sets the default of new plans. */
int rayen_plan_set_lmi_tensor_cores(rayen_plan_t* plan, int mode);

/* LMI forward for samples that carry a kappa of the other families (every set with an LMI and anything else): an exact
 * definiteness filter -- LDL' of (kappa_prior - margin) I - S~(u), all pivots positive <=> the LMI cannot bind -- in
 * front of a one-warp-per-matrix eigen-solver for the samples that fail it (lmi_warp.cuh).  Results are those of the
 * unfiltered path (a passing sample keeps the prior kappa, which is what the merge after a full solve keeps).
 * mode: 0 never (the 8-lanes-per-matrix kernels of lmi.cuh), 1 wherever available, 2 automatic (default: padded LMI
 * sizes 16 and 32).  RAYEN_LMI_WARP=0/1/2 sets the default of new plans; RAYEN_LMI_FILTER=0 keeps the solver but skips
 * the filter. */
int rayen_plan_set_lmi_filter(rayen_plan_t* plan, int mode);

/* Device scratch the forward / backward calls need for a batch of B samples (work lists of the samples
 * that still need the LMI kernels, and d kappa/du of the LMI-bound samples).  0 for plans without an LMI.  The caller owns the buffer; it must
 * not be shared by calls that may run concurrently. */
int64_t rayen_workspace_bytes(const rayen_plan_t* plan, int64_t B);
/* (Layout, for diagnostics: the workspace starts with three int32 counters that a forward call leaves behind -- [0] the
 * number of samples the pruning bound could not settle (the LMI work list), [2] how many of those failed the
 * definiteness filter beyond the filter kernel's own solve budget; [1] is the backward work list's.) */

/*
 * Forward: replaces forwardForRAYEN / forwardForRAYENOld + computeKappa + getyFromz
 * (constraint_module.py:351-474, :512-514).
 *   v      [B, ldv]  (ldv >= n, or >= n+1 for RAYEN_OLD), y [B, k]
 *   kappa  [B] and active [B] receive kappa and (family << 24 | index) of the binding constraint;
 *          they are what backward needs.  They may be NULL only for plans without an LMI.
 *   want_grad  != 0: a backward call will follow.  For plans with an LMI the forward then also computes, for the
 *          samples whose binding constraint is the LMI, d kappa/du (top eigenvector, q' F q) while the
 *          tridiagonal form is still in registers, and leaves it in the workspace for backward.
 *   workspace  rayen_workspace_bytes(plan, B) bytes of device memory (may be NULL when that is 0); the
 *          backward call of the same batch must get the same buffer
 */
int rayen_forward_f32(const rayen_plan_t* plan, const float* v, int64_t ldv, float* y, float* kappa,
                      int32_t* active, int64_t B, int mode, int want_grad, void* workspace, void* cuda_stream);

/*
 * Forward with the layer's mapper fused in (reference constraint_module.py:261 and :525: q = nn.Linear(input_dim, n)(x)
 * followed by forwardForRAYEN): x [B, in_dim] (row stride ldx), weight [n, in_dim] (row stride ldw), bias [n] or NULL.
 * The linear/quadratic/SOC kernel computes v = W x + b itself, writes it to v_out [B, n] (dense; pass it as `v` to
 * rayen_backward_f32, whose g_v is then the gradient w.r.t. the mapper's output) and continues as rayen_forward_f32
 * (mode RAYEN).  RAYEN_ERR_UNSUPPORTED when in_dim / ldx / ldw are not multiples of 4 floats, the tensors are not
 * 16-byte aligned, or the plan has no linear/quadratic/SOC kernel to host the mapper (LMI-only sets): the caller then
 * runs the mapper itself and calls rayen_forward_f32.
 */
int rayen_forward_mapped_f32(const rayen_plan_t* plan, const float* x, int64_t ldx, int32_t in_dim, const float* weight,
                             int64_t ldw, const float* bias, float* v_out, float* y, float* kappa, int32_t* active,
                             int64_t B, int want_grad, void* workspace, void* cuda_stream);

/*
 * Backward: the closed form of what autograd derives from the reference forward.
 *   gy [B, k], kappa/active from the forward call on the same v, gv [B, ldv_g] (ldv_g = n or n+1).
 *   have_dkappa != 0: the forward call ran with want_grad != 0 on this workspace (no LMI kernel is launched);
 *   0: the LMI-bound samples are recomputed by a separate kernel.
 */
int rayen_backward_f32(const rayen_plan_t* plan, const float* v, int64_t ldv, const float* gy,
                       const float* kappa, const int32_t* active, float* gv, int64_t ldgv, int64_t B,
                       int mode, int have_dkappa, void* workspace, void* cuda_stream);

/* Per-kernel launches for profiling and the roofline measurement in bench.py: stage_mask bit 0 = the
 * linear/quadratic/SOC kernel, bit 1 = the LMI kernel (3 = what rayen_forward_f32 / rayen_backward_f32
 * launch).  With stage_mask == 2 the kappa/active buffers and the workspace must already hold the first stage's result. */
int rayen_forward_stage_f32(const rayen_plan_t* plan, const float* v, int64_t ldv, float* y, float* kappa,
                            int32_t* active, int64_t B, int mode, int want_grad, int stage_mask, void* workspace,
                            void* cuda_stream);
int rayen_backward_stage_f32(const rayen_plan_t* plan, const float* v, int64_t ldv, const float* gy,
                             const float* kappa, const int32_t* active, float* gv, int64_t ldgv, int64_t B,
                             int mode, int have_dkappa, int stage_mask, void* workspace, void* cuda_stream);

/* Host-buffer variant (the end-to-end path): the batch is cut into chunks; host->device copies run on an internal
 * copy-in stream, the kernels on `cuda_stream`, device->host copies on an internal copy-out stream, so that the two
 * directions of the link and the SMs work at the same time.  Everything is ordered after what the caller queued on
 * `cuda_stream`, which is complete (and synchronised) when the call returns.  Host buffers should be pinned.
 * `workspace` is a device buffer of rayen_host_workspace_bytes(plan, B) bytes.  Calls on one plan are serialised
 * (the copy streams belong to the plan); RAYEN_HOST_CHUNKS=1..8 overrides the chunk count. */
int64_t rayen_host_workspace_bytes(const rayen_plan_t* plan, int64_t B);
int rayen_forward_backward_host_f32(const rayen_plan_t* plan, const float* v_host, const float* gy_host,
                                    float* y_host, float* gv_host, int64_t B, void* workspace,
                                    void* cuda_stream);

/* The same step, queued without blocking: several steps can be in flight, so that the copy-in of step i+1 overlaps the
 * kernels and the copy-out of step i (a double-buffered input pipeline).  `slot` (0..3) names the step for
 * rayen_forward_backward_host_wait, which blocks until its y_host / gv_host are complete.  Each slot in flight needs
 * its own `workspace` and host buffers; a slot may be submitted again only after it has been waited for.  The caller's
 * stream is NOT ordered after the copy-out of a submitted step. */
int rayen_forward_backward_host_submit_f32(const rayen_plan_t* plan, const float* v_host, const float* gy_host,
                                           float* y_host, float* gv_host, int64_t B, void* workspace, void* cuda_stream,
                                           int slot);
int rayen_forward_backward_host_wait(const rayen_plan_t* plan, int slot);

/*
 * Max constraint residual per sample (<= 0 is feasible), float32, of y [B, ldy] against the ORIGINAL constraints:
 * max(A1 y - b1, |A2 y - b2|, g_i(y), ||M_j y + s_j|| - c_j'y - d_j, relu(-lambda_min(F(y)))).  Replaces the
 * per-sample cvxpy projection distance of constraints.py:549-559 / examples/main.py:176 as the violation metric.
 *   viol [B]
 */
int rayen_violation_f32(const rayen_plan_t* plan, const float* y, int64_t ldy, float* viol, int64_t B,
                        void* cuda_stream);

/*
 * The optional exchange step of the multi-GPU path (all-gather of y when the downstream loss couples the
 * batch) as ONE kernel over peer memory instead of a library collective: the `rows` x `k` block `src` of this rank is
 * stored at row `row_offset` of the gathered [world * rows, k] buffer of every rank.
 *   multicast_base != NULL: the NVSwitch multicast mapping of that buffer (torch symmetric memory: multicast_ptr) -- one
 *                           `multimem.st` per 16 bytes, replicated to all ranks by the switch;
 *   otherwise:              `n_dst` peer-mapped base pointers (host array, the own buffer included), plain stores
 *                           over NVLink (P2P).
 * Launch-and-return on `cuda_stream`; the caller orders the consumers behind it with a cross-rank barrier (the symmetric
 * memory handle's).  Replaces dist.all_gather_into_tensor in rayen_b200/sharding.py::all_gather_outputs.
 */
int rayen_gather_push_f32(const float* src, int64_t rows, int32_t k, float* const* dst_bases, int32_t n_dst,
                          float* multicast_base, int64_t row_offset, void* cuda_stream);

/* Number of kernels this library has launched in the calling process (all plans, all threads). */
int64_t rayen_launch_count(void);
/* Launches `count` empty kernels (148 x 128 threads) on the stream: the launch floor a chain of kernels pays on this
 * GPU, measured by bench.py next to the kernel times (not counted by rayen_launch_count). */
int rayen_launch_empty(int count, void* cuda_stream);

/* Introspection for tests / bench: static shared memory, registers, and chosen launch geometry. */
typedef struct RayenKernelInfo {
  int32_t regs_lqs_fwd, regs_lqs_bwd, regs_lmi_fwd, regs_lmi_bwd;
  int32_t smem_lqs_bytes, smem_lmi_bytes;
  int32_t sm_count;
  int32_t reserved;
} RayenKernelInfo;
int rayen_plan_kernel_info(const rayen_plan_t* plan, RayenKernelInfo* out);

#ifdef __cplusplus
}
#endif
#endif /* RAYEN_B200_H */
