"""Host-side plan packer: constraint set -> the z-space constant block the sm_100a kernels read.

One-time work, float64 numpy, then a single float32 blob laid out exactly as
``include/rayen_b200.h`` (struct ``RayenPlanDesc``) describes.  It restates, in the subspace
coordinates z (y = N z + yp), the per-constraint constants that the reference precomputes in
``ConstraintModule.__init__`` (rayen/constraint_module.py:38 D; :43-52 H, L; :99-122 sigma, phi,
delta) or recomputes on every forward call (:383-399: beta, tau and a' of each cone):

    linear     D = A_p / (b_p - A_p z0)                                       kappa_j = D_j . u
    quadratic  phi_z = N' phi,  G'G = N' Delta N   (G upper triangular)        kappa = phi_z.u + ||G u||
    SOC        c_z = N'c, h = (M N)'beta - tau c_z, R'R = (M N)'(M N), A = tau^2 - |beta|^2
               kappa = largest root of  -A kappa^2 + 2 (h.u) kappa + |R u|^2 - (c_z.u)^2
    LMI        F~z_a = sum_i N[i,a] (-L' F_i L),  L = chol(H^-1), H = F(y0)   kappa = lambda_max(sum_a u_a F~z_a)

so the ambient direction rho = N u is never formed on the device and every family costs O(n^2)
per constraint and sample (the triangular factors halve the reference's dense forms).
"""
import ctypes

import numpy as np

ABI_VERSION = 15
MAX_NP = 32      # widest subspace of the register-resident kernels; beyond it the WIDE section / wide.cuh take over
MAX_WIDE_N = 12288   # wide.cuh kWideMaxN
MAX_LMI = 32          # largest LMI of the register-resident kernels (lmi.cuh / lmi_warp.cuh)
MAX_LMI_BIG = 320     # largest LMI of the one-CTA-per-matrix path (lmi_big.cuh kLbMaxR)
LMIBT_MIN_N = 64      # from this subspace dimension on the big-LMI contraction also gets its tensor-core operand (LMIBT)


class PlanError(RuntimeError):
    pass


def _round_up_pow2(x, choices=(4, 8, 16, 32)):
    for c in choices:
        if x <= c:
            return c
    return None


def _bank_stride(words):
    """Smallest stride >= words that is a multiple of 4 and = 4 (mod 32): consecutive items start 16 B
    apart modulo the 128-B bank window, so 8 lanes reading 8 different items do not conflict."""
    s = (words + 3) // 4 * 4
    while s % 32 != 4:
        s += 4
    return s


def _triangular_factor(S, np_):
    """Upper-triangular T (np_ x np_, zero padded) with T'T = S for a symmetric PSD S (may be singular)."""
    n = S.shape[0]
    S = 0.5 * (S + S.T)
    lam, Q = np.linalg.eigh(S)
    G0 = np.sqrt(np.clip(lam, 0.0, None))[:, None] * Q.T
    R = np.linalg.qr(G0, mode="r")
    T = np.zeros((np_, np_))
    T[: R.shape[0], :n] = np.triu(R)
    return T


def _pack_triangular(T):
    """Row i keeps columns 4*floor(i/4) .. np-1."""
    np_ = T.shape[0]
    out = []
    for i in range(np_):
        out.append(T[i, (i // 4) * 4:])
    return np.concatenate(out)


def packed_triangular_words(np_):
    return sum(np_ - (i // 4) * 4 for i in range(np_))


TC_PANEL = 128  # rows of W per tensor-core panel (MMA N)
TC_TABLE_WORDS = 64
TC_BATCH_ITEMS = 12   # items per batch of the tensor-core kernel (lqs_tc.cuh kTcBatchItems)
LMI_TC_PANEL = 128  # entries of the LMI matrix per panel of the contraction GEMM (MMA N)
LMIW_R, LMIW_ROW_STRIDE = 32, 36   # lmi_warp.cuh: kLwR, kLwRowStride
WIDE_MAGIC = 0x57494445
WIDE_VERSION = 3
WIDE_HEADER_WORDS = 16
WIDE_TASK_WORDS = 8
WIDE_ITEM_WORDS = 8
WIDE_LIN, WIDE_QUAD, WIDE_SOC = 1, 2, 3     # linear-row task; kinds of an item
WIDE_FACTOR, WIDE_HDR = 2, 4                # tasks of the items: a group of factor rows / of header rows
WIDE_GROUP_ROWS = 64                        # rows per task: two per lane (one 8-byte load per column)
WIDE_SLOTS = 256        # partial-sum slots (groups of items) per round: wide.cuh kWideSlots
WIDE_ROUND_ITEMS = 64   # items per round: wide.cuh kWideRoundItems


def split_tf32(x):
    """x (float32) = hi + lo with both parts representable in TF32 (10-bit mantissa), round-to-nearest-away
    like ``cvt.rna.tf32.f32``; the residual x - hi - lo is below 2^-22 |x|."""
    def rna(a):
        bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = rna(x)
    lo = rna(x - hi)
    return hi, lo


def operand_layout(tile):
    """[rows, K] float32 -> the K-major, no-swizzle shared-memory layout of a tcgen05 operand:
    [k/4][row/8][row%8][k%4] (8x16-byte core matrices; LBO = rows*16 bytes, SBO = 128 bytes)."""
    rows, kdim = tile.shape
    return np.ascontiguousarray(tile.reshape(rows // 8, 8, kdim // 4, 4).transpose(2, 0, 1, 3)).reshape(-1)


class PackedPlan:
    """Float32 blob + the integers of ``RayenPlanDesc``; also keeps the float64 pieces for tests."""

    def __init__(self):
        self.blob = None
        self.fields = {}
        self.f64 = {}

    def desc(self):
        from ._cabi import RayenPlanDesc
        d = RayenPlanDesc()
        for key, val in self.fields.items():
            setattr(d, key, val)
        d.abi_version = ABI_VERSION
        d.blob_words = int(self.blob.size)
        d.blob = self.blob.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        return d


def build_plan(A_p, b_p, NA_E, yp, z0, qcs=(), socs=(), lmi=None, lin_rows=None):
    """Pack a preprocessed feasible set.

    ``A_p, b_p, NA_E, yp, z0`` are the fields of ``ConvexConstraints`` (reference constraints.py:366-436),
    ``qcs`` = [(P,q,r)], ``socs`` = [(M,s,c,d)], ``lmi`` = [F_0..F_k] or None, all in the ambient space.
    ``lin_rows`` = (A1, b1, A2, b2) of the original LinearConstraint (entries may be None); only the violation
    checker uses them.  Without them the checker tests the same polyhedron through A_p, b_p and N.
    """
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    A_p, N = f64(A_p), f64(NA_E)
    b_p, yp, z0 = f64(b_p).reshape(-1, 1), f64(yp).reshape(-1, 1), f64(z0).reshape(-1, 1)
    k, n = N.shape
    y0 = N @ z0 + yp
    np_ = _round_up_pow2(n)
    wide = np_ is None
    if wide:
        # n > 32: the directions no longer fit a thread's registers; the WIDE section below feeds wide.cuh
        if n > MAX_WIDE_N:
            raise PlanError(f"subspace dimension n={n} > {MAX_WIDE_N} is not covered by the sm_100a kernels yet")
        np_ = (n + 3) // 4 * 4
    k_pad = (k + 3) // 4 * 4
    plan = PackedPlan()
    sections = []
    cursor = 0

    exact = []   # (offset, float32 array) payloads that must reach the blob bit for bit (int tables, TF32 splits)

    def add_f32(arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        off = add(np.zeros(arr.size))
        exact.append((off, arr))
        return off

    def add(arr):
        nonlocal cursor
        arr = np.asarray(arr, dtype=np.float64).reshape(-1)
        pad = (-arr.size) % 4
        if pad:
            arr = np.concatenate((arr, np.zeros(pad)))
        off = cursor
        sections.append(arr)
        cursor += arr.size
        return off

    # ---- linear rows (reference constraint_module.py:38)
    slack = b_p - A_p @ z0
    if np.any(slack <= 0):
        raise PlanError("z0 is not strictly inside the linear constraints (b_p - A_p z0 must be > 0)")
    D = A_p / slack
    m = D.shape[0]
    m_pad = (m + 3) // 4 * 4
    Dp = np.zeros((m_pad, np_))
    Dp[:m, :n] = D
    lin_stride = 4 * np_ + 4
    # (wide plans keep their constants in the WIDE section only: the register-layout sections below stay empty)
    lin = np.zeros((m_pad // 4, lin_stride)) if not wide else np.zeros(4)
    for c in range(m_pad // 4 if not wide else 0):
        # [kk][i][e] = D[4c+i][4kk+e]
        blockv = Dp[4 * c:4 * c + 4, :].reshape(4, np_ // 4, 4).transpose(1, 0, 2)
        lin[c, :4 * np_] = blockv.reshape(-1)
    off_lin = add(lin)

    # ---- quadratic constraints (reference constraint_module.py:99-122)
    tri_words = packed_triangular_words(np_)
    quad_stride = _bank_stride(np_ + tri_words)
    quad = np.zeros((max(len(qcs), 1), quad_stride)) if not wide else None
    quad_f64 = []
    for i, (P, q, r) in enumerate(qcs):
        P, q, r = f64(P), f64(q).reshape(-1, 1), float(np.asarray(r).reshape(-1)[0])
        level = (0.5 * y0.T @ P @ y0 + q.T @ y0 + r).item()
        if level >= 0:
            raise PlanError(f"y0 is not strictly inside quadratic constraint {i} (g(y0)={level})")
        sigma = 2.0 * level
        w = P @ y0 + q
        phi = -w / sigma
        Delta = (w @ w.T - 2.0 * level * P) / sigma ** 2
        phi_z = (N.T @ phi)[:, 0]
        Delta_z = N.T @ Delta @ N
        G = _triangular_factor(Delta_z, np_)
        if not wide:
            quad[i, :n] = phi_z
            quad[i, np_:np_ + tri_words] = _pack_triangular(G)
        quad_f64.append((phi_z, Delta_z, G))
    off_quad = add(quad) if len(qcs) and not wide else add(np.zeros(4))

    # ---- second-order cones (reference constraint_module.py:383-399)
    soc_stride = _bank_stride(2 * np_ + tri_words + 4)
    soc = np.zeros((max(len(socs), 1), soc_stride)) if not wide else None
    soc_f64 = []
    for j, (M, s, c, d) in enumerate(socs):
        M, s, c, d = f64(M), f64(s).reshape(-1, 1), f64(c).reshape(-1, 1), float(np.asarray(d).reshape(-1)[0])
        beta = M @ y0 + s
        tau = (c.T @ y0 + d).item()
        A = tau * tau - (beta.T @ beta).item()
        if not (A > 0 and tau > 0):
            raise PlanError(f"y0 is not strictly inside SOC constraint {j}")
        Mz = M @ N
        cz = (N.T @ c)[:, 0]
        h = (Mz.T @ beta)[:, 0] - tau * cz
        R = _triangular_factor(Mz.T @ Mz, np_)
        if not wide:
            soc[j, :n] = cz
            soc[j, np_:np_ + n] = h
            soc[j, 2 * np_:2 * np_ + tri_words] = _pack_triangular(R)
            soc[j, 2 * np_ + tri_words] = A
        soc_f64.append((cz, h, Mz, A, R))
    off_soc = add(soc) if len(socs) and not wide else add(np.zeros(4))

    # ---- LMI (reference constraint_module.py:43-52 and :412-421, congruence folded into the constants)
    lmi_r = lmi_rp = 0
    lmi_big = False       # LMI beyond the register-resident kernels (r > 32, or any r with n > 32): lmi_big.cuh, section LMIB
    Fz = Fperm = None
    if lmi is not None:
        allF = np.asarray([f64(F) for F in lmi])
        lmi_r = allF.shape[1]
        lmi_big = wide or lmi_r > MAX_LMI
        if lmi_r > MAX_LMI_BIG:
            raise PlanError(f"LMI size r={lmi_r} > {MAX_LMI_BIG} is not covered by the sm_100a kernels yet")
        lmi_rp = 0 if lmi_big else _round_up_pow2(lmi_r)
        H = allF[-1] + np.einsum("a,aij->ij", y0[:, 0], allF[:-1])
        try:
            L = np.linalg.cholesky(np.linalg.inv(H))
        except np.linalg.LinAlgError:
            raise PlanError("y0 is not strictly inside the LMI constraint (F(y0) must be positive definite)")
        Ft = -np.matmul(np.matmul(L.T, allF[:-1]), L)          # batched BLAS: k matrices of r x r (r up to 320)
        Fz = Ft if (k == n and np.array_equal(N, np.eye(k))) else np.tensordot(N.T, Ft, axes=1)
        Fz = 0.5 * (Fz + Fz.transpose(0, 2, 1))
        if not lmi_big:
            lpm = lmi_rp // 4
            Fpad = np.zeros((n, lmi_rp, lmi_rp))
            Fpad[:, :lmi_r, :lmi_r] = Fz
            # [a][i][q][t] with column j = q + lpm*t
            Fperm = Fpad.reshape(n, lmi_rp, 4, lpm).transpose(0, 1, 3, 2)

    # ---- N and y0
    n_is_identity = int(k == n and np.array_equal(N, np.eye(k)))
    nm = np.zeros((k, np_ + 4))
    nm[:, :n] = N
    off_nmat = add(nm) if not (n_is_identity or wide) else add(np.zeros(4))
    y0p = np.zeros(k_pad)
    y0p[:k] = y0[:, 0]
    off_y0 = add(y0p)

    # ---- pruning bound of the LMI: lambda_max(S) <= tr(S)/r + sqrt((r-1)/r) sqrt(|S|_F^2 - tr(S)^2/r)
    # (Wolkowicz-Styan), with tr(S~(u)) = t.u and |S~(u)|_F = |T u|, T'T = [tr(F~z_a F~z_b)]_ab.  Lets the
    # linear/quadratic/SOC kernel prove, for most samples, that the LMI cannot be the binding constraint.
    # The Gram matrix is CENTRED, G_c = [tr(F~z_a F~z_b) - tr F~z_a tr F~z_b / r] (PSD by Cauchy-Schwarz: the Gram matrix
    # of the trace-free parts), so that |S~|_F^2 - tr(S~)^2/r = |T_c u|^2 is a sum of squares: forming it as a difference
    # in float32 cancels whenever S~(u) is close to a multiple of I (e.g. the epigraph form t I - A(y) >= 0) and the
    # "bound" then drops below lambda_max.  `bound_margin` covers the float32 / 3xTF32 rounding of the two dot products
    # (<= 5e-7 (|t|/r + |T_c|_F) for unit u) with 8x headroom; it is an absolute amount added to the bound.
    bound = np.zeros(np_ + tri_words + 4) if not wide else np.zeros(4)
    bound_margin = 0.0
    tr_F = bound_T = None
    if lmi is not None and not lmi_big:   # (the big path evaluates the same bound from the contracted matrix itself)
        tr_F = np.trace(Fz, axis1=1, axis2=2)
        gram = np.einsum("aij,bij->ab", Fz, Fz)
        gram_c = gram - np.outer(tr_F, tr_F) / float(lmi_r)
        gram_c = 0.5 * (gram_c + gram_c.T)
        bound_T = _triangular_factor(gram_c, np_)
        bound_margin = 4e-6 * (float(np.sqrt(max(np.trace(gram_c), 0.0))) + float(np.linalg.norm(tr_F)) / lmi_r)
        bound[:n] = tr_F
        bound[np_:np_ + tri_words] = _pack_triangular(bound_T)
        bound[np_ + tri_words] = float(lmi_r)
        bound[np_ + tri_words + 1] = bound_margin
    off_bound = add(bound)
    off_lmi = add(Fperm) if Fperm is not None else add(np.zeros(4))

    # ---- tensor-core layout of the same linear/quadratic/SOC/bound constants (lqs_tc.cuh): every constraint
    # becomes rows of one [rows x K] matrix W, so that all dot products of a 128-sample tile are ONE tcgen05
    # GEMM D = U W' with the result in tensor memory.  Rows are grouped in panels of 128; operands are split
    # W = W_hi + W_lo (both TF32-representable) for the error-compensated 3xTF32 product.
    # Panel kinds.  LINEAR: 128 rows of D.  BATCH: up to TC_BATCH_ITEMS items (quadratics, cones, the pruning bound) are
    # taken together and their triangular factors cut into blocks of 8 rows: block j (rows 8j..8j+7 of every item of the
    # batch, 8 x items rows in all) is one panel whose columns left of 8j are zero, so its GEMM starts at K step j
    # (4 + 3 + 2 + 1 = 10 K steps per batch at K = 32 instead of 16); block 0 also carries the items' two header rows
    # (phi_z | c_z | t, then h) behind its factor rows.  The per-item sums of squares are accumulated across the blocks.
    kp = max(8, np_)
    nblocks = kp // 8
    Wrows, table = [], []   # table rows: (ints[32], floats[TC_BATCH_ITEMS])

    def tri_dense(T):          # [np_, np_] upper triangular -> [kp, kp]
        out = np.zeros((kp, kp))
        out[:np_, :np_] = T
        return out

    if not wide:   # the tcgen05 panels exist for K <= 32 only
        Dk = np.zeros((m_pad, kp))
        Dk[:m, :n] = D
        for base in range(0, m_pad, TC_PANEL):
            blk = np.zeros((TC_PANEL, kp))
            rows = Dk[base:base + TC_PANEL]
            blk[:rows.shape[0]] = rows
            Wrows.append(blk)
            table.append(([0, base, TC_PANEL, 0] + [0] * 28, [0.0] * TC_BATCH_ITEMS))
        items = []     # (type, index, scalar, header rows [2, kp], factor [kp, kp])
        for i, (phi_z, Delta_z, G) in enumerate(quad_f64):
            hdr = np.zeros((2, kp))
            hdr[0, :n] = phi_z
            items.append((2, i, 0.0, hdr, tri_dense(G)))
        for j, (cz, h, Mz, A, R) in enumerate(soc_f64):
            hdr = np.zeros((2, kp))
            hdr[0, :n] = cz
            hdr[1, :n] = h
            items.append((3, j, A, hdr, tri_dense(R)))
        if lmi is not None and not lmi_big:
            hdr = np.zeros((2, kp))
            hdr[0, :n] = tr_F
            items.append((5, 0, float(lmi_r), hdr, tri_dense(bound_T)))
        n_batches = -(-len(items) // TC_BATCH_ITEMS)
        per_batch = -(-len(items) // n_batches) if n_batches else 0          # balanced batches
        for b0 in range(0, len(items), max(per_batch, 1)):
            batch = items[b0:b0 + per_batch]
            nb = len(batch)
            n_blk = (8 * nb + 15) // 16 * 16          # MMA N of a block panel (multiple of 16)
            hoff = n_blk                              # header rows of block 0 start here
            n0 = (hoff + 2 * nb + 15) // 16 * 16
            assert n0 <= TC_PANEL
            for jb in range(nblocks):
                blk = np.zeros((TC_PANEL, kp))
                # word 7: bit s set <=> item s of the batch is a cone; bits 16.. = 1 + slot of the pruning bound (0: none)
                meta = sum(1 << s_ for s_, it in enumerate(batch) if it[0] == 3)
                meta |= next((s_ + 1 for s_, it in enumerate(batch) if it[0] == 5), 0) << 16
                ints = [1, jb, n0 if jb == 0 else n_blk, jb, nb, int(jb == nblocks - 1), hoff, meta] + [0] * 24
                flts = [0.0] * TC_BATCH_ITEMS
                for s_, (typ, idx, scal, hdr, T) in enumerate(batch):
                    blk[8 * s_:8 * s_ + 8] = T[8 * jb:8 * jb + 8]
                    assert not np.any(T[8 * jb:8 * jb + 8, :8 * jb])          # the K steps the GEMM skips are zero
                    if jb == 0:
                        blk[hoff + 2 * s_:hoff + 2 * s_ + 2] = hdr
                    ints[8 + 2 * s_], ints[9 + 2 * s_] = typ, idx
                    flts[s_] = scal
                Wrows.append(blk)
                table.append((ints, flts))
    tc_panels = len(Wrows)
    tab = np.zeros((tc_panels, TC_TABLE_WORDS), dtype=np.float32)
    for pi, (ints, flts) in enumerate(table):
        tab[pi, :32] = np.asarray(ints, dtype=np.int32).view(np.float32)
        tab[pi, 32:32 + TC_BATCH_ITEMS] = np.asarray(flts, dtype=np.float32)
        # reciprocals of the item scalars (1 / A of a cone, 1 / r of the bound): the kernel multiplies instead of dividing
        tab[pi, 44:44 + TC_BATCH_ITEMS] = np.asarray([1.0 / f if f != 0.0 else 0.0 for f in flts], dtype=np.float32)
    off_tc = add_f32(tab)
    for blk in Wrows:
        hi, lo = split_tf32(blk.astype(np.float32))
        add_f32(operand_layout(hi))
        add_f32(operand_layout(lo))

    # ---- violation checker (rayen_violation_f32): the ORIGINAL constraints in the ambient space
    k4 = (k + 3) // 4 * 4
    if lin_rows is not None:
        A1, b1, A2, b2 = lin_rows
    else:  # same polyhedron from the preprocessed fields: A_p N'(y - yp) <= b_p and (I - N N')(y - yp) = 0
        A1, b1 = A_p @ N.T, b_p + A_p @ N.T @ yp
        proj = np.eye(k) - N @ N.T
        A2, b2 = (proj, proj @ yp) if n < k else (None, None)

    def rows_block(A, b):
        if A is None:
            return np.zeros((0, k4 + 4))
        A, b = f64(A), f64(b).reshape(-1)
        keep = np.any(A != 0, axis=1) | (b < 0)       # drop the reference's "no constraint" placeholder 0 <= 1
        A, b = A[keep], b[keep]
        blk = np.zeros((A.shape[0], k4 + 4))
        blk[:, :k] = A
        blk[:, k4] = b
        return blk

    vin, veq = rows_block(A1, b1), rows_block(A2, b2)
    vparts = [vin.reshape(-1), veq.reshape(-1)]
    for (P, q, r) in qcs:
        Pp = np.zeros((k4, k4))
        Pp[:k, :k] = f64(P)
        qp = np.zeros(k4)
        qp[:k] = f64(q).reshape(-1)
        vparts += [Pp.reshape(-1), qp, np.array([float(np.asarray(r).reshape(-1)[0]), 0.0, 0.0, 0.0])]
    for (M, s_, c, d) in socs:
        M, s_, c = f64(M), f64(s_).reshape(-1), f64(c).reshape(-1)
        cp = np.zeros(k4)
        cp[:k] = c
        rows = np.zeros((M.shape[0], k4 + 4))
        rows[:, :k] = M
        rows[:, k4] = s_
        vparts += [np.array([float(M.shape[0]), float(np.asarray(d).reshape(-1)[0]), 0.0, 0.0]), cp, rows.reshape(-1)]
    off_viol = add(np.concatenate(vparts) if sum(v.size for v in vparts) else np.zeros(4))
    viol_in, viol_eq = vin.shape[0], veq.shape[0]
    off_lmineg = add(np.zeros(4))
    if lmi is not None and not lmi_big:
        # lambda_max(-F(y)) = -lambda_min(F(y)): the same solver on the k+1 matrices -F_0..-F_k with u = (y, 1)
        allF = np.asarray([f64(F) for F in lmi])
        lpm = lmi_rp // 4
        Fneg = np.zeros((k + 1, lmi_rp, lmi_rp))
        Fneg[:, :lmi_r, :lmi_r] = -0.5 * (allF + allF.transpose(0, 2, 1))
        off_lmineg = add(Fneg.reshape(k + 1, lmi_rp, 4, lpm).transpose(0, 1, 3, 2))

    # ---- LMI matrices as the B operand of the contraction GEMM S = U W' (lmi_tc.cuh): W [rp*rp, kp] with
    # row e = i*rp + 4q + t holding F~z_.[i][q + lpm*t] (the LMI section's order), in 128-row panels, TF32 split
    off_lmitc = add(np.zeros(4))
    lmitc_panels = 0
    if lmi is not None and not lmi_big and lmi_rp >= 16:
        Wl = np.zeros((lmi_rp * lmi_rp, kp), dtype=np.float32)
        Wl[:, :n] = np.asarray(Fperm, dtype=np.float64).reshape(n, lmi_rp * lmi_rp).T.astype(np.float32)
        lmitc_panels = lmi_rp * lmi_rp // LMI_TC_PANEL
        for pi in range(lmitc_panels):
            hi, lo = split_tf32(Wl[pi * LMI_TC_PANEL:(pi + 1) * LMI_TC_PANEL])
            off = add_f32(operand_layout(hi))
            add_f32(operand_layout(lo))
            if pi == 0:
                off_lmitc = off

    # ---- LMIW (lmi_warp.cuh, the filter + one-warp-per-matrix solver): F~z_a row-major, zero padded to 32 x 32, with
    # a row stride of 36 words so that 32 lanes reading 16 bytes of 32 different rows do not collide in shared memory
    off_lmiw = 0
    if lmi is not None and not lmi_big:
        Fw = np.zeros((n, LMIW_R, LMIW_ROW_STRIDE))
        Fw[:, :lmi_r, :lmi_r] = Fz
        off_lmiw = add(Fw)

    # ---- LMIB / LMINEGB (lmi_big.cuh): the lower triangles, row-major packed (entry (i, j), j <= i, at i (i + 1) / 2 + j),
    # rows padded to a multiple of 4 words -- F~z_a for the contraction GEMM S~(v) = V . F, and -F_0 .. -F_k of the ambient
    # space (u = (y, 1)) for the violation metric
    off_lmib = off_lminegb = 0
    lmib_p4 = 0
    if lmi_big:
        il = np.tril_indices(lmi_r)
        lmib_p4 = (il[0].size + 3) // 4 * 4
        Fb = np.zeros((n, lmib_p4))
        Fb[:, :il[0].size] = Fz[:, il[0], il[1]]
        off_lmib = add(Fb)
        Fn = -0.5 * (allF + allF.transpose(0, 2, 1))
        Fnb = np.zeros((k + 1, lmib_p4))
        Fnb[:, :il[0].size] = Fn[:, il[0], il[1]]
        off_lminegb = add(Fnb)
    # ---- LMIBT (lmi_big_tc.cuh): the same F~z as the B operand of the tcgen05 contraction GEMM, for subspaces wide enough
    # for the GEMM to matter (n >= 64): F' [lmib_p4 -> panels of 128 entries][n -> slices of 32], every [128 x 32] tile
    # TF32-split (hi, lo) and stored in the K-major no-swizzle operand layout [k/4][row/8][row%8][k%4];
    # tile pair (panel q, slice s) at off_lmibt + (q * slices + s) * 8192 words
    off_lmibt = 0
    lmibt_panels = lmibt_slices = 0
    if lmi_big and n >= LMIBT_MIN_N:
        lmibt_panels = -(-lmib_p4 // 128)
        lmibt_slices = -(-n // 32)
        Ft32 = np.zeros((lmibt_panels * 128, lmibt_slices * 32), dtype=np.float32)
        Ft32[:lmib_p4, :n] = Fb.T.astype(np.float32)
        hi, lo = split_tf32(Ft32)

        def tiles(a):    # [q*128 + rg*8 + r8][s*32 + kc*4 + k4] -> [q][s][kc][rg][r8][k4]
            return a.reshape(lmibt_panels, 16, 8, lmibt_slices, 8, 4).transpose(0, 3, 4, 1, 2, 5)
        both = np.stack((tiles(hi), tiles(lo)), axis=2)      # [q][s][hi/lo][kc][rg][r8][k4]
        off_lmibt = add_f32(np.ascontiguousarray(both).reshape(-1))

    # ---- WIDE section (n > 32, wide.cuh): every constraint as rows of ONE matrix W [R_pad x n], stored transposed
    # (Wt[j][row]) so that a warp's 32 lanes read 32 consecutive rows of a column with one coalesced load.  The unit
    # of work of a warp is a TASK = one group of 32 rows (see rayen_b200.h); the groups of an item leave partial sums
    # in shared-memory slots that a finalize pass adds up in a fixed order, so items are processed in ROUNDS that fit
    # the slot budget.  N is kept twice: transposed for y = y0 + alpha N u (thread per ambient coordinate) and
    # row-major for g_z = N' g_y (thread per subspace coordinate).
    off_wide = 0
    if wide:
        G_ROWS = WIDE_GROUP_ROWS
        cg = lambda x: (x + G_ROWS - 1) // G_ROWS * G_ROWS
        m_g = cg(m)
        lin_block = np.zeros((m_g, n))
        lin_block[:m] = D
        blocks = [lin_block]
        lin_tasks = [(WIDE_LIN, g * G_ROWS, 0, g * G_ROWS, 0, 0) for g in range(m_g // G_ROWS)]
        row = m_g
        groups_per_item = cg(n) // G_ROWS
        raw_items = [(WIDE_QUAD, i, phi_z, None, G, 0.0) for i, (phi_z, Delta_z, G) in enumerate(quad_f64)] + \
                    [(WIDE_SOC, j, cz, h, R, A) for j, (cz, h, Mz, A, R) in enumerate(soc_f64)]
        # rounds: as many whole items as fit WIDE_SLOTS partial-sum slots and WIDE_ROUND_ITEMS headers
        per_round = max(1, min(WIDE_ROUND_ITEMS, WIDE_SLOTS // groups_per_item))
        rounds, tasks = [], []
        itab = np.zeros((len(raw_items) + 1, WIDE_ITEM_WORDS), dtype=np.float32)
        first = True
        for i0 in range(0, max(len(raw_items), 1), per_round):
            chunk = raw_items[i0:i0 + per_round]
            cur_tasks = list(lin_tasks) if first else []
            first = False
            # header rows of the round's items: rows 2*il (phi_z | c_z) and 2*il + 1 (h of a cone, 0 of a quadratic)
            hdr_block = np.zeros((cg(2 * len(chunk)), n))
            hdr_row0 = row
            for il, (kind, fidx, a0, a1, T, A) in enumerate(chunk):
                hdr_block[2 * il] = a0
                if a1 is not None:
                    hdr_block[2 * il + 1] = a1
            if len(chunk):
                blocks.append(hdr_block)
                for hg in range(hdr_block.shape[0] // G_ROWS):
                    cur_tasks.append((WIDE_HDR, row + hg * G_ROWS, 0, hg * (G_ROWS // 2), 0, 0))
                row += hdr_block.shape[0]
            slots = 0
            for il, (kind, fidx, a0, a1, T, A) in enumerate(chunk):
                blk = np.zeros((cg(n), n))
                blk[:n] = np.triu(T[:n, :n])
                blocks.append(blk)
                itab[i0 + il, :5] = np.asarray([row, kind, fidx, slots, groups_per_item], dtype=np.int32).view(np.float32)
                itab[i0 + il, 5] = A
                itab[i0 + il, 6:7] = np.asarray([hdr_row0 + 2 * il], dtype=np.int32).view(np.float32)
                for g in range(groups_per_item):
                    # columns below the group's first row are zero in every row of the group
                    cur_tasks.append((WIDE_FACTOR, row + g * G_ROWS, g * G_ROWS // 4 * 4, il, 0, slots + g))
                slots += groups_per_item
                row += blk.shape[0]
            # heaviest groups first, stable (linear rows stay in ascending order)
            cur_tasks.sort(key=lambda t: t[2])
            rounds.append((len(tasks), len(tasks) + len(cur_tasks), i0, i0 + len(chunk)))
            tasks.extend(cur_tasks)
        r_pad = row
        header = np.zeros(WIDE_HEADER_WORDS, dtype=np.int32)
        off_wide = add_f32(header.view(np.float32))
        header_slot = exact[-1]
        ttab = np.asarray([list(t) + [0, 0] for t in tasks], dtype=np.int32).reshape(len(tasks), WIDE_TASK_WORDS)
        off_tasks = add_f32(ttab.view(np.float32))
        off_rounds = add_f32(np.asarray(rounds, dtype=np.int32).reshape(-1).view(np.float32))
        off_items = add_f32(itab)
        off_wt = add(np.concatenate(blocks).T)                           # [n][r_pad]
        k32 = (k + 31) // 32 * 32
        off_nt = off_nrow = 0
        if not n_is_identity:
            nt = np.zeros((n, k32))
            nt[:, :k] = N.T
            off_nt = add(nt)                                              # [n][k32]: NT[j][i] = N[i][j]
            nr = np.zeros((k, np_))
            nr[:, :n] = N
            off_nrow = add(nr)                                            # [k][np]
        header[:] = 0
        header[:15] = [WIDE_MAGIC, r_pad, len(tasks), off_tasks, off_wt, off_nt, off_nrow, k32, np_, off_items,
                       len(quad_f64), len(soc_f64), len(rounds), off_rounds, WIDE_VERSION]
        header_slot[1][:] = header.view(np.float32)

    blob = np.concatenate(sections).astype(np.float32)
    assert blob.size == cursor and cursor % 4 == 0
    for off, arr in exact:
        blob[off:off + arr.size] = arr
    plan.blob = np.ascontiguousarray(blob)
    plan.fields = dict(n=n, k=k, np=np_, k_pad=k_pad, m=m, m_pad=m_pad, n_quad=len(qcs), n_soc=len(socs),
                       lmi_r=lmi_r, lmi_rp=lmi_rp, n_is_identity=n_is_identity, lmi_big=int(lmi_big), lmib_p4=lmib_p4,
                       off_lmib=off_lmib, off_lminegb=off_lminegb, off_lmibt=off_lmibt, lmibt_panels=lmibt_panels,
                       lmibt_slices=lmibt_slices,
                       lin_chunk_stride=lin_stride, quad_stride=quad_stride, soc_stride=soc_stride,
                       off_lin=off_lin, off_quad=off_quad, off_soc=off_soc, off_nmat=off_nmat,
                       off_y0=off_y0, off_bound=off_bound, off_lmi=off_lmi, lmi_prune=int(lmi is not None and not lmi_big),
                       off_tc=off_tc, tc_panels=tc_panels, tc_kp=kp,
                       off_viol=off_viol, off_lmineg=off_lmineg, viol_in=viol_in, viol_eq=viol_eq,
                       off_lmitc=off_lmitc, lmitc_panels=lmitc_panels, wide=int(wide), off_wide=off_wide,
                       off_lmiw=off_lmiw,
                       lmi_bound_margin=float(np.float32(bound_margin)))
    plan.f64 = dict(D=D, N=N, y0=y0, z0=z0, yp=yp, quads=quad_f64, socs=soc_f64, Fz=Fz,
                    bound=(tr_F, bound_T, lmi_r, bound_margin) if (lmi is not None and not lmi_big) else None)
    return plan


def build_plan_from_constraints(cs):
    """Pack a ``ConvexConstraints``-like object (this package's or the reference's)."""
    qcs = [(qc.P, qc.q, qc.r) for qc in cs.qcs]
    socs = [(sc.M, sc.s, sc.c, sc.d) for sc in cs.socs]
    lmi = list(cs.lmic.all_F) if cs.lmic is not None else None
    lc = getattr(cs, "lc", None)
    lin_rows = (lc.A1, lc.b1, lc.A2, lc.b2) if lc is not None else (None, None, None, None)
    return build_plan(cs.A_p, cs.b_p, cs.NA_E, cs.yp, cs.z0, qcs, socs, lmi, lin_rows=lin_rows)


# --------------------------------------------------------------------------- numpy model of the kernels
def evaluate_plan_numpy(plan, v):
    """Float64 evaluation of kappa and y straight from the PACKED blob (not from the original matrices).
    Narrow plans (n <= 32) only: wide plans are decoded by ``evaluate_wide_numpy``.

    Host-side self-check of the packer: it decodes the same words the kernels decode, so a layout bug
    shows up on the CPU.  Not a fallback -- nothing in the product path calls it.
    """
    f = plan.fields
    assert not f.get("wide"), "wide plans carry only the WIDE section: use evaluate_wide_numpy"
    blob = plan.blob.astype(np.float64)
    n, k, np_ = f["n"], f["k"], f["np"]
    v = np.asarray(v, dtype=np.float64).reshape(-1, n)
    B = v.shape[0]
    s = np.linalg.norm(v, axis=1)
    u = np.zeros((B, np_))
    u[:, :n] = v / np.maximum(s, 1e-12)[:, None]
    best = np.zeros(B)
    act = np.zeros(B, dtype=np.int64)

    def consider(val, tag):
        nonlocal best, act
        better = val > best
        best = np.where(better, val, best)
        act = np.where(better, tag, act)

    for c in range(f["m_pad"] // 4):
        base = f["off_lin"] + c * f["lin_chunk_stride"]
        blk = blob[base:base + 4 * np_].reshape(np_ // 4, 4, 4).transpose(1, 0, 2).reshape(4, np_)
        for i in range(4):
            consider(u @ blk[i], (1 << 24) | (4 * c + i))

    def unpack_tri(words):
        T = np.zeros((np_, np_))
        pos = 0
        for i in range(np_):
            c0 = (i // 4) * 4
            T[i, c0:] = words[pos:pos + np_ - c0]
            pos += np_ - c0
        return T

    tri_words = packed_triangular_words(np_)
    for i in range(f["n_quad"]):
        base = f["off_quad"] + i * f["quad_stride"]
        phi = blob[base:base + np_]
        G = unpack_tri(blob[base + np_:base + np_ + tri_words])
        consider(u @ phi + np.linalg.norm(u @ G.T, axis=1), (2 << 24) | i)
    for j in range(f["n_soc"]):
        base = f["off_soc"] + j * f["soc_stride"]
        cz, h = blob[base:base + np_], blob[base + np_:base + 2 * np_]
        R = unpack_tri(blob[base + 2 * np_:base + 2 * np_ + tri_words])
        A = blob[base + 2 * np_ + tri_words]
        hb, cu = u @ h, u @ cz
        cq = np.sum((u @ R.T) ** 2, axis=1) - cu ** 2
        root = np.sqrt(np.maximum(hb * hb + A * cq, 0.0))
        consider((hb + root) / A, (3 << 24) | j)
    if f["lmi_r"] and f.get("lmi_big"):
        consider(lmi_big_lambda_max_numpy(plan, u[:, :n]), 4 << 24)
    elif f["lmi_r"]:
        rp = f["lmi_rp"]
        lpm = rp // 4
        Fperm = blob[f["off_lmi"]:f["off_lmi"] + n * rp * rp].reshape(n, rp, lpm, 4)
        Fz = Fperm.transpose(0, 1, 3, 2).reshape(n, rp, rp)
        S = np.einsum("ba,aij->bij", u[:, :n], Fz)
        consider(np.linalg.eigvalsh(S)[:, -1], 4 << 24)
    with np.errstate(divide="ignore"):
        alpha = np.minimum(np.where(best > 0, 1.0 / np.where(best > 0, best, 1.0), np.inf), s)
    y0 = blob[f["off_y0"]:f["off_y0"] + k]
    if f["n_is_identity"]:
        rho = u[:, :n]
    else:
        Nm = blob[f["off_nmat"]:f["off_nmat"] + k * (np_ + 4)].reshape(k, np_ + 4)[:, :np_]
        rho = u @ Nm.T
    return y0[None, :] + alpha[:, None] * rho, best, act


def lmi_big_lambda_max_numpy(plan, u):
    """lambda_max(sum_a u_a F~z_a) for every row of u, decoded from the LMIB section (float64)."""
    f = plan.fields
    n, r, p4 = f["n"], f["lmi_r"], f["lmib_p4"]
    Fb = plan.blob[f["off_lmib"]:f["off_lmib"] + n * p4].astype(np.float64).reshape(n, p4)
    il = np.tril_indices(r)
    S = np.zeros((u.shape[0], r, r))
    S[:, il[0], il[1]] = u @ Fb[:, :il[0].size]
    S = S + np.tril(S, -1).transpose(0, 2, 1)
    return np.linalg.eigvalsh(S)[:, -1]


def evaluate_wide_numpy(plan, v):
    """Float64 evaluation of kappa, the binding constraint and y from the WIDE section of the packed blob, round by
    round and task by task as ``wide_forward_kernel`` walks it, including the leading columns a task skips (a CPU
    self-check of the layout; nothing in the product path calls it)."""
    f = plan.fields
    assert f["wide"], "not a wide plan"
    blob32 = plan.blob
    hdr = blob32[f["off_wide"]:f["off_wide"] + WIDE_HEADER_WORDS].view(np.int32)
    assert hdr[0] == WIDE_MAGIC and hdr[14] == WIDE_VERSION
    r_pad, n_tasks, off_tasks, off_wt, off_nt, off_nrow, k32, np_, off_items, n_quad, n_soc, n_rounds, off_rounds = (
        int(x) for x in hdr[1:14])
    n, k = f["n"], f["k"]
    blob = blob32.astype(np.float64)
    Wt = blob[off_wt:off_wt + n * r_pad].reshape(n, r_pad)
    v = np.asarray(v, dtype=np.float64).reshape(-1, n)
    B = v.shape[0]
    s = np.linalg.norm(v, axis=1)
    u = v / np.maximum(s, 1e-12)[:, None]
    best = np.zeros(B)
    act = np.zeros(B, dtype=np.int64)

    def consider(val, tag):
        nonlocal best, act
        better = (val > best) | ((val == best) & (val > 0) & (tag < act))
        best = np.where(better, val, best)
        act = np.where(better, tag, act)

    tasks = blob32[off_tasks:off_tasks + n_tasks * WIDE_TASK_WORDS].view(np.int32).reshape(n_tasks, WIDE_TASK_WORDS)
    rounds = blob32[off_rounds:off_rounds + 4 * n_rounds].view(np.int32).reshape(n_rounds, 4)
    items = blob32[off_items:off_items + (n_quad + n_soc) * WIDE_ITEM_WORDS].reshape(n_quad + n_soc, WIDE_ITEM_WORDS)
    seen_tasks = 0
    for (t0, t1, i0, i1) in rounds:
        assert t0 == seen_tasks and i1 - i0 <= WIDE_ROUND_ITEMS
        seen_tasks = t1
        part = np.zeros((WIDE_SLOTS, B))
        head = np.zeros((WIDE_ROUND_ITEMS, 2, B))
        for kind, row, j0, idx, _, slot, _, _ in tasks[t0:t1]:
            assert j0 % 4 == 0 and row % WIDE_GROUP_ROWS == 0
            assert not np.any(Wt[:j0, row:row + WIDE_GROUP_ROWS]), "a task skips non-zero columns"
            P = u[:, j0:] @ Wt[j0:, row:row + WIDE_GROUP_ROWS]            # [B, 64]: lane l owns rows 2l, 2l + 1
            if kind == WIDE_LIN:
                for r in range(WIDE_GROUP_ROWS):
                    consider(P[:, r], (1 << 24) | (idx + r))
            elif kind == WIDE_HDR:
                for l in range(WIDE_GROUP_ROWS // 2):
                    if idx + l < i1 - i0:
                        head[idx + l, 0] = P[:, 2 * l]
                        head[idx + l, 1] = P[:, 2 * l + 1]
            else:
                assert kind == WIDE_FACTOR and 0 <= slot < WIDE_SLOTS and 0 <= idx < i1 - i0
                part[slot] = (P ** 2).sum(axis=1)
        for ii in range(i0, i1):
            rb, kind, fidx, slot0, nparts = (int(x) for x in items[ii, :5].view(np.int32))
            hdr_row = int(items[ii, 6:7].view(np.int32)[0])
            assert np.array_equal(u @ Wt[:, hdr_row], head[ii - i0, 0]) or np.allclose(u @ Wt[:, hdr_row], head[ii - i0, 0])
            nrm2 = part[slot0:slot0 + nparts].sum(axis=0)
            a0, a1 = head[ii - i0]
            if kind == WIDE_QUAD:
                consider(a0 + np.sqrt(nrm2), (2 << 24) | fidx)
            else:
                A = float(items[ii, 5])
                cq = nrm2 - a0 ** 2
                root = np.sqrt(np.maximum(a1 * a1 + A * cq, 0.0))
                consider((a1 + root) / A, (3 << 24) | fidx)
    assert seen_tasks == n_tasks
    if f["lmi_r"]:
        lam = lmi_big_lambda_max_numpy(plan, u)
        better = lam > best            # ties keep the earlier family
        best = np.where(better, lam, best)
        act = np.where(better, 4 << 24, act)
    with np.errstate(divide="ignore"):
        alpha = np.minimum(np.where(best > 0, 1.0 / np.where(best > 0, best, 1.0), np.inf), s)
    y0 = blob[f["off_y0"]:f["off_y0"] + k]
    if f["n_is_identity"]:
        rho = u
    else:
        NT = blob[off_nt:off_nt + n * k32].reshape(n, k32)[:, :k]
        Nrow = blob[off_nrow:off_nrow + k * np_].reshape(k, np_)[:, :n]
        assert np.array_equal(NT.T, Nrow)
        rho = u @ NT
    return y0[None, :] + alpha[:, None] * rho, best, act
