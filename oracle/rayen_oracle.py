"""CPU ORACLE for the RAYEN ray-shooting layer -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / the timed CPU baseline.  The product (rayen_b200/) never
imports it and fails loudly when its CUDA extension is missing.

What it restates: the method='RAYEN' forward of leggedrobotics/rayen @ 2f007f7c and, through
autograd exactly like the reference, its backward:

    ctor precompute     rayen/constraint_module.py:38 (D), :43-52 (H, L), :99-122 (sigma, phi, delta)
    computeKappa        rayen/constraint_module.py:351-458
    solveSecondOrderEq  rayen/constraint_module.py:339-348
    forwardForRAYEN     rayen/constraint_module.py:468-474
    getyFromz           rayen/constraint_module.py:512-514

Two independent restatements live here:

* ``TorchOracle``   -- the same sequence of torch ops, in y-space, with the per-constraint Python
  loops, ``eigvalsh`` and autograd.  dtype is a parameter (float32 = the reference's README path,
  float64 = its benchmark path, time_analysis.py:25).  This is what ``bench.py`` times as the CPU
  baseline ("port") and what the fp32 parity bar (1e-5 relative) is measured against.
* ``closed_form_numpy`` -- float64 numpy, no autograd: the analytic forward + backward of SURVEY
  §3.3.  Used to cross-check the autograd path and to give per-sample kappa / active index / case.

Parity pin: the reference ships NO golden vectors or value-asserting tests.  The pin is
(1) tests/test_oracle_vs_reference.py, which runs this file against the unmodified reference
imported from /root/reference (build container only), and (2) tests/golden/*.npz, outputs of that
same reference committed together with the script that made them (tests/golden/make_golden.py).
"""
import numpy as np
import torch


# ----------------------------------------------------------------------------- set description
class OracleSet:
    """Frozen float64 numpy description of a preprocessed feasible set (what ConvexConstraints exposes)."""

    FIELDS = ("A_p", "b_p", "NA_E", "yp", "z0", "y0")

    def __init__(self, A_p, b_p, NA_E, yp, z0, y0, qcs=(), socs=(), lmi=None):
        f64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        self.A_p, self.b_p, self.NA_E = f64(A_p), f64(b_p).reshape(-1, 1), f64(NA_E)
        self.yp, self.z0, self.y0 = f64(yp).reshape(-1, 1), f64(z0).reshape(-1, 1), f64(y0).reshape(-1, 1)
        self.qcs = [(f64(P), f64(q).reshape(-1, 1), f64(r).reshape(1, 1)) for (P, q, r) in qcs]
        self.socs = [(f64(M), f64(s).reshape(-1, 1), f64(c).reshape(-1, 1), f64(d).reshape(1, 1)) for (M, s, c, d) in socs]
        self.lmi = [f64(F) for F in lmi] if lmi is not None else None
        self.k, self.n = self.NA_E.shape

    @classmethod
    def from_constraints(cls, cs):
        """Build from any object with the reference's ConvexConstraints attributes."""
        qcs = [(qc.P, qc.q, qc.r) for qc in cs.qcs]
        socs = [(sc.M, sc.s, sc.c, sc.d) for sc in cs.socs]
        lmi = list(cs.lmic.all_F) if cs.lmic is not None else None
        return cls(cs.A_p, cs.b_p, cs.NA_E, cs.yp, cs.z0, cs.y0, qcs, socs, lmi)


# ----------------------------------------------------------------------------- torch restatement
class TorchOracle:
    """Op-for-op torch restatement (CPU) of the reference's RAYEN forward; backward is autograd."""

    def __init__(self, oset, dtype=torch.float32):
        self.set, self.dtype = oset, dtype
        T = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        self.k, self.n = oset.k, oset.n
        # constraint_module.py:38 -- rows of A_p scaled by the slack of z0 (computed in float64, then cast)
        self.D = T(oset.A_p / ((oset.b_p - oset.A_p @ oset.z0) @ np.ones((1, oset.n))))
        self.NA_E, self.yp, self.z0, self.y0 = T(oset.NA_E), T(oset.yp), T(oset.z0), T(oset.y0)
        self.quads = []
        for (P, q, r) in oset.qcs:  # constraint_module.py:99-122, evaluated in `dtype` like the reference
            P, q, r = T(P), T(q), T(r)
            w = self.y0.T @ P + q.T                                 # [1,k]
            level = 0.5 * self.y0.T @ P @ self.y0 + q.T @ self.y0 + r  # g(y0) < 0
            sigma = 2 * level
            phi = -w / sigma
            delta = (w.T @ w - 4 * level * 0.5 * P) / torch.square(sigma)
            self.quads.append((phi, delta))
        self.socs = [(T(M), T(s), T(c), T(d)) for (M, s, c, d) in oset.socs]
        self.F, self.L = None, None
        if oset.lmi is not None:  # constraint_module.py:43-52 (numpy float64, then cast)
            allF = np.asarray(oset.lmi)
            H = allF[-1] + np.einsum("a,aij->ij", oset.y0[:, 0], allF[:-1])
            self.L = T(np.linalg.cholesky(np.linalg.inv(H)))
            self.F = T(allF[:-1])

    def kappa(self, u):
        """computeKappa (constraint_module.py:351-458).  u:[B,n,1] unit (or zero) directions."""
        kap = torch.relu(torch.max(self.D @ u, dim=1, keepdim=True).values)
        if not (self.quads or self.socs or self.F is not None):
            return kap
        rho = self.NA_E @ u
        rhoT = rho.transpose(1, 2)
        pieces = []
        for phi, delta in self.quads:                       # :360-381
            pieces.append(phi @ rho + torch.sqrt(rhoT @ delta @ rho))
        for (M, s, c, d) in self.socs:                      # :383-399 + :339-348
            beta = M @ self.y0 + s
            tau = c.T @ self.y0 + d
            c_p = rhoT @ M.T @ M @ rho - torch.square(c.T @ rho)
            b_p = 2 * rhoT @ M.T @ beta - 2 * (c.T @ rho) @ tau
            a_p = beta.T @ beta - torch.square(tau)
            disc = torch.square(b_p) - 4 * a_p * c_p
            root_minus = (-b_p - torch.sqrt(disc)) / (2 * a_p)
            root_plus = (-b_p + torch.sqrt(disc)) / (2 * a_p)
            pieces.append(torch.relu(torch.maximum(root_minus, root_plus)))
        if self.F is not None:                              # :401-449
            S = torch.einsum("ajk,ial->ijk", self.F, rho)
            sym = self.L.T @ (-S) @ self.L
            lam = torch.linalg.eigvalsh(sym).unsqueeze(2)
            pieces.append(torch.relu(torch.max(lam, dim=1, keepdim=True).values))
        stacked = torch.cat(pieces, dim=1)
        return torch.maximum(kap, torch.max(stacked, dim=1, keepdim=True).values)

    def forward(self, v):
        """forwardForRAYEN + getyFromz (constraint_module.py:468-474, :512-514).  v:[B,n] or [B,n,1] -> y:[B,k,1]."""
        v = v.reshape(v.shape[0], self.n, 1).to(self.dtype)
        u = torch.nn.functional.normalize(v, dim=1)
        kap = self.kappa(u)
        norm_v = torch.linalg.vector_norm(v, dim=(1, 2), keepdim=True)
        alpha = torch.minimum(1 / kap, norm_v)
        return self.NA_E @ (self.z0 + alpha * u) + self.yp

    def forward_old(self, q):
        """forwardForRAYENOld (constraint_module.py:460-466).  q:[B,n+1] (beta last) -> y:[B,k,1]."""
        q = q.reshape(q.shape[0], self.n + 1, 1).to(self.dtype)
        v, beta = q[:, 0:self.n, 0:1], q[:, self.n:self.n + 1, 0:1]
        u = torch.nn.functional.normalize(v, dim=1)
        alpha = 1 / (torch.exp(beta) + self.kappa(u))
        return self.NA_E @ (self.z0 + alpha * u) + self.yp

    def forward_backward(self, v, gy, method="RAYEN"):
        """Returns (y:[B,k], g_v:[B,n(+1)]) for the scalar loss sum(y * gy), via autograd like the reference."""
        cols = self.n if method == "RAYEN" else self.n + 1
        v = v.detach().clone().to(self.dtype).reshape(v.shape[0], cols).requires_grad_(True)
        y = self.forward(v) if method == "RAYEN" else self.forward_old(v)
        (y[:, :, 0] * gy.to(self.dtype).reshape(y.shape[0], self.k)).sum().backward()
        return y.detach()[:, :, 0], v.grad.detach()


# ----------------------------------------------------------------------------- closed form (numpy, fp64)
CASE_ZERO, CASE_INTERIOR, CASE_BOUNDARY = 0, 1, 2
FAM_NONE, FAM_LINEAR, FAM_QUAD, FAM_SOC, FAM_LMI = 0, 1, 2, 3, 4


def closed_form_numpy(oset, v, gy=None):
    """Analytic forward (+ backward if gy is given) in float64 numpy.

    Returns a dict: y[B,k], kappa[B], family[B], index[B], case[B], margin[B] (relative gap between the
    two largest kappas, for near-tie masking), cone_cond[B] (sqrt(disc)/|b'| of a binding cone, 1 otherwise: small
    means a nearly tangent ray, ill-conditioned in float32), lmi_gap[B] (relative gap between the two largest eigenvalues
    when the LMI binds, 1 otherwise: the gradient of lambda_max is conditioned like 1/gap) and, when gy is given, gv[B,n].
    """
    v = np.asarray(v, dtype=np.float64).reshape(-1, oset.n)
    B, n, k = v.shape[0], oset.n, oset.k
    N, z0, yp, y0 = oset.NA_E, oset.z0[:, 0], oset.yp[:, 0], oset.y0[:, 0]
    s = np.linalg.norm(v, axis=1)
    u = v / np.maximum(s, 1e-12)[:, None]
    rho = u @ N.T                                           # [B,k]

    kappas, grads, fams, conds = [], [], [], []             # grads: d kappa / d u, each [B,n]
    D = oset.A_p / (oset.b_p - oset.A_p @ oset.z0)
    lin = u @ D.T                                           # [B,m]
    j = np.argmax(lin, axis=1)
    kappas.append(lin[np.arange(B), j]); grads.append(D[j]); fams.append((FAM_LINEAR, j)); conds.append(np.ones(B))
    for i, (P, q, r) in enumerate(oset.qcs):
        w = P @ y0 + q[:, 0]
        a = 0.5 * y0 @ P @ y0 + q[:, 0] @ y0 + r[0, 0]
        sigma = 2 * a
        phi = -w / sigma
        Delta = (np.outer(w, w) - 2 * a * P) / sigma ** 2
        root = np.sqrt(np.einsum("bi,ij,bj->b", rho, Delta, rho))
        kap = rho @ phi + root
        dk_drho = phi[None, :] + (rho @ Delta) / np.where(root > 0, root, 1.0)[:, None]
        kappas.append(kap); grads.append(dk_drho @ N); fams.append((FAM_QUAD, np.full(B, i))); conds.append(np.ones(B))
    for i, (M, sv, c, d) in enumerate(oset.socs):
        beta = M @ y0 + sv[:, 0]
        tau = c[:, 0] @ y0 + d[0, 0]
        a_p = beta @ beta - tau ** 2
        Mr, cr = rho @ M.T, rho @ c[:, 0]
        b_p = 2 * Mr @ beta - 2 * cr * tau
        c_p = np.sum(Mr * Mr, axis=1) - cr ** 2
        disc = np.maximum(b_p ** 2 - 4 * a_p * c_p, 0.0)
        kap = np.maximum((-b_p - np.sqrt(disc)) / (2 * a_p), (-b_p + np.sqrt(disc)) / (2 * a_p))
        grad_b = 2 * (M.T @ beta) - 2 * tau * c[:, 0]
        grad_c = 2 * Mr @ M - 2 * cr[:, None] * c[:, 0][None, :]
        denom = 2 * a_p * kap + b_p
        dk_drho = -(kap[:, None] * grad_b[None, :] + grad_c) / np.where(denom != 0, denom, 1.0)[:, None]
        kappas.append(kap); grads.append(dk_drho @ N); fams.append((FAM_SOC, np.full(B, i)))
        # sqrt(disc)/|b'|: small for a ray nearly tangent to the cone, where kappa is ill-conditioned in any float32
        # evaluation (d kappa/d b' = -(1 + b'/sqrt(disc))/(2a'))
        conds.append(np.sqrt(disc) / np.maximum(np.abs(b_p), 1e-300))
    if oset.lmi is not None:
        allF = np.asarray(oset.lmi)
        H = allF[-1] + np.einsum("a,aij->ij", y0, allF[:-1])
        L = np.linalg.cholesky(np.linalg.inv(H))
        Ft = -np.einsum("ji,ajk,kl->ail", L, allF[:-1], L)   # F~_a = -L' F_a L
        S = np.einsum("ba,aij->bij", rho, Ft)
        lam, vec = np.linalg.eigh(S)
        qv = vec[:, :, -1]
        dk_drho = np.einsum("bi,aij,bj->ba", qv, Ft, qv)
        kappas.append(lam[:, -1]); grads.append(dk_drho @ N); fams.append((FAM_LMI, np.zeros(B, dtype=int)))
        conds.append(np.ones(B))
        # relative gap between the two largest eigenvalues: d lambda_max/du = q'F q has the condition number ~ 1/gap
        # (the reference's own float32 gradient is off by ~ eps/gap there: SURVEY 3.3, "eigengap-sensitive")
        if S.shape[1] > 1:
            lmi_gap_all = (lam[:, -1] - lam[:, -2]) / np.maximum(np.abs(lam).max(axis=1), 1e-300)
        else:
            lmi_gap_all = np.ones(B)

    K = np.stack(kappas, axis=1)                            # [B, 1+eta+mu+lmi]
    best = np.argmax(K, axis=1)
    kap = np.maximum(K[np.arange(B), best], 0.0)
    srt = np.sort(K, axis=1)
    runner = np.maximum(srt[:, -2], 0.0) if K.shape[1] > 1 else np.zeros(B)
    # also the runner-up inside the linear family
    if lin.shape[1] > 1:
        ls = np.sort(lin, axis=1)
        runner = np.where(best == 0, np.maximum(runner, np.maximum(ls[:, -2], 0.0)), runner)
    margin = np.where(kap > 0, (kap - runner) / np.maximum(kap, 1e-300), 1.0)
    dk_du = np.stack(grads, axis=1)[np.arange(B), best]     # [B,n]
    family = np.array([f[0] for f in fams])[best]
    index = np.stack([f[1] for f in fams], axis=1)[np.arange(B), best]
    family = np.where(kap > 0, family, FAM_NONE)
    cone_cond = np.stack(conds, axis=1)[np.arange(B), best]
    lmi_gap = np.where(family == FAM_LMI, lmi_gap_all, 1.0) if oset.lmi is not None else np.ones(B)

    with np.errstate(divide="ignore"):
        inv_kappa = np.where(kap > 0, 1.0 / np.where(kap > 0, kap, 1.0), np.inf)
    alpha = np.minimum(inv_kappa, s)
    y = (z0[None, :] + alpha[:, None] * u) @ N.T + yp[None, :]
    case = np.where(s == 0, CASE_ZERO, np.where(s <= inv_kappa, CASE_INTERIOR, CASE_BOUNDARY))
    out = dict(y=y, kappa=kap, family=family, index=index, case=case, margin=margin, alpha=alpha, s=s,
               cone_cond=cone_cond, lmi_gap=lmi_gap)
    if gy is not None:
        gz = np.asarray(gy, dtype=np.float64).reshape(B, k) @ N
        safe_k = np.where(kap > 0, kap, 1.0)
        gu = gz / safe_k[:, None] - (np.sum(gz * u, axis=1) / safe_k ** 2)[:, None] * dk_du
        gv_b = (gu - np.sum(gu * u, axis=1)[:, None] * u) / np.where(s > 0, s, 1.0)[:, None]
        gv = np.where((case == CASE_BOUNDARY)[:, None], gv_b, gz)
        gv = np.where((case == CASE_ZERO)[:, None], 0.0, gv)
        out["gv"] = gv
    return out


def max_violation(oset, y, A1=None, b1=None, A2=None, b2=None):
    """Max residual (float64) of y:[B,k] against every constraint of the ORIGINAL set (<= 0 is feasible)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, oset.k)
    worst = -np.inf
    if A1 is not None:
        worst = max(worst, float(np.max(y @ np.asarray(A1).T - np.asarray(b1).reshape(1, -1))))
    if A2 is not None:
        worst = max(worst, float(np.max(np.abs(y @ np.asarray(A2).T - np.asarray(b2).reshape(1, -1)))))
    for (P, q, r) in oset.qcs:
        worst = max(worst, float(np.max(0.5 * np.einsum("bi,ij,bj->b", y, P, y) + y @ q[:, 0] + r[0, 0])))
    for (M, s, c, d) in oset.socs:
        worst = max(worst, float(np.max(np.linalg.norm(y @ M.T + s[:, 0][None, :], axis=1) - y @ c[:, 0] - d[0, 0])))
    if oset.lmi is not None:
        allF = np.asarray(oset.lmi)
        Fy = np.einsum("ba,aij->bij", y, allF[:-1]) + allF[-1]
        worst = max(worst, float(np.max(-np.linalg.eigvalsh(Fy)[:, 0])))
    return worst
