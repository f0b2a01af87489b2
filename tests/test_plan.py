"""Host-side plan packer: the packed float32 block decodes (in float64 numpy) to the oracle's answers."""
import numpy as np
import pytest
import torch

from helpers import golden_names, load_golden
from oracle.rayen_oracle import OracleSet, closed_form_numpy
from rayen_b200 import plan, synthetic


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("old_")])
def test_packed_plan_reproduces_oracle(name):
    g = load_golden(name)
    cs = synthetic.build_constraints(g["spec"])
    p = plan.build_plan_from_constraints(cs)
    ev = plan.evaluate_wide_numpy if p.fields["wide"] else plan.evaluate_plan_numpy
    y, kap, act = ev(p, g["v"])
    cf = closed_form_numpy(OracleSet.from_constraints(cs), g["v"])
    scale = np.abs(cf["y"]).max()
    assert np.abs(y - cf["y"]).max() <= 2e-6 * scale      # float32 rounding of the constants only
    assert np.abs(kap - cf["kappa"]).max() <= 2e-6 * max(1.0, cf["kappa"].max())
    clear = cf["margin"] > 1e-4
    assert np.array_equal((act >> 24)[clear], cf["family"][clear])
    lin = clear & (cf["family"] == 1)
    assert np.array_equal((act & 0xFFFFFF)[lin], cf["index"][lin])


def test_layout_invariants():
    cs = synthetic.build_constraints(synthetic.config_spec("cfg5"))
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    assert p.blob.dtype == np.float32 and p.blob.size % 4 == 0
    offs = [f[k] for k in ("off_lin", "off_quad", "off_soc", "off_nmat", "off_y0", "off_lmi")]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)
    assert f["np"] == 32 and f["lmi_rp"] == 32 and f["n_is_identity"] == 1
    assert f["quad_stride"] % 32 == 4 and f["soc_stride"] % 32 == 4 and f["lin_chunk_stride"] == 4 * 32 + 4
    assert (f["off_lmi"] - f["off_lin"]) * 4 + 64 <= 227 * 1024      # the LQS constants fit one CTA's shared memory
    assert f["n"] * f["lmi_rp"] ** 2 * 4 == 131072                   # F~z: 128 KiB


def test_equality_constraints_give_a_subspace_plan():
    cs = synthetic.build_constraints(synthetic.example_spec("readme"))
    p = plan.build_plan_from_constraints(cs)
    assert (p.fields["n"], p.fields["k"], p.fields["np"], p.fields["n_is_identity"]) == (2, 3, 4, 0)
    assert p.fields["lmi_r"] == 2 and p.fields["lmi_rp"] == 4


def test_rejects_boundary_or_exterior_points():
    spec = synthetic.example_spec(2)       # sphere of radius 2
    spec["y0"] = np.array([[2.0], [0.0], [0.0]])
    cs = synthetic.build_constraints(spec)
    with pytest.raises(plan.PlanError):
        plan.build_plan_from_constraints(cs)
    spec = synthetic.example_spec(0)
    spec["y0"] = np.array([[1.0], [0.0], [0.0]])   # on a face of the cube
    with pytest.raises(plan.PlanError):
        plan.build_plan_from_constraints(synthetic.build_constraints(spec))


def test_unsupported_sizes_fail_loudly():
    with pytest.raises(plan.PlanError, match="n=12300"):
        plan.build_plan_from_constraints(synthetic.build_constraints(synthetic.random_spec(k=12300, m=2)))
    with pytest.raises(plan.PlanError, match="r=321"):
        plan.build_plan_from_constraints(synthetic.build_constraints(synthetic.random_spec(k=2, r=321)))


def test_big_lmi_plans_pack_the_lower_triangles():
    """LMI beyond the register-resident kernels (r > 32, or any r with n > 32): section LMIB, decoded here in float64
    and compared with the oracle's kappa."""
    from oracle.rayen_oracle import OracleSet, closed_form_numpy
    for spec in (synthetic.random_spec(k=4, m=6, eta=1, mu=1, r_M=3, r=40, seed=1),
                 synthetic.wide_spec(40, 30, 1, 1, 8, 2, seed=2, r=6)):
        cs = synthetic.build_constraints(spec)
        p = plan.build_plan_from_constraints(cs)
        f = p.fields
        assert f["lmi_big"] == 1 and f["lmi_rp"] == 0 and f["lmi_prune"] == 0 and f["off_lmiw"] == 0
        r = f["lmi_r"]
        assert f["lmib_p4"] == (r * (r + 1) // 2 + 3) // 4 * 4
        v, _ = synthetic.sample_inputs(64, cs.n, cs.k, seed_v=3)
        ev = plan.evaluate_wide_numpy if f["wide"] else plan.evaluate_plan_numpy
        y, kap, act = ev(p, v.numpy())
        cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy(), np.zeros((64, cs.k)))
        np.testing.assert_allclose(y, cf["y"], rtol=0, atol=2e-6)
        assert ((act >> 24) == 4).any()          # the LMI binds for some samples


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes(seed):
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 33))
    spec = synthetic.random_spec(k=k, m=int(rng.integers(0, 40)), eta=int(rng.integers(0, 4)),
                                 mu=int(rng.integers(0, 4)), r_M=int(rng.integers(1, 2 * k + 1)),
                                 r=int(rng.integers(0, 2)) * int(rng.integers(2, 33)), seed=seed)
    if spec["A1"] is None and not spec["qcs"] and not spec["socs"] and spec["lmi"] is None:
        spec = synthetic.random_spec(k=k, m=5, seed=seed)
    if spec["b1"] is not None:
        spec["b1"] = spec["b1"] * 3
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    v, _ = synthetic.sample_inputs(64, cs.n, cs.k, seed_v=seed, dtype=torch.float64)
    y, kap, _ = plan.evaluate_plan_numpy(p, v.numpy())
    cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy())
    assert np.abs(y - cf["y"]).max() <= 5e-6 * max(1.0, np.abs(cf["y"]).max())


def _decode_operand(words, rows, kdim):
    """Inverse of plan.operand_layout: [k/4][row/8][row%8][k%4] -> [rows, kdim]."""
    return words.reshape(kdim // 4, rows // 8, 8, 4).transpose(1, 2, 0, 3).reshape(rows, kdim)


@pytest.mark.parametrize("cfg", ["cfg3", "cfg5"])
def test_tensor_core_sections_decode_to_the_fp32_sections(cfg):
    """The B operands of the two tcgen05 GEMMs (TC: linear rows + items of the linear/quadratic/SOC kernel; LMITC: the
    LMI matrices) are the SAME constants as the FP32-pipe sections, split hi + lo: hi and lo are TF32-representable
    (13 low mantissa bits zero), hi + lo reproduces the float32 value to 2^-22, panels and tables have the shape the
    kernels assume."""
    cs = synthetic.build_constraints(synthetic.config_spec(cfg))
    p = plan.build_plan_from_constraints(cs)
    f, blob = p.fields, p.blob
    kp, panels = f["tc_kp"], f["tc_panels"]
    tab = blob[f["off_tc"]:f["off_tc"] + panels * plan.TC_TABLE_WORDS].reshape(panels, plan.TC_TABLE_WORDS)
    ints = tab.view(np.int32)
    w0 = f["off_tc"] + panels * plan.TC_TABLE_WORDS
    tile = plan.TC_PANEL * kp
    rows = []
    for pi in range(panels):
        hi = blob[w0 + (2 * pi) * tile: w0 + (2 * pi + 1) * tile]
        lo = blob[w0 + (2 * pi + 1) * tile: w0 + (2 * pi + 2) * tile]
        for part in (hi, lo):
            assert not np.any(part.view(np.uint32) & np.uint32(0x1FFF))            # TF32-representable
        rows.append(_decode_operand(hi.astype(np.float64) + lo.astype(np.float64), plan.TC_PANEL, kp))
    # linear panels first: their rows are D
    D = np.asarray(p.f64["D"])
    m = D.shape[0]
    lin = np.concatenate([rows[pi] for pi in range(panels) if ints[pi, 0] == 0])[:m, :cs.n]
    assert np.abs(lin - D).max() <= 2.0 ** -21 * max(1.0, np.abs(D).max())
    assert [int(ints[pi, 1]) for pi in range(panels) if ints[pi, 0] == 0] == list(range(0, f["m_pad"], plan.TC_PANEL))
    # batch panels: block j of every item's triangular factor (8 rows per item), headers behind block 0; item types in
    # family order; the K steps a block panel skips are zero; the blocks of an item reassemble its factor
    nblocks = kp // 8
    batch_panels = [pi for pi in range(panels) if ints[pi, 0] == 1]
    assert len(batch_panels) % nblocks == 0
    types, n_lmi = [], (1 if cs.has_lmi_constraints else 0)
    for b0 in range(0, len(batch_panels), nblocks):
        grp = batch_panels[b0:b0 + nblocks]
        nb, hoff = int(ints[grp[0], 4]), int(ints[grp[0], 6])
        assert 1 <= nb <= plan.TC_BATCH_ITEMS and hoff % 16 == 0 and hoff >= 8 * nb
        for j, pi in enumerate(grp):
            kind, jb, N, ks0, nbj, last = (int(x) for x in ints[pi, :6])
            assert (jb, ks0, nbj, last) == (j, j, nb, int(j == nblocks - 1)) and N % 16 == 0 and N <= plan.TC_PANEL
            assert N >= (hoff + 2 * nb if j == 0 else 8 * nb)
            assert not np.any(rows[pi][:, :8 * j])                                  # skipped K steps are zero
            assert not np.any(rows[pi][N:])                                         # nothing beyond the MMA's N
        types += [int(ints[grp[0], 8 + 2 * s]) for s in range(nb)]
        T0 = np.concatenate([rows[pi][0:8, :kp] for pi in grp])                     # factor of the batch's first item
        assert np.allclose(np.tril(T0, -1), 0.0)                                    # upper triangular
    assert types == [2] * len(cs.qcs) + [3] * len(cs.socs) + [5] * n_lmi
    if f["lmitc_panels"]:
        rp, n = f["lmi_rp"], cs.n
        assert f["lmitc_panels"] == rp * rp // plan.LMI_TC_PANEL
        ltile = plan.LMI_TC_PANEL * kp
        W = []
        for pi in range(f["lmitc_panels"]):
            o = f["off_lmitc"] + 2 * pi * ltile
            hi, lo = blob[o:o + ltile], blob[o + ltile:o + 2 * ltile]
            assert not np.any(hi.view(np.uint32) & np.uint32(0x1FFF)) and not np.any(lo.view(np.uint32) & np.uint32(0x1FFF))
            W.append(_decode_operand(hi.astype(np.float64) + lo.astype(np.float64), plan.LMI_TC_PANEL, kp))
        W = np.concatenate(W)                                                      # [rp*rp, kp], row e = i*rp + 4q + t
        F32 = blob[f["off_lmi"]:f["off_lmi"] + n * rp * rp].reshape(n, rp * rp).astype(np.float64)
        assert np.abs(W[:, :n].T - F32).max() <= 2.0 ** -21 * np.abs(F32).max()
        assert not np.any(W[:, n:])


# ----------------------------------------------------------------------------- wide plans (n > 32)
@pytest.mark.parametrize("k,m,eta,mu,r_M,eq", [(40, 50, 2, 2, 20, 0), (48, 100, 3, 1, 60, 3), (100, 300, 2, 2, 10, 5),
                                               (65, 0, 1, 1, 5, 0), (33, 7, 0, 0, 0, 0)])
def test_wide_plan_decodes_to_the_oracle(k, m, eta, mu, r_M, eq):
    """n > 32: the WIDE section (transposed row matrix + warp tasks), decoded the way wide_forward_kernel walks it, gives
    the oracle's kappa, binding constraint and y; tasks tile the rows."""
    spec = synthetic.wide_spec(k, m, eta, mu, r_M, eq, seed=k)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    assert f["wide"] == 1 and f["np"] % 4 == 0 and f["np"] >= cs.n and f["tc_panels"] == 0 and f["off_wide"] % 4 == 0
    hdr = p.blob[f["off_wide"]:f["off_wide"] + plan.WIDE_HEADER_WORDS].view(np.int32)
    r_pad, n_tasks, off_tasks = int(hdr[1]), int(hdr[2]), int(hdr[3])
    assert hdr[0] == plan.WIDE_MAGIC and r_pad % plan.WIDE_GROUP_ROWS == 0 and int(hdr[10]) == eta and int(hdr[11]) == mu
    tasks = p.blob[off_tasks:off_tasks + n_tasks * plan.WIDE_TASK_WORDS].view(np.int32).reshape(n_tasks, -1)
    assert int(hdr[14]) == plan.WIDE_VERSION
    # every group of 64 rows is exactly one task; rounds tile the task list; a round fits the kernel's slot budget
    assert sorted(tasks[:, 1].tolist()) == list(range(0, r_pad, plan.WIDE_GROUP_ROWS))
    n_rounds, off_rounds = int(hdr[12]), int(hdr[13])
    rounds = p.blob[off_rounds:off_rounds + 4 * n_rounds].view(np.int32).reshape(n_rounds, 4)
    assert rounds[0, 0] == 0 and rounds[-1, 1] == n_tasks and rounds[0, 2] == 0 and rounds[-1, 3] == eta + mu
    for r in range(n_rounds):
        t0, t1, i0, i1 = rounds[r]
        part = tasks[t0:t1][tasks[t0:t1, 0] == plan.WIDE_FACTOR]
        heads = tasks[t0:t1][tasks[t0:t1, 0] == plan.WIDE_HDR]
        assert i1 - i0 <= plan.WIDE_ROUND_ITEMS and (part[:, 5] < plan.WIDE_SLOTS).all() and (part[:, 3] < i1 - i0).all()
        assert len(heads) == (i1 - i0 + 31) // 32 and sorted(heads[:, 3].tolist()) == list(range(0, i1 - i0, 32))
        assert len(set(part[:, 5].tolist())) == len(part)              # one slot per group
        assert (np.diff(tasks[t0:t1, 2]) >= 0).all()                   # heaviest (fewest skipped columns) first
    lin = tasks[tasks[:, 0] == plan.WIDE_LIN]
    assert (np.diff(lin[:, 1]) > 0).all() and (lin[:, 3] == lin[:, 1]).all()   # linear rows in ascending order
    v, _ = synthetic.sample_inputs(96, cs.n, cs.k, seed_v=k, dtype=torch.float64)
    cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy())
    y, kap, act = plan.evaluate_wide_numpy(p, v.numpy())
    assert np.abs(y - cf["y"]).max() <= 5e-6 * max(1.0, np.abs(cf["y"]).max())
    assert np.abs(kap - cf["kappa"]).max() <= 5e-6 * max(1.0, cf["kappa"].max())
    ok = cf["margin"] > 1e-4
    assert ((act >> 24)[ok] == cf["family"][ok]).all()
    ok &= cf["family"] != 0                       # kappa = 0: no binding constraint, the index means nothing
    assert ((act & 0xFFFFFF)[ok] == cf["index"][ok]).all()
    # the register-layout sections stay empty for a wide plan: the block is the WIDE section + N + the checker's rows
    k4 = (cs.k + 3) // 4 * 4
    viol_words = (m + 2 * eq + 1) * (k4 + 4) + eta * (k4 * k4 + k4 + 4) + mu * (4 + k4 + r_M * (k4 + 4))
    n_words = 0 if f["n_is_identity"] else cs.n * (cs.k + 32) + cs.k * f["np"]
    assert p.blob.size <= cs.n * r_pad + viol_words + n_words + 8 * n_tasks + 8 * (eta + mu) + 2048


def test_narrow_plans_are_not_wide():
    p = plan.build_plan_from_constraints(synthetic.build_constraints(synthetic.config_spec("cfg5")))
    assert p.fields["wide"] == 0 and p.fields["off_wide"] == 0 and p.fields["tc_panels"] == 14


def test_wide_plan_with_more_items_than_a_round_holds():
    """More items than WIDE_ROUND_ITEMS / more groups than WIDE_SLOTS: the items are processed in several rounds."""
    spec = synthetic.wide_spec(34, 10, 40, 40, 3, 0, seed=4)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    hdr = p.blob[p.fields["off_wide"]:p.fields["off_wide"] + plan.WIDE_HEADER_WORDS].view(np.int32)
    assert int(hdr[12]) >= 2
    v, _ = synthetic.sample_inputs(64, cs.n, cs.k, seed_v=2, dtype=torch.float64)
    cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy())
    y, kap, act = plan.evaluate_wide_numpy(p, v.numpy())
    assert np.abs(y - cf["y"]).max() <= 5e-6 * max(1.0, np.abs(cf["y"]).max())


# ----------------------------------------------------------------------------- LMI pruning bound in float32
def _bound_f32(p, u32, tf32=False):
    """The pruning bound exactly as the kernels form it (lqs.cuh lmi_upper_bound), from the PACKED float32 block, with
    float32 dot products accumulated term by term; ``tf32``: operands split hi + lo like the 3xTF32 GEMM of lqs_tc.cuh
    (the lo*lo term dropped)."""
    f = p.fields
    np_, n = f["np"], f["n"]
    tri_words = plan.packed_triangular_words(np_)
    blob = p.blob
    off = f["off_bound"]
    t = blob[off:off + np_]
    T = np.zeros((np_, np_), dtype=np.float32)
    pos = off + np_
    for i in range(np_):
        c0 = (i // 4) * 4
        T[i, c0:] = blob[pos:pos + np_ - c0]
        pos += np_ - c0
    r, margin = blob[off + np_ + tri_words], blob[off + np_ + tri_words + 1]
    assert margin == np.float32(f["lmi_bound_margin"]) and margin > 0
    W = np.concatenate((t[None, :], T), axis=0)                      # [1 + np, np]
    U = np.zeros((u32.shape[0], np_), dtype=np.float32)
    U[:, :n] = u32

    def dots(Wm, Um):
        acc = np.zeros((Um.shape[0], Wm.shape[0]), dtype=np.float32)
        for j in range(np_):                                          # sequential float32 accumulation (fmaf ~ mul+add here)
            acc = (acc + Um[:, j:j + 1] * Wm[None, :, j]).astype(np.float32)
        return acc
    if tf32:
        Wh, Wl = plan.split_tf32(W)
        Uh, Ul = plan.split_tf32(U)
        d = (dots(Wh, Uh) + dots(Wh, Ul) + dots(Wl, Uh)).astype(np.float32)
    else:
        d = dots(W, U)
    mean = (d[:, 0] / r).astype(np.float32)
    ss = np.sum(d[:, 1:] * d[:, 1:], axis=1, dtype=np.float32)
    radius = np.sqrt(((r - np.float32(1)) / r) * ss, dtype=np.float32)
    return (mean + radius + (np.float32(1e-5) * (np.abs(mean) + radius) + margin)).astype(np.float32)


@pytest.mark.parametrize("sign", [1.0, -1.0])
@pytest.mark.parametrize("perturbation", [1e-4, 1e-3, 1e-2, 1.0])
def test_pruning_bound_is_float32_safe(perturbation, sign):
    """on epigraph LMIs (one F_a = +-I, the others small) a bound that forms
    |S|_F^2 - tr(S)^2/r as a difference cancels in float32 and fell BELOW lambda_max in 7.6 % of the cases.  The centred
    form is a sum of squares: the float32 (and 3xTF32) evaluation must stay above the float64 lambda_max for every
    direction, i.e. pruning remains a proof."""
    worst = np.inf
    for seed in range(5):
        for (k, r) in ((8, 32), (5, 12), (32, 32), (3, 7)):
            spec = synthetic.epigraph_lmi_spec(k, r, perturbation, seed=seed, sign=sign)
            p = plan.build_plan_from_constraints(synthetic.build_constraints(spec))
            Fz = p.f64["Fz"]
            rng = np.random.default_rng(100 + seed)
            u = rng.standard_normal((500, k))
            u[:50, 0] = np.abs(u[:50, 0]) * 30.0 * -sign             # directions in which S~ ~ a positive multiple of I
            u32 = (u / np.linalg.norm(u, axis=1, keepdims=True)).astype(np.float32)
            lam = np.linalg.eigvalsh(np.einsum("ba,aij->bij", u32.astype(np.float64), Fz))[:, -1]
            for tf32 in (False, True):
                ub = _bound_f32(p, u32, tf32).astype(np.float64)
                worst = min(worst, float((ub - lam).min()))
                assert (ub >= lam).all(), (perturbation, sign, k, r, tf32, float((ub - lam).min()))
    assert worst >= 0.0


def test_pruning_bound_uncentred_form_would_fail():
    """Negative control for the test above: the round-1 formula (uncentred Gram, dev2 = ss - tr^2/r in float32, relative
    margin 1e-4) does fall below lambda_max on these sets -- the hazard is real and the test can see it."""
    bad = 0
    for seed in range(3):
        spec = synthetic.epigraph_lmi_spec(8, 32, 1e-3, seed=seed)
        p = plan.build_plan_from_constraints(synthetic.build_constraints(spec))
        Fz = p.f64["Fz"]
        tr = np.trace(Fz, axis1=1, axis2=2)
        gram = np.einsum("aij,bij->ab", Fz, Fz)
        T = plan._triangular_factor(gram, 8).astype(np.float32)
        rng = np.random.default_rng(seed)
        u = rng.standard_normal((2000, 8))
        u[:, 0] = -np.abs(u[:, 0]) * 30.0
        u32 = (u / np.linalg.norm(u, axis=1, keepdims=True)).astype(np.float32)
        lam = np.linalg.eigvalsh(np.einsum("ba,aij->bij", u32.astype(np.float64), Fz))[:, -1]
        h0 = (u32 @ tr.astype(np.float32)).astype(np.float32)
        Tu = (u32 @ T.T).astype(np.float32)
        ss = np.sum(Tu * Tu, axis=1, dtype=np.float32)
        mean = h0 / np.float32(32)
        dev2 = np.maximum(ss - h0 * mean, np.float32(0))
        ub = mean + np.sqrt(np.float32(31.0 / 32.0) * dev2)
        ub = ub + np.float32(1e-4) * np.abs(ub)
        bad += int((ub.astype(np.float64) < lam).sum())
    assert bad > 0


def test_big_lmi_tensor_core_operand_decodes_to_the_packed_matrices():
    """Section LMIBT (the B operand of lmi_big_tc.cuh): per (panel of 128 entries, slice of 32 coordinates) a TF32-split
    pair of [128 x 32] tiles in the operand layout [k/4][row/8][row%8][k%4].  Decoded here and compared with section LMIB:
    hi + lo reproduces F~z' to the residual of the split (2^-22 relative), both halves are TF32-representable, and the
    padding (entries beyond lmib_p4, coordinates beyond n) is zero."""
    spec = synthetic.wide_spec(70, 30, 1, 0, 0, 3, seed=4, r=20)       # n = 67: three slices, the last one ragged
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    n, p4, NP, KS = f["n"], f["lmib_p4"], f["lmibt_panels"], f["lmibt_slices"]
    assert NP == -(-p4 // 128) and KS == -(-n // 32) and f["off_lmibt"] > 0
    Fb = p.blob[f["off_lmib"]:f["off_lmib"] + n * p4].reshape(n, p4)
    T = p.blob[f["off_lmibt"]:f["off_lmibt"] + NP * KS * 2 * 4096].reshape(NP, KS, 2, 8, 16, 8, 4)   # [q][s][hl][kc][rg][r8][k4]
    hi = T[:, :, 0].transpose(0, 3, 4, 1, 2, 5).reshape(NP * 128, KS * 32)      # [q, rg, r8][s, kc, k4]
    lo = T[:, :, 1].transpose(0, 3, 4, 1, 2, 5).reshape(NP * 128, KS * 32)
    for part in (hi, lo):
        assert np.all((part.view(np.uint32) & np.uint32(0x1FFF)) == 0)          # 10-bit mantissa: TF32-representable
    full = np.zeros((NP * 128, KS * 32), dtype=np.float64)
    full[:p4, :n] = Fb.T
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - full)
    assert err.max() <= 2.0 ** -21 * np.abs(full).max()
    assert not hi[p4:].any() and not lo[p4:].any() and not hi[:, n:].any() and not lo[:, n:].any()
    # narrow subspaces keep the FP32 GEMM only
    small = plan.build_plan_from_constraints(synthetic.build_constraints(synthetic.random_spec(k=6, r=40, seed=1)))
    assert small.fields["lmibt_panels"] == 0 and small.fields["off_lmibt"] == 0
