"""Deterministic constraint-set generators: the BASELINE.json configs and the reference's canned sets.

Every generator returns a plain ``dict`` of float64 numpy arrays (a "spec")

    {"A1","b1","A2","b2"  (each ndarray or None),
     "qcs":  [(P,q,r), ...], "socs": [(M,s,c,d), ...], "lmi": [F_0..F_k] or None,
     "y0":   strictly interior point [k,1]}

so that the same numbers can be fed to this package (``build_constraints``), to the oracle, and --
in the build container only -- to the unmodified reference (tests/golden/make_golden.py).

* ``config_spec("cfg2".."cfg5")``: the synthetic sets of SURVEY §8d, restated from the random
  generators of the reference's timing sweep (examples/scripts/time_analysis.py:68-69 linear rows,
  :92-98 quadratic, :128-132 SOC, :166-175 LMI) with ``numpy.random.default_rng(seed)``; y0 = 0 is
  strictly interior by construction.
* ``config_spec("cfg1")`` / ``example_spec("readme")``: the README set (readme.md:42-70).
* ``example_spec(0..14)``: the 15 geometries of examples/examples_sets.py:85-200, each with a
  hand-picked strictly interior y0 (the reference finds one with cvxpy, which is absent here).
"""
import numpy as np

from . import constraints

CONFIG_SHAPES = {
    # name: k, m (ineq rows), eta (ellipsoids), mu (cones), r_M, r (LMI size), batch per GPU
    "cfg1": dict(k=3, m=6, eta=0, mu=0, r_M=0, r=0, batch=500),
    "cfg2": dict(k=8, m=64, eta=8, mu=0, r_M=0, r=0, batch=4096),
    "cfg3": dict(k=16, m=128, eta=0, mu=16, r_M=16, r=0, batch=16384),
    "cfg4": dict(k=8, m=0, eta=0, mu=0, r_M=0, r=32, batch=4096),
    "cfg5": dict(k=32, m=256, eta=16, mu=16, r_M=32, r=32, batch=32768),
}


def _col(x):
    return np.asarray(x, dtype=np.float64).reshape(-1, 1)


def random_spec(k, m=0, eta=0, mu=0, r_M=0, r=0, seed=0):
    """Random set around the origin; y0 = 0 is strictly inside every constraint."""
    rng = np.random.default_rng(seed)
    spec = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=None, y0=np.zeros((k, 1)))
    if m > 0:
        spec["A1"] = rng.uniform(-1.0, 1.0, size=(m, k))
        spec["b1"] = rng.uniform(0.1, 1.0, size=(m, 1))
    for _ in range(eta):
        # ellipsoid (y-c)'E(y-c) <= 1 whose centre is rescaled so that the origin sits at level 0.25
        T = rng.uniform(-1.0, 1.0, size=(k, k))
        E = T @ T.T / k + 0.1 * np.eye(k)
        c = rng.uniform(-1.0, 1.0, size=(k, 1))
        c *= 0.5 / np.sqrt((c.T @ E @ c).item())
        spec["qcs"].append((2.0 * E, -2.0 * E @ c, c.T @ E @ c - 1.0))
    for _ in range(mu):
        M = rng.uniform(-1.0, 1.0, size=(r_M, k))
        s = rng.uniform(-1.0, 1.0, size=(r_M, 1))
        c = rng.uniform(-1.0, 1.0, size=(k, 1))
        d = np.array([[np.linalg.norm(s) + 0.5]])
        spec["socs"].append((M, s, c, d))
    if r > 0:
        all_F = []
        for _ in range(k):
            T = rng.uniform(-1.0, 1.0, size=(r, r))
            all_F.append(0.5 * (T + T.T))
        T = rng.uniform(-1.0, 1.0, size=(r, r))
        all_F.append(T @ T.T + 0.5 * np.eye(r))
        spec["lmi"] = all_F
    return spec


def _cube():
    A1 = np.concatenate((np.eye(3), -np.eye(3)), axis=0)
    b1 = _col([1, 1, 1, 0, 0, 0])
    return A1, b1


def _ellipsoid(E, c):
    E, c = np.asarray(E, dtype=np.float64), _col(c)
    return (2.0 * E, -2.0 * E @ c, c.T @ E @ c - 1.0)


def _sphere(radius, dim):
    return _ellipsoid(np.eye(dim) / radius ** 2, np.zeros(dim))


def _paraboloid3d():
    return (np.diag([1.0, 1.0, 0.0]), _col([0, 0, -1]), np.zeros((1, 1)))


def _soc3d():
    return (np.diag([1.0, 1.0, 0.0]), np.zeros((3, 1)), _col([0, 0, 1]), np.zeros((1, 1)))


def _psd_cone3d():
    return [np.array([[1.0, 0.0], [0.0, 0.0]]), np.array([[0.0, 1.0], [1.0, 0.0]]),
            np.array([[0.0, 0.0], [0.0, 1.0]]), np.zeros((2, 2))]


def example_spec(which):
    """The canned sets of the reference (examples/examples_sets.py:85-200) and the README set."""
    s = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=None, y0=None)
    ones_plane = (np.ones((1, 3)), np.ones((1, 1)))
    if which == "readme":  # readme.md:42-70
        s["A1"], s["b1"] = _cube()
        s["A2"], s["b2"] = ones_plane
        s["qcs"] = [(3.125 * np.eye(3), np.zeros((3, 1)), -np.ones((1, 1)))]
        s["socs"] = [_soc3d()]
        s["lmi"] = _psd_cone3d()
        s["y0"] = _col([0.25, 0.05, 0.7])
    elif which == 0:
        s["A1"], s["b1"] = _cube()
        s["A2"], s["b2"] = ones_plane
        s["y0"] = _col([0.3, 0.3, 0.4])
    elif which == 1:
        s["A1"], s["b1"] = _cube()
        s["A2"], s["b2"] = ones_plane
        s["qcs"] = [_sphere(0.8, 3)]
        s["y0"] = _col([0.3, 0.3, 0.4])
    elif which == 2:
        s["qcs"] = [_sphere(2.0, 3)]
        s["y0"] = _col([0.1, -0.2, 0.3])
    elif which == 3:
        s["qcs"] = [_paraboloid3d()]
        s["y0"] = _col([0.1, 0.2, 1.0])
    elif which in (4, 5):
        s["A1"] = np.array([[-1.0, 0.0], [0.0, -1.0], [0.0, 1.0], [0.6, 0.9701]])
        s["b1"] = _col([0, 0, 1, 1.2127])
        if which == 5:
            s["qcs"] = [_sphere(1.25, 2)]
        s["y0"] = _col([0.5, 0.4])
    elif which == 6:
        s["A1"], s["b1"] = _cube()
        s["A2"] = np.array([[1.0, 1.0, 1.0], [-1.0, 1.0, 1.0]])
        s["b2"] = _col([1.0, 0.1])
        s["y0"] = _col([0.45, 0.2, 0.35])
    elif which == 7:
        s["A2"], s["b2"] = ones_plane
        s["y0"] = _col([0.2, 0.3, 0.5])
    elif which == 8:
        s["A1"] = np.array([[0.0, -1.0], [2.0, -4.0], [-2.0, 1.0]])
        s["b1"] = _col([-2.0, 1.0, -5.0])
        s["y0"] = _col([5.0, 3.0])
    elif which == 9:
        s["qcs"] = [_paraboloid3d()]
        s["A2"], s["b2"] = np.array([[1.0, 1.0, 3.0]]), np.ones((1, 1))
        s["y0"] = _col([0.1, 0.0, 0.3])
    elif which == 10:
        s["qcs"] = [_paraboloid3d(), _sphere(2.0, 3)]
        s["y0"] = _col([0.1, 0.2, 1.0])
    elif which == 11:
        s["socs"] = [_soc3d()]
        s["y0"] = _col([0.1, 0.2, 1.0])
    elif which == 12:
        s["lmi"] = _psd_cone3d()
        s["y0"] = _col([1.0, 0.2, 1.0])
    elif which == 13:
        s["A1"], s["b1"] = -np.ones((1, 3)), -np.ones((1, 1))
        s["qcs"] = [_ellipsoid(np.diag([0.1, 1.0, 1.0]), np.zeros(3))]
        s["socs"] = [_soc3d()]
        s["lmi"] = _psd_cone3d()
        s["y0"] = _col([0.5, 0.05, 0.7])
    elif which == 14:
        s["A1"] = np.array([[-1.0, -1.0, -1.0], [-1.0, 2.0, 2.0]])
        s["b1"] = _col([-1.0, 1.0])
        s["qcs"] = [_ellipsoid(np.diag([0.6, 1.0, 1.0]), np.zeros(3))]
        s["y0"] = _col([1.0, 0.2, 0.2])
    else:
        raise Exception("Not implemented yet")
    return s


EXAMPLE_IDS = ["readme"] + list(range(15))


def config_spec(name, seed=0):
    if name == "cfg1":
        s = example_spec(0)
        s["y0"] = _col([1 / 3, 1 / 3, 1 / 3])
        return s
    shp = CONFIG_SHAPES[name]
    return random_spec(shp["k"], shp["m"], shp["eta"], shp["mu"], shp["r_M"], shp["r"], seed=seed)


def wide_spec(k, m=0, eta=0, mu=0, r_M=0, eq=0, seed=0, loosen=3.0, r=0):
    """A set with more than 32 dimensions (the wide.cuh kernels): random_spec with the rows loosened so that every
    family binds for some samples, plus ``eq`` equality rows through the origin (y0 = 0 stays interior, n = k - eq)."""
    spec = random_spec(k=k, m=m, eta=eta, mu=mu, r_M=r_M, r=r, seed=seed)
    if spec["b1"] is not None:
        spec["b1"] = spec["b1"] * loosen
    if eq:
        rng = np.random.default_rng(10_000 + seed)
        spec["A2"], spec["b2"] = rng.uniform(-1.0, 1.0, size=(eq, k)), np.zeros((eq, 1))
    return spec


BIG_LMI_GOLDEN = ("big_r40", "big_wide_r12", "big_r96_lmi_only")


def big_lmi_spec(name):
    """The sets behind the golden files tests/golden/big_*.npz (LMIs beyond the register-resident kernels)."""
    if name == "big_r40":
        spec = random_spec(k=6, m=20, eta=1, mu=1, r_M=5, r=40, seed=41)
        spec["b1"] = spec["b1"] * 3.0
        return spec
    if name == "big_wide_r12":
        return wide_spec(40, 60, 1, 1, 10, 0, seed=42, r=12)
    if name == "big_r96_lmi_only":
        return random_spec(k=4, r=96, seed=43)
    raise KeyError(name)


def epigraph_lmi_spec(k, r, perturbation, delta=None, seed=0, sign=1.0):
    """The epigraph form ``t I - A(y) >= 0`` (t = y_0), the commonest LMI there is: F_0 = sign * I, F_a = perturbation *
    (a random symmetric matrix) for a >= 1, constant term I, y0 = 0.  S~(u) is then a multiple of the identity plus a
    small perturbation -- the case in which a Frobenius-norm bound written as a DIFFERENCE cancels in float32.
    ``delta`` (a list) adds one linear row per entry, -sign * (1 - d) y_0 <= 1, whose kappa lands within d of the LMI's
    kappa for the directions in which the LMI can bind: the competing constraint that a wrong pruning decision needs."""
    rng = np.random.default_rng(seed)
    all_F = [sign * np.eye(r)]
    for _ in range(k - 1):
        T = rng.uniform(-1.0, 1.0, size=(r, r))
        all_F.append(perturbation * 0.5 * (T + T.T))
    all_F.append(np.eye(r))
    spec = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=all_F, y0=np.zeros((k, 1)))
    if delta:
        A1 = np.zeros((len(delta), k))
        A1[:, 0] = [-sign * (1.0 - d) for d in delta]
        spec["A1"], spec["b1"] = A1, np.ones((len(delta), 1))
    return spec


def build_constraints(spec, module=constraints, y0="spec", do_preprocessing_linear=False):
    """Instantiate ``module``'s classes (this package's, or the reference's) from a spec."""
    lc = None
    if spec["A1"] is not None or spec["A2"] is not None:
        lc = module.LinearConstraint(spec["A1"], spec["b1"], spec["A2"], spec["b2"])
    qcs = [module.ConvexQuadraticConstraint(P, q, r, do_checks_P=False) for (P, q, r) in spec["qcs"]]
    socs = [module.SOCConstraint(M, s, c, d) for (M, s, c, d) in spec["socs"]]
    lmic = module.LMIConstraint(list(spec["lmi"])) if spec["lmi"] is not None else None
    if isinstance(y0, str):
        y0 = spec["y0"]
    return module.ConvexConstraints(lc=lc, qcs=qcs, socs=socs, lmic=lmic, y0=y0,
                                    do_preprocessing_linear=do_preprocessing_linear)


def sample_inputs(batch, n, k, seed_v=1, seed_g=7, dtype=None, scale=2.0):
    """Layer inputs v ~ U(-scale, scale)^n and loss gradients g_y ~ N(0,1)^k, as torch CPU tensors."""
    import torch
    dtype = dtype or torch.float32
    gv = torch.Generator().manual_seed(seed_v)
    gg = torch.Generator().manual_seed(seed_g)
    v = (torch.rand(batch, n, generator=gv, dtype=torch.float64) * 2.0 - 1.0) * scale
    g = torch.randn(batch, k, generator=gg, dtype=torch.float64)
    return v.to(dtype), g.to(dtype)
