"""CPU tests of the checker itself: the clean-room restatement (oracle/pnfft_oracle.c) and, where it was built, the
compiled reference (oracle/_ref) against the golden vectors of tests/golden/ (generated from the unmodified reference
by tools/make_golden.py).  Tolerances: values 1e-13 (double) / 1e-5 (float) rel-l2; integer outputs identical."""
import glob
import os

import numpy as np
import pytest

from oracle import checker, refdrv
from tests.util import fixture_kwargs, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "t_*.npz")))


def _impls(single):
    out = [("port", checker.port(single))]
    if refdrv.available(single):
        out.append(("reference", refdrv.get(single)))
    return out


def test_golden_present():
    assert len(CASES) >= 69
    assert os.path.exists(os.path.join(GOLD, "layouts.npz")) and os.path.exists(os.path.join(GOLD, "node_index.npz"))


@pytest.mark.parametrize("case", CASES)
def test_transform_golden(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    N, n, x_max, acc = fixture_kwargs(g)
    m, flags = int(g["m"]), int(g["flags"])
    kw = dict(n=n, m=m, pnfft_flags=flags, c2r=c2r, x_max=x_max)
    for name, impl in _impls(single):
        if acc:
            t = impl.trafo(N, g["x"], g["f_hat"], f=g["f0"], grad_f=g["grad_f0"], compute_flags=3 | 16, **kw)
            a = impl.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], f_hat=g["f_hat0"], compute_flags=3 | 16, **kw)
        else:
            t = impl.trafo(N, g["x"], g["f_hat"], compute_flags=3, **kw)
            a = impl.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=3, **kw)
        assert rel_l2(t["f"], g["out_f"]) <= tol, name
        # float sinc-power gradient: cot(w) - 1/w cancels catastrophically near w = 0 (reference :1897-1917), the
        # float reference itself is only good to ~1e-4 there
        gtol = 1e-4 if (single and "sinc_power" in case) else tol
        assert rel_l2(t["grad_f"], g["out_grad_f"]) <= gtol, name
        assert rel_l2(a["f_hat"], g["out_f_hat"]) <= gtol, name   # the adjoint spreads grad_f with the same dpsi
        psi, dpsi = impl.probe_tensor(g["x"][:16], N, n=n, m=m, x_max=x_max, pnfft_flags=flags & ~(1 << 8))
        assert np.abs(psi - g["psi"]).max() <= tol * max(1.0, np.abs(g["psi"]).max()), name
        assert np.abs(dpsi - g["dpsi"]).max() <= 10 * gtol * max(1.0, np.abs(g["dpsi"]).max()), name


def test_layout_golden():
    """Block decomposition, pruned FFT output size and node borders over 1x1 .. 2x4 meshes, even / ragged / truncated-torus
    sizes: integers identical, borders bit-identical (reference kernel/ndft-parallel.c:734-822, api/api-guru.c:195-220)."""
    L = np.load(os.path.join(GOLD, "layouts.npz"))
    keys = sorted(k[:-4] for k in L.files if k.endswith("_cfg"))
    assert len(keys) == 24
    for key in keys:
        cfg = L[key + "_cfg"]
        N, n, m, c2r, mesh = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6]), bool(cfg[7]), (int(cfg[8]), int(cfg[9]))
        for name, impl in _impls(False):
            r = impl.layout(N, n, m=m, np_mesh=mesh, x_max=tuple(L[key + "_xmax"]), c2r=c2r)
            for k in ("local_N", "local_N_start", "local_no", "local_no_start"):
                assert np.array_equal(r[k], L[key + "_" + k]), (name, key, k)
            assert tuple(r["no"]) == tuple(L[key + "_no"]), (name, key)
            assert np.array_equal(r["lo"], L[key + "_lo"]) and np.array_equal(r["up"], L[key + "_up"]), (name, key)


def test_layout_golden_transposed_interlaced():
    """Round 2: PNFFT_TRANSPOSED_F_HAT blocks (k1 over mesh dim 0, k2 over mesh dim 1, k0 whole) and PNFFT_INTERLACED on a
    truncated torus, 1x1 .. 2x4 meshes: integers identical, borders bit-identical."""
    L = np.load(os.path.join(GOLD, "layouts_r2.npz"))
    keys = sorted(k[:-4] for k in L.files if k.endswith("_cfg"))
    assert len(keys) == 24
    for key in keys:
        cfg = L[key + "_cfg"]
        N, n, m, c2r, mesh, fl = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6]), bool(cfg[7]), (int(cfg[8]), int(cfg[9])), int(cfg[10])
        for name, impl in _impls(False):
            r = impl.layout(N, n, m=m, np_mesh=mesh, x_max=tuple(L[key + "_xmax"]), c2r=c2r, pnfft_flags=fl)
            for k in ("local_N", "local_N_start", "local_no", "local_no_start"):
                assert np.array_equal(r[k], L[key + "_" + k]), (name, key, k)
            assert np.array_equal(r["lo"], L[key + "_lo"]) and np.array_equal(r["up"], L[key + "_up"]), (name, key)


def test_node_index_golden():
    """node -> rank ownership, local grid index u_j and plain index m0, and the sort key: bit-exact."""
    g = np.load(os.path.join(GOLD, "node_index.npz"))
    N, m, x = tuple(int(v) for v in g["N"]), int(g["m"]), g["x"]
    po = checker.port(False)
    for mesh in [(1, 1), (2, 2), (2, 4)]:
        owner, idx = po.node_index(x, N, m=m, np_mesh=mesh)
        assert np.array_equal(owner, g["owner_%dx%d" % mesh])
        assert np.array_equal(idx, g["index_%dx%d" % mesh])
    keys = po.probe_sort_keys(x, N, m=m)
    assert np.array_equal(np.sort(keys, kind="stable"), g["sort_keys"])
    assert np.array_equal(np.argsort(keys, kind="stable"), g["sort_perm"])   # stable LSD radix == stable argsort


def test_port_matches_reference_on_new_inputs():
    """Beyond the fixtures: fresh seeded inputs, when the compiled reference is available in this container."""
    if not refdrv.available(False):
        pytest.skip("oracle/_ref not built here")
    from tests.util import make_inputs
    ref, po = refdrv.get(False), checker.port(False)
    N = (12, 16, 8)
    for c2r in (False, True):
        x, fh, f, g = make_inputs(N, 150, 11, c2r=c2r)
        for flags in (0, 1 << 14, (1 << 13) | 2):
            rt = ref.trafo(N, x, fh, pnfft_flags=flags, compute_flags=3, c2r=c2r, m=4)
            pt = po.trafo(N, x, fh, pnfft_flags=flags, compute_flags=3, c2r=c2r, m=4)
            assert rel_l2(pt["f"], rt["f"]) <= 1e-13 and rel_l2(pt["grad_f"], rt["grad_f"]) <= 1e-13


def test_adjointness():
    """<A f_hat, f> == <f_hat, A^H f>: a size-independent property of the trafo/adj pair (no oracle needed)."""
    from tests.util import make_inputs
    po = checker.port(False)
    N = (8, 8, 8)
    x, fh, f, _ = make_inputs(N, 100, 5)
    t = po.trafo(N, x, fh, compute_flags=1, m=4)["f"]
    a = po.adj(N, x, f=f, compute_flags=1, m=4)["f_hat"]
    lhs, rhs = np.vdot(f, t), np.vdot(a, fh)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)


@pytest.mark.parametrize("m", [4, 6])
def test_fast_kaiser_bessel_taps_match_reference(m):
    """The node-table kernel evaluates the double-precision Kaiser-Bessel taps with one exponential and one division per
    tap (csrc/window.h kb_tap_fast; reference formulas kernel/ndft-parallel.c:2241-2269).  The same arithmetic, run on the
    host through pnfft_b200_kb_taps_host, against the reference's pre_psi / pre_dpsi tensors (:1782-1797, :1800-1953)
    for random nodes, nodes on grid lines and nodes an ulp away from them."""
    import ctypes as C
    from pnfft_b200 import api as A
    N, n = (16, 16, 16), (32, 32, 32)
    rng = np.random.default_rng(11)
    x = rng.uniform(-0.5, 0.5, (4000, 3))
    x[:6] = np.array([[-0.5, 0.25, 0.0], [0.0, 0.0, 0.0], [0.125, -0.125, 0.375], [-0.25, 0.46875, -0.5],
                      [0.03125, 0.0625, 0.09375], [0.4, 0.0, -0.3]])
    x[6:12] = np.nextafter(x[:6], 1.0)
    x[12:18] = np.nextafter(x[:6], -1.0)
    x[18:24] = x[:6] + 1e-5          # small sinh / sin arguments: the library-call fallback
    x = np.clip(x, -0.5, np.nextafter(0.5, 0.0))
    impl = checker.get(False)
    psi_r, dpsi_r = impl.probe_tensor(x, N, m=m, pnfft_flags=0)
    fn = A.lib().pnfft_b200_kb_taps_host
    fn.restype = None
    P = C.POINTER(C.c_double)
    fn.argtypes = [P, C.c_ssize_t, C.POINTER(C.c_ssize_t), P, C.c_int, P, P]
    b = np.array([np.pi * (2.0 - N[t] / n[t]) for t in range(3)])
    psi, dpsi = np.zeros_like(psi_r), np.zeros_like(dpsi_r)
    xs = np.ascontiguousarray(x)
    fn(xs.ctypes.data_as(P), len(x), (C.c_ssize_t * 3)(*n), b.ctypes.data_as(P), m, psi.ctypes.data_as(P), dpsi.ctypes.data_as(P))
    assert np.abs(psi - psi_r).max() <= 2e-14 * np.abs(psi_r).max()
    assert rel_l2(psi, psi_r) <= 1e-15
    # one ulp off a grid line the reference's own derivative formula cancels catastrophically
    far = np.ones(len(x), bool); far[6:18] = False
    assert np.abs(dpsi[far] - dpsi_r[far]).max() <= 2e-13 * np.abs(dpsi_r).max()
    assert rel_l2(dpsi[far], dpsi_r[far]) <= 1e-14


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("window", ["kaiser_bessel", "gaussian", "gaussian_t", "bspline", "sinc_power", "bessel_i0"])
def test_window_fourier_coefficients_match_reference(window, single):
    """phi_hat / 1/phi_hat as the product builds its D tables (Core::upload_window_tables, pnfft_phi_hat, pnfft_inv_phi_hat),
    run on the host through pnfft_b200_phi_hat_host, against the compiled reference's pnfft_phi_hat / pnfft_inv_phi_hat
    (kernel/matrix_D.c:30-132, 191-225) for every window -- PNFFT_WINDOW_GAUSSIAN_T included, whose coefficients carry
    Re erf(m / sqrt(b) + i pi k sqrt(b) / n) (libcerf in the reference, an own quadrature here) -- on even, ragged and
    strongly oversampled sizes."""
    import ctypes as C
    from pnfft_b200 import api as A
    if not refdrv.available(single):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = refdrv.get(single)
    flags = {"kaiser_bessel": 0, "gaussian": 1 << 13, "gaussian_t": (1 << 13) | (1 << 17), "bspline": 1 << 14,
             "sinc_power": 1 << 15, "bessel_i0": 1 << 16}[window]
    fn = getattr(A.lib(), ("pnfftf_" if single else "pnfft_") + "b200_phi_hat_host")
    fn.restype = None
    fn.argtypes = [C.c_uint, C.c_ssize_t, C.c_ssize_t, C.c_float if single else C.c_double, C.c_int, C.c_void_p, C.c_ssize_t,
                   C.c_int, C.c_void_p]
    tol = 1e-5 if single else 1e-13
    for N, n, m in [((16, 16, 16), (32, 32, 32), 6), ((8, 12, 10), (16, 24, 20), 5), ((24, 32, 20), (48, 64, 40), 4),
                    ((16, 16, 16), (20, 24, 64), 8)]:
        for dim in range(3):
            k = np.arange(-N[dim] // 2, N[dim] // 2).astype(np.int64)
            for inverse, name in ((1, "inv_phi_hat"), (0, "phi_hat")):
                want = ref.probe(name, dim, k, N, n=n, m=m, pnfft_flags=flags)
                got = np.zeros(len(k), np.float32 if single else np.float64)
                fn(flags, N[dim], n[dim], 0.0, m, k.ctypes.data, len(k), inverse, got.ctypes.data)
                assert np.all(np.abs(got - want) <= tol * np.abs(want)), (window, N[dim], n[dim], m, name)
    if window == "gaussian_t" and not single:   # the truncation is visible: not the plain Gaussian's coefficients
        k = np.arange(-8, 8).astype(np.int64)
        plain = ref.probe("inv_phi_hat", 0, k, (16, 16, 16), m=6, pnfft_flags=1 << 13)
        trunc = ref.probe("inv_phi_hat", 0, k, (16, 16, 16), m=6, pnfft_flags=flags)
        assert np.abs(trunc / plain - 1).max() > 1e-8


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("window", ["kaiser_bessel", "gaussian", "bspline", "sinc_power", "bessel_i0"])
def test_window_point_values_match_reference(window, single):
    """pnfft_psi / pnfft_dpsi / pnfft_ddpsi (reference kernel/ndft-parallel.c:2288-2336) as the product evaluates them
    (csrc/api.cuh window_at -> window.h window_tap / window_ddtap, the formulas of the kernels, the Hessian path and the
    interpolation tables), on the host through pnfft_b200_psi_host, against the compiled reference at random offsets inside
    the support, at 0, near the edge of the support and on grid lines."""
    import ctypes as C
    from pnfft_b200 import api as A
    if not refdrv.available(single):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = refdrv.get(single)
    flags = {"kaiser_bessel": 0, "gaussian": 1 << 13, "bspline": 1 << 14, "sinc_power": 1 << 15, "bessel_i0": 1 << 16}[window]
    dt = np.float32 if single else np.float64
    fn = getattr(A.lib(), ("pnfftf_" if single else "pnfft_") + "b200_psi_host")
    fn.restype = None
    fn.argtypes = [C.c_uint, C.c_ssize_t, C.c_ssize_t, C.c_float if single else C.c_double, C.c_int, C.c_int, C.c_void_p,
                   C.c_ssize_t, C.c_void_p]
    rng = np.random.default_rng(3)
    for N, n, m in [((16, 16, 16), (32, 32, 32), 6), ((8, 12, 10), (16, 24, 20), 5), ((24, 32, 20), (48, 64, 40), 4)]:
        dim = 1
        x = (rng.uniform(-1, 1, 400) * (m / n[dim])).astype(dt)
        x[:5] = np.array([0, m / n[dim] * 0.999, -m / n[dim] * 0.5, 1.0 / n[dim], -2.0 / n[dim]], dt)
        for which, name in ((0, "psi"), (1, "dpsi"), (2, "ddpsi")):
            # sinc-power derivatives: cot(w) - 1/w (and its square) cancel near w = 0 in the reference's own formula
            # (:1897-1917, :2060-2075): the float reference is only good to ~1e-4 in dpsi and has no correct digit in ddpsi
            # there, the double one to ~1e-10 in ddpsi
            if window == "sinc_power" and single and which == 2:
                continue
            tol = (1e-5 if single else 1e-13) if not (window == "sinc_power" and which) else (2e-4 if single else 1e-9)
            want = ref.probe(name, dim, x, N, n=n, m=m, pnfft_flags=flags)
            got = np.zeros(len(x), dt)
            fn(flags, N[dim], n[dim], 0.0, m, which, x.ctypes.data, len(x), got.ctypes.data)
            assert np.abs(got - want).max() <= tol * np.abs(want).max(), (window, m, name)


HCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "[hi]_*.npz")))


def test_hessian_and_intpol_golden_present():
    assert len([c for c in HCASES if c.startswith("h_")]) == 22 and len([c for c in HCASES if c.startswith("i_")]) == 12


@pytest.mark.parametrize("case", HCASES)
def test_hessian_intpol_fixtures_reproduce(case):
    """Round-2 fixtures (PNFFT_COMPUTE_HESSIAN_F, PNFFT_PRE_*_PSI interpolation): where the compiled reference is available
    it reproduces them bit for bit (the fixtures are its own output; this pins the generator script and the driver's
    hessian_f plumbing).  The clean-room port restates both -- analytic second derivatives of all windows, the ik variant, the
    interpolated window of PNFFT_PRE_{CONST,LIN,CUB}_PSI with the reference's flag promotion -- and is pinned by them here."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single = bool(g["single"])
    po = checker.port(single)
    Np = tuple(int(v) for v in g["N"])
    kwp = dict(m=int(g["m"]), pnfft_flags=int(g["flags"]), c2r=bool(g["c2r"]))
    t = po.trafo(Np, g["x"], g["f_hat"], compute_flags=7, **kwp)
    tol = 1e-5 if single else 1e-13
    assert rel_l2(t["f"], g["out_f"]) <= tol and rel_l2(t["grad_f"], g["out_grad_f"]) <= (1e-4 if (single and "sinc_power" in case) else tol)
    assert rel_l2(t["hessian_f"], g["out_hessian_f"]) <= (2e-4 if single else 1e-13)
    if case.startswith("i_"):      # the interpolation fixtures also hold the adjoint
        a = po.adj(Np, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=3, **kwp)
        assert rel_l2(a["f_hat"], g["out_f_hat"]) <= tol
    if not refdrv.available(single):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = refdrv.get(single)
    t = ref.trafo(tuple(g["N"]), g["x"], g["f_hat"], m=int(g["m"]), pnfft_flags=int(g["flags"]), compute_flags=7, c2r=bool(g["c2r"]))
    assert np.array_equal(t["f"], g["out_f"]) and np.array_equal(t["grad_f"], g["out_grad_f"])
    assert np.array_equal(t["hessian_f"], g["out_hessian_f"])
    # the Hessian of the trafo is symmetric in exact arithmetic: a direct NDFT on a few nodes pins the component order
    if case.startswith("h_kaiser_bessel_ik_c2c"):
        N = tuple(int(v) for v in g["N"])
        k = [np.arange(-n // 2, n // 2) for n in N]
        K = np.meshgrid(*k, indexing="ij")
        x = g["x"][:8].astype(np.float64)
        ph = np.exp(-2j * np.pi * sum(x[:, t, None, None, None] * K[t] for t in range(3)))
        pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for c, (a, b) in enumerate(pairs):
            direct = (-4 * np.pi ** 2 * K[a] * K[b] * ph * g["f_hat"]).sum((1, 2, 3))
            assert rel_l2(t["hessian_f"][:8, c], direct) <= (1e-3 if single else 1e-9)


BCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "b_*.npz")))


def test_set_b_golden_present():
    assert len(BCASES) == 10


@pytest.mark.parametrize("case", BCASES)
def test_set_b_fixtures_reproduce(case):
    """pnfft_set_b fixtures (tools/make_golden.py --set-b): the compiled reference, where it is available, reproduces them
    bit for bit -- this pins the generator and the driver's b plumbing -- and the default shape gives other values."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single = bool(g["single"])
    kw = dict(m=int(g["m"]), pnfft_flags=int(g["flags"]), compute_flags=3, c2r=bool(g["c2r"]))
    N = tuple(int(v) for v in g["N"])
    po = checker.port(single)          # the clean-room port with the same shape parameters
    tol = 1e-5 if single else 1e-13
    t = po.trafo(N, g["x"], g["f_hat"], b=tuple(g["b"]), **kw)
    a = po.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], b=tuple(g["b"]), **kw)
    assert rel_l2(t["f"], g["out_f"]) <= tol and rel_l2(t["grad_f"], g["out_grad_f"]) <= tol and rel_l2(a["f_hat"], g["out_f_hat"]) <= tol
    if not refdrv.available(single):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = refdrv.get(single)
    t = ref.trafo(N, g["x"], g["f_hat"], b=tuple(g["b"]), **kw)
    a = ref.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], b=tuple(g["b"]), **kw)
    assert np.array_equal(t["f"], g["out_f"]) and np.array_equal(t["grad_f"], g["out_grad_f"])
    assert np.array_equal(a["f_hat"], g["out_f_hat"])
    if not single:      # (in float both shapes are accurate to rounding: nothing to tell apart in f)
        t0 = ref.trafo(N, g["x"], g["f_hat"], **kw)
        assert rel_l2(t0["f"], g["out_f"]) > 1e-10


CCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "c_*.npz")))


def test_combination_golden_present():
    assert len(CCASES) == 28


@pytest.mark.parametrize("case", CCASES)
def test_combination_fixtures_reproduce(case):
    """Combination fixtures (tools/make_golden.py --combinations): the compiled reference, where it is available, reproduces
    them bit for bit; the clean-room port agrees with them too (it evaluates the windows on the fly where the reference read its PRE_PSI tables)."""
    from tests.util import fixture_kwargs
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    N, n, x_max, _ = fixture_kwargs(g)
    cf, pre, hess = int(g["cf"]), int(g["pre"]), "out_hessian_f" in g.files
    kw = dict(n=n, m=int(g["m"]), pnfft_flags=int(g["flags"]), c2r=c2r, x_max=x_max)
    if refdrv.available(single):
        ref = refdrv.get(single)
        t = ref.trafo(N, g["x"], g["f_hat"], compute_flags=cf | (4 if hess else 0), precompute_flags=pre, **kw)
        a = ref.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=cf, precompute_flags=pre, **kw)
        assert np.array_equal(t["f"], g["out_f"]) and np.array_equal(a["f_hat"], g["out_f_hat"])
        if hess:
            assert np.array_equal(t["hessian_f"], g["out_hessian_f"])
    tol = 1e-5 if single else 1e-13
    po = checker.port(single)
    t = po.trafo(N, g["x"], g["f_hat"], compute_flags=cf | (4 if hess else 0), **kw)
    a = po.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=cf, **kw)
    assert rel_l2(t["f"], g["out_f"]) <= tol
    if hess:
        assert rel_l2(t["hessian_f"], g["out_hessian_f"]) <= tol
    if cf & 2:
        assert rel_l2(t["grad_f"], g["out_grad_f"]) <= tol
    assert rel_l2(a["f_hat"], g["out_f_hat"]) <= tol


DCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "d_*.npz")))


def test_direct_golden_present():
    assert len(DCASES) == 9


@pytest.mark.parametrize("case", DCASES)
def test_direct_fixtures(case):
    """PNFFT_COMPUTE_DIRECT fixtures (tools/make_golden.py --direct).  (i) The compiled reference, where available, reproduces
    them bit for bit.  (ii) The algebra csrc/direct.cuh implements, restated in numpy, agrees with them: plain sums for
    c2c; for c2r the half spectrum k2 in [-N2/2, 0] with weight 2, the origin once, the self-conjugate / redundant
    coefficients of the planes k2 = 0, -N2/2 left out by the reference's rule applied to the loop variables in MEMORY
    order (kernel/ndft-parallel.c:356-374), outputs overwritten by the trafo, f_hat added to by the adjoint -- and the c2r
    gradient with the sign of the derivative, which is MINUS the reference's (:517-519, :597)."""
    from tests.util import fixture_kwargs
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    N, n, x_max, acc = fixture_kwargs(g)
    flags = int(g["flags"])
    if refdrv.available(single):
        ref = refdrv.get(single)
        kw = dict(n=n, m=int(g["m"]), pnfft_flags=flags, c2r=c2r, x_max=x_max)
        if acc:
            t = ref.trafo(N, g["x"], g["f_hat"], f=g["f0"], grad_f=g["grad_f0"], compute_flags=7 | 8 | 16, **kw)
            a = ref.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], f_hat=g["f_hat0"], compute_flags=3 | 8 | 16, **kw)
        else:
            t = ref.trafo(N, g["x"], g["f_hat"], compute_flags=7 | 8, **kw)
            a = ref.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=3 | 8, **kw)
        assert np.array_equal(t["f"], g["out_f"]) and np.array_equal(t["grad_f"], g["out_grad_f"])
        assert np.array_equal(t["hessian_f"], g["out_hessian_f"]) and np.array_equal(a["f_hat"], g["out_f_hat"])
    # (iii) the clean-room port (oracle/pnfft_oracle.c: direct_trafo / direct_adj) -- the same algebra in C
    po, sg, tolp = checker.port(single), (-1 if c2r else 1), (1e-5 if single else 1e-13)
    kwp = dict(n=n, m=int(g["m"]), pnfft_flags=flags, c2r=c2r, x_max=x_max)
    if acc:
        tp = po.trafo(N, g["x"], g["f_hat"], f=g["f0"], grad_f=g["grad_f0"], compute_flags=7 | 8 | 16, **kwp)
        ap = po.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], f_hat=g["f_hat0"], compute_flags=3 | 8 | 16, **kwp)
    else:
        tp = po.trafo(N, g["x"], g["f_hat"], compute_flags=7 | 8, **kwp)
        ap = po.adj(N, g["x"], f=g["f"], grad_f=g["grad_f"], compute_flags=3 | 8, **kwp)
    assert rel_l2(tp["f"], g["out_f"]) <= tolp and rel_l2(tp["grad_f"], sg * g["out_grad_f"]) <= tolp
    assert rel_l2(tp["hessian_f"], g["out_hessian_f"]) <= tolp and rel_l2(ap["f_hat"], g["out_f_hat"]) <= tolp
    x, fh = g["x"].astype(np.float64), g["f_hat"].astype(np.complex128)
    k = [np.arange(-v // 2, v // 2) for v in N]
    if c2r:
        k[2] = np.arange(-N[2] // 2, 1)
    K = np.meshgrid(*k, indexing="ij")
    v = np.exp(-2j * np.pi * sum(x[:, t, None, None, None] * K[t] for t in range(3))) * fh
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    if c2r:
        ax = (1, 2, 0) if flags & (1 << 11) else (0, 1, 2)
        L, Nn = [K[a] for a in ax], [N[a] for a in ax]
        e = [(L[i] == 0) | (L[i] == -(Nn[i] // 2)) for i in range(3)]
        w = np.full(K[0].shape, 2.0)
        w[e[0] & e[1] & e[2]] = 0
        w[e[2] & (L[1] > 0)] = 0
        w[e[2] & e[1] & (L[0] > 0)] = 0
        w[(K[0] == 0) & (K[1] == 0) & (K[2] == 0)] = 1
        v = v * w
        f = v.sum((1, 2, 3)).real
        grad = np.stack([2 * np.pi * (K[t] * v).sum((1, 2, 3)).imag for t in range(3)], 1)
        hess = np.stack([-4 * np.pi ** 2 * (K[a] * K[b] * v).sum((1, 2, 3)).real for a, b in pairs], 1)
    else:
        f = v.sum((1, 2, 3))
        grad = np.stack([-2j * np.pi * (K[t] * v).sum((1, 2, 3)) for t in range(3)], 1)
        hess = np.stack([-4 * np.pi ** 2 * (K[a] * K[b] * v).sum((1, 2, 3)) for a, b in pairs], 1)
    ff, gg = g["f"].astype(np.complex128), g["grad_f"].astype(np.complex128)
    wj = ff[:, None, None, None] + 2j * np.pi * sum(gg[:, t, None, None, None] * K[t] for t in range(3))
    fa = (wj * np.exp(2j * np.pi * sum(x[:, t, None, None, None] * K[t] for t in range(3)))).sum(0)
    if acc:
        fa = fa + g["f_hat0"]
    tol = 1e-5 if single else 1e-13
    assert rel_l2(g["out_f"], f) <= tol
    assert rel_l2(g["out_grad_f"], (-1 if c2r else 1) * grad) <= tol
    assert rel_l2(g["out_hessian_f"], hess) <= tol
    assert rel_l2(g["out_f_hat"], fa) <= tol
