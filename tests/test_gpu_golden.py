"""GPU parity against the committed golden vectors (tests/golden/, generated from the unmodified reference by
tools/make_golden.py) and size-independent properties at sizes the CPU oracle cannot reach.  Everything goes through the
C ABI of libpnfft_b200.so (pnfft_b200.api is a ctypes mirror of include/pnfft.h).

Bars (BASELINE.json north_star): rel-l2 <= 1e-13 in double, <= 1e-5 in float.
"""
import glob
import os

import numpy as np
import pytest

from pnfft_b200 import api as A
from tests.util import Run1, fixture_kwargs, make_inputs, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "t_*.npz")))
F, G = A.COMPUTE_F, A.COMPUTE_GRAD_F


@pytest.mark.parametrize("variant", [0, 8, 1])   # default (z-march v2 where m allows) / z-march v1 / generic kernels
@pytest.mark.parametrize("case", CASES)
def test_golden_fixture(case, variant):
    """trafo(F|GRAD_F) and adj(F|GRAD_F) of every fixture: c2c and c2r, double and float, AD and ik gradients, all windows;
    round 2: PNFFT_INTERLACED, PNFFT_TRANSPOSED_F_HAT, truncated torus (x_max < 0.5), PNFFT_COMPUTE_ACCUMULATED."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    gtol = 1e-4 if (single and "sinc_power" in case) else tol   # see tests/test_oracle.py: the float reference itself
    N, n, x_max, acc = fixture_kwargs(g)
    m, flags = int(g["m"]), int(g["flags"])
    run = Run1(N, g["x"], n=n, m=m, flags=flags, c2r=c2r, single=single, x_max=x_max, variant=variant)
    if acc:
        f, gr = run.trafo(g["f_hat"], F | G | A.COMPUTE_ACCUMULATED, f0=g["f0"], g0=g["grad_f0"])
        fh = run.adj(g["f"], g["grad_f"], F | G | A.COMPUTE_ACCUMULATED, f_hat0=g["f_hat0"])
    else:
        f, gr = run.trafo(g["f_hat"], F | G)
        fh = run.adj(g["f"], g["grad_f"], F | G)
    run.close()
    assert rel_l2(f, g["out_f"]) <= tol
    assert rel_l2(gr, g["out_grad_f"]) <= gtol
    assert rel_l2(fh, g["out_f_hat"]) <= gtol


@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("N", [(16, 16, 16), (32, 8, 64)])
def test_own_fft_kernels(N, single, monkeypatch):
    """PNFFT_B200_OWN_FFT=1: the x / y / z passes on the library's own radix-8/4/2 shared-memory kernels (fftown.cuh, the z
    pass fused with the embed into / extract from the padded grid) give the transform cuFFT gives."""
    M = 3000
    x, fh, f, g = make_inputs(N, M, 33, c2r=False, single=single)
    res = []
    for own in ("0", "1"):
        monkeypatch.setenv("PNFFT_B200_OWN_FFT", own)
        run = Run1(N, x, m=4, single=single)
        fo, go = run.trafo(fh, F | G)
        fho = run.adj(f, g, F | G)
        run.close()
        res.append((fo, go, fho))
    tol = 2e-5 if single else 1e-13
    for a, b in zip(res[0], res[1]):
        assert rel_l2(a, b) <= tol


HCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "h_*.npz")))


@pytest.mark.parametrize("case", HCASES)
def test_hessian_golden(case):
    """PNFFT_COMPUTE_HESSIAN_F (reference api/api-basic.c:148-166, kernel/assign.c:881-1027): analytic second derivatives
    of every window and ik differentiation, c2c / c2r, against fixtures from the unmodified reference; f and grad_f of the
    same call must not change."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    # float: the analytic second derivative of sinc-power / Kaiser-Bessel amplifies the rounding of psi, dpsi (see the
    # gradient note in tests/test_oracle.py); the double fixtures pin the formulas
    htol = 2e-4 if single else 1e-13
    if "sinc_power" in case and "_ad_" in case and not single:
        # the reference's own formula (kernel/ndft-parallel.c:2048-2052) takes 1/y^2 - 1 - cot(y)^2 with y = pi (u - s) / b:
        # for a node close to a grid line (here |y| ~ 1e-3) six digits cancel, so two correct evaluations of it differ by
        # ~1e-10 (one ulp of tan() is enough).  Conditioning of the reference formula, not of the kernel: the other windows hold 1e-13.
        htol = 1e-8
    run = Run1(tuple(g["N"]), g["x"], m=int(g["m"]), flags=int(g["flags"]), c2r=c2r, single=single)
    f, gr, h = run.trafo_hessian(g["f_hat"], F | G | A.COMPUTE_HESSIAN_F)
    run.close()
    assert rel_l2(f, g["out_f"]) <= tol
    assert rel_l2(gr, g["out_grad_f"]) <= (1e-4 if (single and "sinc_power" in case) else tol)
    assert rel_l2(h, g["out_hessian_f"]) <= htol


ICASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "i_*.npz")))


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("case", ICASES)
def test_intpol_golden(case, variant):
    """PNFFT_PRE_{CONST,LIN,QUAD,CUB}_PSI (reference kernel/ndft-parallel.c:321-353, 1586-1617): window values, first and
    second derivatives from interpolation tables.  The fixtures come from the unmodified reference, including its flag
    promotion that makes PNFFT_PRE_LIN_PSI interpolate with order 0 (api/api-guru.c:152-155)."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    c2r = bool(g["c2r"])
    run = Run1(tuple(g["N"]), g["x"], m=int(g["m"]), flags=int(g["flags"]), c2r=c2r, variant=variant)
    f, gr, h = run.trafo_hessian(g["f_hat"], F | G | A.COMPUTE_HESSIAN_F)
    fh = run.adj(g["f"], g["grad_f"], F | G)
    run.close()
    assert rel_l2(f, g["out_f"]) <= 1e-13
    assert rel_l2(gr, g["out_grad_f"]) <= 1e-13
    assert rel_l2(h, g["out_hessian_f"]) <= 1e-12
    assert rel_l2(fh, g["out_f_hat"]) <= 1e-13


CCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "c_*.npz")))


@pytest.mark.parametrize("variant", [0, 1])   # the family the plan picks / the generic kernels
@pytest.mark.parametrize("case", CCASES)
def test_combination_golden(case, variant):
    """Flag and size combinations against the compiled reference (tools/make_golden.py: combination_cases): Hessians with
    interlacing / truncated torus / transposed f_hat, oversampling factors other than 2 and different per axis, cutoffs
    m = 2, 3, 5, 7, 8, 10, PNFFT_PRE_PSI with interlacing / torus / transposed f_hat, three-flag mixes."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    N, n, x_max, _ = fixture_kwargs(g)
    cf, pre, hess = int(g["cf"]), int(g["pre"]), "out_hessian_f" in g.files
    run = Run1(N, g["x"], n=n, m=int(g["m"]), flags=int(g["flags"]), c2r=c2r, single=single, x_max=x_max, variant=variant)
    if pre:
        run.plan.precompute_psi(run.nodes, pre)
    if hess:
        f, gr, h = run.trafo_hessian(g["f_hat"], cf | A.COMPUTE_HESSIAN_F)
    else:
        f, gr = run.trafo(g["f_hat"], cf)
    fh = run.adj(g["f"], g["grad_f"], cf)
    run.close()
    assert rel_l2(f, g["out_f"]) <= tol
    if cf & G:
        assert rel_l2(gr, g["out_grad_f"]) <= tol
    if hess:
        assert rel_l2(h, g["out_hessian_f"]) <= 10 * tol
    assert rel_l2(fh, g["out_f_hat"]) <= tol


DCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "d_*.npz")))


@pytest.mark.parametrize("case", DCASES)
def test_direct_ndft_golden(case):
    """PNFFT_COMPUTE_DIRECT (reference kernel/ndft-parallel.c:377-722, csrc/direct.cuh): the slow NDFT and its adjoint the
    reference's drivers compare the fast transform with -- f, grad_f, hessian_f, c2c / c2r, transposed f_hat, float,
    PNFFT_COMPUTE_ACCUMULATED (the direct trafo overwrites its outputs whatever the flag says, the adjoint adds to f_hat).
    The c2r gradient of the reference's NDFT has the wrong sign (:517-519, :597); the product returns the derivative of
    the direct f, so that comparison flips the fixture's sign (tests/test_oracle.py::test_direct_fixtures pins the algebra)."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    N, n, x_max, acc = fixture_kwargs(g)
    D = A.COMPUTE_DIRECT
    run = Run1(N, g["x"], n=n, m=int(g["m"]), flags=int(g["flags"]), c2r=c2r, single=single, x_max=x_max)
    before = run.plan.kernel_launches()
    if acc:
        run.f[...] = g["f0"]; run.g[...] = g["grad_f0"]
        f, gr, h = run.trafo_hessian(g["f_hat"], F | G | A.COMPUTE_HESSIAN_F | D | A.COMPUTE_ACCUMULATED)
        fh = run.adj(g["f"], g["grad_f"], F | G | D | A.COMPUTE_ACCUMULATED, f_hat0=g["f_hat0"])
    else:
        f, gr, h = run.trafo_hessian(g["f_hat"], F | G | A.COMPUTE_HESSIAN_F | D)
        fh = run.adj(g["f"], g["grad_f"], F | G | D)
    assert run.plan.kernel_launches() - before == 3      # NDFT kernel + store, adjoint kernel: nothing of the fast path ran
    # the fast transform of the same plan agrees with the slow one to the method's accuracy at m = 4
    ff, _ = run.trafo(g["f_hat"], F)
    run.close()
    assert rel_l2(f, g["out_f"]) <= tol
    assert rel_l2(gr, (-1 if c2r else 1) * g["out_grad_f"]) <= tol
    assert rel_l2(h, g["out_hessian_f"]) <= tol
    assert rel_l2(fh, g["out_f_hat"]) <= tol
    if not c2r:      # (random half spectra are not Hermitian on the planes k2 = 0, -N2/2: the two c2r transforms read them differently)
        assert rel_l2(ff, f) <= 1e-4


BCASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "b_*.npz")))


@pytest.mark.parametrize("variant", [0, 8, 1])
@pytest.mark.parametrize("case", BCASES)
def test_set_b_golden(case, variant):
    """pnfft_set_b (reference api/api-basic.c:587-596): window shape parameters changed after the plan was made -- the D
    tables, fast-Gaussian constants and interpolation tables are rebuilt.  A transform with the DEFAULT shape runs first on
    the same node object, so that everything it leaves behind for unchanged coordinates (bins, the tensor-core kernels'
    node-table rows) meets the new shape: pnfft_adj right after pnfft_set_b must not gather from rows of the old window."""
    g = np.load(os.path.join(GOLD, case + ".npz"))
    single, c2r = bool(g["single"]), bool(g["c2r"])
    tol = 1e-5 if single else 1e-13
    run = Run1(tuple(g["N"]), g["x"], m=int(g["m"]), flags=int(g["flags"]), c2r=c2r, single=single, variant=variant)
    f_def, _ = run.trafo(g["f_hat"], F | G)
    b = tuple(float(v) for v in g["b"])
    run.plan.set_b(*b)
    assert np.allclose(run.plan.get_b(), b, rtol=1e-6 if single else 0, atol=0)
    fh = run.adj(g["f"], g["grad_f"], F | G)
    f, gr = run.trafo(g["f_hat"], F | G)
    run.close()
    if not single:      # the default shape gives other values (the approximation error differs; in float both are at rounding)
        assert rel_l2(f_def, g["out_f"]) > 1e-10
    assert rel_l2(fh, g["out_f_hat"]) <= tol
    assert rel_l2(f, g["out_f"]) <= tol
    assert rel_l2(gr, g["out_grad_f"]) <= tol


@pytest.mark.parametrize("win", ["kaiser_bessel", "gaussian"])
def test_intpol_quadratic_vs_direct(win):
    """PNFFT_PRE_QUAD_PSI has no usable reference output (its flag bit doubles as PNFFT_REAL_F inside the reference's node
    loop, kernel/ndft-parallel.c:2804-2838): the quadratic tables are checked against direct window evaluation instead; with
    6144 table nodes per grid interval the interpolation error is far below 1e-8."""
    N, M = (16, 16, 16), 2000
    x, fh, f, g = make_inputs(N, M, 77)
    flags = A.WINDOW_GAUSSIAN if win == "gaussian" else 0
    res = []
    for fl in (flags, flags | (1 << 4)):
        run = Run1(N, x, m=6, flags=fl)
        fo, go = run.trafo(fh, F | G)
        fho = run.adj(f, g, F | G)
        run.close()
        res.append((fo, go, fho))
    for a, b in zip(res[0], res[1]):
        assert 0 < rel_l2(a, b) <= 1e-8


def test_hessian_only_and_accumulated():
    """COMPUTE_HESSIAN_F alone (no f / grad_f requested) and with PNFFT_COMPUTE_ACCUMULATED: h0 + H."""
    g = np.load(os.path.join(GOLD, "h_kaiser_bessel_ad_c2c_m6_d.npz"))
    run = Run1(tuple(g["N"]), g["x"], m=6, flags=int(g["flags"]))
    _, _, h = run.trafo_hessian(g["f_hat"], A.COMPUTE_HESSIAN_F)
    assert rel_l2(h, g["out_hessian_f"]) <= 1e-13
    run.h[...] = 1.0 + 2.0j
    run.put_f_hat(g["f_hat"])
    run.plan.trafo(run.nodes, A.COMPUTE_HESSIAN_F | A.COMPUTE_ACCUMULATED)
    assert rel_l2(run.h - (1.0 + 2.0j), g["out_hessian_f"]) <= 1e-13
    run.close()


@pytest.mark.parametrize("m", [6, 8])
def test_table_in_column_batches(m, tmp_path):
    """PNFFT_B200_TABLE_GB: above the budget the window rows are built and consumed in column batches (BASELINE config 5 needs
    it at 2^27 nodes).  A budget of 1 MB forces a dozen batches on 30000 nodes; the results must not change.  The switch is
    read once per process, hence the two subprocesses."""
    import subprocess, sys
    res = []
    for gb in (None, "0.001"):
        env = dict(os.environ)
        if gb:
            env["PNFFT_B200_TABLE_GB"] = gb
        out = str(tmp_path / ("r_%s.npz" % (gb or "all")))
        subprocess.run([sys.executable, os.path.join(os.path.dirname(GOLD), os.pardir, "tools", "run_small_case.py"), out, str(m), "0"],
                       env=env, check=True, timeout=300)
        res.append(np.load(out))
    for k in ("f", "g", "fh"):
        assert rel_l2(res[1][k], res[0][k]) <= 1e-14


@pytest.mark.parametrize("c2r", [False, True])
@pytest.mark.parametrize("m", [5, 7, 8])
def test_family3_shared_window_gather(ref, m, c2r):
    """Kernel family 3 (double, m = 5, 7, 8): the tensor-core gather with the window in shared memory and the tensor-core
    scatter on the same bins (zmarch4.cuh), on a node set with a dense cluster, sparse
    surroundings and empty sub-chunks, against the generic kernels (independent code path) and a node subset against the
    compiled reference; F alone, F with the gradient, and accumulation into the outputs."""
    N, M = (32, 32, 32), 30000
    x, fh, f, g = make_inputs(N, M, 52, c2r=c2r)
    rng = np.random.default_rng(5)
    x[: M // 2] = np.clip(rng.normal(0.1, 0.04, (M // 2, 3)), -0.5, np.nextafter(0.5, 0.0))     # the cluster
    x[M // 2: M // 2 + 64, 2] = -0.5                                                          # nodes on the lowest grid plane
    res = []
    for variant in (0, 1):
        run = Run1(N, x, m=m, c2r=c2r, variant=variant)
        f1, _ = run.trafo(fh, F)
        f2, g2 = run.trafo(fh, F | G)
        run.f[...] = 1.0; run.g[...] = 2.0
        run.plan.trafo(run.nodes, F | G | A.COMPUTE_ACCUMULATED)
        f3, g3 = run.f.copy(), run.g.copy()
        fho = run.adj(f, g, F | G)
        fho_f = run.adj(f, None, F)
        fho_acc = run.adj(f, None, F | A.COMPUTE_ACCUMULATED, f_hat0=fh) - fh
        run.close()
        res.append((f1, f2, g2, f3 - 1.0, g3 - 2.0, fho, fho_f, fho_acc))
    for a, b in zip(res[0], res[1]):
        assert rel_l2(a, b) <= 1e-13
    assert rel_l2(res[0][0], res[0][1]) <= 1e-15       # F and F|GRAD instantiations
    sub = slice(0, 2048)
    rt = ref.trafo(N, x[sub], fh, m=m, compute_flags=F | G, c2r=c2r)
    assert rel_l2(res[0][1][sub], rt["f"]) <= 1e-13 and rel_l2(res[0][2][sub], rt["grad_f"]) <= 1e-13
    ra = ref.adj(N, x, f=f, grad_f=g, m=m, compute_flags=F | G, c2r=c2r)
    assert rel_l2(res[0][5], ra["f_hat"]) <= 1e-13


@pytest.mark.parametrize("m", [4, 6])
@pytest.mark.parametrize("single", [False, True])
@pytest.mark.parametrize("c2r", [False, True])
def test_families_agree(c2r, single, m):
    """z-march v2 / v1 / generic kernels on the same 40k nodes (N=32^3): independent code paths, same numbers."""
    N, M = (32, 32, 32), 40000
    x, fh, f, g = make_inputs(N, M, 21, c2r=c2r, single=single)
    tol = 2e-5 if single else 1e-13
    res = []
    for variant in (0, 8, 1):
        run = Run1(N, x, m=m, c2r=c2r, single=single, variant=variant)
        fo, go = run.trafo(fh, F | G)
        fho = run.adj(f, g, F | G)
        run.close()
        res.append((fo, go, fho))
    for r in res[1:]:
        for a, b in zip(res[0], r):
            assert rel_l2(a, b) <= tol


NSUB_T, NSUB_A = 4096, 1 << 16     # node subsets compared with the reference at the BASELINE sizes


def _check_subset_vs_ref(ref, N, x, fh, f, g, fo, go, cf, m=6, flags=0, c2r=False, tol=1e-13):
    """At sizes the CPU checker cannot run in full: trafo on the first NSUB_T nodes with the SAME full-size f_hat, and the
    adjoint of the first NSUB_A nodes (its own small GPU run, same plan size), against the compiled reference."""
    sub = slice(0, NSUB_T)
    mesh = (2, 4) if N[0] >= 256 else (1, 1)     # the checker's virtual ranks are threads: 5x faster at n = 512^3
    rt = ref.trafo(N, x[sub], fh, m=m, pnfft_flags=flags, compute_flags=cf, c2r=c2r, np_mesh=mesh)
    assert rel_l2(fo[sub], rt["f"]) <= tol, "trafo f vs reference on a node subset"
    if cf & G:
        assert rel_l2(go[sub], rt["grad_f"]) <= tol, "trafo grad_f vs reference on a node subset"
    sa = slice(0, NSUB_A)
    run = Run1(N, x[sa], m=m, flags=flags, c2r=c2r)
    ha = run.adj(f[sa], g[sa] if (cf & G) else None, cf)
    run.close()
    ra = ref.adj(N, x[sa], f=f[sa], grad_f=g[sa] if (cf & G) else None, m=m, pnfft_flags=flags, compute_flags=cf, c2r=c2r,
                 np_mesh=mesh)
    assert rel_l2(ha, ra["f_hat"]) <= tol, "adj f_hat vs reference (node subset)"


@pytest.mark.parametrize("cf", [F, F | G])
def test_adjointness_c2_full_size(ref, cf):
    """BASELINE config 2 at full size (N=128^3, n=256^3, M=2^21 uniform nodes, Kaiser-Bessel m=6, double):
    <A f_hat, (f, g)> == <f_hat, A^H (f, g)> for the whole transform pair (D, F and B included), and a node subset of both
    transforms against the reference on the same full-size data."""
    N, M = (128, 128, 128), 1 << 21
    x, fh, f, g = make_inputs(N, M, 31)
    run = Run1(N, x, m=6)
    fo, go = run.trafo(fh, cf)
    fho = run.adj(f, g, cf)
    run.close()
    lhs = np.vdot(f, fo) + (np.vdot(g, go) if cf & G else 0.0)
    rhs = np.vdot(fho, fh)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    _check_subset_vs_ref(ref, N, x, fh, f, g, fo, go, cf)


def test_c3_full_size_properties(ref):
    """BASELINE config 3 (the bench workload) at full size: N=256^3, n=512^3, M=2^24, Kaiser-Bessel m=6, double.
    Adjointness of the pair, trafo(F|GRAD_F) agrees with trafo(F) on f (two different kernel instantiations), and node
    subsets of trafo(F|GRAD_F) / adj(F|GRAD_F) against the reference on the same full-size f_hat."""
    N, M = (256, 256, 256), 1 << 24
    rng = np.random.default_rng(71)
    x = np.clip(rng.uniform(-0.5, 0.5, (M, 3)), -0.5, np.nextafter(0.5, 0.0))
    fh = (rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N))
    f = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
    g = rng.uniform(-1, 1, (NSUB_A, 3)) + 1j * rng.uniform(-1, 1, (NSUB_A, 3))
    run = Run1(N, x, m=6)
    f1, _ = run.trafo(fh, F)
    f2, g2 = run.trafo(fh, F | G)
    fho = run.adj(f, None, F)
    run.close()
    assert rel_l2(f2, f1) <= 1e-13
    assert np.all(np.isfinite(g2.view(np.float64)))
    lhs, rhs = np.vdot(f, f1), np.vdot(fho, fh)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    _check_subset_vs_ref(ref, N, x, fh, f, g, f2, g2, F | G)


def test_c5_large_c2r_float_vs_double(ref):
    """BASELINE config 5 at N=256^3 (n=512^3), M=2^24 nodes: the c2r real-input transforms in double against the reference
    on node subsets, and in SINGLE precision against double on the whole problem (rel-l2 <= 1e-5: the accumulation of
    2^24 float contributions per grid cell neighbourhood is the risk SURVEY section 7 names)."""
    N, M = (256, 256, 256), 1 << 24
    x, _, f, g = make_inputs(N, M, 93, c2r=True)
    rund = Run1(N, x, m=6, c2r=True)
    hd = rund.adj(f, g, F)
    scale = 1.0 / np.abs(hd).max()
    hz = hd * scale
    hz[0, :, :] = 0; hz[:, 0, :] = 0; hz[:, :, 0] = 0      # planes k_t = -N_t/2 have no Hermitian partner
    fd, gd = rund.trafo(hz, F | G)
    rund.close()
    _check_subset_vs_ref(ref, N, x, hz, f, g, fd, gd, F | G, c2r=True)
    runf = Run1(N, x.astype(np.float32), m=6, c2r=True, single=True)
    hf = runf.adj(f.astype(np.float32), g.astype(np.float32), F)
    ff, gf = runf.trafo(hz.astype(np.complex64), F | G)
    runf.close()
    # the float run rounds x to float first: compare with the double transform of the SAME rounded nodes
    rund = Run1(N, x.astype(np.float32).astype(np.float64), m=6, c2r=True)
    hd32 = rund.adj(f, g, F)
    fd32, gd32 = rund.trafo(hz, F | G)
    rund.close()
    assert rel_l2(hf, hd32) <= 1e-5, "adjoint, float vs double at M = 2^24"
    assert rel_l2(ff, fd32) <= 1e-5 and rel_l2(gf, gd32) <= 1e-5, "trafo, float vs double at M = 2^24"


def test_gather_is_deterministic_over_repeats():
    """The warp-autonomous gather hands partial sums between warps through mbarrier-guarded shared-memory stages (and its
    cross-proxy fence after reading the staged cells was dropped in round 1): 50 repeats at BASELINE config 2 size must be
    bit-identical, or a stage is being read before it is complete."""
    N, M = (128, 128, 128), 1 << 21
    x, fh, f, g = make_inputs(N, M, 33)
    run = Run1(N, x, m=6)
    run.nodes.x_static(True)
    f0, g0 = run.trafo(fh, F | G)
    h0 = run.adj(f, g, F | G)
    for _ in range(50):
        f1, g1 = run.trafo(fh, F | G)
        assert np.array_equal(f1, f0) and np.array_equal(g1, g0)
    h1 = run.adj(f, g, F | G)
    run.close()
    assert rel_l2(h1, h0) <= 1e-15     # the scatter's reduce-adds may arrive in any order


def test_x_static_and_device_hash():
    """Node-side reuse: with pnfft_b200_nodes_x_static the upload and the bins of the first call serve the
    following ones until pnfft_set_x; device-resident coordinates are re-binned exactly when their content hash changes."""
    import torch
    N, M = (32, 32, 32), 30000
    x, fh, f, g = make_inputs(N, M, 45)
    x2 = np.clip(x + 0.01, -0.5, np.nextafter(0.5, 0.0))
    fresh = Run1(N, x2, m=6)
    f_x2, _ = fresh.trafo(fh, F)
    fresh.close()
    run = Run1(N, x.copy(), m=6)
    run.nodes.x_static(True)
    f_a, _ = run.trafo(fh, F)
    l0 = run.plan.kernel_launches()
    f_b, _ = run.trafo(fh, F)
    per_call_static = run.plan.kernel_launches() - l0
    assert np.array_equal(f_a, f_b)
    run.x[...] = x2                       # changed behind the library's back: the promise says it may keep the old nodes
    f_c, _ = run.trafo(fh, F)
    assert np.array_equal(f_c, f_a)
    run.nodes.set_x(run.x)                # ... until it is told
    f_d, _ = run.trafo(fh, F)
    assert rel_l2(f_d, f_x2) <= 1e-14
    run.nodes.x_static(False)
    l0 = run.plan.kernel_launches()
    run.trafo(fh, F)
    assert run.plan.kernel_launches() - l0 > per_call_static      # binning is back
    run.close()
    # device-resident x
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    comm = A.create_procmesh_2d(1, 1)
    plan = A.Plan.init_guru(N, tuple(2 * v for v in N), (0.5,) * 3, 6, 0, comm)
    nodes = A.Nodes(M, 0)
    df = torch.zeros((M, 2), dtype=torch.float64, device=dev)
    dfh = torch.from_numpy(np.ascontiguousarray(fh).view(np.float64).reshape(N + (2,))).to(dev)
    nodes.set_x(dx); nodes.set_f(df); plan.set_f_hat(dfh)
    plan.trafo(nodes, F)
    r1 = df.cpu().numpy().view(np.complex128).ravel().copy()
    l0 = plan.kernel_launches(); plan.trafo(nodes, F); same = plan.kernel_launches() - l0
    assert np.array_equal(df.cpu().numpy().view(np.complex128).ravel(), r1)
    dx.copy_(torch.from_numpy(x2).to(dev))          # same pointer, new content
    torch.cuda.synchronize()
    l0 = plan.kernel_launches(); plan.trafo(nodes, F); changed = plan.kernel_launches() - l0
    assert rel_l2(df.cpu().numpy().view(np.complex128).ravel(), f_x2) <= 1e-14
    assert changed > same                 # the hash mismatch brought the binning kernel back
    nodes.free(0); plan.finalize(0)


def test_adj_runs_ahead_of_the_coordinate_check():
    """pnfft_adj on HOST coordinates that were unchanged last time starts on the bins it has while x is uploaded and hashed
    behind the transform (Core::adj / prepare_nodes); coordinates changed in place must still be noticed and the adjoint
    redone on them."""
    N, M = (32, 32, 32), 30000
    x, fh, f, g = make_inputs(N, M, 46)
    x2 = np.clip(x + 0.013, -0.5, np.nextafter(0.5, 0.0))
    fresh = Run1(N, x2, m=6)
    h_x2 = fresh.adj(f, g, F | G)
    f_x2, g_x2 = fresh.trafo(fh, F | G)
    fresh.close()
    run = Run1(N, x.copy(), m=6)
    run.trafo(fh, F | G)
    h_a = run.adj(f, g, F | G)            # checked first: the coordinates are the trafo's
    run.trafo(fh, F | G)
    l0 = run.plan.kernel_launches()
    h_b = run.adj(f, g, F | G)            # runs ahead of the check
    ahead = run.plan.kernel_launches() - l0
    assert rel_l2(h_b, h_a) <= 1e-15
    run.x[...] = x2                       # same pointer, new content, no pnfft_set_x
    l0 = run.plan.kernel_launches()
    h_c = run.adj(f, g, F | G)            # the hash differs: redone on the new coordinates
    redone = run.plan.kernel_launches() - l0
    assert rel_l2(h_c, h_x2) <= 1e-14
    assert redone > ahead
    f_c, g_c = run.trafo(fh, F | G)
    assert rel_l2(f_c, f_x2) <= 1e-14 and rel_l2(g_c, g_x2) <= 1e-14
    h_d = run.adj(f, g, F | G)
    assert rel_l2(h_d, h_x2) <= 1e-14
    run.close()


@pytest.mark.parametrize("win", [A.WINDOW_GAUSSIAN, A.WINDOW_GAUSSIAN | A.FAST_GAUSSIAN, A.WINDOW_BSPLINE])
def test_c4_clustered_m8(ref, win):
    """BASELINE config 4 in miniature: strongly clustered (Gaussian blob, sigma = 0.05) nodes, Gaussian / fast Gaussian /
    B-spline windows with m=8, PRE_PSI against on-the-fly, against the reference on a node subset."""
    N, M = (64, 64, 64), 1 << 17
    rng = np.random.default_rng(81)
    x = np.mod(rng.normal(0.0, 0.05, (M, 3)) + 0.5, 1.0) - 0.5
    x = np.clip(x, -0.5, np.nextafter(0.5, 0.0))
    _, fh, f, g = make_inputs(N, M, 82)
    run = Run1(N, x, m=8, flags=win)
    f0, g0 = run.trafo(fh, F | G)
    h0 = run.adj(f, g, F | G)
    run.plan.precompute_psi(run.nodes, A.PRE_PSI | A.PRE_GRAD_PSI)
    f1, g1 = run.trafo(fh, F | G)
    h1 = run.adj(f, g, F | G)
    run.close()
    assert rel_l2(f1, f0) <= 1e-13 and rel_l2(g1, g0) <= 1e-13 and rel_l2(h1, h0) <= 1e-13
    sub = slice(0, 4096)
    rt = ref.trafo(N, x[sub], fh, m=8, pnfft_flags=win, compute_flags=F | G)
    assert rel_l2(f0[sub], rt["f"]) <= 1e-13 and rel_l2(g0[sub], rt["grad_f"]) <= 1e-13


@pytest.mark.parametrize("single", [False, True])
def test_c5_c2r_variants(single):
    """BASELINE config 5 in miniature (c2r real input, double and single precision, N=64^3, M=2^18): the real-valued
    transforms against the complex ones on the same data, and float against double."""
    N, M = (64, 64, 64), 1 << 18
    x, fh, f, g = make_inputs(N, M, 91, c2r=True, single=single)
    # f_hat = A^H f of a real vector is a Hermitian-consistent half spectrum; scale it to O(1) so that the second
    # transform stays inside the float range (the Kaiser-Bessel window is not normalised: psi ~ 5e10 at m = 6)
    run = Run1(N, x, m=6, c2r=True, single=single)
    h1 = run.adj(f, g, F)
    scale = 1.0 / np.abs(h1).max()
    # c2r against c2c on the same real data (the reference's own check, tests/simple_test_c2r_c2c_compare_*.c): the stored
    # half spectrum k2 in [-N2/2, 0] of the real adjoint equals that part of the complex adjoint of (f + 0i)
    cdt = np.complex64 if single else np.complex128
    runc = Run1(N, x, m=6, c2r=False, single=single)
    hc = runc.adj(f.astype(cdt), g.astype(cdt), F)
    tol = 2e-5 if single else 1e-13
    assert rel_l2(h1, hc[:, :, :N[2] // 2 + 1]) <= tol, "adjoint: c2r half spectrum vs c2c"
    # trafo: the planes k_t = -N_t/2 have no mirror partner inside [-N/2, N/2), so only a spectrum without them is
    # Hermitian in the sense the c2r transform assumes; zero them on both sides, then c2r == Re(c2c), Im(c2c) == 0
    hz, hcz = (h1 * scale).astype(h1.dtype), (hc * scale).astype(hc.dtype)
    for a in (hz, hcz):
        a[0, :, :] = 0; a[:, 0, :] = 0; a[:, :, 0] = 0
    f1, _ = run.trafo(hz, F)
    fc, _ = runc.trafo(hcz, F)
    run.close(); runc.close()
    assert np.all(np.isfinite(f1))
    assert rel_l2(f1, fc.real) <= (1e-4 if single else 1e-12), "trafo: c2r vs Re(c2c)"
    assert np.abs(fc.imag).max() <= (1e-4 if single else 1e-12) * np.abs(fc.real).max(), "trafo: Im(c2c) of a Hermitian spectrum"
    if single:
        xd, _, fd, gd = make_inputs(N, M, 91, c2r=True, single=True)
        rund = Run1(N, xd.astype(np.float64), m=6, c2r=True, single=False)
        h1d = rund.adj(fd.astype(np.float64), gd.astype(np.float64), F)
        rund.close()
        assert rel_l2(h1, h1d) <= 1e-5


def test_linearity_and_accumulate():
    N, M = (32, 32, 32), 20000
    x, fh, f, g = make_inputs(N, M, 41)
    _, fh2, _, _ = make_inputs(N, M, 42)
    run = Run1(N, x, m=6)
    a, _ = run.trafo(fh, F)
    b, _ = run.trafo(fh2, F)
    c, _ = run.trafo(fh + 2.0 * fh2, F)
    assert rel_l2(c, a + 2.0 * b) <= 1e-13
    # PNFFT_COMPUTE_ACCUMULATED adds to what is already in f (reference api/api-basic.c:210-222)
    run.f[...] = a
    run.f_hat[...] = fh2
    run.plan.trafo(run.nodes, F | A.COMPUTE_ACCUMULATED)
    assert rel_l2(run.f, a + b) <= 1e-13
    run.close()


@pytest.mark.parametrize("win", [0, A.WINDOW_GAUSSIAN, A.WINDOW_BSPLINE])
def test_pre_psi_equals_on_the_fly(win):
    """PNFFT_PRE_PSI | PRE_GRAD_PSI tables (reference kernel/ndft-parallel.c:1144-1370) give the on-the-fly result."""
    N, M = (16, 16, 16), 5000
    x, fh, f, g = make_inputs(N, M, 51)
    run = Run1(N, x, m=6, flags=win)
    f0, g0 = run.trafo(fh, F | G)
    h0 = run.adj(f, g, F | G)
    run.plan.precompute_psi(run.nodes, A.PRE_PSI | A.PRE_GRAD_PSI)
    f1, g1 = run.trafo(fh, F | G)
    h1 = run.adj(f, g, F | G)
    run.close()
    assert rel_l2(f1, f0) <= 1e-13 and rel_l2(g1, g0) <= 1e-13 and rel_l2(h1, h0) <= 1e-13


def test_edge_node_sets(ref):
    """Empty node set, a single node, every node in one grid cell (maximal write contention), nodes on the domain
    border and on grid lines."""
    N = (16, 16, 16)
    _, fh, _, _ = make_inputs(N, 10, 61)
    # empty
    run = Run1(N, np.zeros((0, 3)), m=6)
    run.plan.trafo(run.nodes, F)
    fh0 = run.adj(None, None, F)
    run.close()
    assert np.all(fh0 == 0)
    # one node / one cell / border
    rng = np.random.default_rng(62)
    sets = {
        "single": np.array([[0.123, -0.377, 0.4999]]),
        "one_cell": 0.25 + rng.uniform(0, 1.0 / 32, (3000, 3)),
        "border": np.concatenate([np.full((500, 3), -0.5), np.nextafter(0.5, 0) * np.ones((500, 3)),
                                  rng.integers(-16, 16, (500, 3)) / 32.0]),
    }
    for name, x in sets.items():
        M = x.shape[0]
        f = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
        g = rng.uniform(-1, 1, (M, 3)) + 1j * rng.uniform(-1, 1, (M, 3))
        rt = ref.trafo(N, x, fh, compute_flags=F | G)
        ra = ref.adj(N, x, f=f, grad_f=g, compute_flags=F | G)
        run = Run1(N, x, m=6)
        fo, go = run.trafo(fh, F | G)
        fho = run.adj(f, g, F | G)
        run.close()
        assert rel_l2(fo, rt["f"]) <= 1e-13, name
        # nodes one ulp below a grid line: the reference's Kaiser-Bessel derivative n^2 x / d * (psi - b cosh(b r) / pi)
        # (kernel/ndft-parallel.c:2256-2269) cancels catastrophically as d = m^2 - y^2 -> 0 (d ~ 2e-14 here), so the
        # reference itself carries only ~3 digits in that tap; the bar for that set is what its formula can deliver
        gtol = 5e-12 if name == "border" else 1e-13
        assert rel_l2(go, rt["grad_f"]) <= gtol, name
        assert rel_l2(fho, ra["f_hat"]) <= gtol, name


def test_grid_index_bit_exact():
    """floor(n x) - m - local_no_start + gcells_below and the plain index m0 (reference kernel/ndft-parallel.c:1563-1572,
    :2165-2174): integers identical to the reference for nodes including exact grid lines and the domain border."""
    g = np.load(os.path.join(GOLD, "node_index.npz"))
    N, m, x = tuple(int(v) for v in g["N"]), int(g["m"]), g["x"]
    run = Run1(N, x, m=m)
    idx = run.plan.node_grid_index(run.nodes)
    run.close()
    assert np.array_equal(idx, g["index_1x1"])


def test_sort_keys_bit_exact():
    """PNFFT_SORT_NODES key ((floor(n x - m)) mod n, row-major) and the stable permutation (reference
    kernel/ndft-parallel.c:2121-2159, util/util.c:206-277): identical integers."""
    g = np.load(os.path.join(GOLD, "node_index.npz"))
    N, m, x = tuple(int(v) for v in g["N"]), int(g["m"]), g["x"]
    run = Run1(N, x, m=m)
    keys, perm = run.plan.sort_nodes(run.nodes)
    run.close()
    assert np.array_equal(keys, g["sort_keys"])
    assert np.array_equal(perm, g["sort_perm"])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    """Pencil decomposition over 1x2 / 2x2 / 2x4 GPUs (NCCL ghost cells and FFT transposes) against the single-rank CPU
    checker: tools/mgpu_parity.py under torch.distributed.run.  Skipped where fewer GPUs are visible."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    root = os.path.dirname(GOLD.rstrip("/")).rsplit("/", 1)[0]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29540 + world), os.path.join(root, "tools", "mgpu_parity.py")]
    r = subprocess.run(cmd, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-3000:]
    assert json.loads(lines[-1])["ok"]


def _run_driver(name, args, rc_ok=(0,)):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "drivers", name)
    if not os.path.exists(exe):
        pytest.skip("reference driver %s not built (make -C oracle drivers needs /root/reference)" % name)
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, env=env)
    assert r.returncode in rc_ok, r.stdout[-2000:]
    return r.stdout


def test_reference_drivers():
    """The reference's OWN test programs (tests/check_trafo.c, check_adj.c, check_vs_pfft.c, pnfft_test.c), compiled
    unmodified against include/pnfft.h and linked with libpnfft_b200.so, run on the GPU and report the errors they compute
    themselves: NFFT (m) against a higher-accuracy NFFT (reference tests/check_trafo.c:41-57), and the equispaced
    forward/backward round trip, which must be the identity (reference tests/check_vs_pfft.c:138-156)."""
    import re
    one = ["-pnfft_np", "1", "1", "1", "-pnfft_compute_hessian_f", "0", "-pnfft_N", "16", "16", "16"]
    for drv in ("check_trafo", "check_adj"):
        out = _run_driver(drv, one)
        errs = [float(v) for v in re.findall(r"relative error =\s*([0-9.eE+-]+)", out)]
        assert errs, out[-2000:]
        assert max(errs) < 1e-7, out[-2000:]   # truncation error of the method at m=6 (f ~4e-11, AD gradient ~2e-8)
    # the drivers' default (api/api-basic.c:834-836) asks for the Hessian too: f, 3 gradient and 6 Hessian components,
    # each against the m + 2 transform (two more derivatives cost about five digits of the method's accuracy at m = 6)
    out = _run_driver("check_trafo", ["-pnfft_np", "1", "1", "1", "-pnfft_N", "16", "16", "16"])
    errs = [float(v) for v in re.findall(r"relative error =\s*([0-9.eE+-]+)", out)]
    assert len(errs) == 10, out[-2000:]
    assert max(errs) < 1e-3 and max(errs[:4]) < 1e-7, out[-2000:]
    out = _run_driver("check_vs_pfft", ["-pnfft_np", "1", "1", "1", "-pnfft_N", "16", "16", "16"])
    errs = [float(v) for v in re.findall(r"relative maximum error =\s*([0-9.eE+-]+)", out)]
    assert errs and max(errs) < 1e-10, out[-2000:]
    out = _run_driver("pnfft_test", [], rc_ok=(0, 1))   # takes no options: its 2x2x2 mesh is refused with one rank,
    assert "Procmesh" in out or "PNFFT Results" in out   # through pnfft_create_procmesh's non-zero return (util/util.c:23-40)


def test_reference_ndft_drivers():
    """The reference's NDFT comparison programs, unmodified, on the product library (PNFFT_COMPUTE_DIRECT, csrc/direct.cuh):
    c2c NDFT against c2r NDFT on Hermitian input (tests/check_trafo_vs_ndft_c2r.c:157-175), the c2r adjoint NFFT against
    the adjoint NDFT (tests/check_adj_vs_ndft_c2r.c:150-178 -- with the reference's own library this one exposes its c2r
    spreading defect), the float and the transposed-f_hat NFFT against the NDFT."""
    import re
    out = _run_driver("check_trafo_vs_ndft_c2r", ["-pnfft_np", "1", "1", "-pnfft_N", "16", "16", "16"])
    errs = [float(v) for v in re.findall(r"max error between c2c pndft and c2r pndft:\s*([0-9.eE+-]+)", out)]
    assert errs and max(errs) < 1e-9, out[-2000:]
    out = _run_driver("check_adj_vs_ndft_c2r", ["-pnfft_np", "1", "1", "-pnfft_N", "16", "16", "16"])
    errs = [float(v) for v in re.findall(r"relative error =\s*([0-9.eE+-]+)", out)]
    assert errs and max(errs) < 1e-7, out[-2000:]
    out = _run_driver("check_trafo_vs_ndft_transposed_2d", ["-pnfft_np", "1", "1", "-pnfft_N", "16", "16", "16"])
    errs = [float(v) for v in re.findall(r"relative error =\s*([0-9.eE+-]+)", out)]
    assert errs and max(errs) < 1e-7, out[-2000:]
    # (-pnfft_m 8: this driver's own default, a Gaussian with m = 18, is beyond the library's cutoff limit of 16; measured 2.8e-6)
    out = _run_driver("check_trafo_vs_ndft_float", ["-pnfft_np", "1", "1", "1", "-pnfft_N", "16", "16", "16", "-pnfft_m", "8"])
    errs = [float(v) for v in re.findall(r"relative error =\s*([0-9.eE+-]+)", out)]
    assert errs and max(errs) < 1e-3, out[-2000:]
