"""GPU parity: libpnfft_b200.so (through its C ABI) against the compiled reference PNFFT (oracle/_ref).

Bars (BASELINE.json north_star): rel-l2 <= 1e-13 in double, <= 1e-5 in float; integer outputs identical.
"""
import numpy as np
import pytest

from pnfft_b200 import api as A
from tests.util import Run1, make_inputs, rel_l2

pytestmark = pytest.mark.gpu
TOL_D = 1e-13
TOL_F = 1e-5
F, G = A.COMPUTE_F, A.COMPUTE_GRAD_F


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("cf", [F, F | G, G])
def test_trafo_c2c_kaiser_bessel(ref, variant, cf):
    N, M = (16, 16, 16), 3000
    x, fh, _, _ = make_inputs(N, M, 1)
    r = ref.trafo(N, x, fh, compute_flags=cf)
    run = Run1(N, x, variant=variant)
    f, g = run.trafo(fh, cf)
    run.close()
    if cf & F:
        assert rel_l2(f, r["f"]) <= TOL_D
    if cf & G:
        assert rel_l2(g, r["grad_f"]) <= TOL_D


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("cf", [F, F | G, G])
def test_adj_c2c_kaiser_bessel(ref, variant, cf):
    N, M = (16, 16, 16), 3000
    x, _, f, g = make_inputs(N, M, 2)
    r = ref.adj(N, x, f=f if cf & F else None, grad_f=g if cf & G else None, compute_flags=cf)
    run = Run1(N, x, variant=variant)
    fh = run.adj(f, g, cf)
    run.close()
    assert rel_l2(fh, r["f_hat"]) <= TOL_D


WINDOWS = {
    "kaiser_bessel": 0,
    "gaussian": A.WINDOW_GAUSSIAN,
    "fast_gaussian": A.WINDOW_GAUSSIAN | A.FAST_GAUSSIAN,
    "bspline": A.WINDOW_BSPLINE,
    "sinc_power": A.WINDOW_SINC_POWER,
    "bessel_i0": A.WINDOW_BESSEL_I0,
}


@pytest.mark.parametrize("mode", [0, 2])   # fitted per-tap polynomials / exact formulas
@pytest.mark.parametrize("m", [4, 6, 8])
@pytest.mark.parametrize("win", sorted(WINDOWS))
def test_window_tensor(ref, win, m, mode):
    """3*(2m+1) window values and AD-gradient weights per node, as the kernels evaluate them, against the
    reference's pre_psi_tensor / pre_dpsi_tensor (kernel/ndft-parallel.c:1621-1953)."""
    N, M = (16, 16, 16), 2000
    x, _, _, _ = make_inputs(N, M, 3)
    x[:8] = np.array([[-0.5, 0.25, 0.0], [0.0, 0.0, 0.0], [0.125, -0.125, 0.375], [-0.25, 0.46875, -0.5],
                      [0.03125, 0.0625, 0.09375], [-0.5, -0.5, -0.5], [0.4, 0.0, -0.3], [0.1, 0.2, 0.3125]])  # nodes on grid lines
    flags = WINDOWS[win]
    psi_r, dpsi_r = ref.probe_tensor(x, N, m=m, pnfft_flags=flags)
    run = Run1(N, x, m=m, flags=flags, variant=mode)
    psi, dpsi = run.plan.window_tensor(run.nodes, m)
    run.close()
    s0 = np.abs(psi_r).max()
    s1 = np.abs(dpsi_r).max()
    assert np.abs(psi - psi_r).max() <= 2e-14 * s0
    assert np.abs(dpsi - dpsi_r).max() <= 2e-13 * s1


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("m", [4, 6, 8, 5])
@pytest.mark.parametrize("win", sorted(WINDOWS))
def test_trafo_adj_windows(ref, win, m, variant):
    if variant == 1 and m != 5 and win != "kaiser_bessel":
        pytest.skip("generic kernels are covered with m=5")
    N, M = (16, 16, 16), 1500
    x, fh, f, g = make_inputs(N, M, 4)
    flags = WINDOWS[win]
    rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=F | G)
    ra = ref.adj(N, x, f=f, grad_f=g, m=m, pnfft_flags=flags, compute_flags=F | G)
    run = Run1(N, x, m=m, flags=flags, variant=variant)
    fo, go = run.trafo(fh, F | G)
    fho = run.adj(f, g, F | G)
    run.close()
    assert rel_l2(fo, rt["f"]) <= TOL_D
    assert rel_l2(go, rt["grad_f"]) <= TOL_D
    assert rel_l2(fho, ra["f_hat"]) <= TOL_D
