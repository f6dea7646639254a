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
