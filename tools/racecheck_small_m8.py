"""Development aid: a small trafo + adj through the family-3 tensor-core kernels (m = 8), for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnfft_b200 import api as A
N = (24, 24, 24); M = 6000
rng = np.random.default_rng(0)
x = rng.uniform(-0.5, 0.4999, (M, 3)); x[:3000] = np.clip(rng.normal(0.1, 0.04, (3000, 3)), -0.5, 0.4999)
fh = rng.standard_normal(N) + 1j * rng.standard_normal(N)
f = np.zeros(M, np.complex128); g = np.zeros((M, 3), np.complex128)
comm = A.create_procmesh_2d(1, 1)
plan = A.Plan.init_guru(N, tuple(2 * v for v in N), (0.5,) * 3, 8, A.WINDOW_GAUSSIAN, comm)
nd = A.Nodes(M, 0); nd.set_x(x); nd.set_f(f); nd.set_grad_f(g); plan.set_f_hat(fh)
plan.trafo(nd, 3); plan.adj(nd, 1)
print("done", np.abs(f).max())
