"""Development aid: config C4-like timing (clustered Gaussian-blob nodes) for the gridding kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnfft_b200 import api as A
Ns = int(sys.argv[1]); M = int(sys.argv[2]); m = int(sys.argv[3]); flags = int(sys.argv[4]); sigma = float(sys.argv[5])
N = (Ns,) * 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
if sigma > 0:
    x = torch.randn((M, 3), generator=g, device=dev, dtype=torch.float64) * sigma
    x = torch.remainder(x + 0.5, 1.0) - 0.5
else:
    x = torch.rand((M, 3), generator=g, device=dev, dtype=torch.float64) - 0.5
x = x.clamp_(-0.5, 0.5 - 1e-12)
fh = torch.randn(N + (2,), generator=g, device=dev, dtype=torch.float64)
f = torch.zeros((M, 2), device=dev, dtype=torch.float64); gr = torch.zeros((M, 3, 2), device=dev, dtype=torch.float64)
comm = A.create_procmesh_2d(1, 1)
plan = A.Plan.init_guru(N, tuple(2 * v for v in N), (0.5,) * 3, m, flags, comm)
nodes = A.Nodes(M, 0); nodes.set_x(x); nodes.set_f(f); nodes.set_grad_f(gr); plan.set_f_hat(fh)
for it in range(3):
    plan.trafo(nodes, 1); a = plan.stage_ms(False)["b_kernel"]
    plan.adj(nodes, 1); b = plan.stage_ms(True)["b_kernel"]
print("N", Ns, "M", M, "m", m, "flags", flags, "sigma", sigma, "NSEG", os.environ.get("PNFFT_B200_NSEG"), "gather F %.2f ms scatter F %.2f ms" % (a, b))
