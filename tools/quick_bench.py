"""Scratch timing of trafo+adj with device-resident arrays (development aid, not the bench contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnfft_b200 import api as A

Ns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2**24
cf = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m = int(sys.argv[4]) if len(sys.argv) > 4 else 6
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 0
N = (Ns,) * 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
x = (torch.rand((M, 3), generator=g, device=dev, dtype=torch.float64) - 0.5).clamp_(-0.5, 0.5 - 1e-12)
fh = torch.randn(N + (2,), generator=g, device=dev, dtype=torch.float64)
f = torch.zeros((M, 2), device=dev, dtype=torch.float64)
gr = torch.zeros((M, 3, 2), device=dev, dtype=torch.float64)
comm = A.create_procmesh_2d(1, 1)
plan = A.Plan.init_guru(N, tuple(2 * v for v in N), (0.5,) * 3, m, flags, comm)
plan.set_kernel_variant(variant)
nodes = A.Nodes(M, 0); nodes.set_x(x); nodes.set_f(f); nodes.set_grad_f(gr); plan.set_f_hat(fh)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    plan.trafo(nodes, cf); t1 = time.time()
    plan.adj(nodes, cf); t2 = time.time()
    print("iter", it, "trafo %.2f ms adj %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), "pts/s %.3e" % (M / (t2 - t0)))
    print("  trafo", {k: round(v, 3) for k, v in plan.stage_ms(False).items()})
    print("  adj  ", {k: round(v, 3) for k, v in plan.stage_ms(True).items()})
