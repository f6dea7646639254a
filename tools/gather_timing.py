"""Development aid: per-warp cycle accounting of k_gather_zm2 (library built with NVFLAGS+=-DZM2_TIMING)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from pnfft_b200 import api as A
cf = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N, M = (256,) * 3, 1 << 24
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
x = (torch.rand((M, 3), generator=g, device=dev, dtype=torch.float64) - 0.5).clamp_(-0.5, 0.5 - 1e-12)
fh = torch.randn(N + (2,), generator=g, device=dev, dtype=torch.float64)
f = torch.zeros((M, 2), device=dev, dtype=torch.float64); gr = torch.zeros((M, 3, 2), device=dev, dtype=torch.float64)
comm = A.create_procmesh_2d(1, 1)
plan = A.Plan.init_guru(N, tuple(2 * v for v in N), (0.5,) * 3, 6, 0, comm)
nodes = A.Nodes(M, 0); nodes.set_x(x); nodes.set_f(f); nodes.set_grad_f(gr); plan.set_f_hat(fh)
fn = A.lib().pnfft_b200_gather_timing
out = np.zeros(112, np.int64)
fn(None, 1)
plan.trafo(nodes, cf)
fn(None, 1)
plan.trafo(nodes, cf)
fn(out.ctypes.data_as(C.c_void_p), 0)
names = ["wait_full", "wait_pempty", "advance", "node_loop", "arrive+help", "loop_top"]
nblk = (43 if os.environ.get("RPT2") else 52) * 128
print("kernel b_kernel ms", plan.stage_ms(False)["b_kernel"])
print("per-warp cycles per CTA-average (kilo-cycles), columns:", names)
tq = out[:96].reshape(16, 6)
for w in range(6 if os.environ.get("RPT2") else 11):
    itsum, itn, itmin = out[72 + 2 * w], out[72 + 2 * w + 1], out[96 + w]
    print(w, ["%8.1f" % (v / nblk / 1e3) for v in tq[w]], "total %.1f" % (tq[w].sum() / nblk / 1e3),
          "| node iteration: mean %.0f min %d cycles over %d" % (itsum / max(itn, 1), itmin, itn))
