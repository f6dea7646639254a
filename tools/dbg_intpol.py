import os, sys, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnfft_b200 import api as A
from tests.util import Run1, rel_l2
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for p in sorted(glob.glob(os.path.join(G, "i_*.npz"))):
    g = np.load(p)
    for variant in (0, 1):
        run = Run1(tuple(g["N"]), g["x"], m=int(g["m"]), flags=int(g["flags"]), c2r=bool(g["c2r"]), variant=variant)
        f, gr, h = run.trafo_hessian(g["f_hat"], 7)
        fh = run.adj(g["f"], g["grad_f"], 3)
        run.close()
        run0 = Run1(tuple(g["N"]), g["x"], m=int(g["m"]), flags=int(g["flags"]) & ~0x3c, c2r=bool(g["c2r"]), variant=variant)
        f0, _, _ = run0.trafo_hessian(g["f_hat"], 7)
        run0.close()
        print(os.path.basename(p), variant, "f %.2e g %.2e h %.2e fh %.2e | direct-vs-gold %.2e" % (rel_l2(f, g["out_f"]), rel_l2(gr, g["out_grad_f"]), rel_l2(h, g["out_hessian_f"]), rel_l2(fh, g["out_f_hat"]), rel_l2(f0, g["out_f"])))
