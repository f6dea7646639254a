import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnfft_b200 import api as A
from tests.util import Run1, make_inputs, rel_l2
from oracle import checker
ref = checker.get()
N = (16, 16, 16)
for M in [300, 1000, 3000, 10000]:
    x, fh, f, g = make_inputs(N, M, 1)
    r = ref.trafo(N, x, fh, compute_flags=1)
    run = Run1(N, x, variant=0)
    cell = np.floor(32 * x).astype(int) + 16
    tile = ((cell[:, 0] // 16) * 8 + cell[:, 1] // 4) * 4 + cell[:, 2] // 8
    order = np.argsort(tile, kind="stable")
    pos = np.zeros(M, int)
    start = {}
    for p_, j in enumerate(order):
        t = tile[j]
        if t not in start: start[t] = p_
        pos[j] = p_ - start[t]
    cnt = np.bincount(tile, minlength=256)
    for it in range(2):
        fo, go = run.trafo(fh, 1)
        err = np.abs(fo - r["f"]) / np.abs(r["f"]).max()
        bad = np.where(err > 1e-12)[0]
        print("M", M, "rel %.2e" % rel_l2(fo, r["f"]), "bad", len(bad))
        if it == 0:
            for j in bad[:25]:
                c = cell[j]
                print("   j", j, "dx,dy,dz", c[0] % 16, c[1] % 4, c[2] % 8, "tz", c[2] // 8, "pos", pos[j], "batch", pos[j] // 16, "of cnt", cnt[tile[j]], "err %.1e" % err[j])
    run.close()
