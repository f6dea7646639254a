import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnfft_b200 import api as A
from tests.util import Run1, rel_l2, make_inputs
for N in ((8, 12, 10), (16, 16, 16), (8, 16, 16), (16, 12, 16), (16, 16, 10), (16, 16, 12)):
    for flags, nm in ((0, "kb"), (1 << 13, "gauss")):
        for m in (4, 6):
            x, fh, f, g = make_inputs(N, 300, 4)
            out = []
            for variant in (0, 1):
                run = Run1(N, x, m=m, flags=flags, variant=variant)
                fo, go = run.trafo(fh, 3)
                fho = run.adj(f, g, 3)
                run.close()
                out.append((fo, go, fho))
            print(N, nm, m, "f %.1e g %.1e fh %.1e" % tuple(rel_l2(a, b) for a, b in zip(out[0], out[1])))
