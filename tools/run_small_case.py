"""Test helper: one trafo (F|GRAD_F) + adj (F|GRAD_F) on seeded inputs in a fresh process, results to an .npz file.
Used where a switch is read once per process (tests/test_gpu_golden.py::test_table_in_column_batches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pnfft_b200 import api as A
from tests.util import Run1, make_inputs

out, m, c2r = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3]))
N, M = (32, 32, 32), 30000
x, fh, f, g = make_inputs(N, M, 61, c2r=c2r)
x[:M // 2] = np.clip(np.random.default_rng(3).normal(-0.2, 0.05, (M // 2, 3)), -0.5, np.nextafter(0.5, 0.0))
run = Run1(N, x, m=m, c2r=c2r)
fo, go = run.trafo(fh, A.COMPUTE_F | A.COMPUTE_GRAD_F)
fho = run.adj(f, g, A.COMPUTE_F | A.COMPUTE_GRAD_F)
run.close()
np.savez(out, f=fo, g=go, fh=fho)
