cd $GRAFT_REPO_ROOT
python - <<'PY'
import numpy as np
from pnfft_b200 import api as A
import ctypes as C
comm = A.create_procmesh_2d(1, 1)
for m in (4, 6, 8):
    for name, fl in (("kb", 0), ("gauss", A.WINDOW_GAUSSIAN), ("bspline", A.WINDOW_BSPLINE), ("sinc", A.WINDOW_SINC_POWER), ("i0", A.WINDOW_BESSEL_I0)):
        p = A.Plan.init_guru((32,)*3, (64,)*3, (0.5,)*3, m, fl, comm)
        d = p.P.fn("b200_get_poly_degree", C.c_int, [C.c_void_p])(p.h)
        print(m, name, "deg", d)
        p.finalize(0)
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cf in 1 3; do echo "== variant 0 cf $cf"; timeout 300 python tools/quick_bench.py 256 16777216 $cf 6 0 0 | tail -3; done
