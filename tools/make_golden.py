"""Generate the golden vectors of tests/golden/ from the UNMODIFIED reference PNFFT compiled out of /root/reference
(oracle/_ref, built by `make -C oracle ref`).  Run in the build container, where /root/reference exists:

    python tools/make_golden.py

Every case stores its inputs (seeded) and the reference's outputs, so that the fixtures stay valid on machines
without /root/reference (the GPU box) and pin both oracle/pnfft_oracle.c and the CUDA path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refdrv  # noqa: E402

WIN = {"kaiser_bessel": 0, "gaussian": 1 << 13, "fast_gaussian": (1 << 13) | (1 << 1), "bspline": 1 << 14,
       "sinc_power": 1 << 15, "bessel_i0": 1 << 16}
DIFF_IK = 1 << 12


def inputs(N, M, seed, c2r, single):
    rng = np.random.default_rng(seed)
    rdt = np.float32 if single else np.float64
    cdt = np.complex64 if single else np.complex128
    x = np.clip(rng.uniform(-0.5, 0.5, (M, 3)).astype(rdt), -0.5, np.nextafter(rdt(0.5), rdt(0)))
    Nc = (N[0], N[1], N[2] // 2 + 1) if c2r else tuple(N)
    fh = (rng.uniform(-1, 1, Nc) + 1j * rng.uniform(-1, 1, Nc)).astype(cdt)
    if c2r:
        f = rng.uniform(-1, 1, M).astype(rdt)
        g = rng.uniform(-1, 1, (M, 3)).astype(rdt)
    else:
        f = (rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)).astype(cdt)
        g = (rng.uniform(-1, 1, (M, 3)) + 1j * rng.uniform(-1, 1, (M, 3))).astype(cdt)
    return x, fh, f, g


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    cases = []
    seed = 100
    # transforms: every window, both gradient modes, c2c/c2r, double/float, two cutoffs
    for single in (False, True):
        ref = refdrv.get(single)
        for c2r in (False, True):
            for win, wf in WIN.items():
                for ik in (0, DIFF_IK):
                    if c2r and ik:
                        continue   # the reference's ik path is complex-input only (api/api-basic.c:100,249)
                    for m in ((4, 6) if win == "kaiser_bessel" else (5,)):
                        N, M = (8, 12, 10), 120
                        seed += 1
                        x, fh, f, g = inputs(N, M, seed, c2r, single)
                        flags = wf | ik
                        rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
                        ra = ref.adj(N, x, f=f, grad_f=g, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
                        psi, dpsi = ref.probe_tensor(x[:16], N, m=m, pnfft_flags=flags)
                        name = "t_%s_%s_%s_m%d_%s" % (win, "ik" if ik else "ad", "c2r" if c2r else "c2c", m, "f" if single else "d")
                        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, flags=flags, c2r=c2r,
                                            single=single, x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"],
                                            out_f_hat=ra["f_hat"], psi=psi, dpsi=dpsi)
                        cases.append(name)
    # integer work: layouts over process meshes, node -> rank / grid index assignment, sort keys
    ref = refdrv.get(False)
    lay = {}
    for mesh in [(1, 1), (1, 2), (2, 2), (2, 4)]:
        for tag, N, n, xm in [("even", (16, 16, 16), (32, 32, 32), (0.5, 0.5, 0.5)),
                              ("ragged", (20, 12, 16), (48, 26, 32), (0.5, 0.5, 0.5)),
                              ("torus", (16, 16, 16), (32, 32, 32), (0.3, 0.25, 0.5))]:
            for c2r in (False, True):
                L = ref.layout(N, n, m=4 if tag != "even" else 6, np_mesh=mesh, x_max=xm, c2r=c2r)
                key = "%s_%dx%d_%s" % (tag, mesh[0], mesh[1], "c2r" if c2r else "c2c")
                lay[key + "_cfg"] = np.array(list(N) + list(n) + [4 if tag != "even" else 6, int(c2r)] + list(mesh), np.int64)
                lay[key + "_xmax"] = np.array(xm)
                for k in ("local_N", "local_N_start", "local_no", "local_no_start", "lo", "up"):
                    lay[key + "_" + k] = L[k]
                lay[key + "_no"] = np.array(L["no"])
    np.savez_compressed(os.path.join(out, "layouts.npz"), **lay)
    rng = np.random.default_rng(7)
    N, M = (16, 16, 16), 4000
    x = np.clip(rng.uniform(-0.5, 0.5, (M, 3)), -0.5, np.nextafter(0.5, 0))
    x[:64] = np.round(x[:64] * 32) / 32          # nodes exactly on grid lines
    x[:64] = np.clip(x[:64], -0.5, 0.5 - 1 / 32)
    idx = {}
    for mesh in [(1, 1), (2, 2), (2, 4)]:
        r = ref.run(refdrv.OP_TRAFO, N, x=x, f_hat=np.zeros(N, np.complex128), np_mesh=mesh, want_index=True,
                    compute_flags=refdrv.COMPUTE_F | refdrv.OMIT_FFT | refdrv.OMIT_DECONV)
        idx["owner_%dx%d" % mesh] = r["owner"]
        idx["index_%dx%d" % mesh] = r["node_index"]
    keys, perm = ref.probe_sort(x, N)
    np.savez_compressed(os.path.join(out, "node_index.npz"), N=np.array(N), m=6, x=x, sort_keys=keys, sort_perm=perm, **idx)
    print("wrote %d transform cases + layouts.npz + node_index.npz to %s" % (len(cases), out))


if __name__ == "__main__":
    main()
