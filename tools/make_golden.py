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


INTERLACED, TRANSPOSED, ACC = 1 << 8, 1 << 11, 1 << 4


def extra_cases(out):
    """Round-2 fixtures: PNFFT_INTERLACED (all windows; nodes that fold in the shifted pass included), PNFFT_TRANSPOSED_F_HAT
    (f_hat stored here in natural order; the layout is the caller's business), truncated torus x_max < 0.5 (pruned FFT
    output no < n), PNFFT_COMPUTE_ACCUMULATED for trafo and adj.  Optional keys: n, x_max, acc, f0, grad_f0, f_hat0."""
    names = []
    seed = 500

    def one(name, N, M, m, flags, c2r, single, n=None, x_max=(0.5, 0.5, 0.5), acc=False, edge=False):
        nonlocal seed
        seed += 1
        ref = refdrv.get(single)
        x, fh, f, g = inputs(N, M, seed, c2r, single)
        rdt = np.float32 if single else np.float64
        xm = np.array(x_max)
        x = (x * (2 * xm)).astype(rdt)                         # nodes inside [-x_max, x_max)
        if edge:     # within half a mesh width of the upper border: they fold in the shifted pass.  (Not one ulp below the
            # border: there the reference's own Kaiser-Bessel derivative cancels catastrophically)
            for t in range(3):
                x[:6, t] = rdt(0.5) - rdt(0.3) / (2 * N[t])
            x[6:9, 0] = rdt(0.5) - rdt(0.2) / (2 * N[0]); x[9:12, 2] = rdt(0.5) - rdt(0.01) / (2 * N[2])
        n_ = tuple(n) if n is not None else tuple(2 * v for v in N)
        kw = dict(n=n_, m=m, pnfft_flags=flags, c2r=c2r, x_max=tuple(x_max))
        extra = {}
        if acc:
            rng = np.random.default_rng(seed + 7000)
            f0 = (f[::-1] * 0.5).copy(); g0 = (g[::-1] * 0.25).copy()
            fh0 = (fh[::-1, ::-1, ::-1] * 0.125).copy()
            rt = ref.trafo(N, x, fh, f=f0, grad_f=g0, compute_flags=3 | ACC, **kw)
            ra = ref.adj(N, x, f=f, grad_f=g, f_hat=fh0, compute_flags=3 | ACC, **kw)
            extra = dict(acc=True, f0=f0, grad_f0=g0, f_hat0=fh0)
        else:
            rt = ref.trafo(N, x, fh, compute_flags=3, **kw)
            ra = ref.adj(N, x, f=f, grad_f=g, compute_flags=3, **kw)
        psi, dpsi = ref.probe_tensor(x[:16], N, n=n_, m=m, x_max=tuple(x_max), pnfft_flags=flags & ~INTERLACED)
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), n=np.array(n_), x_max=xm, m=m, flags=flags, c2r=c2r,
                            single=single, x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"],
                            out_f_hat=ra["f_hat"], psi=psi, dpsi=dpsi, **extra)
        names.append(name)

    N, M = (8, 12, 10), 120
    for win, wf in WIN.items():
        m = 6 if win == "kaiser_bessel" else 5
        one("t_il_%s_ad_c2c_m%d_d" % (win, m), N, M, m, wf | INTERLACED, False, False, edge=True)
    one("t_il_kaiser_bessel_ad_c2c_m4_d", N, M, 4, INTERLACED, False, False, edge=True)
    one("t_il_kaiser_bessel_ik_c2c_m6_d", N, M, 6, INTERLACED | DIFF_IK, False, False, edge=True)
    one("t_il_gaussian_ik_c2c_m5_d", N, M, 5, WIN["gaussian"] | INTERLACED | DIFF_IK, False, False, edge=True)
    one("t_il_kaiser_bessel_ad_c2r_m6_d", N, M, 6, INTERLACED, True, False, edge=True)
    one("t_il_bspline_ad_c2r_m5_d", N, M, 5, WIN["bspline"] | INTERLACED, True, False, edge=True)
    one("t_il_kaiser_bessel_ad_c2c_m6_f", N, M, 6, INTERLACED, False, True, edge=True)
    one("t_il_kaiser_bessel_ad_c2r_m4_f", N, M, 4, INTERLACED, True, True, edge=True)
    one("t_tr_kaiser_bessel_ad_c2c_m6_d", N, M, 6, TRANSPOSED, False, False)
    one("t_tr_kaiser_bessel_ik_c2c_m6_d", N, M, 6, TRANSPOSED | DIFF_IK, False, False)
    one("t_tr_kaiser_bessel_ad_c2r_m6_d", N, M, 6, TRANSPOSED, True, False)
    one("t_tr_gaussian_ad_c2c_m5_f", N, M, 5, WIN["gaussian"] | TRANSPOSED, False, True)
    one("t_tr_il_kaiser_bessel_ad_c2c_m4_d", N, M, 4, TRANSPOSED | INTERLACED, False, False, edge=True)
    Nt, nt, xm = (24, 32, 20), (48, 64, 40), (0.3, 0.25, 0.5)
    one("t_torus_kaiser_bessel_ad_c2c_m4_d", Nt, 300, 4, 0, False, False, n=nt, x_max=xm)
    one("t_torus_kaiser_bessel_ik_c2c_m4_d", Nt, 300, 4, DIFF_IK, False, False, n=nt, x_max=xm)
    one("t_torus_gaussian_ad_c2r_m4_d", Nt, 300, 4, WIN["gaussian"], True, False, n=nt, x_max=xm)
    one("t_torus_il_kaiser_bessel_ad_c2c_m4_d", Nt, 300, 4, INTERLACED, False, False, n=nt, x_max=xm)
    one("t_torus_tr_kaiser_bessel_ad_c2c_m4_f", Nt, 300, 4, TRANSPOSED, False, True, n=nt, x_max=xm)
    one("t_acc_kaiser_bessel_ad_c2c_m6_d", N, M, 6, 0, False, False, acc=True)
    one("t_acc_kaiser_bessel_ik_c2c_m6_d", N, M, 6, DIFF_IK, False, False, acc=True)
    one("t_acc_bspline_ad_c2r_m5_d", N, M, 5, WIN["bspline"], True, False, acc=True)
    one("t_acc_il_kaiser_bessel_ad_c2c_m4_d", N, M, 4, INTERLACED, False, False, acc=True, edge=True)
    return names


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
    cases += extra_cases(out)
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
    # round 2: the same with PNFFT_TRANSPOSED_F_HAT and PNFFT_INTERLACED (one more ghost cell above does not move a border)
    lay = {}
    for mesh in [(1, 1), (1, 2), (2, 2), (2, 4)]:
        for tag, N, n, xm, fl in [("tr_even", (16, 16, 16), (32, 32, 32), (0.5, 0.5, 0.5), TRANSPOSED),
                                  ("tr_ragged", (20, 12, 16), (48, 26, 32), (0.5, 0.5, 0.5), TRANSPOSED),
                                  ("tr_il_torus", (16, 16, 16), (32, 32, 32), (0.3, 0.25, 0.5), TRANSPOSED | INTERLACED)]:
            for c2r in (False, True):
                L = ref.layout(N, n, m=4, np_mesh=mesh, x_max=xm, c2r=c2r, pnfft_flags=fl)
                key = "%s_%dx%d_%s" % (tag, mesh[0], mesh[1], "c2r" if c2r else "c2c")
                lay[key + "_cfg"] = np.array(list(N) + list(n) + [4, int(c2r)] + list(mesh) + [fl], np.int64)
                lay[key + "_xmax"] = np.array(xm)
                for k in ("local_N", "local_N_start", "local_no", "local_no_start", "lo", "up"):
                    lay[key + "_" + k] = L[k]
                lay[key + "_no"] = np.array(L["no"])
    np.savez_compressed(os.path.join(out, "layouts_r2.npz"), **lay)
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


def hessian_cases(out):
    """PNFFT_COMPUTE_HESSIAN_F (trafo only): analytic window second derivatives for every window, ik differentiation,
    c2c and c2r, double and float; the six components in the reference's order xx, xy, xz, yy, yz, zz."""
    names = []
    seed = 900
    for single in (False, True):
        ref = refdrv.get(single)
        for win, wf in WIN.items():
            for ik in (0, DIFF_IK):
                for c2r in (False, True):
                    if c2r and ik:
                        continue
                    if single and (win not in ("kaiser_bessel", "gaussian") or c2r):
                        continue
                    m = 6 if win == "kaiser_bessel" else 5
                    N, M = (8, 12, 10), 100
                    seed += 1
                    x, fh, _, _ = inputs(N, M, seed, c2r, single)
                    flags = wf | ik
                    rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=7, c2r=c2r)
                    name = "h_%s_%s_%s_m%d_%s" % (win, "ik" if ik else "ad", "c2r" if c2r else "c2c", m, "f" if single else "d")
                    np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, flags=flags, c2r=c2r, single=single,
                                        x=x, f_hat=fh, out_f=rt["f"], out_grad_f=rt["grad_f"], out_hessian_f=rt["hessian_f"])
                    names.append(name)
    print("wrote %d Hessian cases" % len(names))
    return names


def intpol_cases(out):
    """PNFFT_PRE_{CONST,LIN,QUAD,CUB}_PSI: window values (and their first / second derivatives) interpolated from the
    reference's lookup tables; trafo with f, grad_f and hessian_f, adjoint with f and grad_f."""
    names = []
    seed = 1200
    ref = refdrv.get(False)
    # (no PNFFT_PRE_QUAD_PSI fixtures: bit 4 of the plan flags is also read as the node flag PNFFT_REAL_F inside the
    # reference's node loop, kernel/ndft-parallel.c:2804-2838, so its quadratic runs drop the imaginary parts / return NaN)
    for order, fl in (("const", 1 << 2), ("lin", 1 << 3), ("cub", 1 << 5)):
        for win in ("kaiser_bessel", "gaussian", "bspline"):
            for c2r in (False, True):
                if c2r and win != "kaiser_bessel":
                    continue
                m = 6 if win == "kaiser_bessel" else 4
                N, M = (8, 12, 10), 100
                seed += 1
                x, fh, f, g = inputs(N, M, seed, c2r, False)
                flags = WIN[win] | fl
                rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=7, c2r=c2r)
                ra = ref.adj(N, x, f=f, grad_f=g, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
                name = "i_%s_%s_%s_m%d_d" % (order, win, "c2r" if c2r else "c2c", m)
                np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, flags=flags, c2r=c2r, single=False,
                                    x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"],
                                    out_hessian_f=rt["hessian_f"], out_f_hat=ra["f_hat"])
                names.append(name)
    print("wrote %d interpolation cases" % len(names))
    return names


def gauss_t_cases(out):
    """PNFFT_WINDOW_GAUSSIAN_T: the Gaussian window with the Fourier coefficients of its truncation to the cutoff (D matrix
    only, reference kernel/matrix_D.c:37-65, 195-217; libcerf's cerf inside the compiled reference)."""
    GT = WIN["gaussian"] | (1 << 17)
    names = []
    seed = 1500
    for name, m, flags, c2r, single in [("t_gaussian_t_ad_c2c_m5_d", 5, GT, False, False),
                                        ("t_gaussian_t_ik_c2c_m5_d", 5, GT | DIFF_IK, False, False),
                                        ("t_gaussian_t_ad_c2r_m4_d", 4, GT, True, False),
                                        ("t_fast_gaussian_t_ad_c2c_m6_d", 6, GT | (1 << 1), False, False),
                                        ("t_il_gaussian_t_ad_c2c_m5_d", 5, GT | INTERLACED, False, False),
                                        ("t_gaussian_t_ad_c2c_m5_f", 5, GT, False, True)]:
        ref = refdrv.get(single)
        N, M = (8, 12, 10), 120
        seed += 1
        x, fh, f, g = inputs(N, M, seed, c2r, single)
        rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
        ra = ref.adj(N, x, f=f, grad_f=g, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
        psi, dpsi = ref.probe_tensor(x[:16], N, m=m, pnfft_flags=flags & ~INTERLACED)
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, flags=flags, c2r=c2r, single=single, x=x,
                            f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"], out_f_hat=ra["f_hat"], psi=psi, dpsi=dpsi)
        names.append(name)
    print("wrote %d truncated-Gaussian cases" % len(names))
    return names


def set_b_cases(out):
    """pnfft_set_b (api/api-basic.c:587-596): other window shape parameters than the plan's default, set after the plan is
    made -- the reference recomputes 1/phi_hat, the fast-Gaussian constants and the interpolation tables."""
    names = []
    seed = 1700
    for name, win, m, extra, c2r, single, b in [
            ("b_kaiser_bessel_c2c_m6_d", "kaiser_bessel", 6, 0, False, False, (5.0, 5.2, 4.9)),
            ("b_kaiser_bessel_c2c_m8_d", "kaiser_bessel", 8, 0, False, False, (4.0, 4.2, 5.2)),
            ("b_kaiser_bessel_c2r_m4_d", "kaiser_bessel", 4, 0, True, False, (4.4, 4.5, 4.9)),
            ("b_kaiser_bessel_c2c_m6_f", "kaiser_bessel", 6, 0, False, True, (5.0, 5.2, 4.9)),
            ("b_kaiser_bessel_cub_c2c_m6_d", "kaiser_bessel", 6, 1 << 5, False, False, (5.0, 5.2, 4.9)),
            ("b_gaussian_c2c_m5_d", "gaussian", 5, 0, False, False, (2.0, 2.3, 2.6)),
            ("b_fast_gaussian_c2c_m5_d", "fast_gaussian", 5, 0, False, False, (2.0, 2.3, 2.6)),
            ("b_gaussian_t_c2c_m5_d", "gaussian", 5, 1 << 17, False, False, (2.0, 2.3, 2.6)),
            ("b_bessel_i0_c2c_m5_d", "bessel_i0", 5, 0, False, False, (4.5, 4.9, 5.1)),
            ("b_sinc_power_ik_c2c_m5_d", "sinc_power", 5, DIFF_IK, False, False, (6.0, 6.5, 7.0))]:
        ref = refdrv.get(single)
        N, M = (8, 12, 10), 120
        seed += 1
        x, fh, f, g = inputs(N, M, seed, c2r, single)
        flags = WIN[win] | extra
        rt = ref.trafo(N, x, fh, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r, b=b)
        ra = ref.adj(N, x, f=f, grad_f=g, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r, b=b)
        assert tuple(rt["b"]) == tuple(np.asarray(b, np.float32 if single else np.float64).astype(np.float64))
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), m=m, flags=flags, c2r=c2r, single=single, b=np.array(b),
                            x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"], out_f_hat=ra["f_hat"])
        names.append(name)
    print("wrote %d pnfft_set_b cases" % len(names))
    return names


def combination_cases(out):
    """Flag / size combinations the other groups leave out: Hessians with interlacing, truncated torus and transposed
    f_hat; oversampling factors other than 2 (n != 2N, different per axis); cutoffs m = 2, 3, 5, 7, 8, 10 against the
    reference itself; PNFFT_PRE_PSI with interlacing / torus; three-flag mixes.  Optional keys: out_hessian_f (trafo ran
    with PNFFT_COMPUTE_HESSIAN_F), pre (precompute flags), cf (compute flags of trafo and adj, default F | GRAD_F)."""
    names = []
    seed = 2000
    Ns, Nt, nt, xm = (8, 12, 10), (24, 32, 20), (48, 64, 40), (0.3, 0.25, 0.5)
    KB, GA, BS = 0, WIN["gaussian"], WIN["bspline"]
    cases = [
        # name, N, n, x_max, m, flags, c2r, single, hessian, pre, cf
        ("c_hess_il_kaiser_bessel_c2c_m6_d", Ns, None, None, 6, KB | INTERLACED, False, False, True, 0, 3),
        ("c_hess_il_ik_kaiser_bessel_c2c_m6_d", Ns, None, None, 6, KB | INTERLACED | DIFF_IK, False, False, True, 0, 3),
        ("c_hess_torus_kaiser_bessel_c2c_m4_d", Nt, nt, xm, 4, KB, False, False, True, 0, 3),
        ("c_hess_tr_gaussian_c2c_m5_d", Ns, None, None, 5, GA | TRANSPOSED, False, False, True, 0, 3),
        ("c_hess_torus_bspline_c2r_m4_d", Nt, nt, xm, 4, BS, True, False, True, 0, 3),
        ("c_sigma_kaiser_bessel_c2c_m4_d", Ns, (20, 18, 24), None, 4, KB, False, False, False, 0, 3),
        ("c_sigma_kaiser_bessel_c2r_m4_d", Ns, (20, 18, 24), None, 4, KB, True, False, False, 0, 3),
        ("c_sigma_kaiser_bessel_ik_c2c_m6_d", (12, 12, 12), (18, 32, 20), None, 6, KB | DIFF_IK, False, False, False, 0, 3),
        ("c_sigma_bspline_c2c_m5_f", Ns, (20, 26, 24), None, 5, BS, False, True, False, 0, 3),
        ("c_sigma_gaussian_il_c2c_m5_d", Ns, (24, 20, 28), None, 5, GA | INTERLACED, False, False, False, 0, 3),
        ("c_il_torus_kaiser_bessel_c2r_m4_d", Nt, nt, xm, 4, KB | INTERLACED, True, False, False, 0, 3),
        ("c_ik_torus_tr_kaiser_bessel_c2c_m4_d", Nt, nt, xm, 4, KB | DIFF_IK | TRANSPOSED, False, False, False, 0, 3),
        ("c_fast_gaussian_il_c2r_m5_d", Ns, None, None, 5, WIN["fast_gaussian"] | INTERLACED, True, False, False, 0, 3),
        ("c_bessel_i0_tr_ik_c2c_m5_d", Ns, None, None, 5, WIN["bessel_i0"] | TRANSPOSED | DIFF_IK, False, False, False, 0, 3),
        ("c_sinc_power_torus_c2c_m4_d", Nt, nt, xm, 4, WIN["sinc_power"], False, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m2_d", (16, 16, 16), None, None, 2, KB, False, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m3_d", (16, 16, 16), None, None, 3, KB, False, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m5_d", (16, 16, 16), None, None, 5, KB, False, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m7_d", (16, 16, 16), None, None, 7, KB, False, False, False, 0, 3),
        ("c_kaiser_bessel_c2r_m7_d", (16, 16, 16), None, None, 7, KB, True, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m8_d", (16, 16, 16), None, None, 8, KB, False, False, False, 0, 3),
        ("c_gaussian_c2r_m8_d", (16, 16, 16), None, None, 8, GA, True, False, False, 0, 3),
        ("c_bspline_il_c2c_m8_d", (16, 16, 16), None, None, 8, BS | INTERLACED, False, False, False, 0, 3),
        ("c_kaiser_bessel_c2c_m10_d", (16, 16, 16), None, None, 10, KB, False, False, False, 0, 3),
        # (no float Kaiser-Bessel case at m = 8: the float reference returns NaN there -- I0(8 b)^-3 underflows in D while
        # sinh(8 b)^3 overflows in B)
        ("c_gaussian_c2c_m8_f", (16, 16, 16), None, None, 8, GA, False, True, False, 0, 3),
        ("c_pre_il_kaiser_bessel_c2c_m6_d", Ns, None, None, 6, KB | INTERLACED, False, False, False, 2, 1),
        ("c_pre_torus_gaussian_c2r_m4_d", Nt, nt, xm, 4, GA, True, False, False, 2, 1),
        ("c_pre_tr_kaiser_bessel_c2c_m8_d", (16, 16, 16), None, None, 8, KB | TRANSPOSED, False, False, False, 2, 1),
    ]
    for name, N, n, x_max, m, flags, c2r, single, hess, pre, cf in cases:
        ref = refdrv.get(single)
        seed += 1
        M = 150
        x, fh, f, g = inputs(N, M, seed, c2r, single)
        rdt = np.float32 if single else np.float64
        x_max_ = tuple(x_max) if x_max is not None else (0.5, 0.5, 0.5)
        x = (x * (2 * np.array(x_max_))).astype(rdt)
        n_ = tuple(n) if n is not None else tuple(2 * v for v in N)
        kw = dict(n=n_, m=m, pnfft_flags=flags, c2r=c2r, x_max=x_max_, precompute_flags=pre)
        rt = ref.trafo(N, x, fh, compute_flags=cf | (4 if hess else 0), **kw)
        ra = ref.adj(N, x, f=f, grad_f=g, compute_flags=cf, **kw)
        extra = dict(out_hessian_f=rt["hessian_f"]) if hess else {}
        for k in ("f", "grad_f"):
            assert np.all(np.isfinite(rt[k])), (name, k)
        assert np.all(np.isfinite(ra["f_hat"])), name
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), n=np.array(n_), x_max=np.array(x_max_), m=m, flags=flags,
                            c2r=c2r, single=single, pre=pre, cf=cf, x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"],
                            out_grad_f=rt["grad_f"], out_f_hat=ra["f_hat"], **extra)
        names.append(name)
    print("wrote %d combination cases" % len(names))
    return names


def direct_cases(out):
    """PNFFT_COMPUTE_DIRECT (reference kernel/ndft-parallel.c:377-722): the direct NDFT and its adjoint, f / grad_f / hessian_f,
    c2c and c2r, transposed f_hat, float, PNFFT_COMPUTE_ACCUMULATED (ignored by the direct trafo, honoured by the adjoint)."""
    names = []
    seed = 2500
    Ns, Nt, nt, xm = (8, 12, 10), (12, 16, 10), (24, 32, 20), (0.3, 0.25, 0.5)
    for name, N, n, x_max, flags, c2r, single, acc in [
            ("d_c2c_d", Ns, None, None, 0, False, False, False),
            ("d_tr_c2c_d", Ns, None, None, TRANSPOSED, False, False, False),
            ("d_c2r_d", Ns, None, None, 0, True, False, False),
            ("d_tr_c2r_d", Ns, None, None, TRANSPOSED, True, False, False),
            ("d_torus_c2c_d", Nt, nt, xm, 0, False, False, False),
            ("d_acc_c2c_d", Ns, None, None, 0, False, False, True),
            ("d_acc_c2r_d", Ns, None, None, 0, True, False, True),
            ("d_c2c_f", Ns, None, None, 0, False, True, False),
            ("d_c2r_f", Ns, None, None, 0, True, True, False)]:
        ref = refdrv.get(single)
        seed += 1
        M = 60
        x, fh, f, g = inputs(N, M, seed, c2r, single)
        rdt = np.float32 if single else np.float64
        x_max_ = tuple(x_max) if x_max is not None else (0.5, 0.5, 0.5)
        x = (x * (2 * np.array(x_max_))).astype(rdt)
        n_ = tuple(n) if n is not None else tuple(2 * v for v in N)
        kw = dict(n=n_, m=4, pnfft_flags=flags, c2r=c2r, x_max=x_max_)
        extra = {}
        if acc:
            f0 = (f[::-1] * 0.5).copy(); g0 = (g[::-1] * 0.25).copy(); fh0 = (fh[::-1, ::-1, ::-1] * 0.125).copy()
            rt = ref.trafo(N, x, fh, f=f0, grad_f=g0, compute_flags=7 | 8 | ACC, **kw)
            ra = ref.adj(N, x, f=f, grad_f=g, f_hat=fh0, compute_flags=3 | 8 | ACC, **kw)
            extra = dict(acc=True, f0=f0, grad_f0=g0, f_hat0=fh0)
        else:
            rt = ref.trafo(N, x, fh, compute_flags=7 | 8, **kw)
            ra = ref.adj(N, x, f=f, grad_f=g, compute_flags=3 | 8, **kw)
        np.savez_compressed(os.path.join(out, name + ".npz"), N=np.array(N), n=np.array(n_), x_max=np.array(x_max_), m=4, flags=flags,
                            c2r=c2r, single=single, x=x, f_hat=fh, f=f, grad_f=g, out_f=rt["f"], out_grad_f=rt["grad_f"],
                            out_hessian_f=rt["hessian_f"], out_f_hat=ra["f_hat"], **extra)
        names.append(name)
    print("wrote %d direct NDFT cases" % len(names))
    return names


if __name__ == "__main__":
    gold = os.path.join(ROOT, "tests", "golden")
    if len(sys.argv) > 1 and sys.argv[1] == "--hessian":       # round-2 additions: leave the other fixtures untouched
        hessian_cases(gold)
    elif len(sys.argv) > 1 and sys.argv[1] == "--intpol":
        intpol_cases(gold)
    elif len(sys.argv) > 1 and sys.argv[1] == "--gauss-t":
        gauss_t_cases(gold)
    elif len(sys.argv) > 1 and sys.argv[1] == "--set-b":
        set_b_cases(gold)
    elif len(sys.argv) > 1 and sys.argv[1] == "--combinations":
        combination_cases(gold)
    elif len(sys.argv) > 1 and sys.argv[1] == "--direct":
        direct_cases(gold)
    else:
        main()
        hessian_cases(gold)
        intpol_cases(gold)
        gauss_t_cases(gold)
        set_b_cases(gold)
        combination_cases(gold)
        direct_cases(gold)
