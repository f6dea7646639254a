"""CPU tests of the drop-in boundary: libpnfft_b200.so loads without a GPU, exports every entry point that
include/pnfft.h declares (both precisions), keeps the reference's flag values, and answers the layout queries
(pure integer / one-division host work, reference api/api-guru.c:84-107) identically to the golden reference output.
No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pnfft_b200 import api as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pnfft.h")).read()
    body = src[src.index("#define PNFFT_B200_API"):src.index("#define PNFFT_B200_MANGLE_D")]
    names = set(re.findall(r"PNX\((\w+)\)\s*\(", body))
    names -= {"plan_s", "nodes_s"}
    plain = set(re.findall(r"^\w[\w\s\*]*?\b(pnfft_b200_\w+)\s*\(", src, flags=re.M))
    # the stand-in MPI is exported under pnb_MPI_* only (include/mpi.h renames every entry point), never as MPI_*
    mpi = set("pnb_" + n for n in re.findall(r"^\w[\w\s\*]*?\b(MPI_\w+)\s*\(", open(os.path.join(ROOT, "include", "mpi.h")).read(), flags=re.M))
    mpi |= set(re.findall(r"\b(pfftf?_\w+)\s*\(", open(os.path.join(ROOT, "include", "pfft.h")).read()))
    return names, plain, mpi


def test_library_loads_and_exports_every_declared_symbol():
    lib = A.lib()
    names, plain, mpi = declared_symbols()
    assert len(names) >= 70
    missing = [p + n for n in sorted(names) for p in ("pnfft_", "pnfftf_") if not hasattr(lib, p + n)]
    missing += [n for n in sorted(plain | mpi) if not hasattr(lib, n)]
    assert not missing, missing
    assert len([n for n in mpi if n.startswith("pnb_MPI_")]) >= 18
    # no name of a real MPI library is exported: the .so can share a process with libmpi
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert not re.findall(r" T (MPI_\w+)", out)


def test_flag_values_match_reference_header():
    """ABI constants of reference api/pnfft.h:302-418, parsed from include/pnfft.h."""
    src = open(os.path.join(ROOT, "include", "pnfft.h")).read()
    want = {"PNFFT_PRE_PHI_HAT": 1 << 0, "PNFFT_FAST_GAUSSIAN": 1 << 1, "PNFFT_FG_PSI": 1 << 1, "PNFFT_MALLOC_F_HAT": 1 << 6,
            "PNFFT_INTERLACED": 1 << 8, "PNFFT_TRANSPOSED_F_HAT": 1 << 11, "PNFFT_DIFF_AD": 0, "PNFFT_DIFF_IK": 1 << 12,
            "PNFFT_WINDOW_KAISER_BESSEL": 0, "PNFFT_WINDOW_GAUSSIAN": 1 << 13, "PNFFT_WINDOW_BSPLINE": 1 << 14,
            "PNFFT_WINDOW_SINC_POWER": 1 << 15, "PNFFT_WINDOW_BESSEL_I0": 1 << 16, "PNFFT_SORT_NODES": 1 << 18,
            "PNFFT_MALLOC_X": 1, "PNFFT_MALLOC_F": 2, "PNFFT_MALLOC_GRAD_F": 4, "PNFFT_PRE_PSI": 2, "PNFFT_PRE_GRAD_PSI": 4,
            "PNFFT_COMPUTE_F": 1, "PNFFT_COMPUTE_GRAD_F": 2, "PNFFT_COMPUTE_ACCUMULATED": 16, "PNFFT_OMIT_DECONV": 32,
            "PNFFT_OMIT_FFT": 64, "PNFFT_OMIT_CONV": 128, "PNFFT_TIMER_LENGTH": 10}
    for name, val in want.items():
        mm = re.search(r"#define\s+%s\s+\(?\s*([^\n]+?)\s*\)?\s*(/\*.*)?$" % name, src, flags=re.M)
        assert mm, name
        expr = re.sub(r"(\d)[uU]\b", r"\1", mm.group(1)).strip()
        assert eval(expr, {"__builtins__": {}}, {k: v for k, v in want.items()}) == val, (name, expr)


def test_single_rank_layout_matches_golden_reference():
    L = np.load(os.path.join(GOLD, "layouts.npz"))
    comm = A.create_procmesh_2d(1, 1)
    for key in sorted(k[:-4] for k in L.files if k.endswith("_1x1_c2c_cfg") or k.endswith("_1x1_c2r_cfg")):
        cfg = L[key + "_cfg"]
        N, n, m, c2r = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6]), bool(cfg[7])
        lN, lNs, lo, up = A.local_size_guru(N, n, tuple(L[key + "_xmax"]), m, comm, c2r=c2r)
        assert np.array_equal(lN, L[key + "_local_N"][0]) and np.array_equal(lNs, L[key + "_local_N_start"][0])
        assert np.array_equal(lo, L[key + "_lo"][0]) and np.array_equal(up, L[key + "_up"][0])


def test_single_rank_layout_transposed_interlaced():
    L = np.load(os.path.join(GOLD, "layouts_r2.npz"))
    comm = A.create_procmesh_2d(1, 1)
    keys = sorted(k[:-4] for k in L.files if k.endswith("_1x1_c2c_cfg") or k.endswith("_1x1_c2r_cfg"))
    assert len(keys) == 6
    for key in keys:
        cfg = L[key + "_cfg"]
        N, n, m, c2r, fl = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6]), bool(cfg[7]), int(cfg[10])
        lN, lNs, lo, up = A.local_size_guru(N, n, tuple(L[key + "_xmax"]), m, comm, pnfft_flags=fl, c2r=c2r)
        assert np.array_equal(lN, L[key + "_local_N"][0]) and np.array_equal(lNs, L[key + "_local_N_start"][0])
        assert np.array_equal(lo, L[key + "_lo"][0]) and np.array_equal(up, L[key + "_up"][0])


def test_procmesh_mismatch_is_reported():
    with pytest.raises(ValueError):
        A.create_procmesh_2d(2, 2)      # one rank only: reference util/util.c returns non-zero


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    comm = A.create_procmesh_2d(1, 1)
    with pytest.raises(RuntimeError):
        A.Plan.init_guru((16, 16, 16), (32, 32, 32), (0.5,) * 3, 6, 0, comm)   # NULL + message, never a CPU fallback


@pytest.mark.parametrize("c2r", [0, 1])
@pytest.mark.parametrize("mesh", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("N,n,m", [((16, 16, 16), (32, 32, 32), 6), ((12, 20, 16), (24, 40, 32), 4), ((10, 14, 18), (32, 32, 40), 4)])
def test_fft_self_maps_equal_pack_unpack(N, n, m, mesh, c2r):
    """The chunk a rank sends to itself in a pencil re-distribution moves in ONE composed strided copy (csrc/fftpipe.cuh,
    compose_self_map); on index arrays it must land exactly where pack -> chunk -> unpack puts it, forward and backward,
    for every rank of even and ragged meshes (host-only entry point, no GPU)."""
    fn = A.lib().pnfft_b200_check_self_maps
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t)] + [C.c_int] * 6
    Nv, nv = (C.c_ssize_t * 3)(*N), (C.c_ssize_t * 3)(*n)
    total = 0
    for c0 in range(mesh[0]):
        for c1 in range(mesh[1]):
            r = fn(Nv, nv, m, mesh[0], mesh[1], c0, c1, c2r)
            assert r >= 0, "rank (%d,%d): %s" % (c0, c1, "mismatch" if r == -1 else "self transfer did not compose")
            total += r
    assert total > 0


@pytest.mark.parametrize("target", [0, 40, 400])
@pytest.mark.parametrize("dist", ["uniform", "blob", "empty", "one_chunk"])
def test_column_pieces_partition_the_column(dist, target):
    """Work items of a grid column (csrc/zmarch2.cuh: zm_segment, host-only probe): the pieces of every column are disjoint,
    in order and cover all its nodes; with a node target the pieces of a heavy column weigh about the same (one sub-chunk
    is the granule), a light column stays whole unless `fill` asks otherwise, and items beyond the piece count are empty."""
    fn = A.lib().pnfft_b200_column_piece
    fn.restype = None
    fn.argtypes = [C.POINTER(C.c_int)] + [C.c_int] * 6 + [C.POINTER(C.c_int)] * 2
    rng = np.random.default_rng(9)
    nt2, sub = 37, 8
    if dist == "uniform":
        cnt = rng.integers(0, 6, nt2 * sub)
    elif dist == "blob":
        z = np.arange(nt2)[:, None]
        cnt = rng.poisson(60.0 * np.exp(-0.5 * ((z - 17.3) / 2.5) ** 2) * np.ones((1, sub))).ravel()
    elif dist == "empty":
        cnt = np.zeros(nt2 * sub, int)
    else:
        cnt = np.zeros(nt2 * sub, int); cnt[5 * sub + 3] = 1000
    prefix = np.concatenate([[123], 123 + np.cumsum(cnt)]).astype(np.int32)     # a column in the middle of the table
    total = int(cnt.sum())
    per_chunk = cnt.reshape(nt2, sub).sum(1)
    pv = prefix.ctypes.data_as(C.POINTER(C.c_int))
    for nseg, fill in [(1, 1), (4, 1), (4, 4), (32, 2)]:
        pieces, prev_end = [], 0
        for seg in range(nseg):
            a, b = C.c_int(-1), C.c_int(-1)
            fn(pv, nt2, sub, seg, nseg, target, fill, C.byref(a), C.byref(b))
            a, b = a.value, b.value
            assert 0 <= a <= b <= nt2
            n_in = int(per_chunk[a:b].sum())
            if n_in:
                assert a >= prev_end, "pieces overlap"
                assert int(per_chunk[prev_end:a].sum()) == 0, "nodes between two pieces"
                prev_end = b
                pieces.append(n_in)
        assert sum(pieces) == total
        if target > 0 and nseg > 1 and total:
            want = min(nseg, max(fill, -(-total // target)))
            assert len(pieces) <= want
            if want > 1 and dist == "blob":
                assert max(pieces) <= total / want + per_chunk.max()       # equal counts up to one sub-chunk
            if want == 1:
                assert pieces == [total]                                   # a light column stays whole


def test_direct_ndft_blocks_match_golden_layouts():
    """The direct NDFT (csrc/direct.cuh) walks over every rank's f_hat block: broadcast in pnfft_trafo, reduced to its owner in
    pnfft_adj (reference kernel/ndft-parallel.c:413-424, 630-641).  The block geometry it derives for rank pid of a mesh --
    host code, probed through pnfft_b200_direct_block -- equals the reference's local_N / local_N_start of that rank for
    1x1 .. 2x4 meshes, even / ragged sizes, c2c / c2r, natural and PNFFT_TRANSPOSED_F_HAT order (memory order k1, k2, k0)."""
    import ctypes as C
    fn = A.lib().pnfft_b200_direct_block
    fn.restype = None
    I3 = C.c_ssize_t * 3
    fn.argtypes = [I3, I3, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int * 12]
    checked = 0
    for fname in ("layouts.npz", "layouts_r2.npz"):
        L = np.load(os.path.join(GOLD, fname))
        for key in sorted(k[:-4] for k in L.files if k.endswith("_cfg")):
            cfg = L[key + "_cfg"]
            N, n, m, c2r, mesh = [int(v) for v in cfg[0:3]], [int(v) for v in cfg[3:6]], int(cfg[6]), int(cfg[7]), (int(cfg[8]), int(cfg[9]))
            flags = int(cfg[10]) if len(cfg) > 10 else 0
            ax = (1, 2, 0) if flags & A.TRANSPOSED_F_HAT else (0, 1, 2)
            lN, lNs = L[key + "_local_N"].reshape(-1, 3), L[key + "_local_N_start"].reshape(-1, 3)
            for pid in range(mesh[0] * mesh[1]):
                out = (C.c_int * 12)()
                fn(I3(*N), I3(*n), m, mesh[0], mesh[1], pid, flags, c2r, out)
                o = list(out)
                assert o[6:9] == list(ax), key
                assert o[0:3] == [int(lN[pid][a]) for a in ax], (key, pid)
                assert o[3:6] == [int(lNs[pid][a]) for a in ax], (key, pid)
                assert o[9:12] == [N[a] for a in ax], key
                checked += 1
    assert checked >= 48 * 3


def test_input_generators_match_reference():
    """pnfft_init_x_3d, pnfft_init_x_3d_adv and pnfft_init_f (reference api/api-basic.c:681-718, api/api-adv.c:35-85) draw from
    rand(): with the same seed the product's host generators return the compiled reference's arrays bit for bit (the reference's
    drivers seed with srand(myrank) and compare transforms on these inputs)."""
    import ctypes as C
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libpnfft_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    libc, ref, L = C.CDLL("libc.so.6"), C.CDLL(ref_so), A.lib()
    M = 1000
    D3 = C.c_double * 3
    lo, up, xm = D3(-0.3, -0.25, 0.1), D3(0.0, 0.25, 0.5), D3(0.3, 0.25, 0.5)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for seed, call in [(5, lambda lib, out: lib.pnfft_init_x_3d(lo, up, C.c_ssize_t(M), P(out))),
                       (7, lambda lib, out: lib.pnfft_init_x_3d_adv(lo, up, xm, C.c_ssize_t(M), P(out))),
                       (9, lambda lib, out: lib.pnfft_init_f(C.c_ssize_t(M), P(out)))]:
        a, b = np.zeros((M, 3)), np.zeros((M, 3))
        libc.srand(seed); call(L, a)
        libc.srand(seed); call(ref, b)
        assert a.any() and np.array_equal(a, b)


def test_timer_utilities_match_reference():
    """pnfft_timer_add / _copy / _average (reference kernel/timer.c:57-110): same arrays as the compiled reference, including
    the iteration count left in slot 0 by the average and the untouched array when nothing was counted."""
    import ctypes as C
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libpnfft_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    D = C.c_double * 10
    res = []
    for lib in (A.lib(), C.CDLL(ref_so)):
        lib.pnfft_timer_add.restype = C.POINTER(C.c_double)
        lib.pnfft_timer_copy.restype = C.POINTER(C.c_double)
        a, b = D(3, 1.5, 0.25, 7, 8, 9, 1, 2, 3, 4), D(2, 0.5, 0.75, 1, 1, 1, 1, 1, 1, 1)
        s, c = lib.pnfft_timer_add(a, b), lib.pnfft_timer_copy(b)
        out = [s[i] for i in range(10)] + [c[i] for i in range(10)]
        lib.pnfft_timer_average(a)
        z = D(*([0.0] * 10)); z[3] = 5.0
        lib.pnfft_timer_average(z)
        res.append(out + list(a) + list(z))
    assert res[0] == res[1]
    assert res[0][20] == 3.0 and res[0][21] == 0.5


@pytest.mark.parametrize("args", [
    "",
    "-pnfft_N 8 12 10 -pnfft_m 4 -pnfft_window 0 -pnfft_np 1 2 1",
    "-pnfft_window 5 -pnfft_intpol 3 -pnfft_interlaced 1 -pnfft_diff_ik 1 -pnfft_tr_f_hat 1 -pnfft_x_max 0.3 0.25 0.5 -pnfft_local_M 77 -pnfft_n 20 30 24",
    "-pnfft_window 2 -pnfft_fast_gaussian 1 -pnfft_intpol 1 -pnfft_compute_f 0 -pnfft_compute_hessian_f 0 -pnfft_compare_direct 1 -pnfft_debug 1",
    "-pnfft_window 1 -pnfft_intpol 0 -pnfft_compute_grad_f 0",
    "-pnfft_window 7 -pnfft_intpol 9"])
def test_check_init_parameters_matches_reference(args):
    """pnfft_check_init_parameters (reference api/api-basic.c:820-938), the option parser of the reference's test programs: same
    sizes, cutoff, plan flags, compute flags, x_max, process mesh and switches as the compiled reference for the same argv,
    defaults included (window 4 = Kaiser-Bessel, f + grad_f + hessian_f on)."""
    import ctypes as C
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libpnfft_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    words = args.split()
    res = []
    for lib in (A.lib(), C.CDLL(ref_so)):
        argv = (C.c_char_p * (len(words) + 1))(b"prog", *[w.encode() for w in words])
        N, n, lM, m = (C.c_ssize_t * 3)(), (C.c_ssize_t * 3)(), C.c_ssize_t(), C.c_int()
        pf, cf, xm, mesh, cd, dbg = C.c_uint(), C.c_uint(), (C.c_double * 3)(), (C.c_int * 3)(), C.c_int(0), C.c_int(0)
        lib.pnfft_check_init_parameters(len(words) + 1, argv, N, n, C.byref(lM), C.byref(m), C.byref(pf), C.byref(cf), xm, mesh,
                                        C.byref(cd), C.byref(dbg))
        res.append((list(N), list(n), lM.value, m.value, pf.value, cf.value, list(xm), list(mesh), cd.value, dbg.value))
    assert res[0] == res[1]


def test_print_helpers_follow_reference_formats():
    """pnfft_vpr_complex / pnfft_vpr_real (reference api/api-basic.c:705-780): every rank prints "Rank r, name" and its vector,
    four complex numbers as %.2e+%.2ei or eight reals as %e per numbered line, nothing at all for N < 1;
    pnfft_apr_complex_3d walks a PNFFT_TRANSPOSED_F_HAT block in its memory order (k1, k2, k0) (:783-800).  Host only, one rank."""
    import subprocess
    import sys
    code = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, %r)
from pnfft_b200 import api as A
lib = A.lib()
z = (np.arange(6) * 1.5 + 1j * (np.arange(6) - 3.25)).astype(np.complex128)
r = (np.arange(11) * 0.375 - 1).astype(np.float64)
P = lambda a: a.ctypes.data_as(C.c_void_p)
lib.pnfft_vpr_complex(P(z), C.c_ssize_t(6), b"cv", 1)
lib.pnfft_vpr_real(P(r), C.c_ssize_t(11), b"rv", 1)
lib.pnfft_vpr_real(P(r), C.c_ssize_t(0), b"never", 1)
lN, lNs = (C.c_ssize_t * 3)(3, 1, 2), (C.c_ssize_t * 3)(-1, 0, -2)
lib.pnfft_apr_complex_3d(P(z), lN, lNs, C.c_uint(1 << 11), b"tr", 1)
C.CDLL("libc.so.6").fflush(None)
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    z = np.arange(6) * 1.5 + 1j * (np.arange(6) - 3.25)
    r = np.arange(11) * 0.375 - 1
    want = "\nRank 0, cv"
    for k in range(6):
        if k % 4 == 0:
            want += "\n%4d." % (k // 4)
        want += " %.2e+%.2ei," % (z[k].real, z[k].imag)
    want += "\n\nRank 0, rv"
    for k in range(11):
        if k % 8 == 0:
            want += "\n%4d." % (k // 8)
        want += " %e," % r[k]
    want += "\n"
    assert out.stdout.startswith(want), out.stdout
    rest = out.stdout[len(want):]
    assert "never" not in rest
    # transposed block local_N = (3, 1, 2), starts (-1, 0, -2): memory order (k1, k2, k0) = extents (1, 2, 3), starts (0, -2, -1)
    idx = re.findall(r"\[(-?\d+),(-?\d+),(-?\d+)\]", rest)
    assert [tuple(int(v) for v in t) for t in idx] == [(0, a, b) for a in (-2, -1) for b in (-1, 0, 1)]


def test_timer_report_follows_reference_format(tmp_path):
    """pnfft_write_average_timer / _adv (reference kernel/timer.c:139-367): the Octave-readable report -- legend once per new
    file, one comment line with the plan's flags, index / procs / np / N / n / m, then pnfft_trf / pnfft_adj iteration
    counts and per-iteration maxima, the stage breakdown in the _adv variant -- through the host probe of the same code."""
    import ctypes as C
    comm = A.create_procmesh_2d(1, 1)
    fn = A.lib().pnfft_b200_timer_report_host
    I3, D = C.c_ssize_t * 3, C.c_double * 10
    fn.argtypes = [C.c_char_p, C.c_uint, I3, I3, C.c_int, C.c_int * 3, D, D, C.c_int, C.c_int]
    fn.restype = None
    path = str(tmp_path / "timer.m")
    tr, ad = D(4, 2.0, 0.5, 0.1, 0.2, 0.8, 0.9, 0.3, 0, 0), D(2, 1.0, 0.25, 0.05, 0.1, 0.4, 0.45, 0.15, 0, 0)
    fn(path.encode(), (1 << 13) | (1 << 1) | (1 << 6) | (1 << 12), I3(16, 16, 16), I3(32, 32, 32), 6, (C.c_int * 3)(1, 1, 1), tr, ad, 0, comm)
    fn(path.encode(), 1 << 11, I3(8, 12, 10), I3(16, 24, 20), 4, (C.c_int * 3)(1, 1, 1), tr, ad, 1, comm)
    legend = ("%% N  - NFFT size\n%% n  - FFT size\n%% np - process grid\n%% procs - number of processes\n%% pnfft - PNFFT runtime\n"
              "%% pfft  - PFFT runtime\n%% index(i) = log(procs(i)) + 1\n").replace("%%", "%")

    def run(flagtext, N, n, m):
        return ("%% pnfft_flags == %s\n" % flagtext).replace("%%", "%") + \
            "index(1) = 1;  procs(1) = 1;  np_pnfft(1, 1:3) = [1 1 1];  N_pnfft(1, 1:3) = [%d %d %d ];  n_pnfft(1, 1:3) = [%d %d %d ];  m_pnfft(1) = %d;\n" % (N + n + (m,))

    def basic(prefix, t):
        return "%s_iter(1)    = %d;  %s(1)   = %.3e;\n" % (prefix, int(t[0]), prefix, t[1] / t[0])

    def adv(prefix, t):
        a = [v / t[0] for v in t]
        return ("%s_matrix_D(1)   = %.3e;  %s_matrix_F(1)   = %.3e;\n%s_matrix_B(1)   = %.3e;  %s_gcells(1)     = %.3e;\n"
                "%s_sort_nodes(1) = %.3e;  %s_loop_B(1)     = %.3e;\n%s_shift_in(1)   = %.3e;  %s_shift_out(1)  = %.3e;\n"
                % (prefix, a[7], prefix, a[6], prefix, a[5], prefix, a[4], prefix, a[3], prefix, a[2], prefix, a[8], prefix, a[9]))

    want = legend + run("PNFFT_WINDOW_GAUSSIAN | PNFFT_FAST_GAUSSIAN | PNFFT_FFT_OUT_OF_PLACE | PNFFT_DIFF_IK | PNFFT_MALLOC_F_HAT",
                        (16, 16, 16), (32, 32, 32), 6) + basic("pnfft_trf", list(tr)) + basic("pnfft_adj", list(ad))
    want += run("PNFFT_WINDOW_KAISER_BESSEL | PNFFT_FFT_OUT_OF_PLACE | PNFFT_TRANSPOSED_F_HAT | PNFFT_DIFF_AD", (8, 12, 10), (16, 24, 20), 4)
    want += basic("pnfft_trf", list(tr)) + basic("pnfft_adj", list(ad)) + adv("pnfft_trf", list(tr)) + adv("pnfft_adj", list(ad))
    assert open(path).read() == want
    # the flags shown are the reference plan's after its planner promoted them (api/api-guru.c:150-155; probed from the
    # compiled reference: PRE_LIN_PSI -> + PRE_CONST_PSI + FAST_GAUSSIAN), as pnfft_get_pnfft_flags returns them
    path2 = str(tmp_path / "timer2.m")
    fn(path2.encode(), 1 << 3, I3(8, 12, 10), I3(16, 24, 20), 4, (C.c_int * 3)(1, 1, 1), tr, ad, 0, comm)
    assert "% pnfft_flags == PNFFT_WINDOW_KAISER_BESSEL | PNFFT_FAST_GAUSSIAN | PNFFT_PRE_CONST_PSI | PNFFT_PRE_LIN_PSI | " in open(path2).read()
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libpnfft_ref.so")
    if os.path.exists(ref_so):
        from oracle import refdrv
        got = refdrv.get(False).probe("plan_flags", 0, np.zeros(1), (16, 16, 16), m=4, pnfft_flags=1 << 3)
        assert int(got[0]) == (1 << 3) | (1 << 2) | (1 << 1)


def test_init_f_hat_3d_is_the_reference_drivers_formula():
    """pnfft_init_f_hat_3d forwards to PFFT's generator in the reference (api/api-basic.c:663-679), which is not available; the
    product uses the in-tree formula of the reference's own check program (tests/check_vs_pfft.c:167-181):
    data[k] = 1000 / (2 g + 1) + i 1000 / (2 g + 2) with g the row-major index of k + N/2."""
    import ctypes as C
    N, lN, lNs = (6, 4, 8), (3, 4, 5), (-1, -2, -3)
    out = np.zeros(lN, np.complex128)
    I3 = C.c_ssize_t * 3
    A.lib().pnfft_init_f_hat_3d(I3(*N), I3(*lN), I3(*lNs), C.c_uint(0), out.ctypes.data_as(C.c_void_p))
    k = np.meshgrid(*[np.arange(lNs[t], lNs[t] + lN[t]) for t in range(3)], indexing="ij")
    g = ((k[0] + N[0] // 2) * N[1] + (k[1] + N[1] // 2)) * N[2] + (k[2] + N[2] // 2)
    assert np.array_equal(out, 1000.0 / (2 * g + 1) + 1j * (1000.0 / (2 * g + 2)))
