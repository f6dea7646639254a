"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libpnfft{,f}_ref.so.

The library is the UNMODIFIED reference PNFFT (compiled from /root/reference by
oracle/Makefile against the MPI/PFFT/GSL shims) plus oracle/ref_driver.c.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may
import this module; the product (pnfft_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- flag values of the reference's public header (api/pnfft.h:302-390) ----
PRE_PHI_HAT = 1 << 0
FAST_GAUSSIAN = 1 << 1
MALLOC_F_HAT = 1 << 6
FFT_IN_PLACE = 1 << 7
INTERLACED = 1 << 8
TRANSPOSED_F_HAT = 1 << 11
DIFF_IK = 1 << 12
WINDOW_KAISER_BESSEL = 0
WINDOW_GAUSSIAN = 1 << 13
WINDOW_BSPLINE = 1 << 14
WINDOW_SINC_POWER = 1 << 15
WINDOW_BESSEL_I0 = 1 << 16
SORT_NODES = 1 << 18
PRE_PSI = 1 << 1
PRE_GRAD_PSI = 1 << 2
COMPUTE_F = 1 << 0
COMPUTE_GRAD_F = 1 << 1
COMPUTE_HESSIAN_F = 1 << 2
COMPUTE_DIRECT = 1 << 3
COMPUTE_ACCUMULATED = 1 << 4
OMIT_DECONV = 1 << 5
OMIT_FFT = 1 << 6
OMIT_CONV = 1 << 7
TIMER_LENGTH = 10
TIMER_NAMES = ["iter", "whole", "loop_b", "sort_nodes", "gcells", "matrix_b", "matrix_f",
               "matrix_d", "shift_input", "shift_output"]

OP_TRAFO, OP_ADJ, OP_LAYOUT = 0, 1, 2

INT = C.c_ssize_t


def _job_struct(real):
    RP = C.POINTER(real)

    class Job(C.Structure):
        _fields_ = [
            ("np0", C.c_int), ("np1", C.c_int),
            ("N", INT * 3), ("n", INT * 3),
            ("x_max", C.c_double * 3),
            ("m", C.c_int),
            ("pnfft_flags", C.c_uint), ("precompute_flags", C.c_uint), ("compute_flags", C.c_uint),
            ("c2r", C.c_int), ("op", C.c_int), ("repeat", C.c_int),
            ("M", INT),
            ("x", RP), ("f_hat", RP), ("f", RP), ("grad_f", RP), ("grid", RP), ("g1", RP),
            ("set_grid", C.c_int), ("get_grid", C.c_int), ("set_g1", C.c_int), ("get_g1", C.c_int),
            ("owner", C.POINTER(C.c_int)),
            ("layout", C.POINTER(INT)),
            ("borders", RP),
            ("timers", C.POINTER(C.c_double)),
            ("node_index", C.POINTER(INT)),
            ("sort_perm", C.POINTER(INT)),
            ("b_out", C.c_double * 3),
            ("hessian_f", RP),
            ("b_in", C.c_double * 3),
        ]

    class Probe(C.Structure):
        _fields_ = [("N", INT * 3), ("n", INT * 3), ("x_max", C.c_double * 3), ("m", C.c_int),
                    ("pnfft_flags", C.c_uint), ("c2r", C.c_int)]

    return Job, Probe


class RefLib:
    """One precision of the compiled reference."""

    def __init__(self, single=False):
        name = "libpnfftf_ref.so" if single else "libpnfft_ref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(path)
        self.single = single
        self.real = C.c_float if single else C.c_double
        self.rdt = np.float32 if single else np.float64
        self.cdt = np.complex64 if single else np.complex128
        self.pre = "refdrvf_" if single else "refdrv_"
        self.Job, self.Probe = _job_struct(self.real)
        self._run = getattr(self.lib, self.pre + "run")
        self._run.argtypes = [C.POINTER(self.Job)]
        self._run.restype = C.c_int

    # ---------------- helpers ----------------
    def _rp(self, a):
        return a.ctypes.data_as(C.POINTER(self.real)) if a is not None else None

    def n2c(self, N, c2r):
        return N[2] // 2 + 1 if c2r else N[2]

    def run(self, op, N, n=None, m=6, x=None, f_hat=None, f=None, grad_f=None, np_mesh=(1, 1),
            x_max=(0.5, 0.5, 0.5), pnfft_flags=0, precompute_flags=0, compute_flags=COMPUTE_F, c2r=False,
            grid=None, g1=None, set_grid=False, get_grid=False, set_g1=False, get_g1=False,
            want_index=False, want_sort=False, repeat=1, b=None):
        """Run trafo / adj / layout query of the reference on np_mesh[0] x np_mesh[1] virtual ranks.

        Arrays are global: x [M,3]; f_hat [N0,N1,N2c] complex; f [M] (complex, or real for c2r);
        grad_f [M,3].  Inputs are not modified; results are returned in a dict.
        """
        N = tuple(int(v) for v in N)
        n = tuple(int(v) for v in (n if n is not None else [2 * v for v in N]))
        P = np_mesh[0] * np_mesh[1]
        M = 0 if x is None else int(x.shape[0])
        J = self.Job()
        J.np0, J.np1 = np_mesh
        J.N[:] = N
        J.n[:] = n
        J.x_max[:] = [float(v) for v in x_max]
        J.m = m
        J.pnfft_flags, J.precompute_flags, J.compute_flags = pnfft_flags, precompute_flags, compute_flags
        J.c2r, J.op, J.repeat, J.M = int(c2r), op, repeat, M
        if b is not None:          # pnfft_set_b after the plan is made: other window shape parameters than the default
            J.b_in[:] = [float(v) for v in b]
        keep = {}
        ftype = self.rdt if c2r else self.cdt

        xs = np.ascontiguousarray(x, dtype=self.rdt) if x is not None else np.zeros((0, 3), self.rdt)
        keep["x"] = xs
        J.x = self._rp(xs)
        fh_shape = (N[0], N[1], self.n2c(N, c2r))
        if op == OP_TRAFO:
            fh = np.ascontiguousarray(f_hat, dtype=self.cdt).reshape(fh_shape).copy() if f_hat is not None \
                else np.zeros(fh_shape, self.cdt)
            fo = np.zeros(M, ftype) if f is None else np.ascontiguousarray(f, dtype=ftype).copy()
            go = np.zeros((M, 3), ftype) if grad_f is None else np.ascontiguousarray(grad_f, dtype=ftype).copy()
        else:
            fh = np.zeros(fh_shape, self.cdt) if f_hat is None \
                else np.ascontiguousarray(f_hat, dtype=self.cdt).reshape(fh_shape).copy()
            fo = np.zeros(M, ftype) if f is None else np.ascontiguousarray(f, dtype=ftype).copy()
            go = np.zeros((M, 3), ftype) if grad_f is None else np.ascontiguousarray(grad_f, dtype=ftype).copy()
        J.f_hat, J.f, J.grad_f = self._rp(fh), self._rp(fo), self._rp(go)
        keep.update(fh=fh, fo=fo, go=go)
        ho = None
        if op == OP_TRAFO and (compute_flags & COMPUTE_HESSIAN_F):
            ho = np.zeros((M, 6), ftype)
            J.hessian_f = self._rp(ho)

        gtype = self.rdt if c2r else self.cdt
        grid_a = None
        if grid is not None or get_grid:
            # the no-array extent is only known after the layout query
            lay = self.layout(N, n, m, np_mesh, x_max, pnfft_flags, c2r)
            no = lay["no"]
            grid_a = np.zeros(no, gtype) if grid is None else np.ascontiguousarray(grid, dtype=gtype).reshape(no).copy()
            J.grid = self._rp(grid_a)
        g1_a = None
        if g1 is not None or get_g1:
            g1_a = np.zeros(fh_shape, self.cdt) if g1 is None else np.ascontiguousarray(g1, dtype=self.cdt).reshape(fh_shape).copy()
            J.g1 = self._rp(g1_a)
        J.set_grid, J.get_grid, J.set_g1, J.get_g1 = int(set_grid), int(get_grid), int(set_g1), int(get_g1)

        owner = np.full(M, -1, np.int32)
        layout = np.zeros((P, 15), np.int64)
        borders = np.zeros((P, 6), self.rdt)
        timers = np.zeros((P, 2, TIMER_LENGTH), np.float64)
        J.owner = owner.ctypes.data_as(C.POINTER(C.c_int))
        J.layout = layout.ctypes.data_as(C.POINTER(INT))
        J.borders = self._rp(borders)
        J.timers = timers.ctypes.data_as(C.POINTER(C.c_double))
        node_index = sort_perm = None
        if want_index:
            node_index = np.zeros((M, 4), np.int64)
            J.node_index = node_index.ctypes.data_as(C.POINTER(INT))
        if want_sort:
            sort_perm = np.zeros(M, np.int64)
            J.sort_perm = sort_perm.ctypes.data_as(C.POINTER(INT))

        rc = self._run(C.byref(J))
        if rc:
            raise RuntimeError("reference driver failed (process mesh does not match)")
        return dict(f_hat=fh, f=fo, grad_f=go, hessian_f=ho, grid=grid_a, g1=g1_a, owner=owner,
                    local_N=layout[:, 0:3], local_N_start=layout[:, 3:6], local_no=layout[:, 6:9],
                    local_no_start=layout[:, 9:12], no=tuple(int(v) for v in layout[0, 12:15]),
                    lo=borders[:, 0:3], up=borders[:, 3:6], timers=timers, node_index=node_index,
                    sort_perm=sort_perm, b=tuple(J.b_out))

    def layout(self, N, n=None, m=6, np_mesh=(1, 1), x_max=(0.5, 0.5, 0.5), pnfft_flags=0, c2r=False):
        return self.run(OP_LAYOUT, N, n, m, np_mesh=np_mesh, x_max=x_max, pnfft_flags=pnfft_flags, c2r=c2r)

    def trafo(self, N, x, f_hat, **kw):
        return self.run(OP_TRAFO, N, x=x, f_hat=f_hat, **kw)

    def adj(self, N, x, f=None, grad_f=None, **kw):
        return self.run(OP_ADJ, N, x=x, f=f, grad_f=grad_f, **kw)

    # ---------------- scalar probes ----------------
    def _probe_cfg(self, N, n, m, x_max, pnfft_flags, c2r):
        P = self.Probe()
        P.N[:] = [int(v) for v in N]
        P.n[:] = [int(v) for v in (n if n is not None else [2 * v for v in N])]
        P.x_max[:] = [float(v) for v in x_max]
        P.m, P.pnfft_flags, P.c2r = m, pnfft_flags, int(c2r)
        return P

    def probe(self, which, dim, arg, N, n=None, m=6, x_max=(0.5, 0.5, 0.5), pnfft_flags=0, c2r=False):
        """which in {'psi','dpsi','ddpsi','inv_phi_hat','phi_hat'}; arg = x values or integer k values."""
        code = {"psi": 0, "dpsi": 1, "inv_phi_hat": 2, "phi_hat": 3, "ddpsi": 4, "plan_flags": 5}[which]
        a = np.ascontiguousarray(arg, dtype=self.rdt)
        out = np.zeros_like(a)
        fn = getattr(self.lib, self.pre + "probe")
        fn.argtypes = [C.POINTER(self.Probe), C.c_int, C.c_int, INT, C.POINTER(self.real), C.POINTER(self.real)]
        fn.restype = None
        P = self._probe_cfg(N, n, m, x_max, pnfft_flags, c2r)
        fn(C.byref(P), code, dim, a.size, self._rp(a), self._rp(out))
        return out

    def probe_tensor(self, x, N, n=None, m=6, x_max=(0.5, 0.5, 0.5), pnfft_flags=0, grad=True):
        """[M,3,2m+1] window values (and derivatives) exactly as the hot loop evaluates them."""
        xs = np.ascontiguousarray(x, dtype=self.rdt)
        M = xs.shape[0]
        c = 2 * m + 1
        psi = np.zeros((M, 3, c), self.rdt)
        dpsi = np.zeros((M, 3, c), self.rdt) if grad else None
        fn = getattr(self.lib, self.pre + "probe_tensor")
        fn.argtypes = [C.POINTER(self.Probe), INT, C.POINTER(self.real), C.POINTER(self.real), C.POINTER(self.real)]
        fn.restype = None
        P = self._probe_cfg(N, n, m, x_max, pnfft_flags, False)
        fn(C.byref(P), M, self._rp(xs), self._rp(psi), self._rp(dpsi))
        return psi, dpsi

    def probe_sort(self, x, N, n=None, m=6):
        """(keys, perm) of sort_nodes_for_better_cache_handle (kernel/ndft-parallel.c:2121-2159)."""
        xs = np.ascontiguousarray(x, dtype=self.rdt)
        M = xs.shape[0]
        kp = np.zeros((M, 2), np.int64)
        fn = getattr(self.lib, self.pre + "probe_sort")
        fn.argtypes = [C.POINTER(self.Probe), INT, C.POINTER(self.real), C.POINTER(INT)]
        fn.restype = None
        P = self._probe_cfg(N, n, m, (0.5, 0.5, 0.5), 0, False)
        fn(C.byref(P), M, self._rp(xs), kp.ctypes.data_as(C.POINTER(INT)))
        return kp[:, 0].copy(), kp[:, 1].copy()

    def bessel_i0(self, x):
        fn = getattr(self.lib, self.pre + "bessel_i0")
        fn.argtypes = [C.c_double]
        fn.restype = C.c_double
        return np.array([fn(float(v)) for v in np.atleast_1d(x)])

    def bessel_i1(self, x):
        fn = getattr(self.lib, self.pre + "bessel_i1")
        fn.argtypes = [C.c_double]
        fn.restype = C.c_double
        return np.array([fn(float(v)) for v in np.atleast_1d(x)])


_cache = {}


def get(single=False):
    if single not in _cache:
        _cache[single] = RefLib(single)
    return _cache[single]


def available(single=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libpnfftf_ref.so" if single else "libpnfft_ref.so"))
