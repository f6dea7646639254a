"""Host-side mirror of PNFFT's C interface (reference api/pnfft.h) over libpnfft_b200.so.

Every method forwards 1:1 to the C-ABI entry point of the same name (``pnfft_*`` for double,
``pnfftf_*`` for float), so tests and benchmarks read like the reference's own C drivers
(reference tests/simple_test.c:27-90):

    comm = create_procmesh_2d(np0, np1)
    local_N, local_N_start, lo, up = local_size_guru(N, n, x_max, m, comm, flags)
    plan = Plan.init_guru(N, n, x_max, m, flags, comm)
    nodes = Nodes(local_M, MALLOC_X | MALLOC_F)
    plan.trafo(nodes, COMPUTE_F)

There is no CPU implementation behind this module: without the CUDA library (or without a GPU for
anything beyond layout queries) calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNFFT_B200_LIB") or os.path.join(HERE, "lib", "libpnfft_b200.so")   # (override: development builds)

# ---- flag values (include/pnfft.h == reference api/pnfft.h:302-390) ----
PRE_PHI_HAT = 1 << 0
FAST_GAUSSIAN = FG_PSI = 1 << 1
MALLOC_F_HAT = 1 << 6
FFT_IN_PLACE = 1 << 7
INTERLACED = 1 << 8
TRANSPOSED_F_HAT = 1 << 11
DIFF_AD = 0
DIFF_IK = 1 << 12
WINDOW_KAISER_BESSEL = 0
WINDOW_GAUSSIAN = 1 << 13
WINDOW_BSPLINE = 1 << 14
WINDOW_SINC_POWER = 1 << 15
WINDOW_BESSEL_I0 = 1 << 16
SORT_NODES = 1 << 18
MALLOC_X, MALLOC_F, MALLOC_GRAD_F, MALLOC_HESSIAN_F = 1, 2, 4, 8
FREE_X, FREE_F, FREE_GRAD_F = 1, 2, 4
PRE_FULL, PRE_PSI, PRE_GRAD_PSI = 1, 2, 4
COMPUTE_F, COMPUTE_GRAD_F, COMPUTE_HESSIAN_F, COMPUTE_DIRECT = 1, 2, 4, 8
COMPUTE_ACCUMULATED, OMIT_DECONV, OMIT_FFT, OMIT_CONV = 16, 32, 64, 128
TIMER_LENGTH = 10
MPI_COMM_WORLD = 1

INT = C.c_ssize_t
INT3 = INT * 3

_lib = None


def lib():
    """The C-ABI library; raises if it was not built (python __graft_entry__.py / make -C pnfft_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libpnfft_b200.so is missing (%s): build it with `make -C pnfft_b200/csrc`; "
                               "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.pnb_MPI_Init(None, None)
    return _lib


def measure_fp64_tflops():
    """FP64 FMA rate of the current CUDA device in TFLOP/s (pnfft_b200_measure_fp64_tflops, include/pnfft.h)."""
    f = lib().pnfft_b200_measure_fp64_tflops
    f.restype = C.c_double
    f.argtypes = []
    return float(f())


def _int3(v):
    return INT3(*[int(a) for a in v])


def mpi_rank_size(comm=MPI_COMM_WORLD):
    r, s = C.c_int(), C.c_int()
    lib().pnb_MPI_Comm_rank(comm, C.byref(r))
    lib().pnb_MPI_Comm_size(comm, C.byref(s))
    return r.value, s.value


def mpi_barrier(comm=MPI_COMM_WORLD):
    lib().pnb_MPI_Barrier(comm)


def create_procmesh_2d(np0, np1, comm=MPI_COMM_WORLD):
    """pnfft_create_procmesh_2d (reference util/util.c:32-40); raises where the C call returns non-zero."""
    out = C.c_int(0)
    rc = lib().pnfft_create_procmesh_2d(comm, int(np0), int(np1), C.byref(out))
    if rc:
        raise ValueError("process mesh %dx%d does not match the number of ranks" % (np0, np1))
    return out.value


class _Prec:
    def __init__(self, single):
        self.single = bool(single)
        self.pre = "pnfftf_" if single else "pnfft_"
        self.real = C.c_float if single else C.c_double
        self.rdt = np.float32 if single else np.float64
        self.cdt = np.complex64 if single else np.complex128

    def fn(self, name, restype=None, argtypes=None):
        f = getattr(lib(), self.pre + name)
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f


def _ptr(a):
    """Raw address of a numpy array, a torch tensor (host or CUDA) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def local_size_guru(N, n, x_max, m, comm, pnfft_flags=0, c2r=False, single=False):
    """pnfft_local_size_guru[_c2r] (reference api/api-guru.c:31-62): (local_N, local_N_start, lo, up)."""
    P = _Prec(single)
    R3 = P.real * 3
    lN, lNs, lo, up = INT3(), INT3(), R3(), R3()
    f = P.fn("local_size_guru_c2r" if c2r else "local_size_guru", None,
             [C.c_int, INT3, INT3, R3, C.c_int, C.c_int, C.c_uint, INT3, INT3, R3, R3])
    f(3, _int3(N), _int3(n), R3(*[float(v) for v in x_max]), int(m), comm, pnfft_flags, lN, lNs, lo, up)
    return (np.array(lN[:], np.int64), np.array(lNs[:], np.int64), np.array(lo[:], P.rdt), np.array(up[:], P.rdt))


def local_size_3d(N, comm, pnfft_flags=0, c2r=False, single=False):
    """pnfft_local_size_3d[_c2r] (reference api/api-basic.c:36-65)."""
    P = _Prec(single)
    R3 = P.real * 3
    lN, lNs, lo, up = INT3(), INT3(), R3(), R3()
    f = P.fn("local_size_3d_c2r" if c2r else "local_size_3d", None, [INT3, C.c_int, C.c_uint, INT3, INT3, R3, R3])
    f(_int3(N), comm, pnfft_flags, lN, lNs, lo, up)
    return (np.array(lN[:], np.int64), np.array(lNs[:], np.int64), np.array(lo[:], P.rdt), np.array(up[:], P.rdt))


class Nodes:
    """pnfft_nodes (reference api/api-basic.c:402-447).  Arrays are attached with set_x / set_f / set_grad_f
    and may be numpy arrays (host) or CUDA torch tensors (device resident, no copies)."""

    def __init__(self, local_M, malloc_flags=0, single=False):
        self.P = _Prec(single)
        self.local_M = int(local_M)
        f = self.P.fn("init_nodes", C.c_void_p, [INT, C.c_uint])
        self.h = C.c_void_p(f(self.local_M, malloc_flags))
        self.malloc_flags = malloc_flags
        self._keep = {}

    def set_x(self, x):
        self._keep["x"] = x
        self.P.fn("set_x", None, [C.c_void_p, C.c_void_p])(_ptr(x), self.h)

    def set_f(self, f):
        self._keep["f"] = f
        self.P.fn("set_f", None, [C.c_void_p, C.c_void_p])(_ptr(f), self.h)

    def set_grad_f(self, g):
        self._keep["g"] = g
        self.P.fn("set_grad_f", None, [C.c_void_p, C.c_void_p])(_ptr(g), self.h)

    def set_hessian_f(self, h):
        """6 values per node (xx, xy, xz, yy, yz, zz; complex for c2c plans), written by trafo(COMPUTE_HESSIAN_F)"""
        self._keep["h"] = h
        self.P.fn("set_hessian_f", None, [C.c_void_p, C.c_void_p])(_ptr(h), self.h)

    def x_static(self, on=True):
        """pnfft_b200_nodes_x_static: promise that x stays unchanged until the next set_x (upload and binning are reused)."""
        self.P.fn("b200_nodes_x_static", None, [C.c_void_p, C.c_int])(self.h, int(bool(on)))

    def free(self, flags=0):
        if self.h:
            self.P.fn("free_nodes", None, [C.c_void_p, C.c_uint])(self.h, flags)
            self.h = None


class Plan:
    """pnfft_plan (reference api/api-guru.c:64-165, api/api-basic.c:199-378)."""

    def __init__(self, handle, P, N, c2r, comm):
        self.h, self.P, self.N, self.c2r, self.comm = handle, P, tuple(int(v) for v in N), c2r, comm
        self._keep = {}

    @classmethod
    def init_guru(cls, N, n, x_max, m, pnfft_flags, comm, pfft_flags=0, c2r=False, single=False):
        P = _Prec(single)
        R3 = P.real * 3
        f = P.fn("init_guru_c2r" if c2r else "init_guru", C.c_void_p,
                 [C.c_int, INT3, INT3, R3, C.c_int, C.c_uint, C.c_uint, C.c_int])
        h = f(3, _int3(N), _int3(n), R3(*[float(v) for v in x_max]), int(m), pnfft_flags, pfft_flags, comm)
        if not h:
            raise RuntimeError("pnfft_init_guru returned NULL (see stderr)")
        return cls(C.c_void_p(h), P, N, c2r, comm)

    # ---- reference API ----
    def set_f_hat(self, f_hat):
        self._keep["f_hat"] = f_hat
        self.P.fn("set_f_hat", None, [C.c_void_p, C.c_void_p])(_ptr(f_hat), self.h)

    def set_b(self, b0, b1, b2):
        self.P.fn("set_b", None, [self.P.real] * 3 + [C.c_void_p])(b0, b1, b2, self.h)

    def get_b(self):
        b = [self.P.real() for _ in range(3)]
        self.P.fn("get_b", None, [C.c_void_p] + [C.POINTER(self.P.real)] * 3)(self.h, *[C.byref(v) for v in b])
        return tuple(v.value for v in b)

    def trafo(self, nodes, compute_flags):
        self.P.fn("trafo", None, [C.c_void_p, C.c_void_p, C.c_uint])(self.h, nodes.h if nodes else None, compute_flags)

    def adj(self, nodes, compute_flags):
        self.P.fn("adj", None, [C.c_void_p, C.c_void_p, C.c_uint])(self.h, nodes.h if nodes else None, compute_flags)

    def precompute_psi(self, nodes, precompute_flags):
        self.P.fn("precompute_psi", None, [C.c_void_p, C.c_void_p, C.c_uint])(self.h, nodes.h, precompute_flags)

    def inv_phi_hat(self, dim, k):
        return self.P.fn("inv_phi_hat", self.P.real, [C.c_void_p, C.c_int, INT])(self.h, dim, int(k))

    def phi_hat(self, dim, k):
        return self.P.fn("phi_hat", self.P.real, [C.c_void_p, C.c_int, INT])(self.h, dim, int(k))

    def psi(self, dim, x):
        return self.P.fn("psi", self.P.real, [C.c_void_p, C.c_int, self.P.real])(self.h, dim, float(x))

    def dpsi(self, dim, x):
        return self.P.fn("dpsi", self.P.real, [C.c_void_p, C.c_int, self.P.real])(self.h, dim, float(x))

    def timer(self, adjoint=False):
        f = self.P.fn("get_timer_adj" if adjoint else "get_timer_trafo", C.POINTER(C.c_double), [C.c_void_p])
        p = f(self.h)
        out = np.array([p[i] for i in range(TIMER_LENGTH)])
        self.P.fn("timer_free", None, [C.POINTER(C.c_double)])(p)
        return out

    def reset_timer(self):
        self.P.fn("reset_timer", None, [C.c_void_p])(self.h)

    def finalize(self, flags=0):
        if self.h:
            self.P.fn("finalize", None, [C.c_void_p, C.c_uint])(self.h, flags)
            self.h = None

    # ---- extensions (pnfft_b200_*) ----
    def local_no(self):
        a, b, c = INT3(), INT3(), INT3()
        self.P.fn("b200_get_local_no", None, [C.c_void_p, INT3, INT3, INT3])(self.h, a, b, c)
        return np.array(a[:], np.int64), np.array(b[:], np.int64), np.array(c[:], np.int64)

    def set_grid(self, grid):
        self.P.fn("b200_set_grid", None, [C.c_void_p, C.c_void_p])(self.h, _ptr(grid))

    def get_grid(self):
        lno, _, _ = self.local_no()
        out = np.zeros(tuple(lno), self.P.rdt if self.c2r else self.P.cdt)
        self.P.fn("b200_get_grid", None, [C.c_void_p, C.c_void_p])(self.h, _ptr(out))
        return out

    def set_g1(self, g1):
        self.P.fn("b200_set_g1", None, [C.c_void_p, C.c_void_p])(self.h, _ptr(g1))

    def get_g1(self, shape):
        out = np.zeros(shape, self.P.cdt)
        self.P.fn("b200_get_g1", None, [C.c_void_p, C.c_void_p])(self.h, _ptr(out))
        return out

    def node_grid_index(self, nodes):
        out = np.zeros((nodes.local_M, 4), np.int64)
        self.P.fn("b200_node_grid_index", None, [C.c_void_p, C.c_void_p, C.c_void_p])(self.h, nodes.h, _ptr(out))
        return out

    def sort_nodes(self, nodes):
        keys = np.zeros(nodes.local_M, np.int64)
        perm = np.zeros(nodes.local_M, np.int64)
        self.P.fn("b200_sort_nodes", None, [C.c_void_p] * 4)(self.h, nodes.h, _ptr(keys), _ptr(perm))
        return keys, perm

    def window_tensor(self, nodes, m, grad=True):
        psi = np.zeros((nodes.local_M, 3, 2 * m + 1), self.P.rdt)
        dpsi = np.zeros_like(psi) if grad else None
        self.P.fn("b200_window_tensor", None, [C.c_void_p] * 4)(self.h, nodes.h, _ptr(psi), _ptr(dpsi))
        return psi, dpsi

    def set_kernel_variant(self, v):
        self.P.fn("b200_set_kernel_variant", None, [C.c_void_p, C.c_int])(self.h, int(v))

    def stage_ms(self, adjoint=False):
        out = (C.c_double * 8)()
        self.P.fn("b200_get_stage_ms", None, [C.c_void_p, C.c_int, C.c_double * 8])(self.h, int(adjoint), out)
        return dict(zip(["b_kernel", "binning", "halo", "fft", "deconv", "h2d", "d2h", "whole"], out[:]))

    def kernel_launches(self):
        return int(self.P.fn("b200_kernel_launches", C.c_longlong, [C.c_void_p])(self.h))

    def library_calls(self):
        return int(self.P.fn("b200_library_calls", C.c_longlong, [C.c_void_p])(self.h))

    def stream(self):
        """Raw cudaStream_t of the plan (everything trafo/adj launches goes to this stream)."""
        return int(self.P.fn("b200_get_stream", C.c_void_p, [C.c_void_p])(self.h) or 0)
