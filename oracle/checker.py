"""TEST INFRASTRUCTURE ONLY -- one front end for the two CPU checkers of the PNFFT hot path.

  kind "reference": oracle/_ref/libpnfft{,f}_ref.so, the UNMODIFIED reference compiled from /root/reference
                    (oracle/refdrv.py); runs P virtual ranks as threads, so it can use all host cores.
  kind "port"     : oracle/liboracle.so, the clean-room restatement oracle/pnfft_oracle.c (single thread).

get() prefers the compiled reference and falls back to the port.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module; the product (pnfft_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INT = C.c_ssize_t


class _Cfg(C.Structure):
    _fields_ = [("N", INT * 3), ("n", INT * 3), ("x_max", C.c_double * 3), ("m", C.c_int), ("flags", C.c_uint),
                ("c2r", C.c_int), ("b", C.c_double * 3)]


class PortLib:
    """ctypes front end of oracle/liboracle.so with the call signatures of refdrv.RefLib (global arrays in and out)."""
    kind, threads = "port", False

    def __init__(self, single=False):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle oracle`)")
        self.lib = C.CDLL(path)
        self.single = single
        self.pre = "oraclef_" if single else "oracle_"
        self.real = C.c_float if single else C.c_double
        self.rdt = np.float32 if single else np.float64
        self.cdt = np.complex64 if single else np.complex128
        self.name = "oracle/liboracle.so (restatement, %s)" % ("float" if single else "double")

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.pre + name)
        f.restype = restype
        return f

    def _cfg(self, N, n, m, x_max, pnfft_flags, c2r, b=None):
        c = _Cfg()
        if b is not None:          # pnfft_set_b
            c.b[:] = [float(self.rdt(v)) for v in b]
        c.N[:] = [int(v) for v in N]
        c.n[:] = [int(v) for v in (n if n is not None else [2 * v for v in N])]
        c.x_max[:] = [float(v) for v in x_max]
        c.m, c.flags, c.c2r = int(m), int(pnfft_flags), int(bool(c2r))
        return c

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def layout(self, N, n=None, m=6, np_mesh=(1, 1), x_max=(0.5, 0.5, 0.5), pnfft_flags=0, c2r=False):
        P = np_mesh[0] * np_mesh[1]
        lay = np.zeros((P, 15), np.int64)
        bd = np.zeros((P, 6), self.rdt)
        c = self._cfg(N, n, m, x_max, pnfft_flags, c2r)
        self._fn("layout")(C.byref(c), int(np_mesh[0]), int(np_mesh[1]), self._p(lay), self._p(bd))
        return dict(local_N=lay[:, 0:3], local_N_start=lay[:, 3:6], local_no=lay[:, 6:9], local_no_start=lay[:, 9:12],
                    no=tuple(int(v) for v in lay[0, 12:15]), lo=bd[:, 0:3], up=bd[:, 3:6],
                    b=tuple(self._fn("shape_b", C.c_double)(C.byref(c), t) for t in range(3)))

    def node_index(self, x, N, n=None, m=6, np_mesh=(1, 1), x_max=(0.5, 0.5, 0.5), pnfft_flags=0):
        xs = np.ascontiguousarray(x, self.rdt)
        M = xs.shape[0]
        owner = np.zeros(M, np.int32)
        idx = np.zeros((M, 4), np.int64)
        c = self._cfg(N, n, m, x_max, pnfft_flags, False)
        self._fn("node_index")(C.byref(c), int(np_mesh[0]), int(np_mesh[1]), INT(M), self._p(xs), self._p(owner), self._p(idx))
        return owner, idx

    def probe_sort_keys(self, x, N, n=None, m=6):
        xs = np.ascontiguousarray(x, self.rdt)
        keys = np.zeros(xs.shape[0], np.int64)
        c = self._cfg(N, n, m, (0.5,) * 3, 0, False)
        self._fn("sort_keys")(C.byref(c), INT(xs.shape[0]), self._p(xs), self._p(keys))
        return keys

    def probe_tensor(self, x, N, n=None, m=6, x_max=(0.5, 0.5, 0.5), pnfft_flags=0, grad=True):
        xs = np.ascontiguousarray(x, self.rdt)
        M, cut = xs.shape[0], 2 * m + 1
        psi = np.zeros((M, 3, cut), self.rdt)
        dpsi = np.zeros((M, 3, cut), self.rdt) if grad else None
        c = self._cfg(N, n, m, x_max, pnfft_flags, False)
        self._fn("window_tensor")(C.byref(c), INT(M), self._p(xs), self._p(psi), self._p(dpsi))
        return psi, dpsi

    def probe(self, which, dim, arg, N, n=None, m=6, x_max=(0.5, 0.5, 0.5), pnfft_flags=0, c2r=False):
        if which != "inv_phi_hat":
            raise NotImplementedError(which)
        c = self._cfg(N, n, m, x_max, pnfft_flags, c2r)
        f = self._fn("inv_phi_hat", C.c_double)
        return np.array([f(C.byref(c), int(dim), INT(int(k))) for k in np.atleast_1d(arg)], self.rdt)

    def trafo(self, N, x, f_hat, n=None, m=6, np_mesh=(1, 1), x_max=(0.5, 0.5, 0.5), pnfft_flags=0, compute_flags=1,
              c2r=False, f=None, grad_f=None, b=None, **_):
        xs = np.ascontiguousarray(x, self.rdt)
        M = xs.shape[0]
        ft = self.rdt if c2r else self.cdt
        fh = np.ascontiguousarray(f_hat, self.cdt)
        fo = np.zeros(M, ft) if f is None else np.ascontiguousarray(f, ft).copy()
        go = np.zeros((M, 3), ft) if grad_f is None else np.ascontiguousarray(grad_f, ft).copy()
        c = self._cfg(N, n, m, x_max, pnfft_flags, c2r, b=b)
        ho = np.zeros((M, 6), ft) if (compute_flags & 4) else None     # PNFFT_COMPUTE_HESSIAN_F: xx, xy, xz, yy, yz, zz
        self._fn("trafo_h")(C.byref(c), INT(M), self._p(xs), self._p(fh), self._p(fo), self._p(go),
                            self._p(ho) if ho is not None else None, C.c_uint(compute_flags))
        return dict(f=fo, grad_f=go, hessian_f=ho, timers=None)

    def adj(self, N, x, f=None, grad_f=None, n=None, m=6, np_mesh=(1, 1), x_max=(0.5, 0.5, 0.5), pnfft_flags=0,
            compute_flags=1, c2r=False, f_hat=None, b=None, **_):
        xs = np.ascontiguousarray(x, self.rdt)
        M = xs.shape[0]
        ft = self.rdt if c2r else self.cdt
        N = tuple(int(v) for v in N)
        shape = (N[0], N[1], N[2] // 2 + 1 if c2r else N[2])
        fh = np.zeros(shape, self.cdt) if f_hat is None else np.ascontiguousarray(f_hat, self.cdt).reshape(shape).copy()
        fi = np.zeros(M, ft) if f is None else np.ascontiguousarray(f, ft)
        gi = np.zeros((M, 3), ft) if grad_f is None else np.ascontiguousarray(grad_f, ft)
        c = self._cfg(N, n, m, x_max, pnfft_flags, c2r, b=b)
        self._fn("adj")(C.byref(c), INT(M), self._p(xs), self._p(fi), self._p(gi), self._p(fh), C.c_uint(compute_flags))
        return dict(f_hat=fh, timers=None)


_cache = {}


def port(single=False):
    key = ("port", single)
    if key not in _cache:
        _cache[key] = PortLib(single)
    return _cache[key]


def get(single=False, prefer="reference"):
    """The strongest checker available: the compiled reference if oracle/_ref was built, else the port."""
    from oracle import refdrv
    if prefer == "reference" and refdrv.available(single):
        r = refdrv.get(single)
        r.kind, r.threads = "reference", True
        r.name = "oracle/_ref (unmodified reference PNFFT, %s)" % ("float" if single else "double")
        return r
    return port(single)
