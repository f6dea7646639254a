"""Shared helpers of the parity tests: seeded inputs and single-rank runs through the C ABI."""
import numpy as np

from pnfft_b200 import api as A


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (d if d > 0 else 1.0))


def make_inputs(N, M, seed, c2r=False, single=False, lo=-0.5, up=0.5):
    rng = np.random.default_rng(seed)
    rdt = np.float32 if single else np.float64
    x = rng.uniform(lo, up, (M, 3)).astype(rdt)
    x = np.clip(x, -0.5, np.nextafter(rdt(0.5), rdt(0)))
    Nc = (N[0], N[1], N[2] // 2 + 1) if c2r else tuple(N)
    fh = (rng.uniform(-1, 1, Nc) + 1j * rng.uniform(-1, 1, Nc)).astype(np.complex64 if single else np.complex128)
    if c2r:
        f = rng.uniform(-1, 1, M).astype(rdt)
        g = rng.uniform(-1, 1, (M, 3)).astype(rdt)
    else:
        cdt = np.complex64 if single else np.complex128
        f = (rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)).astype(cdt)
        g = (rng.uniform(-1, 1, (M, 3)) + 1j * rng.uniform(-1, 1, (M, 3))).astype(cdt)
    return x, fh, f, g


def fixture_kwargs(g):
    """Optional keys of a golden fixture (tools/make_golden.py: extra_cases) with the defaults of the round-1 files."""
    N = tuple(int(v) for v in g["N"])
    n = tuple(int(v) for v in g["n"]) if "n" in g.files else tuple(2 * v for v in N)
    x_max = tuple(float(v) for v in g["x_max"]) if "x_max" in g.files else (0.5, 0.5, 0.5)
    return N, n, x_max, ("acc" in g.files and bool(g["acc"]))


class Run1:
    """One-rank plan + node set with host (numpy) arrays.  f_hat goes in and comes out in natural (k0, k1, k2) order;
    with PNFFT_TRANSPOSED_F_HAT the array handed to the library is its (k1, k2, k0) transpose (reference
    kernel/matrix_D.c:331-341)."""

    def __init__(self, N, x, n=None, m=6, flags=0, c2r=False, single=False, x_max=(0.5, 0.5, 0.5), variant=0):
        self.N = tuple(N)
        self.n = tuple(n) if n is not None else tuple(2 * v for v in N)
        self.c2r, self.single, self.m = c2r, single, m
        self.rdt = np.float32 if single else np.float64
        self.cdt = np.complex64 if single else np.complex128
        self.comm = A.create_procmesh_2d(1, 1)
        self.plan = A.Plan.init_guru(self.N, self.n, x_max, m, flags, self.comm, c2r=c2r, single=single)
        self.plan.set_kernel_variant(variant)
        self.M = x.shape[0]
        self.x = np.ascontiguousarray(x, self.rdt)
        self.nodes = A.Nodes(self.M, 0, single=single)
        self.nodes.set_x(self.x)
        ft = self.rdt if c2r else self.cdt
        self.f = np.zeros(self.M, ft)
        self.g = np.zeros((self.M, 3), ft)
        self.nodes.set_f(self.f)
        self.nodes.set_grad_f(self.g)
        Nc = (N[0], N[1], N[2] // 2 + 1) if c2r else self.N
        self.transposed = bool(flags & A.TRANSPOSED_F_HAT)
        self.f_hat = np.zeros((Nc[1], Nc[2], Nc[0]) if self.transposed else Nc, self.cdt)
        self.plan.set_f_hat(self.f_hat)

    def put_f_hat(self, f_hat):
        self.f_hat[...] = np.transpose(f_hat, (1, 2, 0)) if self.transposed else f_hat

    def get_f_hat(self):
        return np.ascontiguousarray(np.transpose(self.f_hat, (2, 0, 1))) if self.transposed else self.f_hat.copy()

    def trafo(self, f_hat, cf, f0=None, g0=None):
        self.put_f_hat(f_hat)
        if f0 is not None:
            self.f[...] = f0
        if g0 is not None:
            self.g[...] = g0
        self.plan.trafo(self.nodes, cf)
        return self.f.copy(), self.g.copy()

    def trafo_hessian(self, f_hat, cf):
        """trafo with PNFFT_COMPUTE_HESSIAN_F among cf: returns f, grad_f, hessian_f [M, 6]"""
        ft = self.rdt if self.c2r else self.cdt
        self.h = np.zeros((self.M, 6), ft)
        self.nodes.set_hessian_f(self.h)
        self.put_f_hat(f_hat)
        self.plan.trafo(self.nodes, cf)
        return self.f.copy(), self.g.copy(), self.h.copy()

    def adj(self, f, g, cf, f_hat0=None):
        if f is not None:
            self.f[...] = f
        if g is not None:
            self.g[...] = g
        if f_hat0 is not None:
            self.put_f_hat(f_hat0)
        self.plan.adj(self.nodes, cf)
        return self.get_f_hat()

    def close(self):
        self.nodes.free(0)
        self.plan.finalize(0)
