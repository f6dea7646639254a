"""world_size-2 (and 4) CPU tests of the host-side multi-rank logic: the mini-MPI control plane (rendezvous, Cartesian
mesh, host collectives) and the per-rank block decomposition / node borders of the C ABI, checked against the golden
reference layouts.  Ranks are real processes joined by torch.distributed (gloo) for the cross-rank comparison; the
library's own rendezvous runs over TCP on 127.0.0.1.  No GPU and no compute call."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, mesh, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), PNFFT_B200_PORT_OFFSET="1")
    sys.path.insert(0, ROOT)
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pnfft_b200 import api as A
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = A.lib()
        r, s = A.mpi_rank_size()
        assert (r, s) == (rank, world)
        comm = A.create_procmesh_2d(*mesh)
        # host collectives of the control plane (reference kernel/timer.c:79-86 uses MPI_Reduce(MAX))
        v = (C.c_double * 2)(float(rank + 1), float(-rank))
        o = (C.c_double * 2)()
        lib.pnb_MPI_Allreduce(v, o, 2, 6, 1, comm)      # MPI_DOUBLE, MPI_SUM
        assert o[0] == world * (world + 1) / 2 and o[1] == -world * (world - 1) / 2
        lib.pnb_MPI_Allreduce(v, o, 2, 6, 2, comm)      # MPI_MAX
        assert o[0] == world and o[1] == 0
        b = (C.c_int * 1)(1234 if rank == 0 else 0)
        lib.pnb_MPI_Bcast(b, 1, 2, 0, comm)
        assert b[0] == 1234
        # the exchange pattern of the direct NDFT across ranks (csrc/direct.cuh; reference kernel/ndft-parallel.c:424, 714):
        # every rank in turn broadcasts a block of bytes / receives the sum of all ranks' doubles
        for root in range(world):
            blk = (C.c_ubyte * 37)(*([(7 * root + i) % 251 for i in range(37)] if rank == root else [0] * 37))
            lib.pnb_MPI_Bcast(blk, 37, 8, root, comm)                 # MPI_BYTE
            assert list(blk) == [(7 * root + i) % 251 for i in range(37)]
            part = (C.c_double * 3)(rank + 0.5, root, -2.0 * rank)
            tot = (C.c_double * 3)(-1.0, -1.0, -1.0)
            lib.pnb_MPI_Reduce(part, tot, 3, 6, 1, root, comm)        # MPI_DOUBLE, MPI_SUM
            if rank == root:
                assert list(tot) == [world * (world - 1) / 2 + 0.5 * world, float(root * world), -float(world * (world - 1))]
            else:
                assert list(tot) == [-1.0, -1.0, -1.0]
        L = np.load(os.path.join(GOLD, "layouts.npz"))
        res = []
        for tag in ("even", "ragged", "torus"):
            for c2r in (False, True):
                key = "%s_%dx%d_%s" % (tag, mesh[0], mesh[1], "c2r" if c2r else "c2c")
                cfg = L[key + "_cfg"]
                N, n, m = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6])
                lN, lNs, lo, up = A.local_size_guru(N, n, tuple(L[key + "_xmax"]), m, comm, c2r=c2r)
                ok = (np.array_equal(lN, L[key + "_local_N"][rank]) and np.array_equal(lNs, L[key + "_local_N_start"][rank])
                      and np.array_equal(lo, L[key + "_lo"][rank]) and np.array_equal(up, L[key + "_up"][rank]))
                res.append(int(ok))
        L2 = np.load(os.path.join(GOLD, "layouts_r2.npz"))     # PNFFT_TRANSPOSED_F_HAT / PNFFT_INTERLACED blocks
        for tag in ("tr_even", "tr_ragged", "tr_il_torus"):
            for c2r in (False, True):
                key = "%s_%dx%d_%s" % (tag, mesh[0], mesh[1], "c2r" if c2r else "c2c")
                cfg = L2[key + "_cfg"]
                N, n, m, fl = tuple(cfg[0:3]), tuple(cfg[3:6]), int(cfg[6]), int(cfg[10])
                lN, lNs, lo, up = A.local_size_guru(N, n, tuple(L2[key + "_xmax"]), m, comm, pnfft_flags=fl, c2r=c2r)
                ok = (np.array_equal(lN, L2[key + "_local_N"][rank]) and np.array_equal(lNs, L2[key + "_local_N_start"][rank])
                      and np.array_equal(lo, L2[key + "_lo"][rank]) and np.array_equal(up, L2[key + "_up"][rank]))
                res.append(int(ok))
        t = torch.tensor(res, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        A.mpi_barrier(comm)
        if rank == 0:
            q.put(t.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mesh", [(1, 2), (2, 2), (2, 4)])
def test_layout_and_control_plane_over_ranks(mesh):
    import torch.multiprocessing as mp
    world = mesh[0] * mesh[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, mesh, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("rank did not finish")
        assert p.exitcode == 0
    res = q.get(timeout=5)
    assert res and all(v == 1 for v in res)
