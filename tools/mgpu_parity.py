"""Multi-rank GPU parity (run under torch.distributed.run, one rank per GPU): the pencil-decomposed library -- block
layout, NCCL ghost-cell exchange / reduce, all-to-all FFT transposes, per-rank gridding -- against the single-rank CPU
checker on the same global problem.  Every rank builds the same seeded global inputs, keeps the nodes its [lo, up) owns
(reference kernel/ndft-parallel.c:734-775) and its f_hat block; rank 0 gathers f, grad_f and the f_hat blocks and compares
them with the oracle.  Prints one JSON line; exit code 1 on a parity failure.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_parity.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pnfft_b200 import api as A  # noqa: E402

MESH = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


def rel_l2(a, b):
    d = np.linalg.norm(np.ravel(b))
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / (d if d > 0 else 1.0))


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mesh = MESH[world]
    comm = A.create_procmesh_2d(*mesh)
    results = {}
    ok = True
    cases = [((32, 32, 32), 20000, 6, 0, False), ((32, 48, 40), 15000, 4, A.WINDOW_GAUSSIAN, False),
             ((32, 32, 32), 20000, 6, A.DIFF_IK, False), ((32, 32, 32), 20000, 6, 0, True),
             ((32, 48, 40), 15000, 6, A.TRANSPOSED_F_HAT, False), ((32, 32, 32), 20000, 4, A.TRANSPOSED_F_HAT, True),
             ((32, 32, 32), 20000, 6, A.INTERLACED, False), ((32, 48, 40), 15000, 4, A.INTERLACED | A.TRANSPOSED_F_HAT | A.DIFF_IK, False),
             # kernel family 3 (tensor-core gather / scatter with the window in shared memory): m = 8, 5, 7
             ((32, 32, 32), 20000, 8, A.WINDOW_GAUSSIAN, False), ((32, 48, 40), 15000, 5, A.WINDOW_BSPLINE, True),
             ((32, 32, 32), 20000, 7, A.INTERLACED, False)]
    if len(sys.argv) > 1:       # a larger problem: N^3 with M nodes, Kaiser-Bessel m=6 (python tools/mgpu_parity.py N M)
        cases = [((int(sys.argv[1]),) * 3, int(sys.argv[2]), 6, 0, False)]
    for ci, (N, M, m, flags, c2r) in enumerate(cases):
        n = tuple(2 * v for v in N)
        rng = np.random.default_rng(77 + ci)
        x = np.clip(rng.uniform(-0.5, 0.5, (M, 3)), -0.5, np.nextafter(0.5, 0.0))
        Nc = (N[0], N[1], N[2] // 2 + 1) if c2r else N
        fh = rng.uniform(-1, 1, Nc) + 1j * rng.uniform(-1, 1, Nc)
        if c2r:
            f = rng.uniform(-1, 1, M)
            g = rng.uniform(-1, 1, (M, 3))
        else:
            f = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
            g = rng.uniform(-1, 1, (M, 3)) + 1j * rng.uniform(-1, 1, (M, 3))
        lN, lNs, lo, up = A.local_size_guru(N, n, (0.5,) * 3, m, comm, c2r=c2r)
        mine = np.all((x >= lo) & (x < up), axis=1)
        idx = np.nonzero(mine)[0]
        xl = np.ascontiguousarray(x[idx])
        # my f_hat block: k_t = local_N_start[t] + i_t, global range [-N_t/2, N_t/2)  (c2r: k2 = 0 .. N2/2)
        lN, lNs, lo, up = A.local_size_guru(N, n, (0.5,) * 3, m, comm, pnfft_flags=flags, c2r=c2r)
        off = [int(lNs[t] + N[t] // 2) for t in range(3)]
        sl = tuple(slice(off[t], off[t] + int(lN[t])) for t in range(3))
        tr = bool(flags & A.TRANSPOSED_F_HAT)       # the library's block is then the (k1, k2, k0) transpose of this slice
        fh_l = np.ascontiguousarray(np.transpose(fh[sl], (1, 2, 0)) if tr else fh[sl])
        plan = A.Plan.init_guru(N, n, (0.5,) * 3, m, flags, comm, c2r=c2r)
        nodes = A.Nodes(len(idx), 0)
        ft = np.float64 if c2r else np.complex128
        fo, go = np.zeros(len(idx), ft), np.zeros((len(idx), 3), ft)
        nodes.set_x(xl); nodes.set_f(fo); nodes.set_grad_f(go)
        fhw = fh_l.copy()
        plan.set_f_hat(fhw)
        plan.trafo(nodes, 3)
        f_t, g_t = fo.copy(), go.copy()
        fo[...] = f[idx]; go[...] = g[idx]
        plan.adj(nodes, 3)
        nodes.free(0); plan.finalize(0)
        # gather on rank 0
        pack = (idx, f_t, g_t, sl, np.transpose(fhw, (2, 0, 1)) if tr else fhw)
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, pack)
        else:
            allp = [pack]
        if rank == 0:
            from oracle import checker
            ref = checker.get()
            rt = ref.trafo(N, x, fh, n=n, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
            ra = ref.adj(N, x, f=f, grad_f=g, n=n, m=m, pnfft_flags=flags, compute_flags=3, c2r=c2r)
            F_all, G_all = np.zeros(M, ft), np.zeros((M, 3), ft)
            H_all = np.zeros(Nc, np.complex128)
            count = 0
            for (i_, f_, g_, sl_, h_) in allp:
                F_all[i_] = f_; G_all[i_] = g_; H_all[sl_] = h_; count += len(i_)
            e = dict(f=rel_l2(F_all, rt["f"]), grad_f=rel_l2(G_all, rt["grad_f"]), f_hat=rel_l2(H_all, ra["f_hat"]),
                     nodes_owned_once=bool(count == M))
            results["case%d N=%s m=%d flags=%d c2r=%d" % (ci, N, m, flags, int(c2r))] = e
            ok = ok and count == M and max(e["f"], e["grad_f"], e["f_hat"]) <= 1e-13
    if rank == 0:
        print(json.dumps({"world": world, "mesh": "%dx%d" % mesh, "ok": bool(ok), "rel_l2": results}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
