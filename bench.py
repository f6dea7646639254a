#!/usr/bin/env python
"""Headline benchmark of pnfft-b200 (contract: README of the build driver).

Default workload = BASELINE.json's metric (config C3): trafo+adj nonuniform points/s at N=256^3, sigma=2 (n=512^3),
Kaiser-Bessel m=6, double, c2c, M=2^24 uniform random nodes, 1/2/4/8 B200 (process mesh 1x1 / 1x2 / 2x2 / 2x4, strong
scaling: the problem is fixed).  `--config C2 | C4 | C5` runs the other BASELINE configurations through the same code
(they are parity-test cases for the driver, measured here for the record under profiles/):
  C2  N=128^3, M=2^21 uniform nodes, Kaiser-Bessel m=6, double, F only
  C4  N=256^3, M=2^24 strongly clustered nodes (Gaussian blob, sigma 0.05), m=8, --window gaussian|fast_gaussian|bspline,
      --pre-psi 0|1 (PNFFT_PRE_PSI|PRE_GRAD_PSI tables against on-the-fly window evaluation)
  C5  N=512^3 (n=1024^3), M=2^27 uniform nodes, Kaiser-Bessel m=6, --variant c2r (real input, double) | float (c2c single)

One "step" = pnfft_trafo(plan, nodes, cf_trafo) + pnfft_adj(plan, nodes, cf_adj) through the C ABI of libpnfft_b200.so.
  value : device-resident arrays (x, f, grad_f, f_hat are CUDA pointers): M_total / step time
  e2e   : HOST (pinned) arrays handed to the same C-ABI calls: the H2D copies of x, f_hat (trafo) / x, f (adj) and the
          D2H copies of f, grad_f (trafo) / f_hat (adj) happen inside the calls and inside the timed region
  roofline : the dominant gridding kernel against max(FP64 FMA time, HBM time) (north_star)
  parity : OUTSIDE the timed region, rank 0 compares a node subset of the last trafo and a small adjoint through the same
           (multi-rank) plan with the CPU checker (oracle/) and prints the rel-l2 errors
  cpu_baseline : the compiled reference PNFFT (oracle/_ref) on the host cores, bounded sample (rank 0, N=1 only)

`--impl reference` times the reference's own CPU implementation (oracle/_ref, else the oracle port) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

MESH = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4), 16: (4, 4)}
UNIT = "pts/s"
F_FAST_GAUSSIAN, F_WIN_GAUSSIAN, F_WIN_BSPLINE = 1 << 1, 1 << 13, 1 << 14
CF_F, CF_GRAD = 1, 2


def workload(args):
    """Everything that defines the timed work, from --config (and the overrides --N --log2M --m --flags)."""
    c = args.config
    w = dict(config=c, c2r=False, single=False, flags=args.flags, cf_trafo=CF_F | CF_GRAD, cf_adj=CF_F, dist="uniform",
             pre_psi=False, window="kaiser_bessel")
    if c == "C3":
        Ns, l2M, m = 256, 24, 6
        w["metric"] = "trafo+adj nonuniform pts/s, N=256^3 m=6 double"
    elif c == "C2":
        Ns, l2M, m = 128, 21, 6
        w["cf_trafo"] = CF_F
        w["metric"] = "trafo+adj nonuniform pts/s, N=128^3 m=6 double, F only (BASELINE config 2)"
    elif c == "C4":
        Ns, l2M, m = 256, 24, 8
        w["dist"] = "gaussian_blob_0.05"
        w["window"] = args.window
        w["flags"] |= {"gaussian": F_WIN_GAUSSIAN, "fast_gaussian": F_WIN_GAUSSIAN | F_FAST_GAUSSIAN, "bspline": F_WIN_BSPLINE}[args.window]
        w["pre_psi"] = bool(args.pre_psi)
        w["metric"] = "trafo+adj nonuniform pts/s, N=256^3 m=8 double, clustered nodes, %s window, %s (BASELINE config 4)" % (
            args.window, "PRE_PSI" if args.pre_psi else "on-the-fly")
    else:
        Ns, l2M, m = 512, 27, 6
        w["c2r"] = args.variant == "c2r"
        w["single"] = args.variant == "float"
        w["metric"] = "trafo+adj nonuniform pts/s, N=512^3 m=6 %s (BASELINE config 5)" % (
            "c2r double" if w["c2r"] else "c2c single")
    Ns = args.N or Ns
    l2M = args.log2M if args.log2M is not None else l2M
    m = args.m or m
    if (args.N or args.log2M is not None or args.m) and c == "C3":
        w["metric"] = "trafo+adj nonuniform pts/s, N=%d^3 m=%d double" % (Ns, m)
    w.update(N=(Ns,) * 3, n=(2 * Ns,) * 3, m=m, M_total=1 << l2M, log2M=l2M)
    return w


def config_dict(w, world, extra=None):
    ncomp, rb = (1 if w["c2r"] else 2), (4 if w["single"] else 8)
    mesh = MESH[world]
    grid_gb = np.prod([w["n"][0] / mesh[0] + 2 * w["m"], w["n"][1] / mesh[1] + 2 * w["m"], w["n"][2] + 2 * w["m"]]) * ncomp * rb / 1e9
    node_b = rb * (3 + ncomp + (3 * ncomp if w["cf_trafo"] & CF_GRAD else 0))
    c = {
        "workload": "%s: N=%d^3, n=%d^3 (sigma=2), M=2^%d %s nodes, %s m=%d, %s %s, step = pnfft_trafo(%s) + pnfft_adj(COMPUTE_F)%s"
                    % (w["config"], w["N"][0], w["n"][0], w["log2M"], "uniform random" if w["dist"] == "uniform" else "clustered (Gaussian blob sigma=0.05)",
                       w["window"], w["m"], "c2r" if w["c2r"] else "c2c", "single" if w["single"] else "double",
                       "COMPUTE_F|COMPUTE_GRAD_F, analytic gradient" if w["cf_trafo"] & CF_GRAD else "COMPUTE_F",
                       ", PNFFT_PRE_PSI|PRE_GRAD_PSI tables" if w["pre_psi"] else ""),
        "N": list(w["N"]), "n": list(w["n"]), "m": w["m"], "M_total": w["M_total"],
        "process_mesh": "%dx%d" % mesh, "window": w["window"], "precision": "single" if w["single"] else "double",
        "cache": "inputs larger than L2 (padded grid %.2f GB + nodes %.2f GB per rank vs 126 MB L2); no flush needed"
                 % (grid_gb, w["M_total"] / world * node_b / 1e9),
        "nodes_per_step": ("fixed (PNFFT_PRE_PSI tables belong to one set of coordinates)" if w["pre_psi"] else
                           "new coordinates every step (two node sets alternate, pnfft_set_x per step): upload, binning and "
                           "window table are redone in every pnfft_trafo; pnfft_adj of the same step reuses them"),
    }
    if extra:
        c.update(extra)
    return c


def global_f_hat(w):
    """The same global spectrum on every rank (seeded), so that any rank can hand its block to the library and rank 0 can
    hand the whole array to the CPU checker."""
    N = w["N"]
    Nc = (N[0], N[1], N[2] // 2 + 1) if w["c2r"] else N
    rng = np.random.default_rng(2000)
    fh = np.empty(Nc, np.complex128)
    fh.real = rng.uniform(-1, 1, Nc)
    fh.imag = rng.uniform(-1, 1, Nc)
    if w["c2r"]:            # planes k_t = -N_t/2 have no Hermitian partner inside [-N/2, N/2)
        fh[0, :, :] = 0; fh[:, 0, :] = 0; fh[:, :, 0] = 0
    return fh


def local_nodes(w, rank, M, lo, up):
    rng = np.random.default_rng(1000 + rank)
    if w["dist"] == "uniform":
        xv = rng.uniform(0.0, 1.0, (M, 3)) * (up - lo) + lo
        return np.minimum(np.maximum(xv, lo), np.nextafter(up, -1.0))
    # clustered: every rank draws the SAME global blob and keeps what falls into its [lo, up) (reference node ownership,
    # kernel/ndft-parallel.c:734-775): the ranks around the origin own almost everything -- that is the point of config 4
    out, rng = [], np.random.default_rng(4000)
    left = w["M_total"]
    while left > 0:
        k = min(left, 1 << 22)
        xg = np.mod(rng.normal(0.0, 0.05, (k, 3)) + 0.5, 1.0) - 0.5
        xg = np.clip(xg, -0.5, np.nextafter(0.5, 0.0))
        out.append(xg[np.all((xg >= lo) & (xg < up), axis=1)])
        left -= k
    return np.ascontiguousarray(np.concatenate(out))


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.rows, self.proc, self.thr = [], None, None
        cmd = ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"]
        if uuid:
            cmd += ["-i", uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            p = [v.strip() for v in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax = max(smax, float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_mesh(cores):
    p = 1
    while p * 2 <= min(cores, 16):
        p *= 2
    return MESH[p]


def reference_step(w, sample_log2M, seed=0):
    """One trafo+adj of the reference on a bounded sample: the full N, n, m plan and 2^sample_log2M of the nodes.
    Returns (pts/s extrapolated to the full node count, detail dict).  Extrapolation: the node loop (LOOP_B timer)
    scales linearly with the node count, D / F / ghost cells do not depend on it."""
    from oracle import checker
    ref = checker.get(w["single"])
    cores = os.cpu_count() or 1
    mesh = cpu_mesh(cores) if ref.threads else (1, 1)
    Ms = min(1 << sample_log2M, w["M_total"])
    rng = np.random.default_rng(seed)
    if w["dist"] == "uniform":
        x = rng.uniform(-0.5, 0.5, (Ms, 3))
    else:
        x = np.mod(rng.normal(0.0, 0.05, (Ms, 3)) + 0.5, 1.0) - 0.5
    x = np.clip(x, -0.5, np.nextafter(0.5, 0.0))
    fh = global_f_hat(w)
    # the reference cannot run PNFFT_PRE_PSI together with COMPUTE_GRAD_F: it never allocates pre_dpsi (its PNFFT_DIFF_AD
    # test is always false, SURVEY 8a defect 2) and dereferences it in the node loop; its on-the-fly path is timed then
    ref_pre = bool(w["pre_psi"]) and not (w["cf_trafo"] & CF_GRAD)
    kw = dict(n=w["n"], m=w["m"], np_mesh=mesh, pnfft_flags=w["flags"], c2r=w["c2r"],
              precompute_flags=6 if ref_pre else 0)
    t0 = time.time()
    rt = ref.trafo(w["N"], x, fh, compute_flags=w["cf_trafo"], **kw)
    t1 = time.time()
    ra = ref.adj(w["N"], x, f=rt["f"], compute_flags=w["cf_adj"], **kw)
    t2 = time.time()
    scale = w["M_total"] / Ms
    det = {"wall_trafo_s": t1 - t0, "wall_adj_s": t2 - t1}
    if rt.get("timers") is not None:
        names = ["iter", "whole", "loop_b", "sort_nodes", "gcells", "matrix_b", "matrix_f", "matrix_d"]
        T = dict(zip(names, rt["timers"][:, 0, :8].max(0)))
        Aj = dict(zip(names, ra["timers"][:, 1, :8].max(0)))
        full = 0.0
        for tm in (T, Aj):
            full += tm["whole"] - tm["loop_b"] + tm["loop_b"] * scale
        det.update(trafo={k: float(v) for k, v in T.items()}, adj={k: float(v) for k, v in Aj.items()})
        sample_s = T["whole"] + Aj["whole"]
    else:
        sample_s = (t2 - t0)
        full = sample_s * scale
    det.update(sample_s=float(sample_s), extrapolated_full_s=float(full), cores=mesh[0] * mesh[1], kind=ref.kind,
               sample="N=%d^3 n=%d^3 m=%d plan, 2^%d of 2^%d nodes, %dx%d ranks (one thread each); node loop scaled x%d, "
                      "D/F/ghost cells as measured (F by the oracle shim's host FFT, FFTW/PFFT are absent)%s"
                      % (w["N"][0], w["n"][0], w["m"], sample_log2M, w["log2M"], mesh[0], mesh[1], int(scale),
                         "; window factors on the fly (the reference segfaults with PNFFT_PRE_PSI and COMPUTE_GRAD_F: pre_dpsi "
                         "is never allocated)" if (w["pre_psi"] and not ref_pre) else ""))
    return w["M_total"] / full, det


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w = workload(args)
    steps = args.steps if args.steps is not None else 2
    warm = args.warmup if args.warmup is not None else 1
    vals, det = [], None
    t_begin = time.time()
    for it in range(warm + steps):
        t0 = time.time()
        v, det = reference_step(w, args.cpu_log2M, seed=it)
        if it >= warm:
            vals.append((v, det["extrapolated_full_s"], time.time() - t0))
    value = len(vals) / sum(1.0 / v for v, _, _ in vals)       # total points / total (extrapolated) time
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm,
        # what one step of THIS run took on the wall (a bounded sample of the workload) ...
        "ms_per_step": 1e3 * sum(t for _, _, t in vals) / len(vals),
        # ... and what `value` is computed from: the node loop of the sample scaled to all 2^log2M nodes
        "ms_per_step_full_workload_extrapolated": 1e3 * sum(t for _, t, _ in vals) / len(vals),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if w["single"] else "f64", "data": "synthetic (seeded nodes, random f_hat)",
        "config": config_dict(w, max(world, 1), {"process_mesh": "host: %d ranks" % det["cores"]}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
                         "stage_s_trafo": det.get("trafo"), "stage_s_adj": det.get("adj")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t_begin,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def bind_rank_to_gpu_numa_node(torch, local_rank):
    """What `mpirun --bind-to numa` does for a PNFFT caller: run this rank on the cores of the NUMA node its GPU hangs off,
    so that the pinned host arrays it allocates next (first touch) are local to that GPU's PCIe root.  torchrun binds
    nothing; with 8 ranks writing f / grad_f back at once, arrays that all sit on one socket cap the D2H side
    (round 1: 10 GB/s per GPU at N = 8).  PNFFT_B200_BENCH_NUMA=0 switches it off.  Returns what was done."""
    if os.environ.get("PNFFT_B200_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return "no NUMA information for %s" % bus
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        mine = cpus & os.sched_getaffinity(0)
        if not mine:
            return "node %d of %s has no core in this process's cpuset" % (node, bus)
        os.sched_setaffinity(0, mine)
        return "node %d (%d cores) for GPU %s" % (node, len(mine), bus)
    except Exception as e:      # the bench must run where sysfs says nothing
        return "unavailable (%s)" % (str(e)[:80],)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pnfft_b200 import api as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d"
                             % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices: pnfft_b200 has no CPU path")
    if world not in MESH:
        raise SystemExit("unsupported number of GPUs %d" % world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_rank_to_gpu_numa_node(torch, local_rank) if world > 1 else "not bound (one rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps = args.steps if args.steps is not None else 10
    warm = max(args.warmup if args.warmup is not None else 3, 3)

    w = workload(args)
    N, n, m = w["N"], w["n"], w["m"]
    c2r, single = w["c2r"], w["single"]
    rdt, tdt = (np.float32, torch.float32) if single else (np.float64, torch.float64)
    rb = 4 if single else 8
    NC = 1 if c2r else 2
    mesh = MESH[world]
    comm = A.create_procmesh_2d(*mesh)
    lN, lNs, lo, up = A.local_size_guru(N, n, (0.5,) * 3, m, comm, pnfft_flags=w["flags"], c2r=c2r, single=single)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    # ---- host (pinned) arrays: what a PNFFT caller owns ----
    if w["dist"] == "uniform":
        M = w["M_total"] // world
        xv = local_nodes(w, rank, M, lo.astype(np.float64), up.astype(np.float64))
    else:
        xv = local_nodes(w, rank, 0, lo.astype(np.float64), up.astype(np.float64))
        M = xv.shape[0]
    hx = pinned((max(M, 1), 3), tdt)[:M]
    hx.numpy()[...] = xv.astype(rdt)
    if single:      # float rounding may push a node onto the upper border
        np.minimum(hx.numpy(), np.nextafter(up.astype(np.float32), np.float32(-1)), out=hx.numpy())
    del xv
    fh_glob = global_f_hat(w)
    off = [int(lNs[t] + N[t] // 2) for t in range(3)]
    blk = tuple(slice(off[t], off[t] + int(lN[t])) for t in range(3))
    transposed = bool(w["flags"] & (1 << 11))     # PNFFT_TRANSPOSED_F_HAT: the block is stored (k1, k2, k0)
    order = (1, 2, 0) if transposed else (0, 1, 2)
    h_fhat_in = pinned(tuple(int(lN[t]) for t in order) + (2,), tdt)
    h_fhat_in.numpy()[...] = np.ascontiguousarray(np.transpose(fh_glob[blk], order)).view(np.float64).reshape(h_fhat_in.shape).astype(rdt)
    if not (rank == 0 and not args.no_parity):
        del fh_glob
    h_fhat_out = pinned(h_fhat_in.shape, tdt)
    fshape, gshape = ((M, NC) if NC == 2 else (M,)), ((M, 3, NC) if NC == 2 else (M, 3))
    hf, hg = pinned(fshape, tdt), pinned(gshape, tdt)
    # a second set of coordinates (the same nodes in reverse order): the timed loops alternate between the two, so that
    # every step meets coordinates the library has not seen in the previous call (nothing derived from x - upload, bins,
    # window table - survives from step to step; pnfft_adj may reuse what pnfft_trafo of the SAME step derived)
    hx2 = pinned((max(M, 1), 3), tdt)[:M]
    hx2.numpy()[...] = hx.numpy()[::-1]
    # ---- device-resident twins ----
    dx, d_fhat_in = hx.to(dev), h_fhat_in.to(dev)
    dx2 = hx2.to(dev)
    d_fhat_out = torch.zeros_like(d_fhat_in)
    df = torch.zeros(fshape, dtype=tdt, device=dev)
    dg = torch.zeros(gshape, dtype=tdt, device=dev)

    plan = A.Plan.init_guru(N, n, (0.5,) * 3, m, w["flags"], comm, c2r=c2r, single=single)
    nd_dev = A.Nodes(M, 0, single=single); nd_dev.set_x(dx); nd_dev.set_f(df); nd_dev.set_grad_f(dg)
    nd_host = A.Nodes(M, 0, single=single); nd_host.set_x(hx); nd_host.set_f(hf); nd_host.set_grad_f(hg)
    if w["pre_psi"]:
        plan.precompute_psi(nd_dev, A.PRE_PSI | A.PRE_GRAD_PSI)
        plan.precompute_psi(nd_host, A.PRE_PSI | A.PRE_GRAD_PSI)
    CF_T, CF_A = w["cf_trafo"], w["cf_adj"]
    stream = torch.cuda.ExternalStream(plan.stream(), device=dev)

    def step(nodes, f_in, f_out, x=None):
        if x is not None:
            nodes.set_x(x)
        plan.set_f_hat(f_in)
        plan.trafo(nodes, CF_T)
        st_t = plan.stage_ms(False)
        plan.set_f_hat(f_out)
        plan.adj(nodes, CF_A)
        return st_t, plan.stage_ms(True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nodes, f_in, f_out, k, xs=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stages = []
        barrier()
        t0 = time.time()
        e0.record(stream)
        for i in range(k):
            stages.append(step(nodes, f_in, f_out, None if xs is None else xs[i % 2]))
        e1.record(stream)
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), stages, (t0, t1)

    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        pass
    clocks = Clocks(uuid) if rank == 0 else None

    pre = bool(w["pre_psi"])       # PNFFT_PRE_PSI tables belong to ONE set of coordinates: those runs keep x fixed
    for i in range(warm):
        step(nd_dev, d_fhat_in, d_fhat_out, None if pre else (dx, dx2)[i % 2])
    l0, c0 = plan.kernel_launches(), plan.library_calls()
    ms_dev, stages, win = timed(nd_dev, d_fhat_in, d_fhat_out, steps, None if pre else (dx, dx2))
    l1, c1 = plan.kernel_launches(), plan.library_calls()
    # the same loop on unchanged coordinates (an iterative solver): bins and window table of the first call serve all others
    if not pre:
        nd_dev.set_x(dx)
    for _ in range(2):
        step(nd_dev, d_fhat_in, d_fhat_out)
    ms_dev_static, _, _ = timed(nd_dev, d_fhat_in, d_fhat_out, steps)
    for i in range(2):
        step(nd_host, h_fhat_in, h_fhat_out, None if pre else (hx, hx2)[i % 2])
    ms_e2e, stages_e2e, _ = timed(nd_host, h_fhat_in, h_fhat_out, steps, None if pre else (hx, hx2))
    # the same end-to-end step when the caller promises unchanged coordinates (pnfft_b200_nodes_x_static): x is uploaded
    # and binned once, not twice per step
    if not pre:
        nd_host.set_x(hx)
    nd_host.x_static(True)
    for _ in range(2):
        step(nd_host, h_fhat_in, h_fhat_out)
    ms_e2e_static, _, _ = timed(nd_host, h_fhat_in, h_fhat_out, steps)
    nd_host.x_static(False)
    clk = None
    if clocks:
        time.sleep(0.15)
        clk = clocks.window(*win)
        clocks.stop()

    # ---- per-kernel roofline (durations from CUDA events on the plan's stream around the kernel launch) ----
    c3 = float((2 * m + 1) ** 3)
    t_gather = statistics.mean(s[0]["b_kernel"] for s in stages) * 1e-3
    t_scatter = statistics.mean(s[1]["b_kernel"] for s in stages) * 1e-3
    Mmax = torch.tensor([float(M)], dtype=torch.float64, device=dev)
    red = torch.tensor([t_gather, t_scatter], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(Mmax, op=dist.ReduceOp.MAX)
    t_gather, t_scatter = float(red[0]), float(red[1])
    M_rf = float(Mmax.item())            # the slowest rank's kernel time belongs to the fullest rank
    fp_peak = float(A.measure_fp64_tflops())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    lno = [n[0] // mesh[0], n[1] // mesh[1], n[2]]
    grid_bytes = float(np.prod(lno)) * NC * rb
    grad = bool(CF_T & CF_GRAD)
    fl_g = (16 if grad else 4) * (NC / 2.0) * c3        # c2c 16 c^3 (F+grad) / 4 c^3 (F); r2r half of it
    fl_s = 4 * (NC / 2.0) * c3
    gname = "gather_f_grad" if grad else "gather_f"
    kern = {
        gname: {"flops": fl_g * M_rf, "bytes": grid_bytes + M_rf * rb * (3 + NC + (3 * NC if grad else 0)), "ms": t_gather * 1e3},
        "scatter_f": {"flops": fl_s * M_rf, "bytes": grid_bytes + M_rf * rb * (3 + NC), "ms": t_scatter * 1e3},
    }
    prec_note = ""
    if single:
        # the float kernels run on the FP32 FMA pipe; there is no measured FP32 peak in MEASURED_PEAKS.json, so the FP64
        # figure stays the (conservative) denominator and the fraction can exceed what a double kernel could reach
        prec_note = " (single-precision kernels: FP32 pipe, reported against the FP64 peak for lack of a measured FP32 figure)"
    for k in kern.values():
        k["tflops"] = k["flops"] / (k["ms"] * 1e-3) * 1e-12
        k["gbs"] = k["bytes"] / (k["ms"] * 1e-3) * 1e-9
        k["frac_fp64"] = k["tflops"] / fp_peak
        k["frac_hbm"] = k["gbs"] / hbm_peak
        k["roofline_ms"] = max(k["flops"] / (fp_peak * 1e12), k["bytes"] / (hbm_peak * 1e9)) * 1e3
    dom_name = max(kern, key=lambda q: kern[q]["ms"])
    dom = kern[dom_name]
    traffic = None
    if world == 1:      # ncu runs on one GPU only: the measured DRAM bytes belong to the N=1 launch of this config
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(w["config"], {}).get(dom_name)
        except Exception:
            pass
    bound_fp64 = dom["flops"] / (fp_peak * 1e12) >= dom["bytes"] / (hbm_peak * 1e9)
    roofline = {
        "kernel": dom_name, "bound": "fp64" if bound_fp64 else "hbm",
        "achieved": dom["tflops"] if bound_fp64 else dom["gbs"], "peak": fp_peak if bound_fp64 else hbm_peak,
        "unit": "TFLOP/s" if bound_fp64 else "GB/s", "frac": dom["frac_fp64"] if bound_fp64 else dom["frac_hbm"],
        "traffic": traffic,
        "peak_source": "FP64 FMA: measured live (independent DFMA chains on all SMs, pnfft_b200_measure_fp64_tflops); HBM: "
                       + hbm_src + prec_note,
        "hbm": {"achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac_hbm"]},
        "algorithmic": "flops/node: gather F+grad 16*(2m+1)^3 = the reference's four weighted sums per tap (a separable evaluation needs about half: frac may exceed 1 for large m), gather / scatter F 4*(2m+1)^3 (c2c; r2r half); bytes: grid "
                       "block once + x,f[,grad_f] once (SURVEY.md 8d); the kernel time of the slowest rank against the "
                       "node count of the fullest rank; kernel ms include the node-table kernel of the call "
                       "(new coordinates every step: pnfft_trafo builds the table, pnfft_adj of the same step reuses it)",
        "kernels": kern,
        "gridding_roofline_ms_per_step": sum(k["roofline_ms"] for k in kern.values()),
        "gridding_measured_ms_per_step": sum(k["ms"] for k in kern.values()),
    }
    roofline["gridding_frac"] = roofline["gridding_roofline_ms_per_step"] / roofline["gridding_measured_ms_per_step"]

    h2d = hx.numel() * rb * 2 + h_fhat_in.numel() * rb + hf.numel() * rb           # x twice, f_hat (trafo), f (adj)
    d2h = hf.numel() * rb + (hg.numel() * rb if grad else 0) + h_fhat_out.numel() * rb   # f, grad_f (trafo), f_hat (adj)
    tot = torch.tensor([float(h2d), float(d2h), float(M)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    M_total = int(tot[2].item())
    value = M_total / (ms_dev * 1e-3 / steps)
    e2e = M_total / (ms_e2e * 1e-3 / steps)
    mean_stage = lambda idx, key, S: statistics.mean(s[idx][key] for s in S)   # noqa: E731
    line = {
        "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if single else "f64", "data": "synthetic (seeded nodes in each rank's [lo,up), seeded random f_hat)",
        "config": config_dict(w, world, {"host_numa_binding": numa}),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                "ms_per_step": ms_e2e / steps},
        "value_x_static": {"value": M_total / (ms_dev_static * 1e-3 / steps), "unit": UNIT, "ms_per_step": ms_dev_static / steps,
                           "note": "device-resident loop on UNCHANGED coordinates (content hash): bins and window table of the "
                                   "first call serve every later trafo / adj"},
        "e2e_x_static": {"value": M_total / (ms_e2e_static * 1e-3 / steps), "unit": UNIT, "ms_per_step": ms_e2e_static / steps,
                         "h2d_bytes_per_step": int(tot[0].item()) - world * int(hx.numel() * rb * 2),
                         "note": "same step with pnfft_b200_nodes_x_static(nodes, 1): coordinates uploaded and binned once"},
        "gpu_launches": int(l1 - l0), "library_calls": int(c1 - c0),
        "roofline": roofline,
        "clocks": clk,
        "stage_ms": {"trafo": {k: mean_stage(0, k, stages) for k in stages[0][0]},
                     "adj": {k: mean_stage(1, k, stages) for k in stages[0][1]},
                     "trafo_e2e": {k: mean_stage(0, k, stages_e2e) for k in stages_e2e[0][0]},
                     "adj_e2e": {k: mean_stage(1, k, stages_e2e) for k in stages_e2e[0][1]}},
    }

    # ---- parity, outside the timed region: the last device-resident trafo on a node subset, and a small adjoint through
    #      the same (multi-rank) plan, against the CPU checker on the global problem ----
    if not args.no_parity:
        try:
            line["parity"] = parity_check(args, w, A, plan, comm, rank, world, dev, hx, h_fhat_in, blk, lN,
                                          fh_glob if rank == 0 else None)
        except Exception as e:   # the checker is optional equipment of the bench, never of the product
            line["parity"] = {"error": repr(e)}
    nd_dev.free(0); nd_host.free(0); plan.finalize(0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, det = reference_step(w, args.cpu_log2M)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
                                    "stage_s_trafo": det.get("trafo"), "stage_s_adj": det.get("adj"),
                                    "sample_wall_s": det["wall_trafo_s"] + det["wall_adj_s"]}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_check(args, w, A, plan, comm, rank, world, dev, hx, h_fhat_in, blk, lN, fh_glob):
    """rel-l2 of f, grad_f (trafo) and f_hat (adj) against the CPU checker, with KT / KA nodes per rank, through the plan
    that was just timed (host arrays: the path a PNFFT caller uses)."""
    import torch
    import torch.distributed as dist
    N, c2r, single = w["N"], w["c2r"], w["single"]
    rdt = np.float32 if single else np.float64
    NC = 1 if c2r else 2
    KT, KA = args.parity_nodes, 4 * args.parity_nodes
    M = hx.shape[0]
    kt, ka = min(KT, M), min(KA, M)
    x_t, x_a = np.ascontiguousarray(hx.numpy()[:kt]), np.ascontiguousarray(hx.numpy()[:ka])
    ft = rdt if c2r else (np.complex64 if single else np.complex128)
    # trafo on the first kt local nodes
    nd = A.Nodes(kt, 0, single=single)
    f_t, g_t = np.zeros(kt, ft), np.zeros((kt, 3), ft)
    nd.set_x(x_t); nd.set_f(f_t); nd.set_grad_f(g_t)
    plan.set_f_hat(h_fhat_in)
    plan.trafo(nd, w["cf_trafo"])
    nd.free(0)
    # adjoint of the first ka local nodes
    rng = np.random.default_rng(3000 + rank)
    f_a = rng.uniform(-1, 1, ka).astype(rdt) if c2r else (rng.uniform(-1, 1, ka) + 1j * rng.uniform(-1, 1, ka)).astype(ft)
    nd = A.Nodes(ka, 0, single=single)
    fa_buf = f_a.copy()
    nd.set_x(x_a); nd.set_f(fa_buf)
    transposed = bool(w["flags"] & (1 << 11))
    order = (1, 2, 0) if transposed else (0, 1, 2)
    h_out = np.zeros(tuple(int(lN[t]) for t in order), np.complex64 if single else np.complex128)
    plan.set_f_hat(h_out)
    plan.adj(nd, w["cf_adj"])
    nd.free(0)
    if transposed:
        h_out = np.ascontiguousarray(np.transpose(h_out, (2, 0, 1)))
    pack = (x_t, f_t, g_t, x_a, f_a, blk, h_out)
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, pack)
    else:
        allp = [pack]
    if rank != 0:
        return None
    from oracle import checker
    ref = checker.get(single)
    cores = os.cpu_count() or 1
    mesh = cpu_mesh(cores) if getattr(ref, "threads", False) else (1, 1)
    kw = dict(n=w["n"], m=w["m"], pnfft_flags=w["flags"], c2r=c2r, np_mesh=mesh)
    t0 = time.time()
    X = np.concatenate([p[0] for p in allp]); Fo = np.concatenate([p[1] for p in allp]); Go = np.concatenate([p[2] for p in allp])
    rt = ref.trafo(N, X, fh_glob, compute_flags=w["cf_trafo"], **kw)
    XA = np.concatenate([p[3] for p in allp]); FA = np.concatenate([p[4] for p in allp])
    ra = ref.adj(N, XA, f=FA, compute_flags=w["cf_adj"], **kw)
    H = np.zeros(ra["f_hat"].shape, np.complex128)
    for p in allp:
        H[p[5]] = p[6]

    def rl2(a, b):
        d = np.linalg.norm(np.ravel(b))
        return float(np.linalg.norm(np.ravel(a).astype(np.complex128) - np.ravel(b)) / (d if d > 0 else 1.0))
    out = {"f": rl2(Fo, rt["f"]), "f_hat": rl2(H, ra["f_hat"]), "checker": ref.name, "trafo_nodes": int(X.shape[0]),
           "adj_nodes": int(XA.shape[0]), "ranks": world, "bar": 1e-5 if single else 1e-13, "checker_wall_s": None}
    if w["cf_trafo"] & CF_GRAD:
        out["grad_f"] = rl2(Go, rt["grad_f"])
    out["parity_rel_l2"] = max(v for k, v in out.items() if k in ("f", "grad_f", "f_hat"))
    out["ok"] = bool(out["parity_rel_l2"] <= out["bar"])
    out["checker_wall_s"] = time.time() - t0
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--window", default="gaussian", choices=["gaussian", "fast_gaussian", "bspline"], help="C4 only")
    ap.add_argument("--pre-psi", dest="pre_psi", type=int, default=0, help="C4 only: PNFFT_PRE_PSI|PRE_GRAD_PSI tables")
    ap.add_argument("--variant", default="c2r", choices=["c2r", "float"], help="C5 only")
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--log2M", type=int, default=None)
    ap.add_argument("--m", type=int, default=None)
    ap.add_argument("--flags", type=int, default=0, help="extra pnfft plan flags (default: analytic gradient)")
    ap.add_argument("--cpu-log2M", dest="cpu_log2M", type=int, default=20,
                    help="node sample of the CPU legs (cpu_baseline and --impl reference use the same one)")
    ap.add_argument("--parity-nodes", dest="parity_nodes", type=int, default=2048, help="trafo nodes per rank in the parity leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
