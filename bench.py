#!/usr/bin/env python
"""Headline benchmark of pnfft-b200 (contract: README of the build driver).

Metric (BASELINE.json): trafo+adj nonuniform points/s at N=256^3, sigma=2 (n=512^3), Kaiser-Bessel m=6, double, c2c,
M=2^24 uniform random nodes, 1/2/4/8 B200 (process mesh 1x1 / 1x2 / 2x2 / 2x4, strong scaling: the problem is fixed).

One "step" = pnfft_trafo(plan, nodes, PNFFT_COMPUTE_F|PNFFT_COMPUTE_GRAD_F) + pnfft_adj(plan, nodes, PNFFT_COMPUTE_F)
through the C ABI of libpnfft_b200.so.
  value : device-resident arrays (x, f, grad_f, f_hat are CUDA pointers): M_total / step time
  e2e   : HOST (pinned) arrays handed to the same C-ABI calls: the H2D copies of x, f_hat (trafo) / x, f (adj) and the
          D2H copies of f, grad_f (trafo) / f_hat (adj) happen inside the calls and inside the timed region
  roofline : the dominant gridding kernel against max(FP64 FMA time, HBM time) (north_star)
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
METRIC = "trafo+adj nonuniform pts/s, N=256^3 m=6 double"
UNIT = "pts/s"


def workload(args):
    N = (args.N,) * 3
    return dict(N=N, n=tuple(2 * v for v in N), m=args.m, M_total=1 << args.log2M)


def config_dict(args, world, extra=None):
    w = workload(args)
    c = {
        "workload": "C3: N=%d^3, n=%d^3 (sigma=2), M=2^%d uniform random nodes, Kaiser-Bessel m=%d, c2c double, "
                    "step = pnfft_trafo(COMPUTE_F|COMPUTE_GRAD_F, analytic gradient) + pnfft_adj(COMPUTE_F)"
                    % (args.N, 2 * args.N, args.log2M, args.m),
        "N": list(w["N"]), "n": list(w["n"]), "m": args.m, "M_total": w["M_total"],
        "process_mesh": "%dx%d" % MESH[world], "window": "kaiser_bessel", "precision": "double",
        "cache": "inputs larger than L2 (padded grid %.2f GB + nodes %.2f GB per rank vs 126 MB L2); no flush needed"
                 % (np.prod([2 * args.N / MESH[world][0] + 2 * args.m, 2 * args.N / MESH[world][1] + 2 * args.m,
                             2 * args.N + 2 * args.m]) * 16 / 1e9, w["M_total"] / world * 88 / 1e9),
    }
    if extra:
        c.update(extra)
    return c


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


def reference_step(args, sample_log2M, seed=0):
    """One trafo(F|GRAD)+adj(F) of the reference on a bounded sample: the full N, n, m plan and 2^sample_log2M of the
    2^log2M nodes.  Returns (pts/s extrapolated to the full node count, detail dict).  Extrapolation: the node loop
    (LOOP_B timer) scales linearly with the node count, D / F / ghost cells do not depend on it."""
    from oracle import checker
    ref = checker.get()
    w = workload(args)
    cores = os.cpu_count() or 1
    mesh = cpu_mesh(cores) if ref.threads else (1, 1)
    Ms = min(1 << sample_log2M, w["M_total"])
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.5, 0.5, (Ms, 3))
    x = np.clip(x, -0.5, np.nextafter(0.5, 0.0))
    fh = rng.uniform(-1, 1, w["N"]) + 1j * rng.uniform(-1, 1, w["N"])
    t0 = time.time()
    rt = ref.trafo(w["N"], x, fh, n=w["n"], m=w["m"], np_mesh=mesh, compute_flags=3)
    t1 = time.time()
    ra = ref.adj(w["N"], x, f=rt["f"], n=w["n"], m=w["m"], np_mesh=mesh, compute_flags=1)
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
                      "D/F/ghost cells as measured (F by the oracle shim's host FFT, FFTW/PFFT are absent)"
                      % (args.N, 2 * args.N, args.m, sample_log2M, args.log2M, mesh[0], mesh[1], int(scale)))
    return w["M_total"] / full, det


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps = args.steps if args.steps is not None else 2
    warm = args.warmup if args.warmup is not None else 1
    vals, det = [], None
    t_begin = time.time()
    for it in range(warm + steps):
        v, det = reference_step(args, args.ref_log2M, seed=it)
        if it >= warm:
            vals.append((v, det["extrapolated_full_s"]))
    value = len(vals) / sum(1.0 / v for v, _ in vals)       # total points / total time
    ms = 1e3 * sum(t for _, t in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (seeded uniform nodes, random f_hat)",
        "config": config_dict(args, max(world, 1), {"process_mesh": "host: %d ranks" % det["cores"]}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
                         "stage_s_trafo": det.get("trafo"), "stage_s_adj": det.get("adj")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t_begin,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps = args.steps if args.steps is not None else 10
    warm = max(args.warmup if args.warmup is not None else 3, 3)

    w = workload(args)
    N, n, m = w["N"], w["n"], w["m"]
    mesh = MESH[world]
    comm = A.create_procmesh_2d(*mesh)
    lN, lNs, lo, up = A.local_size_guru(N, n, (0.5,) * 3, m, comm)
    M = w["M_total"] // world
    rng = np.random.default_rng(1000 + rank)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    # ---- host (pinned) arrays: what a PNFFT caller owns ----
    hx = pinned((M, 3), torch.float64)
    xv = rng.uniform(0.0, 1.0, (M, 3)) * (up - lo) + lo
    xv = np.minimum(np.maximum(xv, lo), np.nextafter(up, -1.0))
    hx.numpy()[...] = xv
    del xv
    h_fhat_in = pinned(tuple(int(v) for v in lN) + (2,), torch.float64)
    h_fhat_in.numpy()[...] = rng.uniform(-1, 1, h_fhat_in.shape)
    h_fhat_out = pinned(h_fhat_in.shape, torch.float64)
    hf = pinned((M, 2), torch.float64)
    hg = pinned((M, 3, 2), torch.float64)
    # ---- device-resident twins ----
    dx, d_fhat_in = hx.to(dev), h_fhat_in.to(dev)
    d_fhat_out = torch.zeros_like(d_fhat_in)
    df = torch.zeros((M, 2), dtype=torch.float64, device=dev)
    dg = torch.zeros((M, 3, 2), dtype=torch.float64, device=dev)

    plan = A.Plan.init_guru(N, n, (0.5,) * 3, m, args.flags, comm)
    nd_dev = A.Nodes(M, 0); nd_dev.set_x(dx); nd_dev.set_f(df); nd_dev.set_grad_f(dg)
    nd_host = A.Nodes(M, 0); nd_host.set_x(hx); nd_host.set_f(hf); nd_host.set_grad_f(hg)
    CF_T, CF_A = A.COMPUTE_F | A.COMPUTE_GRAD_F, A.COMPUTE_F
    stream = torch.cuda.ExternalStream(plan.stream(), device=dev)

    def step(nodes, f_in, f_out):
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

    def timed(nodes, f_in, f_out, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stages = []
        barrier()
        t0 = time.time()
        e0.record(stream)
        for _ in range(k):
            stages.append(step(nodes, f_in, f_out))
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

    for _ in range(warm):
        step(nd_dev, d_fhat_in, d_fhat_out)
    l0, c0 = plan.kernel_launches(), plan.library_calls()
    ms_dev, stages, win = timed(nd_dev, d_fhat_in, d_fhat_out, steps)
    l1, c1 = plan.kernel_launches(), plan.library_calls()
    for _ in range(2):
        step(nd_host, h_fhat_in, h_fhat_out)
    ms_e2e, stages_e2e, _ = timed(nd_host, h_fhat_in, h_fhat_out, steps)
    clk = None
    if clocks:
        time.sleep(0.15)
        clk = clocks.window(*win)
        clocks.stop()

    # ---- per-kernel roofline (durations from CUDA events on the plan's stream around the kernel launch) ----
    c3 = float((2 * m + 1) ** 3)
    t_gather = statistics.mean(s[0]["b_kernel"] for s in stages) * 1e-3
    t_scatter = statistics.mean(s[1]["b_kernel"] for s in stages) * 1e-3
    red = torch.tensor([t_gather, t_scatter], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    t_gather, t_scatter = float(red[0]), float(red[1])
    fp64_peak = float(A.measure_fp64_tflops())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    lno = [n[0] // mesh[0], n[1] // mesh[1], n[2]]
    grid_bytes = float(np.prod(lno)) * 16
    kern = {
        "gather_f_grad": {"flops": 16 * c3 * M, "bytes": grid_bytes + M * (24 + 16 + 48), "ms": t_gather * 1e3},
        "scatter_f": {"flops": 4 * c3 * M, "bytes": grid_bytes + M * (24 + 16), "ms": t_scatter * 1e3},
    }
    for k in kern.values():
        k["tflops"] = k["flops"] / (k["ms"] * 1e-3) * 1e-12
        k["gbs"] = k["bytes"] / (k["ms"] * 1e-3) * 1e-9
        k["frac_fp64"] = k["tflops"] / fp64_peak
        k["frac_hbm"] = k["gbs"] / hbm_peak
        k["roofline_ms"] = max(k["flops"] / (fp64_peak * 1e12), k["bytes"] / (hbm_peak * 1e9)) * 1e3
    dom_name = max(kern, key=lambda q: kern[q]["ms"])
    dom = kern[dom_name]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom_name)
    except Exception:
        pass
    bound_fp64 = dom["flops"] / (fp64_peak * 1e12) >= dom["bytes"] / (hbm_peak * 1e9)
    roofline = {
        "kernel": dom_name, "bound": "fp64" if bound_fp64 else "hbm",
        "achieved": dom["tflops"] if bound_fp64 else dom["gbs"], "peak": fp64_peak if bound_fp64 else hbm_peak,
        "unit": "TFLOP/s" if bound_fp64 else "GB/s", "frac": dom["frac_fp64"] if bound_fp64 else dom["frac_hbm"],
        "traffic": traffic,
        "peak_source": "FP64 FMA: measured live (independent DFMA chains on all SMs, pnfft_b200_measure_fp64_tflops); HBM: "
                       + hbm_src,
        "hbm": {"achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac_hbm"]},
        "algorithmic": "flops/node: gather F+grad 16*(2m+1)^3, scatter F 4*(2m+1)^3; bytes: grid block once + x,f[,grad_f] "
                       "once (SURVEY.md 8d)",
        "kernels": kern,
        "gridding_roofline_ms_per_step": sum(k["roofline_ms"] for k in kern.values()),
        "gridding_measured_ms_per_step": sum(k["ms"] for k in kern.values()),
    }
    roofline["gridding_frac"] = roofline["gridding_roofline_ms_per_step"] / roofline["gridding_measured_ms_per_step"]

    h2d = hx.numel() * 8 * 2 + h_fhat_in.numel() * 8 + hf.numel() * 8           # x twice, f_hat (trafo), f (adj)
    d2h = hf.numel() * 8 + hg.numel() * 8 + h_fhat_out.numel() * 8              # f, grad_f (trafo), f_hat (adj)
    value = w["M_total"] / (ms_dev * 1e-3 / steps)
    e2e = w["M_total"] / (ms_e2e * 1e-3 / steps)
    mean_stage = lambda idx, key, S: statistics.mean(s[idx][key] for s in S)   # noqa: E731
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (seeded uniform nodes in each rank's [lo,up), random f_hat)",
        "config": config_dict(args, world),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(l1 - l0), "library_calls": int(c1 - c0),
        "roofline": roofline,
        "clocks": clk,
        "stage_ms": {"trafo": {k: mean_stage(0, k, stages) for k in stages[0][0]},
                     "adj": {k: mean_stage(1, k, stages) for k in stages[0][1]},
                     "trafo_e2e": {k: mean_stage(0, k, stages_e2e) for k in stages_e2e[0][0]},
                     "adj_e2e": {k: mean_stage(1, k, stages_e2e) for k in stages_e2e[0][1]}},
    }
    nd_dev.free(0); nd_host.free(0); plan.finalize(0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, det = reference_step(args, args.cpu_log2M)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": det["cores"], "kind": det["kind"], "sample": det["sample"],
                                    "stage_s_trafo": det.get("trafo"), "stage_s_adj": det.get("adj"),
                                    "sample_wall_s": det["wall_trafo_s"] + det["wall_adj_s"]}
        except Exception as e:   # the checker is optional equipment of the bench, never of the product
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--N", type=int, default=256)
    ap.add_argument("--log2M", type=int, default=24)
    ap.add_argument("--m", type=int, default=6)
    ap.add_argument("--flags", type=int, default=0, help="pnfft plan flags (default: Kaiser-Bessel, analytic gradient)")
    ap.add_argument("--cpu-log2M", dest="cpu_log2M", type=int, default=20, help="node sample of the cpu_baseline leg")
    ap.add_argument("--ref-log2M", dest="ref_log2M", type=int, default=18, help="node sample per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
