"""Host-side pieces of bench.py that need no GPU: the workload table (BASELINE configs 2-5), the config description,
the NUMA helper's behaviour where sysfs / CUDA say nothing, and the reference arm on a tiny problem (the compiled
reference is the checker, oracle/_ref; here it is the thing timed, which is the one other place bench.py may execute it)."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=None, warmup=None, impl="b200", config="C3", window="gaussian", pre_psi=0, variant="c2r", N=None,
             log2M=None, m=None, flags=0, cpu_log2M=20, parity_nodes=2048, no_cpu_baseline=False, no_parity=False)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("cfg,kw,N,l2M,m", [("C3", {}, 256, 24, 6), ("C2", {}, 128, 21, 6), ("C4", dict(window="bspline", pre_psi=1), 256, 24, 8),
                                           ("C5", dict(variant="c2r"), 512, 27, 6), ("C5", dict(variant="float"), 512, 27, 6)])
def test_workloads_are_the_baseline_configs(cfg, kw, N, l2M, m):
    w = bench.workload(_args(config=cfg, **kw))
    assert w["N"] == (N,) * 3 and w["n"] == (2 * N,) * 3 and w["m"] == m and w["M_total"] == 1 << l2M
    assert w["c2r"] == (kw.get("variant") == "c2r" and cfg == "C5")
    assert w["single"] == (kw.get("variant") == "float" and cfg == "C5")
    assert (w["dist"] != "uniform") == (cfg == "C4")
    assert bool(w["cf_trafo"] & bench.CF_GRAD) == (cfg != "C2")
    for world in (1, 2, 4, 8):
        c = bench.config_dict(w, world, {"host_numa_binding": "x"})
        assert c["workload"].startswith(cfg + ":") and c["M_total"] == w["M_total"] and c["host_numa_binding"] == "x"
        assert c["process_mesh"] == "%dx%d" % bench.MESH[world]
        assert ("PRE_PSI" in c["workload"]) == bool(w["pre_psi"])


def test_numa_binding_is_harmless_without_information():
    class NoCuda:
        class cuda:
            @staticmethod
            def get_device_properties(i):
                raise RuntimeError("no CUDA device")
    before = os.sched_getaffinity(0)
    msg = bench.bind_rank_to_gpu_numa_node(NoCuda, 0)
    assert msg.startswith("unavailable") and os.sched_getaffinity(0) == before
    os.environ["PNFFT_B200_BENCH_NUMA"] = "0"
    try:
        assert bench.bind_rank_to_gpu_numa_node(NoCuda, 0) == "off"
    finally:
        del os.environ["PNFFT_B200_BENCH_NUMA"]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on a small problem: one JSON line with the arm's keys, the metric and config of the GPU arm."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--N", "16", "--log2M", "10", "--cpu-log2M", "10",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "pts/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["M_total"] == 1 << 10 and d["n_gpus"] == 1
