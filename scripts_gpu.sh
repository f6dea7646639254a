cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g ms %.2f e2e %.4g e2e_ms %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), "frac", d["roofline"]["frac"], d["roofline"].get("gridding_frac"), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
