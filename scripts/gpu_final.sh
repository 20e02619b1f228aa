#!/bin/bash
# Round-end validation on one GPU: tests, smoke, both workloads, the reference arm, ncu launch lists.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"), d.get("stages"), "roofline", (d.get("roofline") or {}).get("kernel","")[:30], (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
except Exception as e: print(f, "ERR", e)
PY
}
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r01_bench_dhe1m.json 2> gpurun_out/r01_bench_dhe1m.err; echo rc=$?; tail -2 gpurun_out/r01_bench_dhe1m.err; show gpurun_out/r01_bench_dhe1m.json
timeout 600 python bench.py --workload lsh10m --steps 10 > gpurun_out/r01_bench_lsh10m.json 2> gpurun_out/r01_bench_lsh10m.err; echo rc=$?; show gpurun_out/r01_bench_lsh10m.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref_dhe1m.json 2> gpurun_out/r01_bench_ref.err; echo rc=$?; show gpurun_out/r01_bench_ref_dhe1m.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_dhe1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bench_lsh10m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload lsh10m > gpurun_out/ncu_bench_lsh.log 2>&1; echo rc=$?
