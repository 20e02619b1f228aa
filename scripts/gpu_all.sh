#!/bin/bash
# all GPU tests (one process per file) + both bench workloads
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests/test_gpu_*.py > gpurun_out/gpu_tests_summary.txt 2>&1; echo "tests rc=$?"; grep -E "^== |passed|failed|error" gpurun_out/gpu_tests_summary.txt | tail -n 20
for w in ${BENCH_WORKLOADS:-lsh10m dhe1m}; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 ${BENCH_EXTRA:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["value"]), "q/s", round(d["ms_per_step"],3), "ms/step  e2e", round(d["e2e"]["value"]), d.get("stages"))
    r=d.get("roofline") or {}
    print("  roofline", r.get("kernel","")[:40], r.get("frac"))
except Exception as e:
    print("no bench line:", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
