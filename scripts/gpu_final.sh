#!/bin/bash
# Round-end validation on one GPU: build check, smoke, tests, the default bench (three workload blocks), the reference arm,
# the ncu launch list of the same bench command and one full capture of the scoring main pass.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND:-r02}
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    def one(n, d):
        print(f, n, "value", round(d["value"]), d.get("unit"), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"),
              d.get("stages"), "roofline", (d.get("roofline") or {}).get("kernel","")[:30], (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
    one(d["config"]["workload"], d)
    for n, v in (d.get("workloads") or {}).items(): one(n, v)
except Exception as e: print(f, "ERR", e)
PY
}
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo rc=$?; tail -2 gpurun_out/${R}_bench.err; show gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; echo rc=$?; tail -c 600 gpurun_out/${R}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_score_main2 -s 3 -c 1 -o gpurun_out/${R}_score_main2 -f python scripts/prof_score_10m.py > gpurun_out/ncu_score.log 2>&1; echo rc=$?
