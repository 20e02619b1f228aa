#!/bin/bash
# smoke + all GPU tests + the default bench line (all workload blocks)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo rc=$?; tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench.json").read().strip().splitlines()[-1])
def one(n, d):
    print(n, "value", round(d["value"]), d.get("unit"), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"),
          d.get("stages"), "roofline", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
one(d["config"]["workload"], d)
for n, v in (d.get("workloads") or {}).items(): one(n, v)
PY
