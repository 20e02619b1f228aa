#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], d.get("stages"))
except Exception as e: print(f, "ERR", e)
PY
}
bash scripts/gpu_check.sh tests/test_gpu_dhe_context.py tests/test_gpu_tc.py
timeout 300 python scripts/prof_dhe.py 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err; echo rc=$?; tail -3 gpurun_out/b1.err; show gpurun_out/b1.json
