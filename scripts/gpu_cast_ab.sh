#!/bin/bash
# A/B of the LSH side cast (OOV_LSH_FUSE_CAST=0: the cast as its own launch in front of the LSH kernel)
set -x
python -m pytest tests/test_gpu_tc.py -x -q -k "side_cast or tc_lsh or graphed" 2>&1 | tail -5
for v in 1 0; do
  OOV_LSH_FUSE_CAST=$v python bench.py --single --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/cast_ab_$v.json 2> gpurun_out/cast_ab_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/cast_ab_$v.json").read().strip().splitlines()[-1])
print("FUSE_CAST=$v", d["config"]["workload"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "stages", d.get("stages"))
PY
done
