#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dhe1m', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['stages'], d['roofline']['kernel'][:40], d['roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear2_kernel -s 6 -c 1 -o gpurun_out/r01_linear2 -f python scripts/prof_dhe.py > gpurun_out/ncu_linear2.log 2>&1; echo rc=$?
