#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -m gpu -k "graphed or sharded" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/prof_shard_step.py 8 dhe1m 2>&1 | tail -7
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dhe1m', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['gpu_launches'])"
