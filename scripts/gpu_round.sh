#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests/test_gpu_tc.py
timeout 600 python scripts/prof_shard_step.py 8 dhe1m 2>&1 | tail -8 | tee gpurun_out/shard_step_dhe.log
