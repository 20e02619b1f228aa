#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests/test_gpu_tc.py tests/test_gpu_dhe_context.py
timeout 300 python scripts/prof_dhe.py 2>&1 | tail -1
timeout 300 python scripts/prof_linear_sweep.py 2>&1 | grep "N=256 K=512"
