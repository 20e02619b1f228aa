#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests/test_gpu_tc.py tests/test_gpu_retrieval.py
timeout 300 python scripts/prof_score.py 2>&1 | tee gpurun_out/score_sweep.log
timeout 300 python scripts/prof_small.py 2>&1 | tee gpurun_out/score_small.log
timeout 300 python scripts/prof_dhe.py 2>&1 | tee gpurun_out/dhe.log
timeout 300 python scripts/prof_lsh.py 2000000 2>&1 | tee gpurun_out/lsh.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_score_one.csv python scripts/prof_score.py one > /dev/null 2>&1
