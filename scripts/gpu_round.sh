#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash scripts/gpu_check.sh tests/test_gpu_tc.py tests/test_gpu_retrieval.py
timeout 300 python scripts/prof_score.py 2>&1 | tee gpurun_out/score_sweep.log
timeout 300 python scripts/prof_small.py 2>&1 | tee gpurun_out/score_small.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_score_one.csv python scripts/prof_score.py one > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_score_topk_kernel -c 2 -o gpurun_out/r01_score_v2 -f python scripts/prof_score.py one > gpurun_out/ncu_score_v2.log 2>&1
