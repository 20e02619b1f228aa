#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_score_topk_kernel -c 2 -o gpurun_out/r01_score_v4 -f python scripts/prof_score_10m.py 10000000 > gpurun_out/ncu_score_v4.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_linear2_kernel -s 4 -c 1 -o gpurun_out/r01_linear2_gelu -f python scripts/prof_linear_gelu.py > gpurun_out/ncu_linear2_gelu.log 2>&1; echo rc=$?
