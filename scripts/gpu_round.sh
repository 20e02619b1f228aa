#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python scripts/prof_linear_gelu.py | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear2_kernel -s 4 -c 1 -o gpurun_out/r01_linear2_gelu -f python scripts/prof_linear_gelu.py > gpurun_out/ncu_linear2_gelu.log 2>&1; echo rc=$?
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
