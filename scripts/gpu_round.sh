#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^tc_linear_kernel$ -s 2 -c 1 -o gpurun_out/r01_linear64 -f python scripts/prof_dhe.py > gpurun_out/ncu_linear64.log 2>&1; echo rc=$?
