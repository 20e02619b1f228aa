#!/bin/bash
# LSH tensor-core kernel: parity tests, timing, one ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "lsh" --timeout 300 -p no:cacheprovider > gpurun_out/lsh_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/lsh_tests.log
timeout 300 python scripts/prof_lsh.py 5000000 > gpurun_out/lsh_time.log 2>&1; echo "time rc=$?"; cat gpurun_out/lsh_time.log
timeout 300 python scripts/prof_lsh.py 1000000 >> gpurun_out/lsh_time.log 2>&1; tail -n 3 gpurun_out/lsh_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_lsh_embed_kernel -s 2 -c 1 -o gpurun_out/r02_lsh -f python scripts/prof_lsh.py 5000000 > gpurun_out/ncu_lsh.log 2>&1; echo "ncu rc=$?"
