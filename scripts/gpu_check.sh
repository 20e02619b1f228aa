#!/bin/bash
# Run on the GPU box: every GPU test file in its own process (a CUDA fault in one file cannot poison the
# others), each under a hard timeout; logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 600 -p no:cacheprovider -x -s > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "== $f -> rc=$rc"
  tail -n 25 "gpurun_out/${name}.log"
  [ $rc -ne 0 ] && status=1
done
exit $status
