#!/bin/bash
# timeline of CTA 0 of the LSH kernel: rebuild tc_lsh.cu with the trace hooks on the box (the snapshot's product .so has none)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
touch improving-inductive-oov-recsys_b200/csrc/tc_lsh.cu
OOV_NVCC_EXTRA="-DOOV_LSH_TRACE ${TRACE_DEFS:-}" python -c "import importlib; b=importlib.import_module('improving-inductive-oov-recsys_b200.build'); b.build()" || exit 1
python scripts/trace_lsh.py ${1:-900} > gpurun_out/lsh_trace.txt 2>&1
tail -n 3 gpurun_out/lsh_trace.txt
