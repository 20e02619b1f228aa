#!/bin/bash
# timing experiment: rebuild tc_lsh.cu on the box with extra defines ($1), time the kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
touch improving-inductive-oov-recsys_b200/csrc/tc_lsh.cu
OOV_NVCC_EXTRA="$1" python -c "import importlib; b=importlib.import_module('improving-inductive-oov-recsys_b200.build'); b.build()" || exit 1
python scripts/prof_lsh.py 5000000 2>&1 | tail -n 1
