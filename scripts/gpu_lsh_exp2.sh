#!/bin/bash
# timing experiment over kernel variants kept under scripts/exp/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp improving-inductive-oov-recsys_b200/csrc/tc_lsh.cu /tmp/tc_lsh_cur.cu
for v in v6c v7; do
  for d in "" "-DOOV_LSH_NORARE"; do
    cp scripts/exp/tc_lsh_$v.cu improving-inductive-oov-recsys_b200/csrc/tc_lsh.cu
    OOV_NVCC_EXTRA="$d" python -c "import importlib; b=importlib.import_module('improving-inductive-oov-recsys_b200.build'); b.build()" || exit 1
    echo "== $v $d"; python scripts/prof_lsh.py 5000000 2>&1 | tail -n 1
  done
done
cp /tmp/tc_lsh_cur.cu improving-inductive-oov-recsys_b200/csrc/tc_lsh.cu
