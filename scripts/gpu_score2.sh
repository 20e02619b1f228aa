#!/bin/bash
# main-pass A/B (MODE 0 vs column-split) at 10M and 1M, parity tests of the retrieval path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_tc.py tests/test_gpu_scale.py -q -x 2>&1 | tail -3
OOV_SCORE_MAIN2=0 python scripts/prof_score_10m.py 2>&1 | tail -1
OOV_SCORE_MAIN2=1 python scripts/prof_score_10m.py 2>&1 | tail -1
OOV_SCORE_MAIN2=0 python scripts/prof_score_10m.py 1000000 2>&1 | tail -1
OOV_SCORE_MAIN2=1 python scripts/prof_score_10m.py 1000000 2>&1 | tail -1
