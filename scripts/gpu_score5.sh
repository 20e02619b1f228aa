#!/bin/bash
mkdir -p gpurun_out
for st in 4 16; do for n in 1000000 2000000 4000000; do
  OOV_SCORE_STRIDE=$st ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/sc_${st}_${n}.csv python scripts/prof_score_10m.py $n > /dev/null 2>&1
  echo "stride $st N $n: $(grep -h 'main2\|topk_kernel<1>\|threshold\|merge_keys' gpurun_out/sc_${st}_${n}.csv | tail -4 | awk -F'","' '{printf "%s=%s  ", substr($5,1,28), $NF}' | tr -d '"')"
done; done
