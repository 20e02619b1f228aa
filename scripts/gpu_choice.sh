#!/bin/bash
for a in "8 lsh10m" "4 lsh10m" "2 lsh10m" "1 lsh10m" "1 dhe1m"; do
  set -- $a
  echo "== $2 / $1: $(OOV_SCORE_DEBUG=32 python scripts/prof_shard_step.py $1 $2 2>&1 | grep -E "score_choose" | sort | uniq -c | head -2 | tr '\n' ' ') :: $(python scripts/prof_shard_step.py $1 $2 2>&1 | grep -E "topk \(CSR" | cut -c60-80)"
done
