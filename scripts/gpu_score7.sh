#!/bin/bash
for m in 0 1; do
  OOV_SCORE_MAIN2=$m python bench.py --steps 10 --no-cpu-baseline --workload dhe1m --single 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main2=$m dhe1m', d['ms_per_step'], d['stages']['score_topk_ms'])"
  OOV_SCORE_MAIN2=$m python bench.py --steps 10 --no-cpu-baseline --workload lsh10m --single 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main2=$m lsh10m', d['ms_per_step'], d['stages']['score_topk_ms'])"
done
