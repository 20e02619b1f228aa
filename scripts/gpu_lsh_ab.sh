#!/bin/bash
cp improving-inductive-oov-recsys_b200/liboov_b200.so /tmp/lib_orig.so
for tag in lshold lshnew lshold lshnew; do
  cp build/variants/lib_$tag.so improving-inductive-oov-recsys_b200/liboov_b200.so
  echo "== $tag: $(python bench.py --steps 10 --no-cpu-baseline --single 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stages']['lsh_embed_oov_ms'])")"
done
cp /tmp/lib_orig.so improving-inductive-oov-recsys_b200/liboov_b200.so
