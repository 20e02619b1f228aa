#!/bin/bash
# A/B of build variants (scripts/build_variant.sh) of the score kernel at 10M and 1M items
cp improving-inductive-oov-recsys_b200/liboov_b200.so /tmp/lib_orig.so
for tag in "$@"; do
  echo "== $tag"
  cp build/variants/lib_$tag.so improving-inductive-oov-recsys_b200/liboov_b200.so
  python scripts/prof_score_10m.py 2>&1 | tail -1
  python scripts/prof_score_10m.py 1000000 2>&1 | tail -1
done
cp /tmp/lib_orig.so improving-inductive-oov-recsys_b200/liboov_b200.so
