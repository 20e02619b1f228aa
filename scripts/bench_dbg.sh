for d in 4 6; do echo "== debug $d"; OOV_SCORE_DEBUG=$d python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "^cta" | sort | uniq -c | sort -rn | awk '{print}' | head -5; done
