#!/bin/bash
for st in 4 8 16; do OOV_SCORE_STRIDE=$st python scripts/prof_score_10m.py 2>&1 | tail -1; done
for st in 2 4 8; do OOV_SCORE_STRIDE=$st python scripts/prof_score_10m.py 1000000 2>&1 | tail -1; done
