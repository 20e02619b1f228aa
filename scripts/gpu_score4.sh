#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/score1m_launches.csv python scripts/prof_score_10m.py 1000000 > gpurun_out/score1m_ncu.log 2>&1
