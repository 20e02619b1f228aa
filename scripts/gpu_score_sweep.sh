#!/bin/bash
# scoring + top-k over item-table sizes: live time and the per-kernel launch times (ncu, cold) per size
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for N in 312500 625000 1250000 2500000 5000000 10000000; do
  python scripts/prof_score_10m.py $N 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/score_sweep_$N.csv python scripts/prof_score_10m.py $N > /dev/null 2>&1
  python - $N <<'PY'
import csv, sys
lines=[l for l in open(f"gpurun_out/score_sweep_{sys.argv[1]}.csv") if not l.startswith("==")]
seq=[]
for row in csv.DictReader(lines):
    if row.get("Metric Name")!="gpu__time_duration.sum": continue
    v=float(row["Metric Value"].replace(",","")); u=row["Metric Unit"]
    v = v/1000 if u=="ns" else (v*1000 if u=="ms" else v)
    seq.append((row["Kernel Name"][:48], v))
# one call = the launches after the last pairs_to_csr... print the last call's launches
last=[]
for n,v in reversed(seq):
    last.append((n,v))
    if "tc_score_topk_kernel<1>" in n or len(last)>9: break
print("   ", " | ".join(f"{n.split('(')[0].split('::')[-1]} {v:.1f}" for n,v in reversed(last)))
PY
done
