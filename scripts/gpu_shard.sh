#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_shard_step.py 8 lsh10m 2>&1 | tail -9
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/shard8_launches.csv python scripts/prof_shard_step.py 8 lsh10m > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/shard8_launches.csv')))
hdr=None; seq=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); seq.append((d['Kernel Name'][:70], float(d['Metric Value'])/1e3))
idx=[i for i,(n,t) in enumerate(seq) if 'merge_keys' in n]
a,b=idx[-3],idx[-2]
for n,t in seq[a+1:b+1]: print(f"{t:8.1f} us  {n}")
print('sum', sum(t for n,t in seq[a+1:b+1]))
PY
