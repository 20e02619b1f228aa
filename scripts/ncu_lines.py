#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of an Nsight Compute report (needs -lineinfo + --import-source on).

    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
fname, hdr, data = None, None, []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "File Path":
        fname = r[1]; hdr = None
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)          # first "Source" column = the CUDA line
        data.append((fname, d))
if not data:
    print("no source rows"); sys.exit(0)
keys = list(data[0][1].keys())
tot = sum(int(d.get("# Samples") or 0) for _, d in data)
print(f"total samples {tot}")
for f, d in sorted(data, key=lambda x: -int(x[1].get("# Samples") or 0))[:top]:
    n = int(d.get("# Samples") or 0)
    st = sorted(((k[6:], int(d[k] or 0)) for k in keys if k.startswith("stall_") and "Not" not in k), key=lambda x: -x[1])[:3]
    print(f"{n:8d} {100.0*n/max(tot,1):5.1f}%  {f.split('/')[-1]}:{d['Line No']:>4}  ex={d.get('Instructions Executed','?'):>10}  {st}  | {d['Source'].strip()[:80]}")
