#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, no GPU needed) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_<name>.txt ["free-text note"]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    lines = [f"# ncu summary of {rep}", note, ""]
    if len(raw) > 2:
        hdr, units = raw[0], raw[1]
        for row in raw[2:]:
            d = dict(zip(hdr, row))
            lines.append(f"## kernel: {d.get('Kernel Name', '?')}  grid {d.get('launch__grid_size', '?')} block {d.get('launch__block_size', '?')}")
            for k in WANT:
                if k in d and d[k] != "":
                    lines.append(f"  {k} = {d[k]} {units[hdr.index(k)]}")
            tensor_keys = [k for k in hdr if "tensor" in k and "pct" in k and d.get(k, "") not in ("", "0")]
            for k in tensor_keys[:8]:
                if k not in WANT:
                    lines.append(f"  {k} = {d[k]} {units[hdr.index(k)]}")
            lines.append("")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    hdr = None
    rows = []
    for r in src:
        if r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            rows.append(r)
    if hdr and rows:
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = defaultdict(int)
        for r in rows:
            for h in stalls:
                tot[h] += int(r[ix[h]] or 0)
        total = sum(int(r[ix["# Samples"]] or 0) for r in rows)
        lines.append(f"## warp-state samples (all launches in the report): {total}")
        for h, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
            lines.append(f"  {h}: {v} ({100.0 * v / max(total, 1):.1f} %)")
        lines.append("")
        lines.append("## top instructions by samples")
        for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]] or 0))[:15]:
            lines.append(f"  {r[ix['# Samples']]:>7} samples  {r[ix['Instructions Executed']]:>10} exec  {r[1].strip()[:70]}")
        sass = " ".join(r[1] for r in rows)
        lines.append("")
        lines.append("## Blackwell-native evidence in SASS: " + ", ".join(
            f"{m}={'yes' if m in sass else 'no'}" for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS")))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
