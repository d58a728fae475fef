#!/usr/bin/env python3
"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
    python scripts/ncu_lines.py cs.csv [top_n]

Prints, per (file, line): warp-level instructions executed, stall samples and the share of each.
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    inst = defaultdict(float)
    samp = defaultdict(float)
    text = {}
    stall_cols = {}
    stalls = defaultdict(lambda: defaultdict(float))
    cur_file = "?"
    cur_line = None
    hdr = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "File Name":
                cur_file = row[1].split("/")[-1]
                continue
            if row[0] == "Line No":
                hdr = row
                stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
                continue
            if hdr is None or len(row) < 8:
                continue
            if row[0] != "":
                cur_line = (cur_file, int(row[0]))
                text[cur_line] = row[1].strip()
            if row[2] == "" or cur_line is None:
                continue
            try:
                i_samples = hdr.index("# Samples")
                i_inst = hdr.index("Instructions Executed")
                samp[cur_line] += float(row[i_samples] or 0)
                inst[cur_line] += float(row[i_inst] or 0)
                for i, h in stall_cols.items():
                    v = row[i]
                    if v:
                        stalls[cur_line][h] += float(v)
            except (ValueError, IndexError):
                pass
    tot_i = sum(inst.values()) or 1
    tot_s = sum(samp.values()) or 1
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    print("--- by stall samples")
    for k in sorted(samp, key=lambda k: -samp[k])[:top]:
        st = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{h[6:]}={v:.0f}" for h, v in st)
        print(f"{100 * samp[k] / tot_s:5.1f}% smp {100 * inst[k] / tot_i:5.1f}% ins  {k[0]}:{k[1]:<5d} {text.get(k, '')[:70]}  [{sts}]")
    print("--- by instructions")
    for k in sorted(inst, key=lambda k: -inst[k])[:top]:
        print(f"{100 * inst[k] / tot_i:5.1f}% ins {100 * samp[k] / tot_s:5.1f}% smp  {k[0]}:{k[1]:<5d} {text.get(k, '')[:80]}")


if __name__ == "__main__":
    main()
