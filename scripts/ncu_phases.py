#!/usr/bin/env python3
"""Per-phase breakdown of an `ncu --page source --csv --print-source cuda,sass` export.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
    python scripts/ncu_phases.py cs.csv [path/to/sim_core.hpp]

Everything in the per-cell simulator is inlined into one kernel, so the SASS rows are sorted by
address and each is attributed to the member function of sim_core.hpp whose line range holds the
row's source line; rows that come from other files (cta.hpp primitives, CUDA intrinsics) inherit
the function of the closest preceding sim_core.hpp row in address order. Prints, per function:
warp instructions executed, stall samples (share of the kernel's time) and the top stall reasons.
"""
import csv
import re
import sys
from collections import defaultdict


def function_ranges(path):
    starts = []
    pat = re.compile(r"^\s*(?:template <[^>]*>\s*)?MB_FN(?:_NOINLINE)?\s+(?:static\s+)?[\w:<>\*& ]+?\s+(\w+)\(")
    with open(path) as f:
        for no, line in enumerate(f, 1):
            m = pat.match(line)
            if m:
                starts.append((no, m.group(1)))
    return starts


def func_of(starts, line):
    name = "(file scope)"
    for no, fn in starts:
        if no <= line:
            name = fn
        else:
            break
    return name


def main():
    path = sys.argv[1]
    core = sys.argv[2] if len(sys.argv) > 2 else "modle_b200/csrc/sim_core.hpp"
    starts = function_ranges(core)
    rows = []  # (address, file, line, samples, inst, stalls)
    cur_file, cur_line, hdr, stall_cols = "?", None, None, {}
    with open(path, newline="", errors="replace") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "File Path":
                cur_file = row[1].split("/")[-1]
                continue
            if row[0] == "Line No":
                hdr = row
                i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
                stall_cols = {i: h for i, h in enumerate(hdr)
                              if h.startswith("stall_") and "Not Issued" not in h and "(" not in h}
                continue
            if hdr is None or len(row) < 8:
                continue
            if row[0] != "":
                cur_line = int(row[0])
                continue
            if not row[2].startswith("0x"):
                continue
            st = {h: float(row[i]) for i, h in stall_cols.items() if i < len(row) and row[i] not in ("", "0")}
            rows.append((int(row[2], 16), cur_file, cur_line, float(row[i_s] or 0),
                         float(row[i_i] or 0), st))
    rows.sort()
    by_fn_s, by_fn_i = defaultdict(float), defaultdict(float)
    by_fn_st = defaultdict(lambda: defaultdict(float))
    cur_fn = "(prologue)"
    for _, fl, ln, s, i, st in rows:
        if fl == "sim_core.hpp":
            cur_fn = func_of(starts, ln)
        elif fl == "kernels.cu":
            cur_fn = "(kernel body)"
        by_fn_s[cur_fn] += s
        by_fn_i[cur_fn] += i
        for h, v in st.items():
            by_fn_st[cur_fn][h] += v
    ts, ti = sum(by_fn_s.values()) or 1, sum(by_fn_i.values()) or 1
    print(f"total warp instructions {ti:.3e}, stall samples {ts:.0f}")
    print(f"{'function':42s} {'samples%':>8s} {'inst%':>7s}  top stalls")
    for fn in sorted(by_fn_s, key=lambda k: -by_fn_s[k]):
        top = sorted(by_fn_st[fn].items(), key=lambda kv: -kv[1])[:4]
        tops = " ".join(f"{h[6:]}={100 * v / max(by_fn_s[fn], 1):.0f}%" for h, v in top)
        print(f"{fn:42s} {100 * by_fn_s[fn] / ts:8.1f} {100 * by_fn_i[fn] / ti:7.1f}  {tops}")


if __name__ == "__main__":
    main()
