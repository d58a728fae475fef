#!/usr/bin/env python3
"""Heuristic lint of modle_b200/csrc/sim_core.hpp for the one class of CTA race the thread-order
replay of the emulation cannot see: a scalar of CellShared (`S.<field>`) that every thread reads
in uniform code -- i.e. outside a region -- while the region right after it (no CTA barrier in
between) lets the leader overwrite it; or that is read right after such a region without a
barrier; or that is read by all threads inside the very region the leader writes it in.
(`S.rng_pos` in next_barrier_states / extrude_and_release was exactly that: DESIGN.md section 3.)

The scan is textual and per member function: events are barriers (`cta.sync()`, the block-wide
scans / reductions, which contain barriers), region begin / end (`MB_REGION(cta, tid) {` and its
matching brace) and reads / writes of `S.<field>`. `S.fault` is exempt (any writer wins, by
design), as are the cycle counters.

    python scripts/lint_shared_state.py [path/to/sim_core.hpp]     exit status 1 on findings
"""
import os
import re
import sys

EXEMPT = {"fault", "phase_cycles", "scratch", "move_bound_hit"}
SYNC = re.compile(r"cta\.(sync|exscan_\w+|reduce_\w+)\s*\(|__syncthreads\s*\(")
REGION = re.compile(r"MB_REGION\s*\(\s*cta\s*,\s*tid\s*\)\s*\{")
ACCESS = re.compile(r"\bS\.([a-z_0-9]+)((?:\[[^\]]*\])?)\s*(\+\+|--|(?:[+\-|&^]|<<|>>)?=(?!=))?")
PRE_INC = re.compile(r"(\+\+|--)\s*S\.([a-z_0-9]+)")
FUNC = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?MB_FN(?:_NOINLINE)?\s+[^;(]*?\b(\w+)\s*\([^;]*$")


def strip_comments(line):
    line = re.sub(r"//.*", "", line)
    # accesses through MB_SHARED_* are shared by design (cta.hpp) and checked by the SPMD
    # emulation under ThreadSanitizer (tests/test_emulation_races.py), not by this scan
    return re.sub(r"MB_SHARED_\w+\s*\(\s*&\s*S\.[a-z_0-9]+", "MB_SHARED(", line)


def lint(path):
    findings = []
    lines = [strip_comments(l) for l in open(path).read().split("\n")]
    func = "?"
    depth = 0
    region_depth = None      # brace depth at which the current region closes
    leader_depths = []       # brace depths of enclosing `if (cta.leader(tid))` blocks
    uniform_reads = {}       # field -> line of a read in uniform code since the last barrier
    region_writes = {}       # field -> line, writes inside the current region
    region_reads = {}        # field -> line, non-leader reads inside the current region
    pending_writes = {}      # field -> line, region writes not yet followed by a barrier
    for no, line in enumerate(lines, 1):
        m = FUNC.match(line)
        if m and region_depth is None:
            func = m.group(1)
            uniform_reads, pending_writes = {}, {}
        # barriers first (a scan call both reads its argument and synchronises)
        sync_here = SYNC.search(line) is not None and region_depth is None
        starts_region = REGION.search(line) is not None
        in_region = region_depth is not None or starts_region
        leader_here = "cta.leader(tid)" in line
        pre = {mm.group(2) for mm in PRE_INC.finditer(line)}
        for mm in ACCESS.finditer(line):
            field, op = mm.group(1), mm.group(3)
            if field in EXEMPT:
                continue
            is_write = op is not None or field in pre
            is_read = op is None or op not in ("=",) or field in pre  # compound ops also read
            in_leader = bool(leader_depths) or leader_here
            if in_region:
                if is_write:
                    region_writes.setdefault(field, no)
                    if field in uniform_reads:
                        findings.append((func, field, uniform_reads[field], no,
                                         "read in uniform code, then written by the next region "
                                         "without a barrier in between"))
                if is_read and not in_leader:
                    region_reads.setdefault(field, no)
            else:
                if is_read and not is_write:
                    if field in pending_writes and not sync_here:
                        findings.append((func, field, pending_writes[field], no,
                                         "written in a region, then read without a barrier"))
                    uniform_reads.setdefault(field, no)
        # brace bookkeeping
        for ch in line:
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
                if leader_depths and depth < leader_depths[-1]:
                    leader_depths.pop()
                if region_depth is not None and depth < region_depth:
                    for f, wl in region_writes.items():
                        if f in region_reads:
                            findings.append((func, f, region_reads[f], wl,
                                             "read by all threads and written inside one region"))
                    pending_writes.update(region_writes)
                    region_writes, region_reads = {}, {}
                    region_depth = None
                    uniform_reads = {}  # a new region follows a barrier in the bulk-synchronous style
        if starts_region:
            region_depth = depth  # depth after the opening brace of the region
        if leader_here and "{" in line[line.index("cta.leader(tid)"):]:
            leader_depths.append(depth)
        if sync_here:
            uniform_reads, pending_writes = {}, {}
    return findings


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "modle_b200", "csrc",
                                                              "sim_core.hpp")
    findings = lint(path)
    for func, field, l1, l2, what in findings:
        print(f"{os.path.basename(path)}:{l1}/{l2}: {func}: S.{field}: {what}")
    print(f"{len(findings)} finding(s)")
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
