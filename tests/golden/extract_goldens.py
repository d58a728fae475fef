#!/usr/bin/env python3
"""Extracts the reference's own hot-path golden vectors into tests/golden/reference_goldens.json.

Source: /root/reference/test/units/simulation_cpu/{simulation_simple_unit_test.cpp,
simulation_complex_unit_test.cpp} (Catch2 cases with hand-written LEF / barrier layouts and the
expected collision words, moves and ranks). Only the test DATA is lifted; the calls each case
makes are recorded as a list of step names so tests/ can replay them against the oracle and the
kernel emulation. Run here (the reference tree is not available on the GPU box):

    python tests/golden/extract_goldens.py
"""
import json
import os
import re
import sys

REF = "/root/reference/test/units/simulation_cpu"
FILES = ["simulation_simple_unit_test.cpp", "simulation_complex_unit_test.cpp"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

EVENTS = {"CHROM_BOUNDARY": 0x18, "LEF_BAR": 0x14, "LEF_LEF_PRIMARY": 0x12,
          "LEF_LEF_SECONDARY": 0x11}

# reference test shim -> oracle step names (src/libmodle/cpu/include/modle/simulation.hpp:413-567)
CALLS = {
    "test_adjust_and_clamp_moves": ["adjust", "clamp"],
    "test_adjust_moves": ["adjust"],
    "test_detect_units_at_interval_boundaries": ["boundaries"],
    "test_detect_lef_bar_collisions": ["lef_bar"],
    "test_correct_moves_for_lef_bar_collisions": ["correct_lef_bar"],
    "test_detect_primary_lef_lef_collisions": ["primary"],
    "test_adjust_moves_for_primary_lef_lef_collisions": ["correct_primary"],
    "test_process_lef_lef_collisions": ["primary", "correct_primary", "secondary"],
    "test_process_collisions": ["boundaries", "lef_bar", "primary", "correct_lef_bar",
                                "correct_primary", "secondary"],
    "test_fix_secondary_lef_lef_collisions": ["fix_secondary"],
}


def ints(body):
    return [int(x) for x in re.findall(r"-?\d+", body)]


def array(block, name):
    m = re.search(r"\b" + re.escape(name) + r"\s*\{([^;]*?)\}\s*;", block, re.S)
    return None if m is None else m.group(1)


def collisions(block, name):
    body = array(block, name)
    if body is None:
        return None
    out = []
    for m in re.finditer(r"CollisionT\{\s*(?:(\d+)\s*,\s*(\w+))?\s*\}", body):
        if m.group(1) is None:
            out.append(0)
        else:
            out.append((EVENTS[m.group(2)] << 24) | int(m.group(1)))
    return out


def parse_case(name, block, line):
    case = {"name": name, "source_line": line}
    m = re.search(r"init_config\((\d+),\s*(\d+)", block)
    case["prob_bypass"] = 0.0
    case["pblock_major"], case["pblock_minor"] = 1.0, 0.0
    for field, key in (("probability_of_extrusion_unit_bypass", "prob_bypass"),
                       ("lef_bar_major_collision_pblock", "pblock_major"),
                       ("lef_bar_minor_collision_pblock", "pblock_minor")):
        mm = re.search(r"c\." + field + r"\s*=\s*([\d.]+)", block)
        if mm:
            case[key] = float(mm.group(1))
    m = re.search(r'init_interval\("(\w+)",\s*(\d+)(?:,\s*(\d+))?(?:,\s*(\d+))?\)', block)
    if m:
        size = int(m.group(2))
        case["start"] = int(m.group(3)) if m.group(3) else 0
        case["end"] = min(int(m.group(4)), size) if m.group(4) else size
    m = re.search(r"random::PRNG\((\d+)ULL\)", block)
    case["rng_seed"] = int(m.group(1)) if m else 10556020843759504871  # DEFAULT_PRNG
    # one or more LEF arrays (ranking tests have lefs1 / lefs2)
    lef_sets = []
    for m in re.finditer(r"std::(?:array<Lef,\s*\w+>|vector<Lef>)\s+(\w+)\s*\{(.*?)\}\s*;", block,
                         re.S):
        lefs = [[int(a), int(b), int(e)] for a, b, e in
                re.findall(r"construct_lef\((\d+),\s*(\d+),\s*(\d+)\)", m.group(2))]
        lef_sets.append((m.group(1), lefs))
    case["lef_sets"] = {k: v for k, v in lef_sets}
    case["released"] = [int(i) for i in re.findall(r"lefs\[(\d+)\]\.release\(\)", block)]
    bars = re.findall(r"ExtrusionBarrier\{(\d+),\s*[\d.]+,\s*[\d.]+,\s*'([+-])'\}", block)
    case["barriers"] = [[int(p), 1 if s == "+" else 2] for p, s in bars]  # blocking direction
    case["barriers_active"] = 0 if "State::INACTIVE" in block else 1
    for key in ("rev_ranks", "fwd_ranks", "rev_moves", "fwd_moves", "rev_moves_expected",
                "fwd_moves_expected", "rev_moves_adjusted", "fwd_moves_adjusted",
                "rev_ranks_expected1", "fwd_ranks_expected1", "rev_ranks_expected2",
                "fwd_ranks_expected2"):
        body = array(block, key)
        if body is not None and body.strip():
            case[key] = ints(body)
    for key in ("rev_collisions_expected", "fwd_collisions_expected"):
        c = collisions(block, key)
        if c is not None:
            case[key] = c
    steps = []
    for m in re.finditer(r"\b(test_\w+)\s*\(", block):
        if m.group(1) in CALLS:
            steps += CALLS[m.group(1)]
    case["steps"] = steps
    case["rank_calls"] = len(re.findall(r"test_rank_lefs\(", block))
    final = {}
    for m in re.finditer(r"CHECK\((rev|fwd)_ranks\[(\d+)\]\s*==\s*(\d+)\)", block):
        final.setdefault(m.group(1) + "_ranks_final", {})[m.group(2)] = int(m.group(3))
    case.update(final)
    return case


def main():
    cases = []
    for f in FILES:
        path = os.path.join(REF, f)
        if not os.path.exists(path):
            sys.exit(f"reference test file not found: {path}")
        text = open(path).read()
        starts = [(m.start(), m.group(1)) for m in re.finditer(r'TEST_CASE\("([^"]+)"', text)]
        for i, (pos, name) in enumerate(starts):
            end = starts[i + 1][0] if i + 1 < len(starts) else len(text)
            block = text[pos:end]
            line = text.count("\n", 0, pos) + 1
            case = parse_case(name, block, line)
            case["source_file"] = "test/units/simulation_cpu/" + f
            wanted = (case["steps"] and "lefs" in case["lef_sets"]) or case["rank_calls"]
            if wanted:
                cases.append(case)
    with open(OUT, "w") as fh:
        json.dump({"generated_by": "tests/golden/extract_goldens.py",
                   "reference": "paulsengroup/modle v1.1.0, test/units/simulation_cpu",
                   "cases": cases}, fh, indent=1)
    print(f"wrote {len(cases)} cases to {OUT}")
    for c in cases:
        print(" ", c["name"], c["steps"] or "rank", "n=", len(c["lef_sets"].get("lefs", [])))


if __name__ == "__main__":
    main()
