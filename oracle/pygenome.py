"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the reference's genome import, used to
check modle_b200_genome_import (modle_b200/csrc/genome.cpp). Never imported by the product.

Follows (paths relative to the reference checkout):
  chrom_sizes::Parser::parse_all             src/libmodle_io/chrom_sizes.cpp:24-66
  bed::Parser / bed::BED                     src/libmodle_io/bed.cpp:44-330,411-590
  BED_tree::find_overlaps over IITree        src/libmodle_io/bed_impl.hpp:166-180,
                                             src/interval_tree/interval_tree_impl.hpp:150-215
  Genome::Genome and helpers                 src/libmodle/internal/genome.cpp:255-488
  override_extrusion_barrier_occupancy       src/libmodle/cpu/simulation.cpp:51-60
Pinned by the reference's own in-source vectors (bed_parser_test.cpp:71-123: quote stripping, the
two malformed records, CRLF input) and by the example data set (tests/golden/genome_goldens.json).
"""
import re

STRANDS = {}
for k in ("+", "plus", "fwd", "Fwd", "forward", "Forward", "FWD", "FORWARD"):
    STRANDS[k] = "+"
for k in ("-", "minus", "rev", "Rev", "reverse", "Reverse", "REV", "REVERSE"):
    STRANDS[k] = "-"
for k in (".", "", "none", "None", "NONE", "unknown", "Unknown", "unk", "Unk", "UNK"):
    STRANDS[k] = "."


class ParseError(ValueError):
    pass


def strip_quote_pairs(s):  # src/common/utils_impl.hpp:204-214
    if len(s) >= 2 and s[0] in "'\"" and s[-1] in "'\"":
        return s[1:-1]
    return s


def _from_chars_u64(tok):
    """std::from_chars(u64) + the reference's "throw only if not fully consumed AND error"."""
    m = re.match(r"[0-9]+", tok)
    if not m:
        raise ParseError(f"Unable to convert field \"{tok}\" to a number")
    v = int(m.group(0))
    if v >= 1 << 64:  # result_out_of_range: ptr is past the digits
        if m.end() != len(tok):
            raise ParseError(f"Unable to convert field \"{tok}\" to a number")
        return 0  # field left untouched (value-initialised)
    return v


def _from_chars_f64(tok):
    m = re.match(r"-?(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][-+]?[0-9]+)?|-?(?:inf(?:inity)?|nan)",
                 tok, re.IGNORECASE)
    if not m:
        raise ParseError(f"Unable to convert field \"{tok}\" to a number")
    return float(m.group(0))


def parse_bed_record(line, dialect):
    toks = [t for t in re.split(r"[\t ]", line.rstrip(" \t\r\n\v\f")) if t]
    if len(toks) < 3:
        raise ParseError(f"expected at least 3 fields, got {len(toks)}")
    detected = len(toks) if len(toks) in (3, 4, 5, 6, 9, 12) else 254
    if detected < dialect:
        raise ParseError(f"Expected BED record with at least {dialect} fields, got {len(toks)}")
    rec = dict(chrom=strip_quote_pairs(toks[0]), start=_from_chars_u64(toks[1]),
               end=_from_chars_u64(toks[2]), name="", score=0.0, strand=".")
    if rec["start"] > rec["end"]:
        raise ParseError("chrom_start > chrom_end")
    if dialect == 3:
        return rec
    rec["name"] = strip_quote_pairs(toks[3])
    rec["score"] = _from_chars_f64(toks[4])
    if rec["score"] < 0 or rec["score"] > 1000:
        raise ParseError("score field should be between 0.0 and 1000.0")
    s = strip_quote_pairs(toks[5])
    if s not in STRANDS:
        raise ParseError(f"unrecognized strand \"{s}\"")
    rec["strand"] = STRANDS[s]
    return rec


def parse_bed_lines(lines, dialect):
    i = 0
    while i < len(lines):  # skip_header
        l = lines[i]
        if l == "" or l[0] == "#" or "track" in l or "browser" in l:
            i += 1
            continue
        break
    out, seen = [], {}
    for k in range(i, len(lines)):
        if lines[k] == "":
            continue
        r = parse_bed_record(lines[k], dialect)
        key = (r["chrom"], r["start"], r["end"])
        if key in seen:
            raise ParseError(f"Detected duplicate record at line {k + 1}")
        seen[key] = k + 1
        out.append(r)
    return out


def parse_chrom_sizes_lines(lines):
    out, seen = [], set()
    for l in lines:
        buff = l.rstrip(" \t\r\n\v\f")
        if not buff:
            continue
        toks = buff.split("\t")
        if len(toks) != 2:
            raise ParseError(f"expected exactly 2 fields, found {len(toks)}")
        name = strip_quote_pairs(toks[0])
        if name in seen:
            raise ParseError(f"found multiple records for chrom \"{name}\"")
        if toks[1] == "0":
            raise ParseError(f"chrom \"{name}\" has a length of 0bp")
        r = parse_bed_record(f"{name}\t0\t{toks[1]}", 3)
        seen.add(name)
        out.append((r["chrom"], r["end"]))
    if not out:
        raise ParseError("Unable to import any chromosome")
    return out


def find_overlaps(records, chrom, start, end):
    srt = sorted((r for r in records if r["chrom"] == chrom), key=lambda r: (r["start"], r["end"]))
    hits = [i for i, r in enumerate(srt) if r["start"] < end and start < r["end"]]
    return srt[hits[0]:hits[-1] + 1] if hits else []


def stp_active_from_occupancy(stp_inactive, occupancy):  # extrusion_barriers_impl.hpp:106-116
    if occupancy == 0:
        return 0.0
    to_active = 1.0 - stp_inactive
    to_inactive = (to_active - (occupancy * to_active)) / occupancy
    return max(0.0, min(1.0, 1.0 - to_inactive))


def _lines(path):
    with open(path, "rb") as f:
        data = f.read().decode()
    lines = data.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def import_genome(path_chrom_sizes, path_barriers, bin_size, pbb, puu, path_intervals="",
                  override_occupancy=False, interpret_name_field_as_puu=False):
    """Returns [dict(chrom_name, chrom_id, chrom_size, start, end, barriers=[(pos, stp_active,
    stp_inactive, blocking_direction)], bin_offset)] in the reference's processing order."""
    chroms = parse_chrom_sizes_lines(_lines(path_chrom_sizes))
    first_bin, b = [], 0
    for _, size in chroms:
        first_bin.append(b)
        b += (size + bin_size - 1) // bin_size
    intervals = []
    if not path_intervals:
        intervals = [(c, 0, size) for c, (_, size) in enumerate(chroms)]
    else:
        recs = parse_bed_lines(_lines(path_intervals), 3)
        for c, (name, size) in enumerate(chroms):
            intervals += [(c, r["start"], r["end"]) for r in find_overlaps(recs, name, 0, size)]
        if not intervals:
            raise ParseError("unable to import any interval")
    bars = parse_bed_lines(_lines(path_barriers), 6)
    out = []
    for c, start, end in intervals:
        lst = []
        for r in find_overlaps(bars, chroms[c][0], start, end):
            if r["strand"] == ".":
                continue
            if r["score"] < 0 or r["score"] > 1:
                raise ParseError("invalid score field: expected a score between 0 and 1")
            if interpret_name_field_as_puu:
                try:
                    v = _from_chars_f64(r["name"])
                except ParseError:
                    v = -1.0
                if v < 0 or v > 1:
                    raise ParseError("invalid name field")
            pos = (r["start"] + r["end"] + 1) // 2
            if r["score"] != 0.0 and not override_occupancy:
                stp_a, stp_i = stp_active_from_occupancy(puu, r["score"]), puu
            else:
                stp_a, stp_i = pbb, puu
            lst.append((pos, stp_a, stp_i, 1 if r["strand"] == "+" else 2))
        lst.sort(key=lambda t: t[0])  # stable
        out.append(dict(chrom_name=chroms[c][0], chrom_id=c, chrom_size=chroms[c][1], start=start,
                        end=end, barriers=lst, bin_offset=first_bin[c] + start // bin_size))
    return out
