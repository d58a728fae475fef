"""Genome import: modle_b200_genome_import (C++ host layer) against the Python
restatement of the reference's parsers (oracle/pygenome.py), the reference's in-source vectors
(test/units/libmodle_io/bed_parser_test.cpp:71-123) and goldens of its example data set."""
import json
import lzma
import os

import numpy as np
import pytest

from modle_b200 import abi, host
from oracle import pygenome

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/examples/data"

CHROM_SIZES = "chrA\t1000000\n\"chrB\"\t500000\r\nchrC\t250000\n\n"
BARRIERS = (
    "# a comment\n"
    "track name=barriers\n"
    "\n"
    "chrA\t1000\t1021\tb0\t0.9\t+\n"
    "chrA\t5000 5019\tb1\t0.75\t-\n"            # blanks separate fields as well as tabs
    "chrA\t7000\t7020\tb2\t0\tplus\n"           # score 0 -> default stp; strand alias
    "chrA\t9000\t9020\tb3\t0.8\t.\n"            # no strand -> skipped
    "chrA\t400\t420\tb4\t1\t\"-\"\textra\n"     # 7 fields pass as "none"; quoted strand
    "chrB\t499990\t500010\t0.5\t0.6\tREV\r\n"   # CRLF; pos 500000 == chrom size (kept by the ref)
    "\n"
    "chrB\t100\t100\t0.25\t0.61\tFwd\n"         # zero-length record
    "chrZ\t10\t20\tbz\t0.9\t+\n"                # chromosome not in chrom.sizes -> unused
    "chrC\t12abc\t30\tbq\t0.7\t-\n"             # from_chars stops at 'a': start = 12
)
INTERVALS = "chrA\t0\t6000\nchrA\t6500\t1000000\nchrC\t0\t250000\n"


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text.encode())
    return str(p)


def params(**kw):
    p = host.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    host.transform_params(p, False, False, "extrusion_barrier_occupancy" in kw)
    return p


def assert_same(cpp, py):
    assert len(cpp) == len(py)
    for a, b in zip(cpp, py):
        for k in ("chrom_name", "chrom_id", "chrom_size", "start", "end", "bin_offset"):
            assert a[k] == b[k], k
        got = [(int(x["pos"]), float(x["stp_active"]), float(x["stp_inactive"]),
                int(x["blocking_direction"])) for x in a["barriers"]]
        assert got == b["barriers"]


@pytest.mark.parametrize("with_intervals", [False, True])
@pytest.mark.parametrize("override", [False, True])
def test_import_matches_python_restatement(product_lib, tmp_path, with_intervals, override):
    cs = write(tmp_path, "g.chrom.sizes", CHROM_SIZES)
    bed = write(tmp_path, "barriers.bed", BARRIERS)
    ivs = write(tmp_path, "intervals.bed", INTERVALS) if with_intervals else ""
    p = params(extrusion_barrier_occupancy=0.9) if override else params()
    assert bool(p.override_extrusion_barrier_occupancy) == override
    cpp = host.import_genome(cs, bed, p, ivs)
    py = pygenome.import_genome(cs, bed, int(p.bin_size), p.barrier_occupied_stp,
                                p.barrier_not_occupied_stp, ivs, override_occupancy=override)
    assert_same(cpp, py)
    if not with_intervals:
        assert [g["chrom_name"] for g in cpp] == ["chrA", "chrB", "chrC"]
        assert [len(g["barriers"]) for g in cpp] == [4, 2, 1]
        a = cpp[0]["barriers"]
        # pos = (start + end + 1) / 2; '+' motifs block REV-moving units; sorted by position
        assert a["pos"].tolist() == [410, 1011, 5010, 7010]
        assert a["blocking_direction"].tolist() == [abi.DIR_FWD, abi.DIR_REV, abi.DIR_FWD,
                                                    abi.DIR_REV]
        assert cpp[2]["barriers"]["pos"].tolist() == [21]
        assert cpp[1]["bin_offset"] == 200 and cpp[2]["bin_offset"] == 300
        if not override:
            # score 0 -> the default self-transition probabilities (genome.cpp:255-271)
            assert a["stp_active"][3] == p.barrier_occupied_stp
            assert a["stp_active"][1] == host.lib().modle_b200_stp_active_from_occupancy(
                p.barrier_not_occupied_stp, 0.9)
        else:
            assert set(a["stp_active"].tolist()) == {p.barrier_occupied_stp}
    else:
        assert [(g["chrom_name"], g["start"], g["end"]) for g in cpp] == [
            ("chrA", 0, 6000), ("chrA", 6500, 1000000), ("chrC", 0, 250000)]
        assert cpp[1]["bin_offset"] == 1


def test_imported_barriers_equal_the_record_path(product_lib, tmp_path):
    """The importer and host.barriers_from_records (used by the workloads) agree."""
    cs = write(tmp_path, "g.chrom.sizes", "chrA\t1000000\n")
    bed = write(tmp_path, "b.bed", "chrA\t1000\t1021\tx\t0.9\t+\nchrA\t300\t320\ty\t0.65\t-\n")
    p = params()
    g = host.import_genome(cs, bed, p)
    want = host.barriers_from_records([(1011, "+", 0.9), (310, "-", 0.65)], p)
    assert np.array_equal(g[0]["barriers"], want)


BAD_BEDS = {
    "duplicate": "chrA\t10\t20\ta\t0.5\t+\nchrA\t10\t20\tb\t0.7\t-\n",
    "strand": "chrA\t10\t20\ta\t0.5\tsideways\n",
    "score_above_1": "chrA\t10\t20\ta\t1.5\t+\n",
    "score_above_1000": "chrA\t10\t20\ta\t1001\t+\n",
    "quoted_start": "chrA\t\"0\"\t1\ta\t0.5\t+\n",          # bed_parser_test.cpp:100
    "quoted_score": "chrA\t0\t1\t.\t\"0.0\"\t+\n",          # bed_parser_test.cpp:101
    "start_after_end": "chrA\t30\t20\ta\t0.5\t+\n",
    "too_few_fields": "chrA\t10\t20\ta\t0.5\n",
    "comment_after_header": "chrA\t10\t20\ta\t0.5\t+\n# late comment\n",
}


@pytest.mark.parametrize("name", sorted(BAD_BEDS))
def test_malformed_barrier_files_are_rejected(product_lib, tmp_path, name):
    cs = write(tmp_path, "g.chrom.sizes", "chrA\t1000000\n")
    bed = write(tmp_path, "b.bed", BAD_BEDS[name])
    p = params()
    with pytest.raises(pygenome.ParseError):
        pygenome.import_genome(cs, bed, 5000, 0.0, 0.7)
    with pytest.raises(host.ModleB200Error) as e:
        host.import_genome(cs, bed, p)
    assert e.value.code == abi.ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("text", ["chrA\t100\textra\n", "chrA\t100\nchrA\t200\n", "chrA\t0\n", "",
                                  "chrA 100\n"])
def test_malformed_chrom_sizes_are_rejected(product_lib, tmp_path, text):
    cs = write(tmp_path, "g.chrom.sizes", text)
    bed = write(tmp_path, "b.bed", "chrA\t10\t20\ta\t0.5\t+\n")
    with pytest.raises(pygenome.ParseError):
        pygenome.import_genome(cs, bed, 5000, 0.0, 0.7)
    with pytest.raises(host.ModleB200Error):
        host.import_genome(cs, bed, params())


def test_reference_in_source_vectors():
    # bed_parser_test.cpp:71-98 ("BED: strip quotes", valid section)
    r = pygenome.parse_bed_record("chr1\t0\t10\t\"name\"\t0.0\t\"+\"\t0\t1\t\"0,0,0\"", 6)
    assert (r["chrom"], r["start"], r["end"], r["name"], r["score"], r["strand"]) == \
        ("chr1", 0, 10, "name", 0.0, "+")
    assert pygenome.parse_bed_record("\"chr1\t0\t1", 3)["chrom"] == "\"chr1"
    # :105-123 ("BED Parser CRLF")
    recs = pygenome.parse_bed_lines([f"chr{i}\t0\t1\r" for i in range(3)], 3)
    assert [(r["chrom"], r["start"], r["end"]) for r in recs] == [(f"chr{i}", 0, 1) for i in range(3)]


def test_name_field_as_puu_is_validated_but_unused(product_lib, tmp_path):
    cs = write(tmp_path, "g.chrom.sizes", "chrA\t1000000\n")
    ok = write(tmp_path, "ok.bed", "chrA\t10\t20\t0.3\t0.5\t+\n")
    bad = write(tmp_path, "bad.bed", "chrA\t10\t20\tctcf\t0.5\t+\n")
    p = params()
    a = host.import_genome(cs, ok, p, interpret_name_field_as_puu=True)
    b = host.import_genome(cs, ok, p)
    assert np.array_equal(a[0]["barriers"], b[0]["barriers"])  # genome.cpp:456-457
    with pytest.raises(host.ModleB200Error):
        host.import_genome(cs, bad, p, interpret_name_field_as_puu=True)
    host.import_genome(cs, bad, p)


@pytest.mark.skipif(not os.path.exists(REF_DATA), reason="reference example data not present")
def test_example_data_goldens(product_lib, tmp_path):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "genome_goldens.json")))
    bed = tmp_path / "barriers.bed"
    bed.write_bytes(lzma.open(os.path.join(REF_DATA, "hg38_extrusion_barriers.bed.xz")).read())
    ivs = write(tmp_path, "intervals.bed", gold["sub_intervals_bed"])
    p = params()
    for key, path_iv in (("whole_genome", ""), ("sub_intervals", ivs)):
        g = host.import_genome(os.path.join(REF_DATA, "hg38.chrom.sizes"), str(bed), p, path_iv)
        assert len(g) == len(gold[key])
        for iv, s in zip(g, gold[key]):
            pos = iv["barriers"]["pos"].astype(np.uint64)
            assert (iv["chrom_name"], iv["chrom_size"], iv["start"], iv["end"],
                    iv["bin_offset"]) == (s["chrom"], s["size"], s["start"], s["end"],
                                          s["bin_offset"])
            assert len(pos) == s["n"]
            assert int((iv["barriers"]["blocking_direction"] == abi.DIR_REV).sum()) == \
                s["n_block_rev"]
            assert int(pos.sum()) == s["pos_sum"]
            assert int(np.bitwise_xor.reduce(pos)) == s["pos_xor"]
            assert (int(pos[0]), int(pos[-1])) == (s["first"], s["last"])
            assert abs(float(iv["barriers"]["stp_active"].sum()) - s["stp_active_sum"]) < 1e-6
    assert sum(s["n"] for s in gold["whole_genome"]) == 38815
