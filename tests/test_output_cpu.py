"""The per-locus consumers and the fix application (libpilonb200.so's pb_out_*, host code) against the literal Python
transliteration of GenomeRegion.identifyAndFixIssues / fixFixList / fixIssues / writeChanges / writeVcf, Vcf.writeRecord
and GenomeFile's output rules (oracle/pilon_output_oracle.py).  The per-locus results fed to pb_out_* come from the C
oracle here (same pb_region_result layout as the engine fills): no GPU is needed; tests/test_output_gpu.py repeats the
comparison with the CUDA engine's results."""
import random

import numpy as np
import pytest

from oracle import pilon_oracle as po
from oracle import pilon_output_oracle as oo
from pilon_b200 import output as out
from pilon_b200.packing import pack_records
from tests import helpers as H


def oracle_outputs(contig, start, stop, groups, cfg=None, ocfg=None, name="ctg|1"):
    gr = oo.GenomeRegionOut(contig, start, stop, cfg, name, ocfg)
    gr.initializePileUps(oob_drop=True)
    for reads, frag in groups:
        gr.processBam(reads, "frags" if frag else "jumps")
    gr.postProcess()
    gr.identifyAndFixIssues()
    return gr


def compare(res, contig, start, stop, gr, ocfg, name="ctg|1"):
    ro = out.RegionOutput(res, contig, name, start, stop, out.OutputConfig(ocfg.fixSnps, ocfg.fixIndels, ocfg.iupac, ocfg.diploid,
                                                                           ocfg.vcfQE, ocfg.longread))
    try:
        st = ro.stats
        for k, v in gr.stats.items():
            key = {"nonN": "non_n", "insBases": "ins_bases", "delBases": "del_bases"}.get(k, k)
            assert st[key] == v, (k, st[key], v)
        assert ro.bases == bytes(gr.bases)
        assert np.array_equal(ro.copyNumber, np.array(gr.copyNumber, np.int16))
        assert [l for l in ro.log() if not l.startswith("Fix mismatch")] == [l for l in gr.loglines if not l.startswith("Fix mismatch")]
        assert st["fix_mismatches"] == sum(1 for l in gr.loglines if l.startswith("Fix mismatch"))
        new = oo.pilonName(name)
        assert out.pilonName(name) == new
        assert ro.writeChanges(new, 17) == gr.writeChanges(new, 17)
        assert ro.writeChanges() == gr.writeChanges()
        v = oo.Vcf(ocfg)
        gr.writeVcf(v)
        got = ro.writeVcf(3).splitlines()
        assert len(got) == len(v.lines)
        for a, b in zip(got, v.lines):
            assert a == b
        return ro.stats
    finally:
        ro.close()


@pytest.mark.parametrize("seed", range(12))
def test_outputs_match_the_oracle_on_random_cases(seed):
    contig, start, stop, reads = H.random_case(seed, contig_len=900, n_reads=400)
    rng = random.Random(seed)
    groups = H.split_batches(reads, rng)
    cfg = po.Config(fixAmb=seed % 3 == 0, iupac=seed % 4 == 1)
    ocfg = oo.OutConfig(fixSnps=seed % 5 != 4, fixIndels=seed % 7 != 6, iupac=cfg.iupac, diploid=seed % 6 == 2, vcfQE=seed % 2 == 1)
    gr = oracle_outputs(contig, start, stop, groups, cfg, ocfg)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(g), f) for g, f in groups], cfg)
    compare(res, contig, start, stop, gr, ocfg)


def test_outputs_on_a_clean_region_with_every_kind_of_change():
    contig, start, stop, reads = H.clean_case(3, n=30000, start=2001, stop=26000, depth=12, n_sites=40)
    ocfg = oo.OutConfig()
    gr = oracle_outputs(contig, start, stop, [(reads, True)], None, ocfg)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)], indels_cap=1 << 18, bytes_cap=1 << 22)
    st = compare(res, contig, start, stop, gr, ocfg)
    assert st["snps"] > 0 and st["amb"] > 0 and st["ins"] > 0 and st["dels"] > 0 and st["n_fixes"] > 20


def test_duplication_events_and_copy_number():
    """A 30 kb stretch at three times the coverage of the rest: copy number 3, one <DUP> record in the VCF."""
    rng = random.Random(8)
    n = 80000
    contig = H.random_contig(rng, n, lower_frac=0.0, n_runs=0)
    L = 100
    reads = []
    for lo, hi, depth in ((1, n - L, 10), (25000, 55000 - L, 20)):
        for _ in range(depth * (hi - lo) // L):
            p = rng.randint(lo, hi)
            reads.append(po.Read(pos=p, cigar=[("M", L)], bases=contig[p - 1:p - 1 + L].upper(), quals=bytes([30]) * L, mapq=60, paired=False))
    reads.sort(key=lambda r: r.pos)
    ocfg = oo.OutConfig()
    gr = oracle_outputs(contig, 1, n, [(reads, True)], None, ocfg, name="dup")
    assert len(gr.duplicationEvents()) == 1
    res, _ = H.run_c_oracle(contig, 1, n, [(pack_records(reads), True)])
    st = compare(res, contig, 1, n, gr, ocfg, name="dup")
    assert st["n_dups"] == 1


def test_fix_list_overlaps_keep_the_larger_fix():
    # GenomeRegion.scala:557-595 through the oracle alone (hand-made lists), then the same rule in the library through a
    # region where a deletion call overlaps a SNP inside the deleted stretch is impossible by construction -- so the
    # hand-made case pins the oracle, and the random cases above pin the library against the oracle
    gr = oo.GenomeRegionOut(b"ACGT" * 50, 1, 200)
    fixes = [(50, "ACG", ""), (51, "C", "T"), (10, "", "GG"), (10, "A", "C"), (120, "T", "A")]
    assert gr.fixFixList(fixes) == [(10, "", "GG"), (50, "ACG", ""), (120, "T", "A")]


def test_genome_level_helpers():
    assert out.pilonName("chr1") == "chr1_pilon" and out.pilonName("gi|123|") == "gi|123|pilon" and out.pilonName("gi|123|x") == "gi|123|x|pilon"
    for n in (0, 1, 79, 80, 81, 400):
        seq = bytes(random.Random(n).choice(b"ACGT") for _ in range(n))
        assert out.fastaElement("h", seq).splitlines() == oo.fastaElement("h", seq.decode())
    v = oo.Vcf(oo.OutConfig(vcfQE=True))
    v.writeHeader("20260101", "Pilon version 1.24", "--genome g.fa --frags f.bam", "file:/tmp/g.fa", [("a", 10), ("b", 20)])
    assert out.vcfHeader("20260101", "Pilon version 1.24", "--genome g.fa --frags f.bam", "file:/tmp/g.fa", [("a", 10), ("b", 20)],
                         out.OutputConfig(vcfQE=True)).splitlines() == v.lines
    assert out.coverageSummary([("frags", 1000), ("jumps", 300), ("frags", 500)], 10) == oo.coverageSummary([("frags", 1000), ("jumps", 300), ("frags", 500)], 10)


def test_java_format_of_allele_fractions():
    # "%.2f".format: HALF_UP on exact ties (1/8 -> 0.13), where C's printf would say 0.12
    assert oo.java_fmt2(0.125) == "0.13" and oo.java_fmt2(0.375) == "0.38" and oo.java_fmt2(float(np.float32(1) / np.float32(3))) == "0.33"


def test_wiggle_tracks_are_per_locus_maps_of_the_planes():
    contig, start, stop, reads = H.clean_case(5, n=6000, start=501, stop=5500, depth=10, n_sites=8)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)])
    ro = out.RegionOutput(res, contig, "c", start, stop)
    try:
        fl = res["flags"]
        for track, want in (("Changes", (fl & 2) // 2), ("Unconfirmed", 1 - (fl & 1)), ("Coverage", res["coverage_arr"]),
                            ("Bad Coverage", res["bad_pair"]), ("Physical Coverage", res["phys_cov"]), ("Weighted MQ", res["weighted_mq"]),
                            ("Copy Number", ro.copyNumber.astype(np.int64) - 1)):
            lines = ro.wig(track).splitlines()
            assert lines[0] == "fixedStep chrom=c start=%d step=1" % start
            assert [int(x) for x in lines[1:]] == [int(x) for x in want]
    finally:
        ro.close()


def minimal_fix_result(res):
    """What a `--fix snps,indels --changes` caller downloads: flags, frag_coverage, the sparse call entries and the
    indel evidence -- the same pb_region_result with every other plane NULL."""
    from pilon_b200.packing import ResultBuffers
    m = ResultBuffers(res.size, ["flags", "frag_coverage"], indels_cap=res.indels_cap, indel_bytes_cap=res.indel_bytes_cap,
                      calls_cap=max(1, int(res.c.n_calls)))
    m["flags"][:] = res["flags"]
    m["frag_coverage"][:] = res["frag_coverage"]
    m._calls[:int(res.c.n_calls)] = res.calls()
    m.c.n_calls = res.c.n_calls
    m._indels = res._indels if res.indels_cap else None
    if res.indels_cap:
        import ctypes as C
        m.c.indels = C.addressof(res._indels)
    if res.indel_bytes_cap:
        m.indel_bytes = res.indel_bytes
        m.c.indel_bytes = res.indel_bytes.ctypes.data
    for f in ("size", "base_count", "coverage", "aligned_bases", "read_count", "min_depth", "unknown_ops", "dropped_oob", "n_indels",
              "n_indel_bytes", "n_batches"):
        setattr(m.c, f, getattr(res.c, f))
    return m


@pytest.mark.parametrize("seed", range(6))
def test_fix_outputs_from_the_minimal_result_equal_those_from_every_plane(seed):
    """flags + frag_coverage + sparse call entries are all the fix path reads (GenomeRegion.scala:275-283, 307-380): the
    FASTA bases, the change list, the statistics and the log are identical to what the full planes give."""
    if seed < 4:
        contig, start, stop, reads = H.random_case(seed, contig_len=900, n_reads=400)
        groups = H.split_batches(reads, random.Random(seed))
    else:
        contig, start, stop, reads = H.clean_case(seed, n=20000, start=1001, stop=18000, depth=12, n_sites=30)
        groups = [(reads, True)]
    cfg = po.Config(iupac=seed % 2 == 1)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(g), f) for g, f in groups], cfg, indels_cap=1 << 18, bytes_cap=1 << 22)
    H.assert_calls_match_planes(res, res)
    ocfg = out.OutputConfig(iupac=cfg.iupac)
    full = out.RegionOutput(res, contig, "c|1", start, stop, ocfg)
    m = minimal_fix_result(res)
    mini = out.RegionOutput(m, contig, "c|1", start, stop, ocfg)
    try:
        assert mini.stats == full.stats and mini.bases == full.bases and mini.log() == full.log()
        assert np.array_equal(mini.copyNumber, full.copyNumber)
        assert mini.writeChanges() == full.writeChanges()
        if seed >= 4:
            assert full.stats["n_fixes"] > 10
    finally:
        full.close(); mini.close()


def test_truncated_sparse_calls_are_refused():
    contig, start, stop, reads = H.clean_case(4, n=20000, start=1001, stop=18000, depth=12, n_sites=30)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)], indels_cap=1 << 18, bytes_cap=1 << 22)
    m = minimal_fix_result(res)
    assert m.c.n_calls > 2
    m.c.calls_cap = int(m.c.n_calls) - 1
    with pytest.raises(Exception):
        out.RegionOutput(m, contig, "c|1", start, stop)


def test_facade_reads_call_records_from_the_sparse_list():
    from pilon_b200.engine import _call_record
    contig, start, stop, reads = H.clean_case(4, n=20000, start=1001, stop=18000, depth=12, n_sites=30)
    res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)], indels_cap=1 << 18, bytes_cap=1 << 22)
    m = minimal_fix_result(res)
    ent = res.calls()
    assert len(ent) > 10
    for k in (0, len(ent) // 2, len(ent) - 1):
        i = int(ent["locus_index"][k])
        assert _call_record(m, i) == _call_record(res, i) == int(res["call"][i])
    quiet = int(np.flatnonzero((res["flags"] & 6) == 0)[0])
    with pytest.raises(KeyError):
        _call_record(m, quiet)
