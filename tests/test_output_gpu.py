"""`--fix snps,indels --changes --vcf` end to end: CUDA engine through the C ABI -> pb_out_* -> .fasta / .changes / .vcf
text, against the literal Python transliteration of the whole chain (oracle/pilon_oracle.py + pilon_output_oracle.py)."""
import random

import pytest

from oracle import pilon_oracle as po
from oracle import pilon_output_oracle as oo
from pilon_b200 import output as out
from pilon_b200.engine import Engine
from pilon_b200.packing import pack_records
from tests import helpers as H
from tests.test_output_cpu import compare, oracle_outputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 4, 9, 16])
def test_engine_results_feed_the_consumers(seed):
    contig, start, stop, reads = H.random_case(seed, contig_len=1200, n_reads=600)
    groups = H.split_batches(reads, random.Random(seed))
    ocfg = oo.OutConfig(vcfQE=seed % 2 == 0)
    gr = oracle_outputs(contig, start, stop, groups, None, ocfg)
    e = Engine(0)
    try:
        res, _ = e.run_region(contig, start, stop, [(pack_records(g), f) for g, f in groups])
    finally:
        e.close()
    compare(res, contig, start, stop, gr, ocfg)


def test_two_chunk_contig_fasta_changes_and_vcf():
    """One contig cut into two chunks (GenomeFile.scala:67-74): the changes file carries the running offset of the fixed
    sequence across chunks (:150-153), the FASTA is the concatenation of the fixed chunks 80 columns wide (:155-161)."""
    contig, _, _, reads = H.clean_case(7, n=24000, start=1, stop=24000, depth=12, n_sites=30)
    name = "scaffold_7"
    chunks = [(1, 12000), (12001, 24000)]
    ocfg = oo.OutConfig()
    e = Engine(0)
    outs, grs = [], []
    try:
        for a, b in chunks:
            mine = [r for r in reads if a - 10000 <= r.pos <= b + 10000]          # BamFile.scala:118-119 window
            gr = oo.GenomeRegionOut(contig, a, b, None, name, ocfg)
            gr.initializePileUps(oob_drop=True)
            gr.processBam(mine, "frags")
            gr.postProcess()
            gr.identifyAndFixIssues()
            grs.append(gr)
            res, _ = e.run_region(contig, a, b, [(pack_records(mine), True)], indels_cap=1 << 16, bytes_cap=1 << 20)
            outs.append(out.RegionOutput(res, contig, name, a, b))
        vcf = oo.Vcf(ocfg)
        want_changes, want_fasta = oo.writeContig(name, grs, vcf, True)
        got_changes, got_fasta, got_vcf = out.writeContig(name, outs, vcf=True, changes=True)
        assert got_changes == want_changes and len(want_changes) > 10
        assert got_fasta.splitlines() == want_fasta
        assert got_vcf.splitlines() == vcf.lines
        assert any(" ." in c for c in want_changes)                               # at least one indel moved the offset
    finally:
        for o in outs:
            o.close()
        e.close()


def test_fasta_and_bams_in_fasta_changes_vcf_out(tmp_path):
    """The whole replaced chain on files: genome.fasta + frags.bam + jumps.bam (written by the synthetic generator) ->
    native BAM ingest -> CUDA engine -> consumers -> pilon.fasta / .changes / .vcf, against the literal oracle chain fed with
    the same records.  This is what tools/run_real_pilon.sh diffs against the JVM where one exists."""
    import os
    import subprocess
    import sys
    from pilon_b200 import bamio, synth
    sys.path.insert(0, os.path.join(H.ROOT, "tools"))
    import make_pilon_inputs
    wl = synth.workload("C2", 0.0008)                      # 20 contigs of 20 kb: 60x frags + 10x jumps
    wl.contig_lens = wl.contig_lens[:3]
    names, paths = make_pilon_inputs.write_inputs(wl, str(tmp_path / "in"))
    subprocess.check_call([sys.executable, os.path.join(H.ROOT, "tools", "pilon_b200_run.py"), "--genome", str(tmp_path / "in" / "genome.fasta"),
                           "--frags", paths["frags"], "--jumps", paths["jumps"], "--changes", "--vcf", "--outdir", str(tmp_path / "out"),
                           "--chunksize", "12000"], stdout=subprocess.DEVNULL)
    want_fa, want_ch, vcf = [], [], oo.Vcf()
    for ci, name in enumerate(names):
        seq = wl.contig_bases(ci).tobytes()
        grs = []
        for a, b in synth.chunks_of(len(seq), 12000):
            gr = oo.GenomeRegionOut(seq, a, b, None, name)
            gr.initializePileUps(oob_drop=True)
            for kind in ("frags", "jumps"):
                bf = bamio.BamFile(paths[kind], kind)
                gr.processBam(H.unpack_batch(bf.process(name, a, b)), kind)
                bf.close()
            gr.postProcess()
            gr.identifyAndFixIssues()
            grs.append(gr)
        ch, fa = oo.writeContig(name, grs, vcf, True)
        want_ch += ch
        want_fa += fa
    got = lambda ext: open(str(tmp_path / "out" / ("pilon." + ext))).read().splitlines()
    assert got("fasta") == want_fa
    assert got("changes") == want_ch and len(want_ch) > 5
    assert [l for l in got("vcf") if not l.startswith("#")] == vcf.lines


@pytest.mark.parametrize("seed", [2, 5])
def test_fix_outputs_from_the_minimal_download(seed):
    """The e2e arm's default result set (bench.py FIXMIN_PLANES): flags + frag_coverage + the sparse call entries the
    engine compacts on the device.  FASTA bases, change list, statistics and log are those of the literal oracle chain."""
    contig, start, stop, reads = H.clean_case(seed, n=30000, start=1501, stop=28000, depth=12, n_sites=40)
    ocfg = oo.OutConfig()
    gr = oracle_outputs(contig, start, stop, [(reads, True)], None, ocfg)
    e = Engine(0)
    try:
        packed = [(pack_records(reads), True)]
        full, _ = e.run_region(contig, start, stop, packed, indels_cap=1 << 18, bytes_cap=1 << 22)
        mini, _ = e.run_region(contig, start, stop, packed, planes=["flags", "frag_coverage"], indels_cap=1 << 18, bytes_cap=1 << 22)
    finally:
        e.close()
    assert mini.c.call is None and mini.c.n_calls > 20
    H.assert_calls_match_planes(mini, full, "minimal download")
    ro = out.RegionOutput(mini, contig, "ctg|1", start, stop)
    try:
        st = ro.stats
        for k, v in gr.stats.items():
            key = {"nonN": "non_n", "insBases": "ins_bases", "delBases": "del_bases"}.get(k, k)
            assert st[key] == v, (k, st[key], v)
        assert ro.bases == bytes(gr.bases)
        assert [l for l in ro.log() if not l.startswith("Fix mismatch")] == [l for l in gr.loglines if not l.startswith("Fix mismatch")]
        assert ro.writeChanges() == gr.writeChanges()
        assert st["n_fixes"] > 20
    finally:
        ro.close()


def test_sparse_calls_beyond_the_capacity_are_counted_not_written():
    contig, start, stop, reads = H.clean_case(6, n=20000, start=1, stop=20000, depth=12, n_sites=40)
    e = Engine(0)
    try:
        packed = [(pack_records(reads), True)]
        full, _ = e.run_region(contig, start, stop, packed)
        small, _ = e.run_region(contig, start, stop, packed, calls_cap=5)
    finally:
        e.close()
    assert small.c.n_calls == full.c.n_calls > 5
    assert (small._calls == full.calls()[:5]).all()
    with pytest.raises(ValueError):
        small.calls()
