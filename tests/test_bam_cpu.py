"""BAM out, BAM in: the writer's files are parsed by an independent pure-Python reader (gzip + struct,
straight from the SAM/BAM specification), and the native reader (BGZF inflate, BAI seek, record decode, validateRead, packer)
gives back exactly the batches that were written.  Host code only: no GPU."""
import gzip
import os
import random
import struct

import numpy as np
import pytest

from oracle import pilon_oracle as po
from pilon_b200 import _capi as capi
from pilon_b200 import bamio
from pilon_b200.packing import pack_records
from tests import helpers as H

FIELDS = ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off", "quals", "bases2", "exc_idx", "exc_base", "exc_qual")


def spec_parse(path):
    """(header text, [(name, len)], [record dict]) with nothing but gzip and struct (SAMv1 4.2)."""
    d = gzip.open(path, "rb").read()                   # BGZF = concatenated gzip members
    assert d[:4] == b"BAM\x01"
    l_text, = struct.unpack_from("<i", d, 4)
    text = d[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", d, p); p += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", d, p); p += 4
        name = d[p:p + l_name - 1].decode(); p += l_name
        l_ref, = struct.unpack_from("<i", d, p); p += 4
        refs.append((name, l_ref))
    recs = []
    while p < len(d):
        bs, = struct.unpack_from("<i", d, p); p += 4
        refID, pos, l_rn, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", d, p)
        q = p + 32
        name = d[q:q + l_rn - 1].decode(); q += l_rn
        cig = struct.unpack_from("<%dI" % n_cig, d, q); q += 4 * n_cig
        sq = d[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(sq[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = d[q:q + l_seq]; q += l_seq
        assert q == p + bs
        recs.append(dict(refID=refID, pos=pos, mapq=mapq, bin=bin_, flag=flag, cigar=list(cig), seq=seq, qual=qual, tlen=tlen, nref=nref, name=name))
        p += bs
    return text, refs, recs


def reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    d = tmp_path_factory.mktemp("bam")
    rng = random.Random(12)
    contigs = [("ctgA", H.random_contig(rng, 60000, n_runs=3)), ("ctgB", H.random_contig(rng, 9000))]
    batches = []
    all_reads = []
    for ci, (_, c) in enumerate(contigs):
        reads = [H.random_read(rng, c, 1, len(c) - 200, max_len=120) for _ in range(4000 if ci == 0 else 500)]
        reads.sort(key=lambda r: r.pos)
        all_reads.append(reads)
        batches.append((ci, pack_records(reads)))
    path = str(d / "x.bam")
    bamio.write_bam(path, [(n, len(c)) for n, c in contigs], batches, program_line="@PG\tID:pilon_b200")
    bamio.write_fasta(str(d / "x.fasta"), contigs)
    return path, contigs, batches, all_reads


def test_written_bam_parses_with_an_independent_reader(case):
    path, contigs, batches, all_reads = case
    text, refs, recs = spec_parse(path)
    assert text.startswith("@HD\tVN:1.6\tSO:coordinate\n") and "@SQ\tSN:ctgA\tLN:60000" in text
    assert refs == [(n, len(c)) for n, c in contigs]
    flat = [(ci, r) for ci, reads in enumerate(all_reads) for r in reads]
    assert len(recs) == len(flat)
    for rec, (ci, r) in zip(recs, flat):
        assert rec["refID"] == ci and rec["pos"] == r.pos - 1 and rec["mapq"] == r.mapq and rec["tlen"] == r.tlen
        assert rec["seq"] == r.bases.decode() and rec["cigar"] == [(l << 4) | "MIDNSHP=X".index(op) for op, l in r.cigar]
        assert rec["qual"] == (r.quals if len(r.quals) else b"\xff" * len(r.bases))
        reflen = sum(l for op, l in r.cigar if op in "MDN=X")
        assert rec["bin"] == reg2bin(r.pos - 1, r.pos - 1 + (1 if (r.unmapped or reflen == 0) else reflen))
        assert bool(rec["flag"] & 1) == r.paired and bool(rec["flag"] & 2) == r.proper and bool(rec["flag"] & 4) == r.unmapped
        assert (rec["nref"] == rec["refID"]) == r.mate_same_ref
    # end-of-file marker block (SAMv1 4.1.2)
    assert open(path, "rb").read()[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    fa = open(path[:-4] + ".fasta").read().split(">")[1:]
    assert [x.split("\n", 1)[0] for x in fa] == [n for n, _ in contigs]
    assert fa[0].split("\n", 1)[1].replace("\n", "").encode() == contigs[0][1]


def test_native_reader_round_trips_whole_contigs(case):
    path, contigs, batches, _ = case
    bf = bamio.BamFile(path)
    try:
        assert bf.refs == [(n, len(c)) for n, c in contigs]
        for (ci, want), (name, c) in zip(batches, contigs):
            got, rej = bf.query(name, 0, len(c))
            assert rej == 0
            for f in FIELDS:
                assert np.array_equal(getattr(got, f), getattr(want, f)), f
    finally:
        bf.close()


def test_region_queries_return_exactly_the_overlapping_records(case):
    path, contigs, batches, all_reads = case
    bf = bamio.BamFile(path)
    try:
        name, c = contigs[0]
        for a, b in ((1, 100), (20001, 30000), (16384, 16385), (49000, 60000), (32769, 32769)):
            got, _ = bf.query(name, a, b)
            want = []
            for r in all_reads[0]:
                reflen = sum(l for op, l in r.cigar if op in "MDN=X")
                end = r.pos if (r.unmapped or reflen == 0) else r.pos + reflen - 1      # htsjdk: no alignment end -> one base long
                if r.pos <= b and end >= a:
                    want.append(r)
            wb = pack_records(want)
            for f in FIELDS:
                assert np.array_equal(getattr(got, f), getattr(wb, f)), (a, b, f)
        # BamFile.process applies the +-10 kb of BamFile.scala:118-119
        got = bf.process(name, 30001, 40000)
        assert got.pos.min() <= 20050 and got.pos.max() >= 49900 and got.pos.max() <= 50000
    finally:
        bf.close()


def test_validate_read_filters_qcfail_duplicate_and_secondary(tmp_path):
    rng = random.Random(5)
    c = H.random_contig(rng, 5000)
    reads = sorted([H.random_read(rng, c, 1, 4800, max_len=80) for _ in range(600)], key=lambda r: r.pos)
    rb = pack_records(reads)
    xf = np.zeros(len(reads), np.uint16)
    for i in range(len(reads)):
        xf[i] = rng.choice([0, 0, 0, 0x100, 0x200, 0x400, 0x800, 0x600])
    p = str(tmp_path / "f.bam")
    bamio.write_bam(p, [("c", len(c))], [(0, rb)], extra_flags=[xf])
    for nonpf, dups in ((False, False), (True, False), (False, True), (True, True)):
        bf = bamio.BamFile(p, nonPf=nonpf, duplicates=dups)
        got, rej = bf.query("c", 0, len(c))
        bf.close()
        keep = [r for r, f in zip(reads, xf) if po.validateRead(bool(f & 0x200), bool(f & 0x400), bool(f & 0x100), nonpf, dups)]
        assert rej == len(reads) - len(keep) and got.n_reads == len(keep)
        assert np.array_equal(got.pos, pack_records(keep).pos)           # supplementary (0x800) alignments are kept (BamFile.scala:101-105)


def test_reader_works_without_an_index_and_rejects_garbage(case, tmp_path):
    path, contigs, batches, _ = case
    bf = bamio.BamFile(path, index=str(tmp_path / "missing.bai"))
    got, _ = bf.query("ctgB", 0, 9000)
    bf.close()
    assert np.array_equal(got.pos, batches[1][1].pos)
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"not a bam file at all")
    with pytest.raises(capi.EngineError):
        bamio.BamFile(str(bad))
