"""BASELINE-sized parity: whole C1 / C5 regions, a 10 Mb C2 chunk with frags + jumps and an interior C3 chunk,
each compared plane by plane, scalar by scalar and indel entry by indel entry with the C oracle, under the
engine's OWN kernel selection (no PB_PILEUP override): the scatter kernel with >= 256 tiles and the gather
kernel on the deep amplicon are the code paths bench.py times.  Plus the literal Python oracle on a region wide
enough for the scatter kernel, and directed cases for corners the random generators do not reach."""
import os
import random

import numpy as np
import pytest

from oracle import pilon_oracle as po
from pilon_b200 import _capi as capi
from pilon_b200 import synth
from pilon_b200.engine import Engine, EngineConfig
from pilon_b200.packing import pack_records
from tests import helpers as H

pytestmark = pytest.mark.gpu


class kernel_choice:
    """PB_PILEUP is read by pb_create: engines made inside the block use the forced kernel (None = the engine's choice)."""

    def __init__(self, which):
        self.val = {None: None, "auto": None, "gather": "5", "scatter": "7", "cluster": "8"}[which]

    def __enter__(self):
        self.old = os.environ.get("PB_PILEUP")
        if self.val is None:
            os.environ.pop("PB_PILEUP", None)
        else:
            os.environ["PB_PILEUP"] = self.val

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("PB_PILEUP", None)
        else:
            os.environ["PB_PILEUP"] = self.old


def _region_vs_c_oracle(wl, ci, a, b, which=None):
    contig = wl.contig_bases(ci).tobytes()
    sbs = wl.region_batches(ci, a, b)
    batches = [(sb.as_read_batch(), sb.frag) for sb in sbs]
    n_ops = sum(int(rb.cigar.shape[0]) for rb, _ in batches)
    with kernel_choice(which):
        e = Engine(0)
    try:
        res, ins = e.run_region(contig, a, b, batches, indels_cap=max(1 << 16, n_ops // 8), bytes_cap=1 << 24)
    finally:
        e.close()
    ref, ins_ref = H.run_c_oracle(contig, a, b, batches, indels_cap=max(1 << 16, n_ops // 8), bytes_cap=1 << 24)
    H.assert_results_equal(res, ref, "%s contig %d %d-%d" % (wl.name, ci, a, b))
    for x, y in zip(ins, ins_ref):
        assert np.array_equal(x, y)
    assert res.c.aligned_bases == sum(sb.aligned_bases for sb in sbs)
    return res


def test_c1_whole_genome_5mb_100x():
    wl = synth.workload("C1")
    (ci, a, b), = wl.regions()
    assert b - a + 1 == 5_000_000                        # 2442 scatter tiles
    res = _region_vs_c_oracle(wl, ci, a, b)
    fl = res["flags"]
    snps = int((((fl & capi.PB_FL_CHANGED) != 0) & (((fl >> capi.PB_FL_KIND_SHIFT) & 3) == capi.PB_KIND_SNP)).sum())
    assert 0.8 * 5000 <= snps <= 1.2 * 5000


def test_c2_one_10mb_chunk_frags_and_jumps():
    wl = synth.workload("C2")
    ci, a, b = wl.regions()[0]
    assert b - a + 1 == 10_000_000 and len(wl.libraries) == 2
    res = _region_vs_c_oracle(wl, ci, a, b)
    # jumps do not count toward fragCoverage (GenomeRegion.scala:291,296): it must stay below the depth on average
    assert res["frag_coverage"].astype(np.int64).sum() < res["coverage_arr"].astype(np.int64).sum()


def test_c3_interior_chunk_with_halo_on_both_sides():
    wl = synth.workload("C3")
    regs = wl.regions()
    assert len(regs) == 7
    ci, a, b = regs[3]
    assert a > 1 and b < wl.contig_lens[ci] and b - a + 1 == 9_142_858
    _region_vs_c_oracle(wl, ci, a, b)


@pytest.mark.parametrize("kernel", [None, "cluster"])
def test_c5_whole_amplicon_200kb_5000x(kernel):
    wl = synth.workload("C5")
    (ci, a, b), = wl.regions()
    assert b - a + 1 == 200_000
    res = _region_vs_c_oracle(wl, ci, a, b, kernel)
    assert res.c.coverage > 4000


def test_python_oracle_on_a_region_wide_enough_for_the_scatter_kernel():
    """600 k loci = 293 tiles: under the engine's own choice this is k_pileup7.  The literal Python transliteration
    shares no BaseCall / hetIndelCall code with the device path (the C oracle's restatement is close to it)."""
    contig, start, stop, reads = H.clean_case(41, n=620_000, start=10_001, stop=610_000, depth=8, n_sites=400)
    rng = random.Random(41)
    groups = [([r for i, r in enumerate(reads) if i % 4 != 0], True), ([r for i, r in enumerate(reads) if i % 4 == 0], False)]
    with kernel_choice(None):
        e = Engine(0)
    try:
        packed = [(pack_records(g), f) for g, f in groups]
        res, ins = e.run_region(contig, start, stop, packed, indels_cap=1 << 18, bytes_cap=1 << 22)
    finally:
        e.close()
    py = H.run_py_oracle(contig, start, stop, groups)
    H.assert_matches_py(res, ins, py, "engine (auto) vs python oracle, 600 k loci")
    fl = res["flags"]
    assert (fl & capi.PB_FL_DELETED).any() and (fl & capi.PB_FL_AMBIGUOUS).any()
    del rng


# ---------------------------------------------------------------------------------------------
# directed cases
# ---------------------------------------------------------------------------------------------
def _eng_cfg(cfg):
    return EngineConfig(cfg.minQual, cfg.minMq, cfg.flank, cfg.defaultQual, cfg.minMinDepth, cfg.minDepth,
                        cfg.oldIndel, cfg.iupac, cfg.fixAmb)


@pytest.mark.parametrize("kernel", ["gather", "scatter", "cluster"])
def test_deletion_shift_readds_bases_left_of_the_reads_pos(kernel):
    """PileUpRegion.scala:167-178: the shift walks read offsets backwards across earlier CIGAR elements; after a long
    insertion inside a homopolymer the re-added bases start LEFT of the read's alignment start.  No read of the batch
    has a soft clip, so nothing else raises the batch's backward reach: the window / tile that ends just left of pos
    finds the segment only through the reach k_indel records for it."""
    ref = bytearray(b"ACGT" * 100)
    for l in range(40, 140):
        ref[l - 1] = ord("A")                            # homopolymer over loci 40..139
    ref = bytes(ref)
    pos = 97                                             # region index 96 = first locus of a 32-locus window
    bases = b"AA" + b"A" * 20 + b"AA" + ref[pos + 4:pos + 4 + 30]      # 2M 20I 2M 1D 30M
    rd = po.Read(pos=pos, cigar=[("M", 2), ("I", 20), ("M", 2), ("D", 1), ("M", 30)], bases=bases, quals=bytes([30]) * len(bases), mapq=60)
    other = po.Read(pos=300, cigar=[("M", 40)], bases=ref[299:339], quals=bytes([25]) * 40, mapq=60)
    cfg = po.Config(flank=0)
    groups = [([rd, other], True)]
    py = H.run_py_oracle(ref, 1, 400, groups, cfg)
    cnt = py["base_count4"].sum(axis=1)
    assert cnt[pos - 1 - 5] >= 1 and cnt[:pos - 1].sum() >= 19      # the literal transliteration re-adds left of pos
    with kernel_choice(kernel):
        e = Engine(0, _eng_cfg(cfg))
    try:
        packed = [(pack_records(g), f) for g, f in groups]
        res, ins = e.run_region(ref, 1, 400, packed)
        H.assert_matches_py(res, ins, py, "D-shift re-add left of pos (%s)" % kernel)
        ref_c, _ = H.run_c_oracle(ref, 1, 400, packed, cfg)
        H.assert_results_equal(res, ref_c, "D-shift re-add left of pos, C oracle (%s)" % kernel)
    finally:
        e.close()


@pytest.mark.parametrize("kernel", ["gather", "scatter", "cluster"])
def test_more_than_4064_descriptors_per_tile_with_mixed_mapq(kernel):
    """The scatter kernel folds its 12-bit tile counters into the output planes every 4064 descriptors; reads with
    five different mapping qualities make the folds carry Bq / C terms as well."""
    rng = random.Random(77)
    n = 3000
    contig = H.random_contig(rng, n, n_runs=2)
    reads = []
    for i in range(14000):
        pos = rng.randint(1, n - 120)
        L = rng.randint(60, 110)
        b = bytearray(contig[pos - 1:pos - 1 + L].upper())
        if rng.random() < 0.2:
            b[rng.randrange(L)] = rng.choice(b"ACGTN")
        reads.append(po.Read(pos=pos, cigar=[("M", L)], bases=bytes(b), quals=bytes(rng.randint(2, 60) for _ in range(L)),
                             mapq=rng.choice([0, 13, 37, 60, 255]), paired=rng.random() < 0.5, proper=rng.random() < 0.95,
                             tlen=rng.choice([250, -250])))
    reads.sort(key=lambda r: r.pos)
    groups = [([r for i, r in enumerate(reads) if i % 3], True), ([r for i, r in enumerate(reads) if i % 3 == 0], False)]
    packed = [(pack_records(g), f) for g, f in groups]
    with kernel_choice(kernel):
        e = Engine(0)
    try:
        res, ins = e.run_region(contig, 1, n, packed)
    finally:
        e.close()
    ref, ins_ref = H.run_c_oracle(contig, 1, n, packed)
    H.assert_results_equal(res, ref, "deep tile, mixed MAPQ (%s)" % kernel)
    assert int(res["base_count4"].sum(axis=1).max()) > 300


@pytest.mark.parametrize("kernel", ["gather", "scatter", "cluster"])
def test_int32_wrap_of_mqsum(kernel):
    """mqSum is a JVM Int (PileUp.scala:33): 8.5 M bases of MAPQ 255 at one locus push it past 2^31 and it wraps;
    BaseSum (qualSum) is 64-bit and does not.  score becomes 0 through `mqSum > 0` (PileUp.scala:148)."""
    n_reads = 8_500_000
    contig = (b"ACGT" * 64)[:200]
    pos = 100
    packed = pack_records([])
    z = np.zeros
    packed.pos = np.full(n_reads, pos, np.int32)
    packed.tlen = z(n_reads, np.int32)
    packed.read_len = np.full(n_reads, 2, np.int32)
    packed.mapq = np.full(n_reads, 255, np.uint8)
    packed.flags = np.full(n_reads, capi.PB_F_HAS_QUALS, np.uint8)
    packed.cigar_off = np.arange(n_reads + 1, dtype=np.uint32)
    packed.cigar = np.full(n_reads, (2 << 4) | 0, np.uint32)
    packed.seq_off = (np.arange(n_reads, dtype=np.uint32) * 4)
    q = z(n_reads * 4, np.uint8)
    q[0::4] = 40
    q[1::4] = 41
    packed.quals = q
    c0 = b"ACGT".index(contig[pos - 1]); c1 = b"ACGT".index(contig[pos])
    packed.bases2 = np.full(n_reads, c0 | (c1 << 2), np.uint8)
    cfg = po.Config(flank=0)
    with kernel_choice(kernel):
        e = Engine(0, _eng_cfg(cfg))
    try:
        res, _ = e.run_region(contig, 1, 200, [(packed, True)])
    finally:
        e.close()
    ref, _ = H.run_c_oracle(contig, 1, 200, [(packed, True)], cfg)
    H.assert_results_equal(res, ref, "mqSum wrap (%s)" % kernel)
    i = pos - 1
    assert int(res["mq_sum"][i]) == ((n_reads * 256 + 2 ** 31) % 2 ** 32) - 2 ** 31 < 0
    assert int(res["qual_sum4"][i][c0]) == n_reads * 40 * 256
    assert capi_score(res["call"][i]) == 0


def capi_score(c):
    return int(c) >> 16


@pytest.mark.parametrize("kernel", ["gather", "scatter", "cluster"])
def test_deletion_spill_chain_of_overlapping_candidates(kernel):
    """GenomeRegion.scala:259-264 is sequential: a deletion called INSIDE a deleted span makes no call, so the span it would
    have deleted stays callable.  Three homozygous deletions in a chain: A = 100 (+10) is accepted, B = 105 (+8) lies inside A
    and is rejected, C = 112 (+5) lies inside B's span only and must be accepted.  k_spill resolves chains in parallel from
    their heads; a candidate far to the right is a chain of its own."""
    rng = random.Random(5)
    ref = bytearray(rng.choice(b"ACGT") for _ in range(400))
    for start, dlen in ((100, 10), (105, 8), (112, 5), (300, 6)):
        while ref[start - 2] == ref[start + dlen - 2]:      # base before != last deleted base: nothing shifts left (PileUpRegion.scala:167-178)
            ref[start - 2] = rng.choice(b"ACGT")
    ref = bytes(ref)

    def del_read(pos, left, dlen, right):
        bases = ref[pos - 1:pos - 1 + left] + ref[pos - 1 + left + dlen:pos - 1 + left + dlen + right]
        return po.Read(pos=pos, cigar=[("M", left), ("D", dlen), ("M", right)], bases=bases, quals=bytes([35]) * len(bases), mapq=60)

    reads = []
    reads += [del_read(70, 30, 10, 2) for _ in range(20)]       # A: loci 100..109 deleted, reads end at 111
    reads += [del_read(103, 2, 8, 2) for _ in range(20)]        # B: 105..112, reads cover 103,104 and 113,114
    reads += [del_read(111, 1, 5, 40) for _ in range(20)]       # C: 112..116
    reads += [del_read(270, 30, 6, 30) for _ in range(20)]      # far away: its own chain
    reads.sort(key=lambda r: r.pos)
    cfg = po.Config(flank=0)
    groups = [(reads, True)]
    py = H.run_py_oracle(ref, 1, 400, groups, cfg)
    fl = py["flags"]
    is_del = lambda l: bool(fl[l - 1] & capi.PB_FL_CHANGED) and ((fl[l - 1] >> capi.PB_FL_KIND_SHIFT) & 3) == capi.PB_KIND_DEL   # noqa: E731
    deleted = lambda l: bool(fl[l - 1] & capi.PB_FL_DELETED)                                                                   # noqa: E731
    assert is_del(100) and all(deleted(l) for l in range(101, 110))
    assert not is_del(105) and not deleted(110) and not deleted(111)
    assert is_del(112) and all(deleted(l) for l in range(113, 117)) and not deleted(117)
    assert is_del(300) and deleted(301)
    with kernel_choice(kernel):
        e = Engine(0, _eng_cfg(cfg))
    try:
        packed = [(pack_records(g), f) for g, f in groups]
        res, ins = e.run_region(ref, 1, 400, packed)
        H.assert_matches_py(res, ins, py, "deletion spill chain (%s)" % kernel)
        ref_c, _ = H.run_c_oracle(ref, 1, 400, packed, cfg)
        H.assert_results_equal(res, ref_c, "deletion spill chain, C oracle (%s)" % kernel)
    finally:
        e.close()
