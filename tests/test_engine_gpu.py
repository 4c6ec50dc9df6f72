"""Parity tests proper: the CUDA engine, called through the C ABI, against the oracles.
Bit-exact on every counter, flag and call record (integer path; no floating-point tolerance)."""
import random

import numpy as np
import pytest

from oracle import pilon_oracle as po
from pilon_b200 import _capi as capi
from pilon_b200.engine import Engine, EngineConfig, PileUpRegion
from pilon_b200.packing import pack_records
from tests import helpers as H

pytestmark = pytest.mark.gpu


def eng_cfg(cfg: po.Config) -> EngineConfig:
    return EngineConfig(cfg.minQual, cfg.minMq, cfg.flank, cfg.defaultQual, cfg.minMinDepth, cfg.minDepth,
                        cfg.oldIndel, cfg.iupac, cfg.fixAmb)


@pytest.fixture(scope="module", autouse=True)
def _kernel(pileup_kernel):
    return pileup_kernel


@pytest.fixture(scope="module")
def engine(pileup_kernel):
    e = Engine(0)
    yield e
    e.close()


def run_both(engine, contig, start, stop, groups, cfg=None):
    packed = [(pack_records(g), f) for g, f in groups]
    res, ins = engine.run_region(contig, start, stop, packed)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, packed, cfg)
    H.assert_results_equal(res, ref, "engine vs C oracle")
    for a, b in zip(ins, ins_ref):
        assert np.array_equal(a, b)
    return res, ins


@pytest.mark.parametrize("seed", range(60))
def test_random_cases_match_both_oracles(seed):
    contig, start, stop, reads = H.random_case(seed)
    rng = random.Random(seed * 7 + 1)
    groups = H.split_batches(reads, rng)
    cfg = po.Config()
    if seed % 5 == 1:
        cfg = po.Config(minQual=7, minMq=2, flank=rng.choice([0, 3, 10]), defaultQual=rng.choice([3, 10, 15]))
    if seed % 7 == 3:
        cfg.oldIndel = True
    if seed % 11 == 4:
        cfg.minDepth = 3.0
    if seed % 13 == 5:
        cfg.fixAmb = True
    e = Engine(0, eng_cfg(cfg))
    try:
        res, ins = run_both(e, contig, start, stop, groups, cfg)
        py = H.run_py_oracle(contig, start, stop, groups, cfg)
        H.assert_matches_py(res, ins, py, "engine vs python oracle, seed %d" % seed)
    finally:
        e.close()


def test_kat1_through_c_abi(engine):
    ref = (b"ACGT" * 30)[:100]
    rd = po.Read(pos=11, cigar=[("M", 30)], bases=ref[10:40], quals=bytes([30]) * 30, mapq=60)
    res, ins = engine.run_region(ref, 1, 100, [(pack_records([rd]), True)])
    assert res.c.base_count == 10 and res.c.read_count == 1 and int(ins[0][0]) == 29
    cnt, qs = res["base_count4"], res["qual_sum4"]
    for locus in range(1, 101):
        i = locus - 1
        if 21 <= locus <= 30:
            bi = b"ACGT".index(ref[i])
            assert cnt[i][bi] == 1 and cnt[i].sum() == 1 and qs[i][bi] == 1830
            assert res["mq_sum"][i] == 61 and res["q_sum"][i] == 30
        else:
            assert cnt[i].sum() == 0
        assert res["phys_cov"][i] == (1 if 11 <= locus <= 39 else 0)
        assert res["insert_size"][i] == (29 if 11 <= locus <= 39 else 0)


def test_kat7_deletion_shift_through_c_abi(engine):
    ref = bytearray(b"C" * 300)
    for l in range(101, 105):
        ref[l - 1] = ord("A")
    ref[99], ref[104] = ord("G"), ord("T")
    ref = bytes(ref)
    pos = 101 - 47
    rb = ref[pos - 1:pos - 1 + 50] + ref[104:104 + 50]
    rd = po.Read(pos=pos, cigar=[("M", 50), ("D", 1), ("M", 50)], bases=rb, quals=bytes([30]) * 100, mapq=60)
    res, _ = engine.run_region(ref, 1, 300, [(pack_records([rd]), True)])
    assert [int(res["base_count4"][l - 1][0]) for l in (101, 102, 103, 104)] == [1, 2, 2, 1]
    assert res["deletions"][100] == 1 and res.c.base_count == 83
    ind = res.indels()
    assert len(ind) == 1 and ind[0]["locus_index"] == 100 and ind[0]["kind"] == 2 and ind[0]["list_len"] == 1


def test_empty_region_and_zero_reads(engine):
    contig = H.random_contig(random.Random(3), 500)
    empty = pack_records([])
    res, _ = engine.run_region(contig, 10, 400, [(empty, True)])
    ref, _ = H.run_c_oracle(contig, 10, 400, [(empty, True)])
    H.assert_results_equal(res, ref)
    assert res.c.read_count == 0 and not res["flags"].any()
    res2, _ = engine.run_region(contig, 10, 400, [])
    ref2, _ = H.run_c_oracle(contig, 10, 400, [])
    H.assert_results_equal(res2, ref2)
    assert res.per_bam() == [(0, 0, 0)] and res2.per_bam() == []


def test_unsorted_batch_is_rejected(engine):
    contig = H.random_contig(random.Random(4), 300)
    q = bytes([30]) * 20
    a = po.Read(pos=100, cigar=[("M", 20)], bases=contig[99:119].upper(), quals=q)
    b = po.Read(pos=50, cigar=[("M", 20)], bases=contig[49:69].upper(), quals=q)
    with pytest.raises(capi.EngineError) as ei:
        engine.run_region(contig, 1, 300, [(pack_records([a, b]), True)])
    assert ei.value.code == capi.PB_ERR_UNSORTED
    # the engine must be reusable (and clean) after a failed region
    res, _ = engine.run_region(contig, 1, 300, [(pack_records([b, a]), True)])
    ref, _ = H.run_c_oracle(contig, 1, 300, [(pack_records([b, a]), True)])
    H.assert_results_equal(res, ref)


def test_engine_reuse_across_regions_leaves_no_residue(engine):
    # sparse planes are self-cleaning: a second, different region on the same handle must be exact
    for seed in (101, 102, 103, 104):
        contig, start, stop, reads = H.random_case(seed, contig_len=700, n_reads=300)
        run_both(engine, contig, start, stop, [(reads, True)])


@pytest.mark.parametrize("seed,depth", [(1, 30), (2, 120)])
def test_medium_synthetic_region(engine, seed, depth):
    """A few tens of thousands of loci: many windows, halo reads on both sides, chunk boundaries."""
    rng = random.Random(seed)
    n = 30000
    contig = H.random_contig(rng, n, n_runs=4)
    start, stop = 2001, 26000
    nreads = depth * (stop - start + 2000) // 100
    reads = [H.random_read(rng, contig, start - 1000, stop + 1000, max_len=110) for _ in range(nreads)]
    for _ in range(30):
        reads += H.planted_indel_cluster(rng, contig, start, stop)
        reads += H.planted_snp_cluster(rng, contig, start, stop)
    reads.sort(key=lambda r: r.pos)
    half = [r for i, r in enumerate(reads) if i % 3 != 0], [r for i, r in enumerate(reads) if i % 3 == 0]
    res, _ = run_both(engine, contig, start, stop, [(half[0], True), (half[1], False)])
    assert (res["flags"] & capi.PB_FL_CONFIRMED).any()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_clean_region_with_planted_variants_calls_everything(engine, seed):
    contig, start, stop, reads = H.clean_case(seed)
    res, _ = run_both(engine, contig, start, stop, [(reads, True)])
    fl = res["flags"]
    kind = (fl >> capi.PB_FL_KIND_SHIFT) & 3
    changed = (fl & capi.PB_FL_CHANGED) != 0
    assert (fl & capi.PB_FL_CONFIRMED).sum() > 0.8 * len(fl)
    for k in (capi.PB_KIND_SNP, capi.PB_KIND_INS, capi.PB_KIND_DEL):
        assert (changed & (kind == k)).any(), "no call of kind %d" % k
    assert (fl & capi.PB_FL_DELETED).any() and (fl & capi.PB_FL_AMBIGUOUS).any()


def test_high_depth_single_locus_contention(engine):
    """5000x on a tiny region (config 5's shape): every read hits the same windows."""
    rng = random.Random(9)
    contig = H.random_contig(rng, 600)
    reads = []
    for _ in range(5000):
        pos = rng.randint(1, 450)
        L = min(100, 600 - pos + 1)
        b = bytearray(contig[pos - 1:pos - 1 + L].upper())
        if rng.random() < 0.3:
            b[rng.randrange(L)] = rng.choice(b"ACGT")
        reads.append(po.Read(pos=pos, cigar=[("M", L)], bases=bytes(b), quals=bytes(rng.randint(2, 41) for _ in range(L)),
                             mapq=rng.choice([0, 20, 60]), paired=True, proper=rng.random() < 0.97,
                             tlen=rng.choice([300, -300])))
    reads.sort(key=lambda r: r.pos)
    run_both(engine, contig, 1, 600, [(reads, True)])


def test_determinism(engine):
    contig, start, stop, reads = H.random_case(77, contig_len=2000, n_reads=900)
    packed = [(pack_records(reads), True)]
    a, _ = engine.run_region(contig, start, stop, packed)
    b, _ = engine.run_region(contig, start, stop, packed)
    H.assert_results_equal(a, b)


def test_compute_timed_is_repeatable_and_leaves_engine_clean(engine):
    contig, start, stop, reads = H.random_case(55, contig_len=3000, n_reads=1500)
    packed = pack_records(reads)
    engine.region_begin(contig, start, stop)
    engine.add_batch(packed, True)
    tot, pil, launches = engine.compute_timed(3)
    assert tot > 0 and pil > 0 and launches >= 3 * 6
    from pilon_b200.packing import ResultBuffers
    res = ResultBuffers(stop + 1 - start, indels_cap=1 << 16, indel_bytes_cap=1 << 20)
    engine.finish(res, [np.zeros(packed.n_reads, np.int32)])
    ref, _ = H.run_c_oracle(contig, start, stop, [(packed, True)])
    H.assert_results_equal(res, ref)


def test_facade_mirrors_reference_surface():
    contig, start, stop, reads = H.random_case(12, contig_len=600, n_reads=250)
    pur = PileUpRegion("c", start, stop, contig)
    gr = po.GenomeRegionHot(contig, start, stop)
    gr.initializePileUps(oob_drop=True)
    rets = [pur.addRead(r, contig) for r in reads]
    gr.processBam(reads, "frags")
    assert rets == [x[0] for x in gr.insert_sizes]
    pur.postProcess()
    gr.postProcess()
    o = gr.pileUpRegion
    assert (pur.readCount, pur.baseCount, pur.coverage, pur.minDepth) == (o.readCount, o.baseCount, o.coverage, gr.minDepth)
    for i in range(pur.size):
        a, b = pur[i], o[i]
        assert (a.depth, a.count, a.badPair, a.physCov, a.insertSize, a.clips, a.deletions, a.insertions) == \
               (b.depth, b.count, b.badPair, b.physCov, b.insertSize, b.clips, b.deletions, b.insertions)
        assert (a.weightedQual, a.weightedMq, a.meanQual, a.meanMq, a.insPct, a.delPct) == \
               (b.weightedQual, b.weightedMq, b.meanQual, b.meanMq, b.insPct, b.delPct)
        assert str(a.baseCount) == str(b.baseCount) and a.qualSum.toStringPct() == b.qualSum.toStringPct()
        ca, cb = a.baseCall(), b.baseCall()
        assert (ca.base, ca.altBase, ca.homo, ca.score, ca.q, ca.highConfidence, ca.called, ca.indel, ca.homoIndel) == \
               (cb.base, cb.altBase, cb.homo, cb.score, cb.q, cb.highConfidence, cb.called, cb.indel, cb.homoIndel)
        assert (ca.insertion, ca.deletion, ca.callString(), ca.baseSum, ca.altBaseSum) == \
               (cb.insertion, cb.deletion, cb.callString(), cb.baseSum, cb.altBaseSum)
    kinds = {po.SNP: 0, po.INS: 1, po.DEL: 2, po.AMB: 3}
    assert pur.changes() == sorted((i, kinds[k]) for i, (k, _) in gr.changeMap.items())


@pytest.mark.parametrize("levels", [(2, 12, 25, 37), (2, 6, 9, 12, 17, 22, 25, 27, 30, 33, 37, 40)])
def test_packed_quality_transport_matches_byte_transport(engine, levels):
    """pb_batch.qual_codes: the engine uploads 3- or 4-bit codes and expands them on the device; results must be identical
    to the byte transport (and to the oracle, which always reads `quals`)."""
    from pilon_b200.packing import ResultBuffers
    rng = random.Random(11)
    contig, start, stop, reads = H.clean_case(11, n=8000, start=501, stop=6500, depth=30, n_sites=10)
    for r in reads:                                       # binned qualities
        if r.quals:
            r.quals = bytes(rng.choice(levels) for _ in r.quals)
    packed = pack_records(reads)
    pq = packed.with_packed_quals()
    assert pq.qual_codes is not None and pq.qual_code_bits == (3 if len(levels) <= 6 else 4)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, [(packed, True)])
    res8, ins8 = engine.run_region(contig, start, stop, [(packed, True)])
    H.assert_results_equal(res8, ref, "byte transport vs C oracle")
    only_codes = pq.to_c()
    only_codes.quals = None                                # the engine must not need the bytes
    engine.region_begin(contig, start, stop)
    engine.add_batch(only_codes, True)
    res4 = ResultBuffers(stop + 1 - start, None, 1 << 16, 1 << 20)
    ins4 = [np.zeros(packed.n_reads, np.int32)]
    engine.finish(res4, ins4)
    H.assert_results_equal(res4, ref, "packed transport vs C oracle")
    assert np.array_equal(ins4[0], ins_ref[0])


@pytest.mark.parametrize("seed", [3, 8])
def test_compact_metadata_transport_matches_plain_arrays(engine, seed):
    """pb_batch.meta_codes: the eight per-read arrays travel as 8 bytes per read + listed CIGARs + escapes and are rebuilt
    on the device (one scan + one kernel); with the plain pointers NULL the results equal the oracle's, which reads the
    plain arrays.  Reads with indels and clips, template lengths beyond int16, a gap of more than 65534 loci, two batches,
    together with the packed quality and reference-delta transports."""
    from pilon_b200.packing import ResultBuffers
    contig, start, stop, reads = H.clean_case(seed, n=260000, start=1001, stop=250000, depth=3, n_sites=30)
    reads = [r for r in reads if not (60000 < r.pos < 160000)]          # a hole in the coverage: pos delta > 65534
    rng = random.Random(seed)
    for r in reads[::11]:
        r.tlen = rng.choice([-70000, 40000, -32768, 1 << 29])
    groups = [([r for i, r in enumerate(reads) if i % 4], True), ([r for i, r in enumerate(reads) if i % 4 == 0], False)]
    packed = [(pack_records(g), f) for g, f in groups]
    ref, ins_ref = H.run_c_oracle(contig, start, stop, packed, indels_cap=1 << 18, bytes_cap=1 << 22)
    engine.region_begin(contig, start, stop)
    keep = []
    for rb, f in packed:
        m = rb.with_compact_meta().with_base_deltas(contig, start, stop)
        assert m.meta is not None and m.meta[4] > 0 and m.meta[3] > 0            # escapes and listed CIGARs present
        c = m.to_c()
        for name in ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off", "bases2"):
            setattr(c, name, None)                         # the engine must not need the plain arrays
        keep.append((m, c))
        engine.add_batch(c, f)
    res = ResultBuffers(stop + 1 - start, None, 1 << 18, 1 << 22, calls_cap=1 << 16)
    ins = [np.zeros(rb.n_reads, np.int32) for rb, _ in packed]
    engine.finish(res, ins)
    H.assert_results_equal(res, ref, "compact metadata vs C oracle")
    for a, b in zip(ins, ins_ref):
        assert np.array_equal(a, b)


def test_compact_metadata_that_overruns_its_batch_is_refused(engine):
    """Records that describe more CIGAR ops than n_cigar says must not write past the arrays: the pass is refused."""
    from pilon_b200.packing import ResultBuffers
    contig, start, stop, reads = H.clean_case(5, n=6000, start=501, stop=5500, depth=6, n_sites=6)
    m = pack_records(reads).with_compact_meta()
    c = m.to_c()
    c.n_cigar = c.n_cigar - 3
    engine.region_begin(contig, start, stop)
    engine.add_batch(c, True)
    res = ResultBuffers(stop + 1 - start, ["flags"], 1 << 16, 1 << 20)
    with pytest.raises(capi.EngineError):
        engine.finish(res, [np.zeros(m.n_reads, np.int32)])
    # the engine is usable afterwards
    ok, _ = engine.run_region(contig, start, stop, [(pack_records(reads), True)])
    ref, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)])
    H.assert_results_equal(ok, ref, "after a refused pass")


def test_more_batches_than_the_by_value_table_holds(engine):
    """More than 20 BAMs in one region (the kernels' by-value batch table holds 20): the engine falls back to the
    kernel generation that walks a device-side batch list; results stay bit-identical."""
    contig, start, stop, reads = H.clean_case(21, n=9000, start=301, stop=7000, depth=40, n_sites=12)
    rng = random.Random(21)
    k = 27
    groups = [[] for _ in range(k)]
    for r in reads:
        groups[rng.randrange(k)].append(r)
    run_both(engine, contig, start, stop, [(g, i % 3 != 0) for i, g in enumerate(groups)])


def test_device_resident_batches_match_host_batches(engine):
    """PB_MEM_DEVICE: arrays already on the GPU are used in place (the bench's HBM-resident arm)."""
    import ctypes as C
    import torch
    contig, start, stop, reads = H.clean_case(22, n=9000, start=301, stop=7000, depth=30, n_sites=10)
    packed = pack_records(reads)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, [(packed, True)])
    host = packed.to_c()
    dev = capi.pb_batch()
    dev.n_reads, dev.n_cigar, dev.n_seq, dev.n_exc = host.n_reads, host.n_cigar, host.n_seq, host.n_exc
    keep = []
    for name in ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off", "quals", "bases2",
                 "exc_idx", "exc_base", "exc_qual"):
        a = getattr(packed, name)
        t = torch.zeros(a.nbytes + 64, dtype=torch.uint8, device="cuda:0")      # readable 16 bytes past the end
        if a.nbytes:
            t[:a.nbytes].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
        keep.append(t)
        setattr(dev, name, t.data_ptr())
    dev.mem = capi.PB_MEM_DEVICE
    torch.cuda.synchronize()
    from pilon_b200.packing import ResultBuffers
    engine.region_begin(contig, start, stop)
    engine.add_batch(dev, True)
    res = ResultBuffers(stop + 1 - start, None, 1 << 16, 1 << 20)
    ins = [np.zeros(packed.n_reads, np.int32)]
    engine.finish(res, ins)
    H.assert_results_equal(res, ref, "device-resident batch vs C oracle")
    assert np.array_equal(ins[0], ins_ref[0])


def test_insertion_only_cigars_break_the_event_bound_and_are_retried():
    """The region pass sizes its indel event list by a host-side bound (CIGAR operations minus reads); reads whose CIGAR
    is a single I operation break it, the kernels raise the capacity flag and pb_region_finish repeats the pass at full
    capacity.  Results must match the oracle exactly as if nothing had happened."""
    rng = random.Random(5)
    contig = bytes(rng.choice(b"ACGT") for _ in range(1500))
    reads = []
    for i in range(40):                                  # ordinary coverage
        p = 100 + 20 * i
        reads.append(po.Read(pos=p, cigar=[("M", 80)], bases=contig[p - 1:p + 79], quals=bytes([30]) * 80, mapq=50))
    for i in range(150):                                 # 150 single-op insertions > the 64-entry slack of the bound
        p = 300 + 3 * i
        ins = bytes(rng.choice(b"ACGT") for _ in range(2))
        reads.append(po.Read(pos=p, cigar=[("I", 2)], bases=ins, quals=bytes([25, 25]), mapq=40))
    reads.sort(key=lambda r: r.pos)
    cfg = po.Config(flank=0)
    e = Engine(0, eng_cfg(cfg))
    try:
        res, _ = run_both(e, contig, 1, 1500, [(reads, True)], cfg)
        assert int(res["insertions"].sum()) == 150
    finally:
        e.close()


def test_graph_replayed_passes_leave_the_engine_clean(engine):
    """pb_region_compute: plain pass, captured pass, graph replays -- each must consume and re-zero the sparse planes
    exactly like pb_region_finish's own pass, so the results read afterwards are still bit-exact."""
    contig, start, stop, reads = H.clean_case(31, n=9000, start=301, stop=7000, depth=30, n_sites=12)
    packed = pack_records(reads)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, [(packed, True)])
    from pilon_b200.packing import ResultBuffers
    engine.region_begin(contig, start, stop)
    engine.add_batch(packed, True)
    for _ in range(5):
        engine.compute()
    res = ResultBuffers(stop + 1 - start, None, 1 << 16, 1 << 20)
    ins = [np.zeros(packed.n_reads, np.int32)]
    engine.finish(res, ins)
    H.assert_results_equal(res, ref, "after five async passes")
    assert np.array_equal(ins[0], ins_ref[0])


@pytest.mark.parametrize("seed", [2, 8, 17, 33])
def test_base_delta_transport_matches_byte_transport(engine, seed):
    """pb_batch.base_delta_idx: the engine uploads only the bases that differ from the reference prediction and rebuilds
    bases2 on the device; together with the packed qualities nothing of the per-base arrays is uploaded as is."""
    from pilon_b200.packing import ResultBuffers
    contig, start, stop, reads = H.random_case(seed)
    packed = pack_records(reads)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, [(packed, True)])
    compact = packed.with_base_deltas(contig, start, stop).with_packed_quals()
    cb = compact.to_c()
    cb.bases2 = None                                       # the engine must not need the bytes
    if compact.qual_codes is not None:
        cb.quals = None
    engine.region_begin(contig, start, stop)
    engine.add_batch(cb, True)
    res = ResultBuffers(stop + 1 - start, None, 1 << 16, 1 << 20)
    ins = [np.zeros(packed.n_reads, np.int32)]
    engine.finish(res, ins)
    H.assert_results_equal(res, ref, "base-delta transport vs C oracle")
    assert np.array_equal(ins[0], ins_ref[0])


@pytest.mark.parametrize("seed,long_read", [(s, 1 + s % 2) for s in range(10)])
def test_long_read_batches_match_both_oracles(engine, seed, long_read):
    """--nanopore / --pacbio batches (BamFile.longReadType 1 / 2) beside a short-read batch in the same region: k_long
    accumulates them, the tile kernels the rest, the epilogue merges both (PileUpRegion.scala:120-134,142,160,180-181,190)."""
    rng = random.Random(1000 + seed)
    contig = bytearray(H.random_contig(rng, 900))
    for _ in range(14):
        p = rng.randrange(2, 890)
        contig[p:p + 5] = b"CC" + bytes([rng.choice(b"ACGT")]) + b"GG"
    contig = bytes(contig)
    start = 1 if seed % 3 == 0 else rng.randint(2, 150)
    stop = rng.randint(600, 900)
    reads = [H.random_read(rng, contig, max(1, start - 40), min(900, stop + 20), max_len=150) for _ in range(400)]
    reads += H.planted_indel_cluster(rng, contig, start, stop) + H.planted_snp_cluster(rng, contig, start, stop)
    reads.sort(key=lambda r: r.pos)
    groups = [([r for i, r in enumerate(reads) if i % 3 == 0], True, 0), ([r for i, r in enumerate(reads) if i % 3 == 1], False, long_read),
              ([r for i, r in enumerate(reads) if i % 3 == 2], True, long_read)]
    packed = [(pack_records(g), f, lr) for g, f, lr in groups]
    res, ins = engine.run_region(contig, start, stop, packed)
    ref, ins_ref = H.run_c_oracle(contig, start, stop, packed)
    H.assert_results_equal(res, ref, "long reads, engine vs C oracle")
    H.assert_matches_py(res, ins, H.run_py_oracle(contig, start, stop, groups), "long reads, engine vs python oracle")
    # and the engine is clean afterwards: a plain short-read region on the same handle
    c2, s2, e2, r2 = H.random_case(seed, contig_len=500, n_reads=150)
    run_both(engine, c2, s2, e2, [(r2, True)])
