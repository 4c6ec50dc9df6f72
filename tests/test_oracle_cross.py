"""Two independent restatements of the reference path must agree bit for bit:
the literal Python transliteration (oracle/pilon_oracle.py) and the C restatement
(oracle/pilon_oracle.c) that consumes the engine's packed batches."""
import random

import pytest

from oracle import pilon_oracle as po
from pilon_b200.packing import pack_records
from tests import helpers as H


@pytest.mark.parametrize("seed", range(40))
def test_c_oracle_matches_python_oracle(seed):
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
    py = H.run_py_oracle(contig, start, stop, groups, cfg)
    res, ins = H.run_c_oracle(contig, start, stop, [(pack_records(g), f) for g, f in groups], cfg)
    H.assert_matches_py(res, ins, py, "seed %d" % seed)


def test_cases_exercise_the_interesting_paths():
    """The generator must actually reach indel calls, the deletion spill, drops, clips..."""
    seen = dict(ins_call=0, del_call=0, deleted=0, dropped=0, snp=0, amb=0, confirmed=0, unknown=0)
    for seed in range(40):
        contig, start, stop, reads = H.random_case(seed)
        res, _ = H.run_c_oracle(contig, start, stop, [(pack_records(reads), True)])
        fl = res["flags"]
        kind = (fl >> 4) & 3
        seen["ins_call"] += int(((fl & 2) != 0)[kind == 1].sum())
        seen["del_call"] += int(((fl & 2) != 0)[kind == 2].sum())
        seen["snp"] += int((((fl & 2) != 0) & (kind == 0)).sum())
        seen["amb"] += int(((fl & 4) != 0).sum())
        seen["deleted"] += int(((fl & 8) != 0).sum())
        seen["confirmed"] += int((fl & 1).sum())
        seen["dropped"] += res.c.dropped_oob
        seen["unknown"] += res.c.unknown_ops
    assert all(v > 0 for v in seen.values()), seen


def test_c_oracle_matches_python_oracle_on_a_wide_region():
    """120 k loci, ~25 k paired reads, planted SNP / AMB / INS / DEL sites and the deletion spill: the C restatement
    (whose BaseCall / hetIndelCall code is close to the device code's) is cross-examined by the literal transliteration
    well beyond the toy sizes of the random cases."""
    contig, start, stop, reads = H.clean_case(5, n=130_000, start=5_001, stop=125_000, depth=10, n_sites=150)
    groups = [([r for i, r in enumerate(reads) if i % 3 != 0], True), ([r for i, r in enumerate(reads) if i % 3 == 0], False)]
    py = H.run_py_oracle(contig, start, stop, groups)
    res, ins = H.run_c_oracle(contig, start, stop, [(pack_records(g), f) for g, f in groups], indels_cap=1 << 18, bytes_cap=1 << 22)
    H.assert_matches_py(res, ins, py, "120 k loci")
    fl = res["flags"]
    kind = (fl >> 4) & 3
    for k in (0, 1, 2):
        assert (((fl & 2) != 0) & (kind == k)).any()
    assert (fl & 8).any() and (fl & 4).any()


@pytest.mark.parametrize("seed,long_read", [(s, 1 + s % 2) for s in range(16)])
def test_long_read_branches_c_oracle_matches_python_oracle(seed, long_read):
    """--nanopore (1) / --pacbio (2): indelMq capped at 8, insertions / deletions in homopolymers dropped, nanopore CC.GG
    motifs (PileUpRegion.scala:120-134,142,160,180-181,190), including the reference's habit of indexing the contig with
    REGION indices there -- visible whenever the region does not start at locus 1."""
    rng = random.Random(1000 + seed)
    contig = bytearray(H.random_contig(rng, 700))
    for _ in range(12):                                   # plant CC.GG motifs
        p = rng.randrange(2, 690)
        contig[p:p + 5] = b"CC" + bytes([rng.choice(b"ACGT")]) + b"GG"
    contig = bytes(contig)
    start = 1 if seed % 3 == 0 else rng.randint(2, 150)
    stop = rng.randint(500, 700)
    reads = [H.random_read(rng, contig, max(1, start - 40), min(700, stop + 20)) for _ in range(260)]
    reads += H.planted_indel_cluster(rng, contig, start, stop)
    reads.sort(key=lambda r: r.pos)
    short = [r for i, r in enumerate(reads) if i % 3 == 0]
    longs = [r for i, r in enumerate(reads) if i % 3 != 0]
    groups = [(short, True, 0), (longs, True, long_read)]
    py = H.run_py_oracle(contig, start, stop, groups)
    res, ins = H.run_c_oracle(contig, start, stop, [(pack_records(g), f, lr) for g, f, lr in groups])
    H.assert_matches_py(res, ins, py, "long read %d seed %d" % (long_read, seed))
