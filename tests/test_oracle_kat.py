"""Hand-derived known-answer tests pinning the literal Python oracle.

The reference ships no tests; these vectors were derived by hand from the Scala
(PileUpRegion.scala:102-220, PileUp.scala:75-114,132-247) with flank=10, minQual=0, minMq=0.
"""
import pytest

from oracle import pilon_oracle as po
from oracle.pilon_oracle import Config, PileUp, PileUpRegion, Read


def mkref(n, fill=b"ACGT"):
    return (fill * (n // len(fill) + 1))[:n]


def test_kat1_single_30M_read():
    ref = mkref(100)
    pur = PileUpRegion("c", 1, 100)
    bases = ref[10:40]  # aligned at aStart=11
    rd = Read(pos=11, cigar=[("M", 30)], bases=bases, quals=bytes([30]) * 30, mapq=60)
    ins = pur.addRead(rd, ref)
    for locus in range(1, 101):
        pu = pur[locus - 1]
        if 21 <= locus <= 30:
            bi = po.baseIndex(ref[locus - 1])
            assert pu.baseCount.sums[bi] == 1 and pu.count == 1
            assert pu.qualSum.sums[bi] == 30 * 61 == 1830
            assert pu.mqSum == 61 and pu.qSum == 30
        else:
            assert pu.count == 0 and pu.mqSum == 0
    assert pur.baseCount == 10 and pur.readCount == 1
    assert ins == 29
    pur.postProcess()
    for locus in range(1, 101):
        pu = pur[locus - 1]
        if 11 <= locus <= 39:
            assert pu.physCov == 1 and pu.insertSize == 29
        else:
            assert pu.physCov == 0


def _pu_with(qs, bc, mqSum, qSum=1):
    pu = PileUp(Config())
    pu.qualSum.sums = list(qs)
    pu.baseCount.sums = list(bc)
    pu.mqSum = mqSum
    pu.qSum = qSum
    return pu


def test_kat2_homozygous_call():
    bc = _pu_with([18300, 0, 0, 0], [10, 0, 0, 0], 610, 300).baseCall()
    assert bc.base == "A" and bc.homo and bc.score == 300 and bc.q == 30 and bc.highConfidence
    assert bc.called and not bc.indel


def test_kat3_tie_is_het_and_stable_order():
    bc = _pu_with([9150, 9150, 0, 0], [5, 5, 0, 0], 610, 300).baseCall()
    assert bc.base == "A" and bc.altBase == "C"
    assert not bc.homo and bc.score == 300


def test_kat4_empty_locus():
    bc = PileUp(Config()).baseCall()
    assert bc.n == 0 and bc.base == "N" and bc.altBase == "C" and bc.score == 0
    assert not bc.called and bc.homo and bc.q == 0


def test_kat5_leading_soft_clip():
    ref = mkref(200)
    pur = PileUpRegion("c", 1, 200)
    rd = Read(pos=50, cigar=[("S", 5), ("M", 45)], bases=ref[44:94], quals=bytes([30]) * 50, mapq=60)
    pur.addRead(rd, ref)
    assert [pur[l - 1].clips for l in range(44, 51)] == [0, 1, 0, 0, 0, 1, 0]
    assert [pur[l - 1].badPair for l in range(44, 51)] == [0, 1, 1, 1, 1, 1, 0]
    # adjMq = roundDiv(60*45, 50) = 54 -> mq1 55 on trusted aligned bases (offsets 10..39 -> loci 55..84)
    assert pur[55 - 1].mqSum == 55 and pur[54 - 1].mqSum == 0 and pur[84 - 1].mqSum == 55 and pur[85 - 1].mqSum == 0


def test_kat6_1S_bumps_clips_twice():
    ref = mkref(200)
    pur = PileUpRegion("c", 1, 200)
    rd = Read(pos=50, cigar=[("M", 49), ("S", 1)], bases=ref[49:99], quals=bytes([30]) * 50, mapq=60)
    pur.addRead(rd, ref)
    assert pur[99 - 1].clips == 2 and pur[99 - 1].badPair == 1


def test_kat7_deletion_left_shift_readds_bases():
    # ref: 'A' at loci 101..104; 100 and 105 differ
    ref = bytearray(b"C" * 300)
    for l in range(101, 105):
        ref[l - 1] = ord("A")
    ref[99] = ord("G")
    ref[104] = ord("T")
    ref = bytes(ref)
    # read: 50M1D50M with the deletion reported at 104; offsets 47,48,49 <-> loci 101,102,103
    pos = 101 - 47
    rb = ref[pos - 1:pos - 1 + 50] + ref[104:104 + 50]
    pur = PileUpRegion("c", 1, 300)
    rd = Read(pos=pos, cigar=[("M", 50), ("D", 1), ("M", 50)], bases=rb, quals=bytes([30]) * 100, mapq=60)
    pur.addRead(rd, ref)
    A = 0
    assert [pur[l - 1].baseCount.sums[A] for l in (101, 102, 103, 104)] == [1, 2, 2, 1]
    assert pur[101 - 1].deletions == 1 and pur[101 - 1].deletionList == [b"A"]
    assert pur[104 - 1].deletions == 0
    # region baseCount counts the three re-adds too: 80 trusted M bases + 3
    assert pur.baseCount == 83


def test_kat8_insertion_left_shift_and_unshifted_qual():
    ref = bytearray(b"C" * 200)
    for l in range(60, 63):  # AAA at 60,61,62
        ref[l - 1] = ord("A")
    ref = bytes(ref)
    # read starts at 40, 23M (loci 40..62) then 1I 'A', then 27M from locus 63
    quals = bytearray([30]) * 51
    quals[23] = 17  # qual of the inserted base (unshifted readOffset = 23)
    rb = ref[39:62] + b"A" + ref[62:89]
    pur = PileUpRegion("c", 1, 200)
    rd = Read(pos=40, cigar=[("M", 23), ("I", 1), ("M", 27)], bases=rb, quals=bytes(quals), mapq=60)
    pur.addRead(rd, ref)
    # reported at locus 63; shifts left while ref[iloc-1] == 'A': 63 -> 60
    assert pur[60 - 1].insertions == 1 and pur[60 - 1].insertionList == [b"A"]
    assert pur[60 - 1].insQual == 61
    # qSum at 60 = 30 (aligned base, trusted) + 17 (insertion anchor qual)
    assert pur[60 - 1].qSum == 47
    assert pur[63 - 1].insertions == 0


def test_kat9_physcov_pairs():
    ref = mkref(1000)
    pur = PileUpRegion("c", 1, 1000)
    q = bytes([30]) * 50
    left = Read(pos=100, cigar=[("M", 50)], bases=ref[99:149], quals=q, paired=True, proper=True, tlen=400)
    right = Read(pos=450, cigar=[("M", 50)], bases=ref[449:499], quals=q, paired=True, proper=True, tlen=-400)
    assert pur.addRead(left, ref) == 400
    assert pur.addRead(right, ref) == 0
    pur.postProcess()
    cov = [pur[l - 1].physCov for l in range(1, 1001)]
    assert all(c == 1 for c in cov[99:499]) and cov[98] == 0 and cov[499] == 0
    assert pur[300 - 1].insertSize == 400


def _indel_pu(n_same, n_other, count, mqSum, insQual):
    pu = PileUp(Config())
    pu.baseCount.sums = [count, 0, 0, 0]
    pu.qualSum.sums = [count * 30 * 61, 0, 0, 0]
    pu.mqSum = mqSum
    pu.qSum = 30 * count
    pu.insertions = n_same + n_other
    pu.insQual = insQual
    pu.insertionList = [b"G"] * n_same + [b"T"] * n_other
    return pu


def test_kat10_het_indel_window():
    # 3 of 5 identical 1-bp insertions; pct chosen through count: insPct = max(pct(insQual,mqSum), pct(5,count))
    pu = _indel_pu(3, 2, count=100, mqSum=6100, insQual=3660)  # pct = 60
    assert pu.insPct == 60
    bc = pu.baseCall()
    assert bc.isInsertion and bc.insertion == "G" and not bc.homoIndel  # middle 44, low 22, high 66
    pu = _indel_pu(3, 2, count=100, mqSum=6100, insQual=4270)  # pct = 70
    assert pu.insPct == 70
    bc = pu.baseCall()
    assert bc.isInsertion and bc.homoIndel
    pu = _indel_pu(2, 2, count=100, mqSum=6100, insQual=4270)  # 2-vs-2 tie of 4
    bc = pu.baseCall()
    assert not bc.indel


def test_jvm_arithmetic_helpers():
    assert po.roundDivI(7, 2) == 4 and po.roundDivI(7, 0) == 0 and po.roundDivI(7, -3) == 0
    assert po.roundDivL(-7, 2) == -3  # (-7 + 1) / 2 truncates toward zero
    assert po.pctI(30_000_000, 7) == po.i32(jdiv_ref(po.i32(3_000_000_000) + 3, 7))
    assert po.sbyte(0xFF) == -1 and po.toshort(0x18000) == -32768
    assert po.jround(2.5) == 3 and po.jround(-2.5) == -2


def jdiv_ref(n, d):
    return int(n / d)


def test_insertion_shift_below_region_start_raises_like_jvm():
    # latent crash at PileUpRegion.scala:156-161 (region.start > 1): we surface it as IndexError
    ref = b"C" * 50 + b"A" * 20 + b"C" * 50
    pur = PileUpRegion("c", 61, 120)
    rb = ref[39:62] + b"A" + ref[62:89]
    rd = Read(pos=40, cigar=[("M", 23), ("I", 1), ("M", 27)], bases=rb, quals=bytes([30]) * 51, mapq=60)
    with pytest.raises(IndexError):
        pur.addRead(rd, ref)
