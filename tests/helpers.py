"""Test infrastructure shared by the CPU and GPU suites: the C oracle binding, the adversarial
random read generator, and result comparison.  Nothing here is imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import random
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from oracle import pilon_oracle as po
from pilon_b200 import _capi as capi
from pilon_b200.packing import ReadBatch, ResultBuffers, pack_records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libpilon_oracle.so")

_olib = None


def oracle_lib() -> C.CDLL:
    global _olib
    if _olib is None:
        src = os.path.join(ROOT, "oracle", "pilon_oracle.c")
        if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
        lib = C.CDLL(ORACLE_SO)
        lib.po_region_new.restype = C.c_void_p
        lib.po_region_new.argtypes = [C.POINTER(capi.pb_config), C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
        lib.po_region_add_batch.argtypes = [C.c_void_p, C.POINTER(capi.pb_batch), C.c_int, C.c_int, C.c_void_p]
        lib.po_region_finish.argtypes = [C.c_void_p, C.POINTER(capi.pb_region_result)]
        lib.po_region_free.argtypes = [C.c_void_p]
        _olib = lib
    return _olib


def make_config(cfg: Optional[po.Config] = None) -> capi.pb_config:
    cfg = cfg or po.Config()
    return capi.pb_config(min_qual=cfg.minQual, min_mq=cfg.minMq, flank=cfg.flank, default_qual=cfg.defaultQual,
                          min_min_depth=cfg.minMinDepth, old_indel=int(cfg.oldIndel),
                          fix_amb=int(cfg.iupac or cfg.fixAmb), min_depth=cfg.minDepth)


def run_c_oracle(contig: bytes, start: int, stop: int, batches: Sequence[Tuple[ReadBatch, bool]],
                 cfg: Optional[po.Config] = None, indels_cap: int = 1 << 16, bytes_cap: int = 1 << 20):
    """batches = [(ReadBatch, counts_toward_frag_coverage)].  Returns (ResultBuffers, [insert sizes])."""
    lib = oracle_lib()
    ccfg = make_config(cfg)
    cbuf = np.frombuffer(contig, np.uint8)
    reg = lib.po_region_new(C.byref(ccfg), cbuf.ctypes.data, len(contig), start, stop)
    inserts = []
    try:
        for bt in batches:
            rb, frag, long_read = bt if len(bt) == 3 else (bt[0], bt[1], 0)
            cb = rb.to_c()
            ins = np.zeros(rb.n_reads, np.int32)
            rc = lib.po_region_add_batch(reg, C.byref(cb), int(frag), int(long_read), ins.ctypes.data)
            assert rc == 0
            inserts.append(ins)
        res = ResultBuffers(stop + 1 - start, indels_cap=indels_cap, indel_bytes_cap=bytes_cap, calls_cap=min(stop + 1 - start, 1 << 20))
        assert lib.po_region_finish(reg, C.byref(res.c)) == 0
    finally:
        lib.po_region_free(reg)
    return res, inserts


# ---------------------------------------------------------------------------------------------
# literal Python oracle -> the same result layout
# ---------------------------------------------------------------------------------------------
def run_py_oracle(contig: bytes, start: int, stop: int, batches: Sequence[Tuple[Sequence[po.Read], bool]],
                  cfg: Optional[po.Config] = None, oob_drop: bool = True) -> Dict[str, object]:
    cfg = cfg or po.Config()
    gr = po.GenomeRegionHot(contig, start, stop, cfg)
    gr.initializePileUps(oob_drop=oob_drop)
    for bt in batches:
        reads, frag, long_read = bt if len(bt) == 3 else (bt[0], bt[1], 0)
        gr.processBam(reads, "frags" if frag else "jumps", long_read)
    gr.postProcess()
    pur = gr.pileUpRegion
    S = gr.size
    out: Dict[str, object] = {}
    out["base_count4"] = np.array([p.baseCount.sums for p in pur.pileups], np.int64).astype(np.int32).reshape(S, 4)
    out["qual_sum4"] = np.array([p.qualSum.sums for p in pur.pileups], np.int64).reshape(S, 4)
    for name, attr in (("mq_sum", "mqSum"), ("q_sum", "qSum"), ("phys_cov", "physCov"), ("insert_size", "insertSize"),
                       ("bad_pair", "badPair"), ("deletions", "deletions"), ("del_qual", "delQual"),
                       ("insertions", "insertions"), ("ins_qual", "insQual"), ("clips", "clips")):
        out[name] = np.array([getattr(p, attr) for p in pur.pileups], np.int32)
    out["coverage_arr"] = np.array(gr.coverage, np.int32)
    out["frag_coverage"] = np.array(gr.fragCoverage, np.int32)
    out["weighted_qual"] = np.array(gr.weightedQual, np.int8)
    out["weighted_mq"] = np.array(gr.weightedMq, np.int8)
    kinds = {po.SNP: capi.PB_KIND_SNP, po.INS: capi.PB_KIND_INS, po.DEL: capi.PB_KIND_DEL, po.AMB: capi.PB_KIND_AMB}
    flags = np.zeros(S, np.uint8)
    for i in range(S):
        f = 0
        if gr.confirmed[i]:
            f |= capi.PB_FL_CONFIRMED
        if gr.changed[i]:
            f |= capi.PB_FL_CHANGED
        if gr.ambiguous[i]:
            f |= capi.PB_FL_AMBIGUOUS
        if gr.deleted[i]:
            f |= capi.PB_FL_DELETED
        if i in gr.changeMap:
            f |= kinds[gr.changeMap[i][0]] << capi.PB_FL_KIND_SHIFT
        flags[i] = f
    out["flags"] = flags
    # final-state BaseCall (what Vcf.writeRecord recomputes)
    call = np.zeros(S, np.uint64)
    indel_strings = {}
    for i in range(S):
        bc = pur.pileups[i].baseCall()
        base = "ACGTN".index(bc.base)
        indel = 1 if bc.isInsertion else 2 if bc.isDeletion else 0
        if indel:
            indel_strings[(i, indel)] = (bc.insertion if indel == 1 else bc.deletion).encode("latin1")
        call[i] = (base | (bc.altBaseIndex << 3) | (int(bc.homo) << 5) | (indel << 6) | (int(bc.homoIndel) << 8)
                   | (int(bc.called) << 9) | (int(bc.highConfidence) << 10) | (int(bc.score) << 16))
    out["call"] = call
    out["indel_strings"] = indel_strings
    out["indel_list_len"] = {}
    for i, p in enumerate(pur.pileups):
        if p.insertionList:
            out["indel_list_len"][(i, 1)] = len(p.insertionList)
        if p.deletionList:
            out["indel_list_len"][(i, 2)] = len(p.deletionList)
    out["scalars"] = dict(base_count=pur.baseCount, read_count=pur.readCount, coverage=pur.coverage,
                          min_depth=gr.minDepth, unknown_ops=pur.unknown_ops, dropped_oob=pur.dropped_oob)
    out["insert_sizes"] = [x[0] for x in gr.insert_sizes]
    out["per_bam"] = list(gr.per_bam)
    return out


PLANE_NAMES = [p[0] for p in capi.RESULT_PLANES]


def assert_results_equal(a: ResultBuffers, b: ResultBuffers, what: str = ""):
    """Bit-exact comparison of two engine-layout results (e.g. CUDA engine vs C oracle)."""
    for f in ("size", "base_count", "coverage", "aligned_bases", "read_count", "min_depth", "unknown_ops",
              "dropped_oob", "n_indels"):
        assert getattr(a.c, f) == getattr(b.c, f), "%s scalar %s: %r != %r" % (what, f, getattr(a.c, f), getattr(b.c, f))
    assert a.c.n_batches == b.c.n_batches and a.per_bam() == b.per_bam(), "%s per-BAM deltas: %r != %r" % (what, a.per_bam(), b.per_bam())
    for name in PLANE_NAMES:
        if name in a.arrays and name in b.arrays:
            x, y = a[name], b[name]
            if not np.array_equal(x, y):
                bad = np.argwhere(x != y)[:5]
                raise AssertionError("%s plane %s differs at %s: %s vs %s" % (
                    what, name, bad.tolist(), x[tuple(bad[0])], y[tuple(bad[0])]))
    ia, ib = a.indels(), b.indels()
    assert len(ia) == len(ib)
    for ea, eb in zip(ia, ib):
        assert (ea["locus_index"], ea["kind"], ea["list_len"]) == (eb["locus_index"], eb["kind"], eb["list_len"]), (ea, eb)
        assert _winner(ea) == _winner(eb), (ea, eb)
    # the call plane in sparse form: against the other side's entries and against the planes it abbreviates
    for r in (a, b):
        if r.calls_cap and "flags" in r.arrays and "call" in r.arrays:
            assert_calls_match_planes(r, r, what)
    if a.calls_cap and b.calls_cap:
        assert a.c.n_calls == b.c.n_calls, "%s n_calls: %d != %d" % (what, a.c.n_calls, b.c.n_calls)
        assert np.array_equal(a.calls(), b.calls()), "%s sparse call entries differ" % what


def assert_calls_match_planes(sparse: ResultBuffers, full: ResultBuffers, what: str = ""):
    """pb_region_result.calls == [(i, flags[i], call[i]) for the loci with a changing / ambiguous call], in locus order."""
    fl = full["flags"]
    idx = np.flatnonzero(fl & (capi.PB_FL_CHANGED | capi.PB_FL_AMBIGUOUS))
    got = sparse.calls()
    assert int(sparse.c.n_calls) == len(idx), "%s n_calls %d != %d flagged loci" % (what, sparse.c.n_calls, len(idx))
    assert np.array_equal(got["locus_index"], idx.astype(np.int32)), "%s sparse call loci differ" % what
    assert np.array_equal(got["flags"], fl[idx].astype(np.uint32)), "%s sparse call flags differ" % what
    assert np.array_equal(got["call"], full["call"][idx]), "%s sparse call records differ" % what


def _winner(e):
    """The winning string is only defined (and only ever consumed, PileUp.scala:219-220) when it is a
    strict majority with count >= 2; the engine reports win_count = 0 otherwise."""
    if e["win_count"] >= 2 and e["win_count"] * 2 > e["list_len"]:
        return (e["win_count"], e["win_len"], e["win_has_n"], e["string"])
    return None


def assert_matches_py(res: ResultBuffers, inserts: List[np.ndarray], py: Dict[str, object], what: str = ""):
    """Engine-layout result vs the literal Python oracle."""
    sc = py["scalars"]
    for f in ("base_count", "read_count", "coverage", "min_depth", "unknown_ops", "dropped_oob"):
        assert getattr(res.c, f) == sc[f], "%s scalar %s: %r != %r" % (what, f, getattr(res.c, f), sc[f])
    for name in PLANE_NAMES:
        x, y = res[name], py[name]
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)[:5]
            raise AssertionError("%s plane %s differs at %s: got %s want %s" % (
                what, name, bad.tolist(), x[tuple(bad[0])], y[tuple(bad[0])]))
    got = {(e["locus_index"], e["kind"]): e for e in res.indels()}
    assert {k: v["list_len"] for k, v in got.items()} == py["indel_list_len"]
    for key, s in py["indel_strings"].items():
        assert got[key]["string"] == s, (key, got[key], s)
    flat = [int(v) for arr in inserts for v in arr]
    assert flat == py["insert_sizes"]
    assert res.per_bam() == py["per_bam"], "%s per-BAM deltas: %r != %r" % (what, res.per_bam(), py["per_bam"])


# ---------------------------------------------------------------------------------------------
# adversarial random inputs
# ---------------------------------------------------------------------------------------------
def random_contig(rng: random.Random, n: int, lower_frac: float = 0.02, n_runs: int = 1) -> bytes:
    out = bytearray()
    while len(out) < n:
        if rng.random() < 0.25:
            out += bytes([rng.choice(b"ACGT")]) * rng.randint(2, 9)     # homopolymers: indel left-shift fodder
        elif rng.random() < 0.1:
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(2, 4)))
            out += unit * rng.randint(2, 5)                              # short tandem repeats
        else:
            out += bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 12)))
    out = out[:n]
    for i in range(n):
        if rng.random() < lower_frac:
            out[i] = out[i] | 0x20                                       # lower case (raw compares!)
    for _ in range(n_runs):
        if n > 60:
            s = rng.randrange(0, n - 20)
            for i in range(s, min(n, s + rng.randint(1, 15))):
                out[i] = ord("N")
    return bytes(out)


def random_read(rng: random.Random, contig: bytes, lo: int, hi: int, max_len: int = 60) -> po.Read:
    """A read with an arbitrary (legal-ish) CIGAR whose aligned part starts in [lo, hi]."""
    n = len(contig)
    pos = max(1, min(n, rng.randint(lo, hi)))
    cigar: List[Tuple[str, int]] = []
    bases = bytearray()
    if rng.random() < 0.1:
        cigar.append(("H", rng.randint(1, 5)))
    if rng.random() < 0.25:
        ln = rng.randint(1, 12)
        cigar.append(("S", ln))
        bases += bytes(rng.choice(b"ACGTN") for _ in range(ln))
    refpos = pos  # 1-based locus of next ref base
    nseg = rng.randint(1, 4)
    for s in range(nseg):
        ln = rng.randint(1, max(1, max_len // nseg))
        ln = min(ln, n - refpos + 1)
        if ln <= 0:
            break
        op = rng.choice("MMMMM=X")
        seg = bytearray(contig[refpos - 1:refpos - 1 + ln].upper())
        for i in range(ln):
            u = rng.random()
            if u < 0.03:
                seg[i] = rng.choice(b"ACGT")
            elif u < 0.04:
                seg[i] = rng.choice(b"NRYMK=")
        cigar.append((op, ln))
        bases += seg
        refpos += ln
        if s < nseg - 1 and refpos <= n:
            u = rng.random()
            if u < 0.35:
                il = rng.randint(1, 4)
                # bias inserted bases toward the preceding reference base so the left shift triggers
                prev = contig[refpos - 2:refpos - 1].upper() or b"A"
                ib = bytes(prev[0] if rng.random() < 0.6 else rng.choice(b"ACGTN") for _ in range(il))
                cigar.append(("I", il))
                bases += ib
            elif u < 0.7:
                dl = min(rng.randint(1, 4), n - refpos + 1)
                if dl > 0:
                    cigar.append(("D", dl))
                    refpos += dl
            elif u < 0.8:
                nl = min(rng.randint(1, 30), n - refpos + 1)
                if nl > 0:
                    cigar.append(("N", nl))
                    refpos += nl
            elif u < 0.85:
                cigar.append(("P", rng.randint(1, 3)))
    if not any(op in "M=X" for op, _ in cigar):
        ln = min(5, n - pos + 1)
        cigar.append(("M", ln))
        bases += contig[pos - 1:pos - 1 + ln].upper()
    if rng.random() < 0.25:
        ln = rng.randint(1, 12)
        cigar.append(("S", ln))
        bases += bytes(rng.choice(b"ACGTN") for _ in range(ln))
    if rng.random() < 0.1:
        cigar.append(("H", rng.randint(1, 5)))
    L = len(bases)
    u = rng.random()
    if u < 0.08:
        quals = b""
    elif u < 0.12:
        quals = bytes(rng.choice([0, 2, 30, 41, 93, 127, 128, 200, 255]) for _ in range(L))
        if quals[0] == 255:
            quals = bytes([30]) + quals[1:]
    else:
        quals = bytes(rng.randint(0, 41) for _ in range(L))
    paired = rng.random() < 0.7
    proper = rng.random() < 0.85
    tlen = rng.choice([0, 0, rng.randint(1, 400), -rng.randint(1, 400)]) if paired else rng.choice([0, 0, 123])
    unmapped = rng.random() < 0.02
    return po.Read(pos=pos, cigar=cigar, bases=bytes(bases), quals=quals,
                   mapq=rng.choice([0, 1, 3, 17, 30, 60, 60, 60, 60, 255]), paired=paired, proper=proper,
                   mate_same_ref=rng.random() < 0.95, tlen=tlen, unmapped=unmapped, reverse=rng.random() < 0.5)


def random_case(seed: int, contig_len: int = 400, n_reads: int = 120, start: Optional[int] = None,
                stop: Optional[int] = None, deep_site: bool = True):
    """(contig, start, stop, [reads sorted by pos]).  With deep_site, a cluster of reads carrying the
    same indel is planted so that indel calls (and the pass-1 deletion spill) actually fire."""
    rng = random.Random(seed)
    contig = random_contig(rng, contig_len)
    if start is None:
        start = 1 if rng.random() < 0.5 else rng.randint(2, max(2, contig_len // 4))
    if stop is None:
        stop = contig_len if rng.random() < 0.5 else rng.randint(min(contig_len, start + 20), contig_len)
    reads = [random_read(rng, contig, max(1, start - 40), min(contig_len, stop + 20)) for _ in range(n_reads)]
    if deep_site:
        reads += planted_indel_cluster(rng, contig, start, stop)
        reads += planted_snp_cluster(rng, contig, start, stop)
    reads.sort(key=lambda r: r.pos)
    return contig, start, stop, reads


def planted_indel_cluster(rng: random.Random, contig: bytes, start: int, stop: int,
                          depth_range=(4, 14), fracs=(1.0, 0.9, 0.5, 0.3)) -> List[po.Read]:
    out: List[po.Read] = []
    n = len(contig)
    for _ in range(rng.randint(1, 3)):
        L = 50
        if stop - start < 2 * L + 10:
            break
        site = rng.randint(start + L // 2 + 12, stop - L // 2 - 12)
        kind = rng.choice("ID")
        k = rng.randint(1, 5)
        insb = bytes(rng.choice(b"ACGT") for _ in range(k))
        depth = rng.randint(*depth_range)
        frac = rng.choice(fracs)
        for d in range(depth):
            pos = site - rng.randint(12, L - 14)
            left = site - pos
            if pos < 1:
                continue
            if rng.random() < frac:
                if kind == "I":
                    right = L - left - k
                    if right <= 0 or site + right - 1 > n:
                        continue
                    bases = contig[pos - 1:site - 1].upper() + insb + contig[site - 1:site - 1 + right].upper()
                    cigar = [("M", left), ("I", k), ("M", right)]
                else:
                    right = L - left
                    if site + k + right - 1 > n:
                        continue
                    bases = contig[pos - 1:site - 1].upper() + contig[site - 1 + k:site - 1 + k + right].upper()
                    cigar = [("M", left), ("D", k), ("M", right)]
            else:
                if pos + L - 1 > n:
                    continue
                bases = contig[pos - 1:pos - 1 + L].upper()
                cigar = [("M", L)]
            out.append(po.Read(pos=pos, cigar=cigar, bases=bytes(bases), quals=bytes([rng.randint(20, 40)]) * len(bases),
                               mapq=60, paired=False))
    return out


def planted_snp_cluster(rng: random.Random, contig: bytes, start: int, stop: int,
                        depth_range=(8, 16)) -> List[po.Read]:
    """Depth at one site with two alleles: hom-alt (SNP), ref/alt het, or alt1/alt2 het (AMB)."""
    out: List[po.Read] = []
    n = len(contig)
    L = 40
    if stop - start < 2 * L:
        return out
    site = rng.randint(start + 15, stop - 15)
    ref = contig[site - 1:site].upper()
    alts = [b for b in b"ACGT" if b != ref[0]]
    rng.shuffle(alts)
    mode = rng.choice(["hom", "het_ref", "het_alt"])
    for d in range(rng.randint(*depth_range)):
        pos = site - rng.randint(11, L - 12)
        if pos < 1 or pos + L - 1 > n:
            continue
        b = bytearray(contig[pos - 1:pos - 1 + L].upper())
        if mode == "hom":
            b[site - pos] = alts[0]
        elif mode == "het_ref":
            if d % 2:
                b[site - pos] = alts[0]
        else:
            b[site - pos] = alts[d % 2]
        out.append(po.Read(pos=pos, cigar=[("M", L)], bases=bytes(b), quals=bytes([rng.randint(25, 40)]) * L,
                           mapq=rng.choice([40, 60]), paired=False))
    return out


def split_batches(reads: Sequence[po.Read], rng: random.Random):
    """Split into 1..3 'BAMs' (each sorted), some not counting toward fragCoverage."""
    k = rng.randint(1, 3)
    groups: List[List[po.Read]] = [[] for _ in range(k)]
    for r in reads:
        groups[rng.randrange(k)].append(r)
    return [(g, rng.random() < 0.7) for g in groups]


def clean_case(seed: int, n: int = 30000, start: int = 2001, stop: int = 26000, depth: int = 12, n_sites: int = 40):
    """Paired, mostly-perfect reads over a long region (halo reads on both sides) plus planted
    variant clusters deep enough to be called: exercises SNP / AMB / INS / DEL / deleted-spill."""
    rng = random.Random(seed)
    contig = random_contig(rng, n, lower_frac=0.005, n_runs=3)
    L = 100
    reads: List[po.Read] = []
    nfrag = depth * (stop - start + 1200) // (2 * L)
    for _ in range(nfrag):
        ins = max(2 * L // 2 + 20, int(rng.gauss(300, 30)))
        p1 = rng.randint(max(1, start - 600), min(n - ins, stop + 300))
        p2 = p1 + ins - L
        for pos, tl, rev in ((p1, ins, False), (p2, -ins, True)):
            b = bytearray(contig[pos - 1:pos - 1 + L].upper())
            for i in range(L):
                if rng.random() < 0.003:
                    b[i] = rng.choice(b"ACGTN")
            reads.append(po.Read(pos=pos, cigar=[("M", L)], bases=bytes(b), quals=bytes(rng.randint(15, 40) for _ in range(L)),
                                 mapq=rng.choice([60, 60, 60, 60, 30, 0]), paired=True, proper=rng.random() < 0.98,
                                 tlen=tl, reverse=rev))
    for _ in range(n_sites):
        reads += planted_indel_cluster(rng, contig, start, stop, depth_range=(2 * depth, 3 * depth), fracs=(1.0, 0.95, 0.5))
        reads += planted_snp_cluster(rng, contig, start, stop, depth_range=(2 * depth, 3 * depth))
    reads.sort(key=lambda r: r.pos)
    return contig, start, stop, reads


def unpack_batch(rb: ReadBatch) -> List[po.Read]:
    """The records a packed batch stands for, as the literal oracle wants them (inverse of pack_records)."""
    out: List[po.Read] = []
    exc = {int(i): (int(b), int(q)) for i, b, q in zip(rb.exc_idx, rb.exc_base, rb.exc_qual)}
    for r in range(rb.n_reads):
        L, s0, fl = int(rb.read_len[r]), int(rb.seq_off[r]), int(rb.flags[r])
        bases, quals = bytearray(L), bytearray(L)
        for j in range(L):
            i = s0 + j
            q = int(rb.quals[i])
            if q & 0x80:
                bases[j], quals[j] = exc[i]
            else:
                bases[j] = b"ACGT"[(int(rb.bases2[i >> 2]) >> (2 * (i & 3))) & 3]
                quals[j] = q
        cig = [(capi.CIGAR_OPS[int(e) & 15], int(e) >> 4) for e in rb.cigar[int(rb.cigar_off[r]):int(rb.cigar_off[r + 1])]]
        out.append(po.Read(pos=int(rb.pos[r]), cigar=cig, bases=bytes(bases), quals=bytes(quals) if fl & capi.PB_F_HAS_QUALS else b"",
                           mapq=int(rb.mapq[r]), paired=bool(fl & capi.PB_F_PAIRED), proper=bool(fl & capi.PB_F_PROPER),
                           mate_same_ref=bool(fl & capi.PB_F_MATE_SAME_REF), tlen=int(rb.tlen[r]),
                           unmapped=bool(fl & capi.PB_F_UNMAPPED), reverse=bool(fl & capi.PB_F_REVERSE)))
    return out
