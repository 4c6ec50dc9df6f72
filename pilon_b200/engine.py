"""Host-side mirror of the reference interface for the pileup path, over the C ABI.

`Engine` is a thin binding of include/pilon_b200.h.  `PileUpRegion`, `PileUp` and `BaseCall` mirror
the public surface of the reference classes (PileUpRegion.scala, PileUp.scala) that the Scala driver
and writers use: same member names, argument meaning and return values, but served
from the arrays the GPU engine produced.  The integer helpers below are the host copy of
Utils.scala:22-27 needed to serve the derived per-locus statistics (PileUp.scala:56-72,122-123).

There is no CPU implementation behind these classes: without libpilonb200.so and a CUDA device
they raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from .packing import ReadBatch, ResultBuffers, pack_records


class EngineConfig:
    """The `object Pilon` vars the path reads (Pilon.scala:28-73), snapshotted at engine creation."""

    def __init__(self, minQual=0, minMq=0, flank=10, defaultQual=10, minMinDepth=5, minDepth=0.1,
                 oldIndel=False, iupac=False, fixAmb=False):
        self.minQual, self.minMq, self.flank, self.defaultQual = minQual, minMq, flank, defaultQual
        self.minMinDepth, self.minDepth, self.oldIndel = minMinDepth, minDepth, oldIndel
        self.iupac, self.fixAmb = iupac, fixAmb

    def to_c(self) -> capi.pb_config:
        return capi.pb_config(min_qual=self.minQual, min_mq=self.minMq, flank=self.flank,
                              default_qual=self.defaultQual, min_min_depth=self.minMinDepth,
                              old_indel=int(self.oldIndel), fix_amb=int(self.iupac or self.fixAmb),
                              min_depth=self.minDepth)


class Engine:
    """One engine handle = one CUDA stream on one GPU (not thread-safe)."""

    def __init__(self, device: int = 0, config: Optional[EngineConfig] = None):
        self.lib = capi.load_library()
        self.config = config or EngineConfig()
        self._h = C.c_void_p()
        cfg = self.config.to_c()
        capi.check(self.lib.pb_create(device, C.byref(cfg), C.byref(self._h)))
        self.device = device
        self._keep: list = []

    def close(self):
        if self._h:
            self.lib.pb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def region_begin(self, contig: bytes, start: int, stop: int):
        buf = np.frombuffer(contig, np.uint8)
        self._keep = [buf]
        capi.check(self.lib.pb_region_begin(self._h, buf.ctypes.data, len(contig), start, stop))

    def add_batch(self, batch, frag: bool = True, long_read_type: int = 0):
        cb = batch.to_c() if hasattr(batch, "to_c") else batch
        self._keep.append((batch, cb))
        capi.check(self.lib.pb_region_add_batch(self._h, C.byref(cb), int(frag), long_read_type))

    def finish(self, res: ResultBuffers, insert_sizes: Optional[Sequence[np.ndarray]] = None):
        ptrs = None
        if insert_sizes is not None:
            arr = (C.c_void_p * len(insert_sizes))(*[a.ctypes.data for a in insert_sizes])
            ptrs = C.cast(arr, C.c_void_p)
        capi.check(self.lib.pb_region_finish(self._h, C.byref(res.c), ptrs))
        self._keep = []

    def compute_timed(self, iters: int = 1) -> Tuple[float, float, int]:
        tot, pil, n = C.c_float(), C.c_float(), C.c_int64()
        capi.check(self.lib.pb_region_compute_timed(self._h, iters, C.byref(tot), C.byref(pil), C.byref(n)))
        return tot.value, pil.value, n.value

    def compute(self):
        """Launch the compute pass on the engine's stream without waiting for it (see pb_region_compute)."""
        capi.check(self.lib.pb_region_compute(self._h))

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        capi.check(self.lib.pb_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def run_region(self, contig: bytes, start: int, stop: int, batches: Sequence[Tuple[ReadBatch, bool]],
                   planes: Optional[Sequence[str]] = None, indels_cap: int = 1 << 16,
                   bytes_cap: int = 1 << 20, pinned: bool = False, calls_cap: Optional[int] = None):
        """begin + add every (batch, counts_toward_frag_coverage[, longReadType]) + finish.  calls_cap: entries of the
        sparse call list (pb_region_result.calls) to make room for; default = one per locus up to 2^20."""
        self.region_begin(contig, start, stop)
        inserts = []
        for bt in batches:
            rb, frag, long_read = bt if len(bt) == 3 else (bt[0], bt[1], 0)
            self.add_batch(rb, frag, long_read)
            inserts.append(np.zeros(rb.n_reads, np.int32))
        size = stop + 1 - start
        res = ResultBuffers(size, planes, indels_cap, bytes_cap, pinned, calls_cap=min(size, 1 << 20) if calls_cap is None else calls_cap)
        self.finish(res, inserts)
        return res, inserts


# ---------------------------------------------------------------------------------------------
# Utils.scala:22-27 on the host (JVM semantics: truncating division, 32/64-bit wrap)
# ---------------------------------------------------------------------------------------------
def _i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _jdiv(n: int, d: int) -> int:
    q = abs(n) // abs(d)
    return q if (n >= 0) == (d >= 0) else -q


def roundDiv(n: int, d: int) -> int:
    return _jdiv(n + _jdiv(d, 2), d) if d > 0 else 0


def _roundDivI(n: int, d: int) -> int:
    return _i32(_jdiv(_i32(n + _jdiv(d, 2)), d)) if d > 0 else 0


def _pctI(n: int, d: int) -> int:
    return _roundDivI(_i32(100 * n), d)


class BaseSumView:
    """BaseSum.scala: `sums`, `sum`, `toString`, `toStringPct`."""

    def __init__(self, sums):
        self.sums = [int(x) for x in sums]

    @property
    def sum(self) -> int:
        return sum(self.sums)

    def toStringPct(self) -> str:  # BaseSum.scala:68-71
        div = self.sum
        return ",".join(str(0 if div == 0 else _jdiv(100 * x + _jdiv(div, 2), div)) for x in self.sums)

    def __str__(self) -> str:
        return ",".join(str(x) for x in self.sums)


def _call_record(res, i: int) -> int:
    """The packed call record of locus index i: from the `call` plane, or -- when only the sparse form was downloaded
    (pb_region_result.calls: changed / ambiguous loci, the ones GenomeRegion.scala:321-351 looks at) -- from there."""
    if "call" in res.arrays:
        return int(res["call"][i])
    ent = res.calls()
    k = int(np.searchsorted(ent["locus_index"], i))
    if k < len(ent) and int(ent["locus_index"][k]) == i:
        return int(ent["call"][k])
    raise KeyError("locus index %d has no call record in the sparse list (its call neither changes nor questions the reference); "
                   "download the `call` plane to read every locus" % i)


class BaseCall:
    """PileUp.BaseCall (PileUp.scala:132-173) decoded from the engine's packed call record."""

    def __init__(self, pu: "PileUp"):
        c = _call_record(pu._r, pu._i)
        self._pu = pu
        b = c & 7
        self.base = "ACGTN"[b]
        self.altBase = "ACGT"[(c >> 3) & 3]
        self.homo = bool((c >> 5) & 1)
        kind = (c >> 6) & 3
        self.homoIndel = bool((c >> 8) & 1)
        self.called = bool((c >> 9) & 1)
        self.highConfidence = bool((c >> 10) & 1)
        self.score = c >> 16
        self.n = pu.count
        self.indel = kind != 0
        s = pu._region._indel_string(pu._i, kind) if kind else ""
        self.insertion = s if kind == 1 else ""
        self.deletion = s if kind == 2 else ""
        qs = pu.qualSum.sums
        self.baseSum = qs[b] if b < 4 else qs[self._order0(pu)]
        self.altBaseSum = qs[(c >> 3) & 3]

    @staticmethod
    def _order0(pu: "PileUp") -> int:
        s = pu.qualSum.sums if pu.qSum > 0 else pu.baseCount.sums
        return max(range(4), key=lambda a: (s[a], -a))

    @property
    def isInsertion(self) -> bool:
        return self.insertion != ""

    @property
    def isDeletion(self) -> bool:
        return self.deletion != ""

    @property
    def q(self) -> int:
        return _jdiv(self.score, self.n) if self.n > 0 else 0

    def callString(self, indelOk: bool = True) -> str:
        if indelOk and self.isInsertion:
            return self.insertion
        if indelOk and self.isDeletion:
            return self.deletion
        return self.base


class PileUp:
    """Read-only per-locus view with the members the driver and writers use."""

    def __init__(self, region: "PileUpRegion", i: int):
        self._region, self._r, self._i = region, region.result, i

    def _v(self, name: str) -> int:
        return int(self._r[name][self._i])

    baseCount = property(lambda s: BaseSumView(s._r["base_count4"][s._i]))
    qualSum = property(lambda s: BaseSumView(s._r["qual_sum4"][s._i]))
    mqSum = property(lambda s: s._v("mq_sum"))
    qSum = property(lambda s: s._v("q_sum"))
    physCov = property(lambda s: s._v("phys_cov"))
    insertSize = property(lambda s: s._v("insert_size"))
    badPair = property(lambda s: s._v("bad_pair"))
    deletions = property(lambda s: s._v("deletions"))
    delQual = property(lambda s: s._v("del_qual"))
    insertions = property(lambda s: s._v("insertions"))
    insQual = property(lambda s: s._v("ins_qual"))
    clips = property(lambda s: s._v("clips"))

    @property
    def count(self) -> int:  # PileUp.scala:43
        return int(self._r["base_count4"][self._i].sum())

    @property
    def depth(self) -> int:  # PileUp.scala:44
        return self.count + self.deletions

    @property
    def weightedMq(self) -> int:  # :56-58
        return roundDiv(self.qualSum.sum, self.qSum)

    @property
    def weightedQual(self) -> int:  # :60-62
        return roundDiv(self.qualSum.sum, self.mqSum)

    @property
    def meanQual(self) -> int:  # :64-67
        return roundDiv(self.qualSum.sum, roundDiv(self.mqSum * self.count, self.depth))

    @property
    def meanMq(self) -> int:  # :70-72
        return roundDiv(self.mqSum - self.depth, self.depth)

    @property
    def insPct(self) -> int:  # :122
        return max(_pctI(self.insQual, self.mqSum), _pctI(self.insertions, _i32(self.count)))

    @property
    def delPct(self) -> int:  # :123
        return max(_pctI(self.delQual, self.mqSum), _pctI(self.deletions, _i32(_i32(self.count) + self.deletions)))

    def baseCall(self) -> BaseCall:  # :257
        return BaseCall(self)


class PileUpRegion:
    """Drop-in for `new PileUpRegion(name, start, stop)` (PileUpRegion.scala:26-36).

    addRead buffers records (the reference adds them one at a time, PileUpRegion.scala:102-220);
    postProcess ships them to the GPU, runs the whole path and makes `apply(i)` servable.
    """

    def __init__(self, name: str, start: int, stop: int, contigBases: bytes,
                 config: Optional[EngineConfig] = None, device: int = 0, engine: Optional[Engine] = None):
        self.name, self.start, self.stop = name, start, stop
        self.size = stop + 1 - start
        self.contigBases = contigBases
        self.engine = engine or Engine(device, config)
        self.config = self.engine.config
        self._pending: List = []
        self._pending_frag = True
        self._batches: List[Tuple[ReadBatch, bool]] = []
        self.result: Optional[ResultBuffers] = None
        self.insertSizes: List[np.ndarray] = []
        self._indel_map: Dict[Tuple[int, int], dict] = {}

    # Region.scala:23-28
    def inRegion(self, locus: int) -> bool:
        return self.start <= locus <= self.stop

    def index(self, locus: int) -> int:
        return locus - self.start

    def locus(self, index: int) -> int:
        return self.start + index

    def addRead(self, r, refBases=None, longRead: int = 0) -> int:
        """Buffers the record; returns what the reference returns (physCovIncr, PileUpRegion.scala:62-88)."""
        self._pending_long = longRead                            # BamFile.longReadType: one value per BAM
        self._pending.append(r)
        valid = (r.mapq >= self.config.minMq) and ((not r.paired) or (r.proper and r.mate_same_ref))
        if (not valid) or (r.paired and r.tlen <= 0):
            return 0
        if not r.paired:
            ref_len = sum(l for op, l in r.cigar if op in "MDN=X")
            aEnd = 0 if getattr(r, "unmapped", False) else r.pos + ref_len - 1
            return _i32(max(r.pos, aEnd) - min(r.pos, aEnd))
        return _i32(r.tlen)

    def endBam(self, countsTowardFragCoverage: bool = True):
        """Marks the end of one BAM's reads (GenomeRegion.processBam, GenomeRegion.scala:287-300)."""
        if self._pending:
            self._batches.append((pack_records(self._pending), countsTowardFragCoverage, getattr(self, "_pending_long", 0)))
            self._pending = []

    def addBatch(self, batch: ReadBatch, countsTowardFragCoverage: bool = True):
        self.endBam()
        self._batches.append((batch, countsTowardFragCoverage))

    def postProcess(self, planes: Optional[Sequence[str]] = None):
        """PileUpRegion.postProcess + GenomeRegion.postProcess pass 1 on the GPU."""
        self.endBam()
        n_ops = sum(int(b[0].cigar.shape[0]) for b in self._batches)
        self.result, self.insertSizes = self.engine.run_region(
            self.contigBases, self.start, self.stop, self._batches, planes,
            indels_cap=max(16, n_ops), bytes_cap=max(1024, sum(int(b[0].quals.shape[0]) for b in self._batches)))
        self._indel_map = {(e["locus_index"], e["kind"]): e for e in self.result.indels()}
        self._batches = []

    def _indel_string(self, i: int, kind: int) -> str:
        return self._indel_map[(i, kind)]["string"].decode("latin1")

    @property
    def readCount(self) -> int:
        return int(self.result.c.read_count)

    @property
    def baseCount(self) -> int:
        return int(self.result.c.base_count)

    @property
    def coverage(self) -> int:  # PileUpRegion.scala:36
        return int(self.result.c.coverage)

    @property
    def minDepth(self) -> int:  # GenomeRegion.scala:221-224
        return int(self.result.c.min_depth)

    def __getitem__(self, i: int) -> PileUp:  # apply(i), PileUpRegion.scala:233
        if i < 0 or i >= self.size:
            raise IndexError(i)
        return PileUp(self, i)

    def changes(self) -> List[Tuple[int, int]]:
        """(locus index, PB_KIND_*) of every pass-1 change, ascending: GenomeRegion.changeList (:90)."""
        fl = self.result["flags"]
        idx = np.nonzero(fl & (capi.PB_FL_CHANGED | capi.PB_FL_AMBIGUOUS))[0]
        return [(int(i), int(fl[i] >> capi.PB_FL_KIND_SHIFT) & 3) for i in idx]
