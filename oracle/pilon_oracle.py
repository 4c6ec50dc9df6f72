"""
ORACLE (test infrastructure, NOT product code) -- literal CPU restatement of Pilon's
pileup + BaseCall hot path.

    *** parity unpinned ***
    The reference ships no tests, golden vectors or fixtures for this path and cannot be
    executed in the build container (no JVM).  This file is a line-by-line restatement of
    the Scala; it is pinned only by the hand-derived known-answer tests of SURVEY.md 8(c)
    (tests/test_oracle_kat.py) and cross-checked against an independent C restatement
    (oracle/pilon_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product path (pilon_b200/) never does.

All file:line citations are relative to
/root/reference/src/main/scala/org/broadinstitute/pilon/ .

JVM semantics reproduced on purpose:
  * `Int` fields wrap at 32 bit, `Long` at 64 bit; integer division truncates toward zero.
  * `Byte` is signed: quality bytes >= 128 are negative Ints.
  * Array indexing outside [0, size) throws (we raise IndexError) -- never Python's
    negative-index wrap-around.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


# ----------------------------------------------------------------------------------------
# JVM integer helpers
# ----------------------------------------------------------------------------------------
def i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def i64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - 0x10000000000000000 if x & 0x8000000000000000 else x


def jdiv(n: int, d: int) -> int:
    """JVM integer division (truncates toward zero)."""
    q = abs(n) // abs(d)
    return q if (n >= 0) == (d >= 0) else -q


def sbyte(b: int) -> int:
    """JVM Byte -> Int widening."""
    b &= 0xFF
    return b - 256 if b & 0x80 else b


# Utils.scala:22-27
def roundDivL(n: int, d: int) -> int:  # Long overload, Utils.scala:23
    return i64(jdiv(i64(n + jdiv(d, 2)), d)) if d > 0 else 0


def roundDivI(n: int, d: int) -> int:  # Int overload, Utils.scala:24
    return i32(jdiv(i32(n + jdiv(d, 2)), d)) if d > 0 else 0


def pctL(n: int, d: int) -> int:  # Utils.scala:25
    return roundDivL(i64(100 * n), d)


def pctI(n: int, d: int) -> int:  # Utils.scala:26
    return roundDivI(i32(100 * n), d)


# ----------------------------------------------------------------------------------------
# Global configuration (object Pilon vars read on the hot path; Pilon.scala:28-73)
# ----------------------------------------------------------------------------------------
@dataclass
class Config:
    minQual: int = 0          # Pilon.scala:65
    minMq: int = 0            # Pilon.scala:66
    flank: int = 10           # Pilon.scala:59
    defaultQual: int = 10     # Pilon.scala:53 (Byte)
    minMinDepth: int = 5      # Pilon.scala:62
    minDepth: float = 0.1     # Pilon.scala:64
    oldIndel: bool = False    # Pilon.scala:68
    iupac: bool = False       # Pilon.scala:61
    fixAmb: bool = False      # Pilon.scala:32


NANOPORE_LONG_READ = 1  # BamFile.nanoporeLongRead (BamFile.scala, companion object)
PACBIO_LONG_READ = 2


# ----------------------------------------------------------------------------------------
# The slice of htsjdk.samtools.SAMRecord the path reads (htsjdk 2.23.0, un-vendored;
# semantics restated from the SAM spec / htsjdk API).
# ----------------------------------------------------------------------------------------
CONSUMES_READ = set("MIS=X")   # CigarOperator.consumesReadBases
CONSUMES_REF = set("MDN=X")    # CigarOperator.consumesReferenceBases


@dataclass
class Read:
    pos: int                              # getAlignmentStart (1-based)
    cigar: List[Tuple[str, int]]          # [(op, len)], op in "MIDNSHP=X"
    bases: bytes                          # getReadBases: ASCII, upper-case as htsjdk decodes
    quals: bytes = b""                    # getBaseQualities: empty when BAM stores 0xFF..
    mapq: int = 60                        # getMappingQuality
    paired: bool = False                  # getReadPairedFlag
    proper: bool = False                  # getProperPairFlag
    mate_same_ref: bool = True            # getReferenceIndex == getMateReferenceIndex
    tlen: int = 0                         # getInferredInsertSize
    unmapped: bool = False                # getReadUnmappedFlag
    reverse: bool = False                 # getReadNegativeStrandFlag (BamFile.scala:137 only)

    @property
    def read_length(self) -> int:
        return len(self.bases)

    @property
    def alignment_end(self) -> int:
        # SAMRecord.getAlignmentEnd: 0 for unmapped reads, else start + refLength - 1
        if self.unmapped:
            return 0
        return self.pos + sum(l for op, l in self.cigar if op in CONSUMES_REF) - 1


# ----------------------------------------------------------------------------------------
# BaseSum.scala
# ----------------------------------------------------------------------------------------
class BaseSum:
    __slots__ = ("sums",)

    def __init__(self):
        self.sums = [0, 0, 0, 0]  # Array[Long](4), BaseSum.scala:24

    def add(self, base: int, n: int = 1):  # BaseSum.scala:26
        self.sums[base] = i64(self.sums[base] + n)

    @property
    def sum(self) -> int:  # BaseSum.scala:30
        return i64(sum(self.sums))

    def order(self) -> List[int]:  # BaseSum.scala:57-60: stable sortWith(sums(a) > sums(b))
        return sorted(range(4), key=lambda a: -self.sums[a])  # Python sort is stable

    def toStringPct(self) -> str:  # BaseSum.scala:68-71
        div = self.sum
        return ",".join(str(0 if div == 0 else jdiv(100 * x + jdiv(div, 2), div)) for x in self.sums)

    def __str__(self) -> str:  # BaseSum.scala:73
        return ",".join(str(x) for x in self.sums)


def baseIndex(c: int) -> int:  # PileUp.scala:46-52
    return {65: 0, 67: 1, 71: 2, 84: 3}.get(c, -1)


# ----------------------------------------------------------------------------------------
# PileUp.scala
# ----------------------------------------------------------------------------------------
class PileUp:
    __slots__ = ("cfg", "baseCount", "qualSum", "mqSum", "qSum", "physCov", "insertSize", "badPair",
                 "deletions", "delQual", "insertions", "insQual", "clips", "insertionList", "deletionList")

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.baseCount = BaseSum()   # PileUp.scala:26
        self.qualSum = BaseSum()     # :27
        self.mqSum = 0               # :30  (Int)
        self.qSum = 0                # :31
        self.physCov = 0             # :32
        self.insertSize = 0          # :33
        self.badPair = 0             # :34
        self.deletions = 0           # :35
        self.delQual = 0             # :36
        self.insertions = 0          # :37
        self.insQual = 0             # :38
        self.clips = 0               # :39
        self.insertionList: List[bytes] = []  # :40 (prepend order irrelevant to every consumer)
        self.deletionList: List[bytes] = []   # :41

    @property
    def count(self) -> int:  # :43
        return self.baseCount.sum

    @property
    def depth(self) -> int:  # :44
        return i64(self.baseCount.sum + self.deletions)

    @property
    def weightedMq(self) -> int:  # :56-58
        return roundDivL(self.qualSum.sum, self.qSum)

    @property
    def weightedQual(self) -> int:  # :60-62
        return roundDivL(self.qualSum.sum, self.mqSum)

    @property
    def meanQual(self) -> int:  # :64-67
        return roundDivL(self.qualSum.sum, roundDivL(i64(self.mqSum * self.count), self.depth))

    @property
    def meanMq(self) -> int:  # :70-72
        return roundDivL(i64(self.mqSum - self.depth), self.depth)

    def add(self, base: int, qual: int, mq: int):  # :75-84
        bi = baseIndex(base)
        if bi >= 0 and qual >= self.cfg.minQual:
            mq1 = i32(mq + 1)
            self.baseCount.add(bi)
            self.qualSum.add(bi, i32(qual * mq1))
            self.mqSum = i32(self.mqSum + mq1)
            self.qSum = i32(self.qSum + qual)

    def addInsertion(self, insertion: bytes, qual: int, mq: int):  # :98-105
        mq1 = i32(mq + 1)
        self.insQual = i32(self.insQual + mq1)
        self.qSum = i32(self.qSum + qual)
        self.insertionList.insert(0, insertion)
        self.insertions = i32(self.insertions + 1)

    def addDeletion(self, deletion: bytes, qual: int, mq: int):  # :107-114
        mq1 = i32(mq + 1)
        self.mqSum = i32(self.mqSum + mq1)
        self.delQual = i32(self.delQual + mq1)
        self.qSum = i32(self.qSum + qual)
        self.deletionList.insert(0, deletion)
        self.deletions = i32(self.deletions + 1)

    @property
    def insPct(self) -> int:  # :122  (Int overloads; count.toInt)
        return max(pctI(self.insQual, self.mqSum), pctI(self.insertions, i32(self.count)))

    @property
    def delPct(self) -> int:  # :123
        return max(pctI(self.delQual, self.mqSum),
                   pctI(self.deletions, i32(i32(self.count) + self.deletions)))

    def baseCall(self) -> "BaseCall":  # :257
        return BaseCall(self)

    def __str__(self) -> str:  # :260-264
        return ("<PileUp " + str(BaseCall(self)) + ",b=" + str(self.baseCount) + "/" + self.qualSum.toStringPct()
                + ",c=" + str(self.depth) + "/" + str(self.depth + self.badPair)
                + ",i=" + str(self.insertions) + ",d=" + str(self.deletions) + ",q=" + str(self.weightedQual)
                + ",mq=" + str(self.weightedMq) + ",p=" + str(self.physCov) + ",s=" + str(self.insertSize)
                + ",x=" + str(self.clips) + ">")


class BaseCall:  # PileUp.scala:132-255
    def __init__(self, pu: PileUp):
        self.pu = pu
        self.n = pu.count                                             # :133
        order = pu.qualSum.order() if pu.qSum > 0 else pu.baseCount.order()  # :135
        self.baseIndex, self.altBaseIndex = order[0], order[1]        # :136
        self.base = "ACGT"[self.baseIndex] if self.n > 0 else "N"     # :138
        self.baseSum = pu.qualSum.sums[self.baseIndex]                # :139
        self.altBase = "ACGT"[self.altBaseIndex]                      # :140
        self.altBaseSum = pu.qualSum.sums[self.altBaseIndex]          # :141
        total = pu.qualSum.sum                                        # :143
        homoScore = i64(self.baseSum - (total - self.baseSum))        # :144
        halfTotal = jdiv(total, 2)                                    # :145
        heteroScore = i64(total - abs(halfTotal - self.baseSum) - abs(halfTotal - self.altBaseSum))  # :146
        self.homo = homoScore >= heteroScore                          # :147
        self.score = (jdiv(i64(abs(homoScore - heteroScore) * self.n), pu.mqSum)
                      if pu.mqSum > 0 else 0)                         # :148
        ins, homoIns = self.insertCall()                              # :151-162
        if ins != "":
            self.insertion, self.deletion, self.indel, self.homoIndel = ins, "", True, homoIns
        else:
            dele, homoDel = self.deletionCall()
            if dele != "":
                self.insertion, self.deletion, self.indel, self.homoIndel = "", dele, True, homoDel
            else:
                self.insertion, self.deletion, self.indel, self.homoIndel = "", "", False, True

    @property
    def isInsertion(self) -> bool:  # :163
        return self.insertion != ""

    @property
    def isDeletion(self) -> bool:  # :164
        return self.deletion != ""

    @property
    def called(self) -> bool:  # :165
        return self.base != "N" or self.indel

    @property
    def q(self) -> int:  # :166
        return jdiv(self.score, self.n) if self.n > 0 else 0

    @property
    def highConfidence(self) -> bool:  # :167
        return self.q >= 10

    def callString(self, indelOk: bool = True) -> str:  # :169-173
        if indelOk and self.isInsertion:
            return self.insertion
        if indelOk and self.isDeletion:
            return self.deletion
        return self.base

    def insertCall(self):  # :183-186
        pu = self.pu
        if pu.insertions > 2 and pu.insertions > pu.deletions:
            return self.hetIndelCall(pu.insertionList, pu.insPct)
        return ("", True)

    def deletionCall(self):  # :188-191
        pu = self.pu
        if pu.deletions > 2 and pu.deletions > pu.insertions:
            return self.hetIndelCall(pu.deletionList, pu.delPct)
        return ("", True)

    def hetIndelCall(self, indelList: List[bytes], pct: int):  # :209-247
        pu = self.pu
        cfg = pu.cfg
        if pu.depth < cfg.minMinDepth or pct < 5 or not indelList:  # :213
            return ("", True)
        m: Dict[str, int] = {}
        for indel in indelList:                                      # :215-218
            s = "".join(chr(b & 0xFF) if b < 128 else chr(0xFF00 | b) for b in indel)
            m[s] = m.get(s, 0) + 1
        # :219  map.toSeq.sortBy(_._2).last -- any maximal-count entry; the strict-majority test
        # below makes the choice among ties irrelevant.
        winStr, winCount = max(m.items(), key=lambda kv: kv[1])
        if winCount < 2 or winCount <= jdiv(len(indelList), 2):      # :220
            return ("", True)
        if "N" in winStr:                                            # :222
            return ("", True)
        if cfg.oldIndel:                                             # :223-228
            if pct >= 33 and pct >= 50 - len(winStr):
                return (winStr, True)
            return ("", True)
        middle = max(45 - len(winStr), 10)                           # :232
        low = jdiv(middle, 2)                                        # :234
        high = middle + middle - low                                 # :236
        if pct > high:                                               # :238
            return (winStr, True)
        if pct >= low:                                               # :241
            return (winStr, False)
        return ("", True)                                            # :245

    def __str__(self) -> str:  # :249-254
        pu = self.pu
        if self.isInsertion:
            return "bc=i" + str(self.insertCall()) + ",cq=" + str(jdiv(pu.insQual, pu.insertions))
        if self.isDeletion:
            return "bc=d" + str(self.deletionCall()) + ",cq=" + str(jdiv(pu.delQual, pu.deletions))
        return "bc=" + self.base + ("" if self.homo else "/" + self.altBase) + ",cq=" + str(self.q)


# ----------------------------------------------------------------------------------------
# Region.scala:22-28 + PileUpRegion.scala
# ----------------------------------------------------------------------------------------
class PileUpRegion:
    def __init__(self, name: str, start: int, stop: int, cfg: Optional[Config] = None, oob_drop: bool = False):
        # oob_drop=False is the literal JVM behaviour (IndexError where the JVM throws
        # ArrayIndexOutOfBoundsException).  oob_drop=True is the engine's DEFINED behaviour for that
        # latent crash (PileUpRegion.scala:156-161,167-182 with region.start > 1): an indel whose
        # left shift leaves the region contributes nothing and is counted in dropped_oob.
        self.oob_drop = oob_drop
        self.dropped_oob = 0
        self.cfg = cfg or Config()
        self.name, self.start, self.stop = name, start, stop
        self.size = stop + 1 - start                              # Region.scala:27
        self.pileups = [PileUp(self.cfg) for _ in range(self.size)]  # PileUpRegion.scala:29-30
        self.baseCount = 0                                        # :32 (Long)
        self.readCount = 0                                        # :33
        self.trustedFlank = self.cfg.flank                        # :34
        self.physCovStart = 0                                     # :59
        self.insertSizeStart = 0                                  # :60
        self.unknown_ops = 0                                      # count of the println at :212

    # Region.scala:23-26
    def inRegion(self, locus: int) -> bool:
        return self.start <= locus <= self.stop

    def beforeRegion(self, locus: int) -> bool:
        return locus < self.start

    def index(self, locus: int) -> int:
        return locus - self.start

    def locus(self, index: int) -> int:
        return self.start + index

    def _pu(self, idx: int) -> PileUp:
        if idx < 0 or idx >= self.size:
            raise IndexError("ArrayIndexOutOfBoundsException: %d" % idx)  # JVM array semantics
        return self.pileups[idx]

    @property
    def coverage(self) -> int:  # :36
        return roundDivL(self.baseCount, self.size)

    def add(self, locus: int, base: int, qual: int, mq: int, pair: bool):  # :38-48
        if self.inRegion(locus):
            if pair:
                self._pu(self.index(locus)).add(base, qual, mq)
                self.baseCount = i64(self.baseCount + 1)
            else:
                pu = self._pu(self.index(locus))
                pu.badPair = i32(pu.badPair + 1)

    def remove(self, locus: int, base: int, qual: int, mq: int, pair: bool):  # :50-58
        if self.inRegion(locus):
            if pair:
                pass
            else:
                pu = self._pu(self.index(locus))
                pu.badPair = i32(pu.badPair - 1)

    def physCovIncr(self, aStart: int, aEnd: int, iSize: int, paired: bool, valid: bool) -> int:  # :62-88
        if (not valid) or (paired and iSize <= 0):
            return 0
        if not paired:
            start, end = min(aStart, aEnd), max(aStart, aEnd)
        elif iSize > 0:
            start, end = aStart, aStart + iSize
        else:  # unreachable, kept literal
            start, end = aEnd + 1 + iSize, aEnd + 1
        insertSize = i32(end - start)
        if self.inRegion(start):
            pu = self._pu(self.index(start))
            pu.physCov = i32(pu.physCov + 1)
            pu.insertSize = i32(pu.insertSize + insertSize)
        elif self.beforeRegion(start) and not self.beforeRegion(end):
            self.physCovStart = i32(self.physCovStart + 1)
            self.insertSizeStart = i32(self.insertSizeStart + insertSize)
        if self.inRegion(end):
            pu = self._pu(self.index(end))
            pu.physCov = i32(pu.physCov - 1)
            pu.insertSize = i32(pu.insertSize - insertSize)
        return insertSize

    def computePhysCov(self):  # :90-100
        p = self.pileups
        p[0].physCov = i32(p[0].physCov + self.physCovStart)
        p[0].insertSize = i32(p[0].insertSize + self.insertSizeStart)
        for i in range(1, len(p)):
            p[i].physCov = i32(p[i].physCov + p[i - 1].physCov)
            p[i].insertSize = i32(p[i].insertSize + p[i - 1].insertSize)
        for i in range(len(p)):
            if p[i].physCov > 0:
                p[i].insertSize = jdiv(p[i].insertSize, p[i].physCov)

    def addRead(self, r: Read, refBases: bytes, longRead: int = 0) -> int:  # :102-220
        cfg = self.cfg
        length = r.read_length                                   # :103
        bases = r.bases                                          # :104
        mq = r.mapq                                              # :105
        paired = r.paired                                        # :106
        valid = (mq >= cfg.minMq) and ((not paired) or (r.proper and r.mate_same_ref))  # :107
        insert = r.tlen                                          # :108
        aStart = r.pos                                           # :109
        aEnd = r.alignment_end                                   # :110
        readOffset = 0
        refOffset = 0
        quals = r.quals if len(r.quals) > 0 else bytes([cfg.defaultQual & 0xFF]) * length  # :114-115
        trustedFlank = self.trustedFlank

        def trusted(offset: int) -> bool:                        # :118
            return offset >= trustedFlank and length - trustedFlank > offset

        def homoRun(i0: int) -> int:                             # :120-126
            baseAtLoc = refBases[i0]
            for i in range(i0 + 1, len(refBases)):
                if refBases[i] != baseAtLoc:
                    return i - i0
            return len(refBases) - i0

        def nanoporeExclude(i0: int) -> bool:                    # :128-134
            return (self.inRegion(self.locus(i0 - 2)) and self.inRegion(self.locus(i0 + 2))
                    and refBases[i0 - 2] == 67 and refBases[i0 - 1] == 67
                    and refBases[i0 + 1] == 71 and refBases[i0 + 2] == 71)

        clippedBases = sum(l for op, l in r.cigar if op == "S")  # :139
        adjMq = roundDivI(i32(mq * (length - clippedBases)), length)  # :141
        indelMq = min(adjMq, 8) if longRead > 0 else adjMq       # :142

        for op, ln in r.cigar:                                   # :145
            locus = aStart + refOffset                           # :148
            if op == "I":                                        # :150-162
                insertion = bytes(bases[readOffset:readOffset + ln])
                iloc = locus
                if valid and trusted(readOffset) and self.inRegion(iloc):
                    if self.oob_drop and self._ins_shift_end(refBases, insertion, iloc) < self.start:
                        self.dropped_oob += 1
                        readOffset += ln
                        continue
                    while iloc > 1 and refBases[iloc - 2] == insertion[ln - 1]:
                        iloc -= 1
                        insertion = insertion[ln - 1:ln] + insertion[0:ln - 1]
                    if not (longRead > 0 and homoRun(iloc) >= 4):
                        self._pu(self.index(iloc)).addInsertion(insertion, sbyte(quals[readOffset]), indelMq)
            elif op == "D":                                      # :163-183
                dloc = locus
                rloc = readOffset
                if valid and trusted(readOffset) and self.inRegion(dloc) and self.inRegion(dloc + ln - 1):
                    if self.oob_drop and self._del_shift_end(refBases, dloc, rloc, ln) < self.start:
                        self.dropped_oob += 1
                        refOffset += ln
                        continue
                    while dloc > 1 and rloc > 0 and refBases[dloc - 2] == refBases[dloc + ln - 2]:
                        dloc -= 1
                        rloc -= 1
                        base = bases[rloc]
                        qual = sbyte(quals[rloc])
                        if trusted(rloc) and self.inRegion(dloc):
                            self.remove(dloc, base, qual, adjMq, valid)
                            if self.inRegion(dloc + ln):
                                self.add(dloc + ln, base, qual, adjMq, valid)
                    if not (longRead > 0 and (homoRun(self.index(dloc)) >= 4
                                              or (longRead == NANOPORE_LONG_READ
                                                  and nanoporeExclude(self.index(dloc))))):
                        self._pu(self.index(dloc)).addDeletion(
                            bytes(refBases[dloc - 1:dloc + ln - 1]), sbyte(quals[readOffset]), indelMq)
            elif op in ("M", "=", "X"):                          # :184-193
                for i in range(ln):
                    rOff = readOffset + i
                    if trusted(rOff):
                        locusPlus = locus + i
                        base = bases[rOff]
                        if longRead == NANOPORE_LONG_READ and nanoporeExclude(self.index(locusPlus)):
                            qual = 0
                        else:
                            qual = sbyte(quals[rOff])
                        self.add(locusPlus, base, qual, adjMq, valid)
            elif op == "S":                                      # :194-206
                clipStart = locus - ln if readOffset == 0 else locus
                clipEnd = clipStart + ln - 1
                if self.inRegion(clipStart):
                    pu = self._pu(self.index(clipStart))
                    pu.clips = i32(pu.clips + 1)
                if self.inRegion(clipEnd):
                    pu = self._pu(self.index(clipEnd))
                    pu.clips = i32(pu.clips + 1)
                for i in range(ln):
                    rOff = readOffset + i
                    locusPlus = clipStart + i
                    if self.inRegion(locusPlus):
                        self.add(locusPlus, bases[rOff], sbyte(quals[rOff]), adjMq, False)
            elif op in ("H", "N"):                               # :207-210
                pass
            else:                                                # :211-212 (println, continue)
                self.unknown_ops += 1
            if op in CONSUMES_READ:                              # :214
                readOffset += ln
            if op in CONSUMES_REF:                               # :215
                refOffset += ln

        self.readCount += 1                                      # :218
        return self.physCovIncr(aStart, aEnd, insert, paired, valid)  # :219

    def postProcess(self):  # :226-229
        self.computePhysCov()

    @staticmethod
    def _ins_shift_end(refBases, insertion, iloc):
        ln = len(insertion)
        while iloc > 1 and refBases[iloc - 2] == insertion[ln - 1]:
            iloc -= 1
            insertion = insertion[ln - 1:ln] + insertion[0:ln - 1]
        return iloc

    @staticmethod
    def _del_shift_end(refBases, dloc, rloc, ln):
        while dloc > 1 and rloc > 0 and refBases[dloc - 2] == refBases[dloc + ln - 2]:
            dloc -= 1
            rloc -= 1
        return dloc

    def __getitem__(self, i: int) -> PileUp:  # :233
        return self._pu(i)


# ----------------------------------------------------------------------------------------
# GenomeRegion.scala: the hot part only (:214-272 pass 1, :287-300 fragCoverage)
# ----------------------------------------------------------------------------------------
SNP, INS, DEL, AMB = "SNP", "INS", "DEL", "AMB"


def validateRead(flag_qcfail: bool, flag_dup: bool, flag_secondary: bool,
                 nonPf: bool = False, duplicates: bool = False) -> bool:
    """BamFile.scala:101-105 (supplementary alignments are kept)."""
    return (nonPf or not flag_qcfail) and (duplicates or not flag_dup) and not flag_secondary


class GenomeRegionHot:
    """Just enough of GenomeRegion to drive the engine the way the Scala driver does."""

    def __init__(self, contigBases: bytes, start: int, stop: int, cfg: Optional[Config] = None,
                 name: str = "contig"):
        assert stop <= len(contigBases)                          # GenomeRegion.scala:34
        self.cfg = cfg or Config()
        self.contigBases = contigBases
        self.name, self.start, self.stop = name, start, stop
        self.size = stop + 1 - start
        n = self.size
        self.minDepth = self.cfg.minMinDepth                     # :40
        self.confirmed = [False] * n                             # :45-49
        self.ambiguous = [False] * n
        self.changed = [False] * n
        self.deleted = [False] * n
        self.badCoverage = [0] * n                               # :54-62
        self.clips = [0] * n
        self.coverage = [0] * n
        self.insertSize = [0] * n
        self.physCoverage = [0] * n
        self.fragCoverage = [0] * n
        self.weightedQual = [0] * n
        self.weightedMq = [0] * n
        self.changeMap: Dict[int, Tuple[str, PileUp]] = {}       # :82
        self.pileUpRegion: Optional[PileUpRegion] = None
        self.insert_sizes: List[Tuple[int, bool]] = []           # what BamFile.addInsert receives
        self.per_bam: List[Tuple[int, int, int]] = []            # (nReads, baseCount delta, meanCoverage) per processBam
        self.log: List[str] = []

    def refBase(self, locus: int) -> str:                        # :783-787 (upper-cased)
        assert self.start <= locus <= self.stop
        return chr(self.contigBases[locus - 1]).upper()

    def initializePileUps(self, oob_drop: bool = False):         # :149-151
        self.pileUpRegion = PileUpRegion(self.name, self.start, self.stop, self.cfg, oob_drop)

    def processBam(self, reads: Sequence[Read], bamType: str = "frags", longReadType: int = 0):
        """GenomeRegion.scala:287-300 around BamFile.process (BamFile.scala:108-148).
        `reads` is what queryOverlapping(+-10 kb) returned and validateRead kept."""
        pur = self.pileUpRegion
        readsBefore = pur.readCount                              # BamFile.scala:120-122
        baseCountBefore = pur.baseCount
        covBeforeBam = pur.coverage
        covBefore = [0] * self.size
        if bamType != "jumps":
            for i in range(self.size):
                covBefore[i] = i32(pur.pileups[i].depth)
        for rd in reads:
            insertSize = pur.addRead(rd, self.contigBases, longReadType)   # BamFile.scala:133
            self.insert_sizes.append((insertSize, rd.reverse))             # BamFile.scala:137
        if bamType != "jumps":
            for i in range(self.size):
                self.fragCoverage[i] = i32(self.fragCoverage[i] + i32(pur.pileups[i].depth) - covBefore[i])
        meanCoverage = pur.coverage - covBeforeBam               # BamFile.scala:142
        nReads = pur.readCount - readsBefore                     # :143
        self.per_bam.append((nReads, pur.baseCount - baseCountBefore, meanCoverage))   # :146 feeds BamFile.baseCount
        return meanCoverage                                      # :147

    def postProcess(self):                                       # :214-272
        cfg = self.cfg
        pur = self.pileUpRegion
        pur.postProcess()                                        # :215
        meanCoverage = pur.coverage                              # :216
        nReads = pur.readCount                                   # :217
        if cfg.minDepth >= 1:                                    # :221-224
            self.minDepth = int(cfg.minDepth)
        else:
            self.minDepth = max(jround(cfg.minDepth * meanCoverage), cfg.minMinDepth)
        self.log.append("Total Reads: %d, Coverage: %d, minDepth: %d" % (nReads, meanCoverage, self.minDepth))
        if nReads == 0:                                          # :229-231
            return
        fixamb = cfg.iupac or cfg.fixAmb                         # :235
        for i in range(self.size):                               # :237-272
            pu = pur[i]
            n = pu.depth
            bc = pu.baseCall()
            b = bc.base
            homo = bc.homo
            r = self.refBase(i + self.start)
            self.coverage[i] = i32(n)                            # :247
            self.badCoverage[i] = pu.badPair
            self.physCoverage[i] = pu.physCov
            self.insertSize[i] = pu.insertSize
            self.weightedQual[i] = sbyte(pu.weightedQual)        # :251 .toByte
            self.weightedMq[i] = sbyte(pu.weightedMq)            # :252 .toByte
            self.clips[i] = toshort(pu.clips)                    # :253 .toShort
            if n >= self.minDepth and r != "N" and not self.deleted[i] and bc.called:  # :255
                if homo and b == r and bc.highConfidence and not bc.indel:
                    self.confirmed[i] = True
                elif bc.isInsertion and bc.homoIndel:
                    self._addChange(i, INS, pu)
                elif bc.isDeletion and bc.homoIndel:
                    self._addChange(i, DEL, pu)
                    for j in range(1, len(bc.deletion)):
                        if i + j >= self.size:
                            raise IndexError("ArrayIndexOutOfBoundsException")
                        self.deleted[i + j] = True
                        pj = pur[i + j]
                        pj.deletions = i32(pj.deletions + pu.deletions)
                elif b != r and bc.score > 0:
                    if homo:
                        self._addChange(i, SNP, pu)
                    elif fixamb or bc.altBase != r:
                        self._addChange(i, AMB, pu)

    def _addChange(self, loc: int, kind: str, pu: PileUp):       # :84-88
        if kind == AMB:
            self.ambiguous[loc] = True
        else:
            self.changed[loc] = True
        self.changeMap[loc] = (kind, pu)


def jround(x: float) -> int:
    """scala Double.round == java.lang.Math.round: floor(x + 0.5)."""
    import math
    return int(math.floor(x + 0.5))


def toshort(x: int) -> int:
    x &= 0xFFFF
    return x - 0x10000 if x & 0x8000 else x
