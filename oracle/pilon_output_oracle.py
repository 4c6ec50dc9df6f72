"""TEST INFRASTRUCTURE ONLY -- literal Python transliteration of the reference's per-locus CONSUMERS of the pileup
path and of its fix application, for `--fix snps,indels [--changes] [--vcf]`.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product (pilon_b200/) never does.

Parity is UNPINNED, as for oracle/pilon_oracle.py: the reference ships no tests or vectors and cannot run here (Scala /
JVM).  Every function cites the Scala it follows, relative to
/root/reference/src/main/scala/org/broadinstitute/pilon/.

What is restated:
  GenomeRegion.postProcess pass 2 (copy number)          GenomeRegion.scala:275-283, smooth :188-208
  GenomeRegion.summaryRegions / nearEdge / duplicationEvents   :742-763, :690, :735-741
  GenomeRegion.identifyAndFixIssues (snps, indels, amb)  :307-380, 413
  GenomeRegion.fixFixList / fixIssues                    :557-621
  GenomeRegion.writeChanges / writeVcf                   :623-657
  Vcf.writeHeader / writeRecord / writeDup               Vcf.scala:28-68, 74-176, 193-201
  GenomeFile output naming, FASTA layout, changes offsets, coverageSummary   GenomeFile.scala:79-82, 135-162, 178-187
  Region.regionString                                    Region.scala:42
Not restated (out of scope, north_star): gap filling / local reassembly (bigFixList stays empty), writeFixRecord.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from decimal import ROUND_HALF_UP, Decimal
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .pilon_oracle import AMB, DEL, INS, SNP, GenomeRegionHot, i32, jdiv, pctI, roundDivL, toshort

Fix = Tuple[int, str, str]                     # GenomeRegion.Fix = (locus, ref, patch)


@dataclass
class OutConfig:
    fixSnps: bool = True          # Pilon.scala: --fix snps
    fixIndels: bool = True        # --fix indels
    iupac: bool = False           # Pilon.iupac
    diploid: bool = False         # Pilon.diploid
    vcfQE: bool = False           # Pilon.vcfQE
    longread: bool = False        # Pilon.longread


def regionString(name: str, start: int, stop: int) -> str:        # Region.scala:27,42
    size = stop + 1 - start
    return name + ":" + str(start) + ("" if size < 2 else "-" + str(stop))


def toIUPAC(base1: str, base2: str) -> str:                        # Bases.scala:62-88
    bit = {"A": 1, "C": 2, "G": 4, "T": 8}
    table = {1: "A", 2: "C", 4: "G", 8: "T", 1 | 2: "M", 1 | 4: "R", 1 | 8: "W", 2 | 4: "S", 2 | 8: "Y", 4 | 8: "K",
             1 | 2 | 4: "V", 1 | 2 | 8: "H", 1 | 4 | 8: "D", 2 | 4 | 8: "B", 15: "N"}
    return table[bit[base1] | bit[base2]]


def java_fmt2(x: float) -> str:
    """"%.2f".format(x: Double): java.util.Formatter rounds the shortest decimal representation HALF_UP."""
    return str(Decimal(repr(float(x))).quantize(Decimal("0.01"), rounding=ROUND_HALF_UP))


def smooth(inp: Sequence[int], window: int) -> List[int]:         # GenomeRegion.scala:188-208
    inputSize = len(inp)
    result = [0] * inputSize
    half = window // 2
    accum = 0
    for i in range(inputSize):
        accum = i32(accum + inp[i])
        if i > window:
            accum = i32(accum - inp[i - window])
            smoothed = jdiv(i32(accum + half), window)
            result[i - half] = smoothed
    if inputSize > window:
        for i in range(0, window - half):
            result[i] = result[window - half]
        for i in range(inputSize - half, inputSize):
            result[i] = result[inputSize - half - 1]
    else:
        for i in range(inputSize):
            result[i] = jdiv(accum, inputSize)
    return result


class GenomeRegionOut(GenomeRegionHot):
    """GenomeRegionHot + everything downstream of pass 1 that `--fix snps,indels --changes --vcf` runs."""

    def __init__(self, contigBases: bytes, start: int, stop: int, cfg=None, name: str = "contig", out: Optional[OutConfig] = None):
        super().__init__(contigBases, start, stop, cfg, name)
        self.out = out or OutConfig()
        self.originalBases = bytes(contigBases[start - 1:stop])       # :36, refBases :789-791
        self.bases = bytearray(self.originalBases)                    # :37
        self.copyNumber = [0] * self.size                              # :56
        self.excluded = [False] * self.size                            # :49 (only set for long-read-only runs, :219)
        self.snpFixList: List[Fix] = []                                # :303-305 (Scala lists: newest first)
        self.smallFixList: List[Fix] = []
        self.bigFixList: List[Fix] = []
        self.loglines: List[str] = []

    def locus(self, i: int) -> int:
        return self.start + i

    def index(self, locus: int) -> int:
        return locus - self.start

    # ---- postProcess pass 2 ------------------------------------------------------------------
    def postProcess(self):
        super().postProcess()
        if self.pileUpRegion.readCount == 0:                           # :229-231 returns before pass 2
            return
        baseCov = float(sum(float(v) for v in self.fragCoverage)) / self.size      # NormalDistribution.mean
        smoothCov = smooth(self.fragCoverage, 200)                     # :279
        for i in range(self.size):
            n = smoothCov[i]
            cn = toshort(int(math.floor(n / baseCov + 0.5))) if baseCov > 0 else 0   # (n / baseCov).round.toShort
            self.copyNumber[i] = cn

    # ---- summaries ---------------------------------------------------------------------------
    def nearEdge(self, r: Tuple[int, int], radius: int = 100) -> bool:               # :690
        return r[0] - self.start < radius or self.stop - r[1] < radius

    def summaryRegions(self, positionTest, slop: int = 100) -> List[Tuple[int, int]]:   # :742-763
        regions: List[Tuple[int, int]] = []
        first = last = -1
        for i in range(self.size):
            if positionTest(i):
                last = i
                if first < 0:
                    first = i
            else:
                if last >= 0 and i > last + slop:
                    regions.append((self.locus(first), self.locus(last)))
                    first = last = -1
        if last >= 0:
            regions.append((self.locus(first), self.locus(last)))
        return [r for r in regions if not self.nearEdge(r)]

    def duplicationEvents(self) -> List[Tuple[int, int]]:                           # :735-741
        regions = self.summaryRegions(lambda i: self.copyNumber[i] > 1, 2000)
        return [r for r in regions if r[1] + 1 - r[0] > 10000]

    # ---- identifyAndFixIssues for --fix snps,indels -------------------------------------------
    def identifyAndFixIssues(self):                                                 # :307-414
        o = self.out
        snps = ins = dels = insBases = delBases = amb = 0
        for i in sorted(self.changeMap):                                            # changeList :90
            kind, pu = self.changeMap[i]
            loc = self.locus(i)
            rBase = self.refBase(loc)
            bc = pu.baseCall()
            cBase = bc.base
            if not self.excluded[i]:
                if kind == SNP:
                    if o.fixSnps:
                        self.snpFixList.insert(0, (loc, rBase, cBase))
                    snps += 1
                elif kind == AMB:
                    if o.fixSnps and not o.longread:
                        if o.iupac:
                            self.smallFixList.insert(0, (loc, rBase, toIUPAC(cBase, bc.altBase)))
                        else:
                            self.snpFixList.insert(0, (loc, rBase, cBase))
                        amb += 1
                elif kind == INS:
                    insert = bc.insertion
                    if o.fixIndels:
                        self.smallFixList.insert(0, (loc, "", insert))
                    ins += 1
                    insBases += len(insert)
                elif kind == DEL:
                    deletion = bc.deletion
                    if o.fixIndels:
                        self.smallFixList.insert(0, (loc, deletion, ""))
                    dels += 1
                    delBases += len(deletion)
        nConfirmed = sum(1 for x in self.confirmed if x)
        nonN = sum(1 for x in self.originalBases if x != ord("N"))
        pctc = "%.2f" % (nConfirmed * 100.0 / nonN) if nonN else "NaN"
        self.loglines.append("Confirmed %d of %d bases (%s%%)" % (nConfirmed, nonN, pctc))
        line = "Corrected " if o.fixSnps else "Found "
        line += ("%d snps" % (snps + amb)) if o.diploid else ("%d snps; %d ambiguous bases" % (snps, amb))
        line += "; corrected " if o.fixIndels else "; found "
        line += "%d small insertions totaling %d bases, %d small deletions totaling %d bases" % (ins, insBases, dels, delBases)
        self.loglines.append(line)
        for d in self.duplicationEvents():
            self.loglines.append("Large collapsed region: %s size %d" % (regionString(self.name, d[0], d[1]), d[1] + 1 - d[0]))
        self.stats = dict(confirmed=nConfirmed, nonN=nonN, snps=snps, amb=amb, ins=ins, dels=dels, insBases=insBases, delBases=delBases)
        self.fixIssues(self.snpFixList)                                             # :380
        self.fixIssues(self.smallFixList + self.bigFixList)                         # :413

    def fixFixList(self, inList: List[Fix]) -> List[Fix]:                           # :557-595
        fixes = sorted(inList, key=lambda x: x[0])                                  # sortWith: stable
        outList: List[Fix] = []
        while fixes:
            if len(fixes) >= 2:
                fix1, fix2, tail = fixes[0], fixes[1], fixes[2:]
                r1 = (fix1[0], fix1[0] + max(len(fix1[1]) - 1, 0))
                r2 = (fix2[0], fix2[0] + max(len(fix2[1]) - 1, 0))
                if r2[0] <= r1[1] and r2[1] >= r1[0]:                               # Region.overlaps :33-34
                    fix1len = len(fix1[1]) + len(fix1[2])
                    fix2len = len(fix2[1]) + len(fix2[2])
                    fixes = ([fix1] if fix1len >= fix2len else [fix2]) + tail
                else:
                    fixes = [fix2] + tail
                    outList.append(fix1)
            else:
                outList.append(fixes[0])
                fixes = []
        return outList

    def fixIssues(self, fixList: List[Fix]):                                        # :597-621
        newBases = bytearray(self.bases)
        for locus, was, patch in reversed(self.fixFixList(fixList)):
            start = self.index(locus)
            if len(was) == len(patch):
                for i in range(len(was)):
                    ref = chr(self.originalBases[start + i]).upper()
                    if ref != was[i]:
                        self.loglines.append("Fix mismatch: loc=%d ref=%s was=%s" % (locus + i, ref, was[i]))
                    newBases[start + i] = ord(patch[i])
            else:
                ref = self.originalBases[start:start + len(was)].decode("latin1").upper()
                if ref != was:
                    self.loglines.append("Fix mismatch: loc=%d ref=%s was=%s" % (locus, ref, was))
                newBases = newBases[:start] + bytearray(patch.encode("latin1")) + newBases[start + len(was):]
        self.bases = newBases

    # ---- writers -----------------------------------------------------------------------------
    def writeChanges(self, newName: Optional[str] = None, offset: int = 0) -> List[str]:      # :646-657
        newName = self.name if newName is None else newName
        fixes = self.fixFixList(self.snpFixList + self.smallFixList + self.bigFixList)
        delta = 0
        out = []
        for loc, frm, to in fixes:
            newLoc = loc + delta
            out.append(regionString(self.name, loc, loc + len(frm) - 1) + " " +
                       regionString(newName, newLoc + offset, newLoc + offset + len(to) - 1) + " " +
                       (frm if frm else ".") + " " + (to if to else "."))
            delta += len(to) - len(frm)
        return out

    def writeVcf(self, vcf: "Vcf"):                                                 # :623-643
        fixes = self.fixFixList(self.snpFixList + self.smallFixList + self.bigFixList)
        dupes = self.duplicationEvents()
        for i in range(self.size):
            loc = self.locus(i)
            if dupes and dupes[0][0] == loc:
                vcf.writeDup(self, dupes[0])
                dupes = dupes[1:]
            if fixes and fixes[0][0] == loc:
                fixes = fixes[1:]                                                   # (big fixes only would write a fix record)
            vcf.writeRecord(self, i, self.deleted[i])


class Vcf:
    """Vcf.scala; lines are collected instead of printed."""

    def __init__(self, out: Optional[OutConfig] = None):
        self.out = out or OutConfig()
        self.lines: List[str] = []

    def writeHeader(self, date: str, version: str, commandArgs: str, reference: str, contigsWithSizes: Sequence[Tuple[str, int]]):   # :28-68
        w = self.lines.append
        w("##fileformat=VCFv4.1")
        w("##fileDate=" + date)
        w("##source=\"" + version + "\"")
        w("##PILON=\"" + commandArgs + "\"")
        w("##reference=" + reference)
        for c, s in contigsWithSizes:
            w("##contig=<ID=" + c + ",length=" + str(s) + ">")
        w("##FILTER=<ID=LowCov,Description=\"Low Coverage of good reads at location\">")
        w("##FILTER=<ID=Amb,Description=\"Ambiguous evidence in haploid genome\">")
        w("##FILTER=<ID=Del,Description=\"This base is in a deletion or change event from another record\">")
        w("##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Valid read depth; some reads may have been filtered\">")
        w("##INFO=<ID=TD,Number=1,Type=Integer,Description=\"Total read depth including bad pairs\">")
        w("##INFO=<ID=PC,Number=1,Type=Integer,Description=\"Physical coverage of valid inserts across locus\">")
        w("##INFO=<ID=BQ,Number=1,Type=Integer,Description=\"Mean base quality at locus\">")
        w("##INFO=<ID=MQ,Number=1,Type=Integer,Description=\"Mean read mapping quality at locus\">")
        w("##INFO=<ID=QD,Number=1,Type=Integer,Description=\"Variant confidence/quality by depth\">")
        w("##INFO=<ID=BC,Number=4,Type=Integer,Description=\"Count of As, Cs, Gs, Ts at locus\">")
        if self.out.vcfQE:
            w("##INFO=<ID=QE,Number=4,Type=Integer,Description=\"Evidence for As, Cs, Gs, Ts weighted by Q & MQ at locus\">")
        else:
            w("##INFO=<ID=QP,Number=4,Type=Integer,Description=\"Percentage of As, Cs, Gs, Ts weighted by Q & MQ at locus\">")
        w("##INFO=<ID=IC,Number=1,Type=Integer,Description=\"Number of reads with insertion here\">")
        w("##INFO=<ID=DC,Number=1,Type=Integer,Description=\"Number of reads with deletion here\">")
        w("##INFO=<ID=XC,Number=1,Type=Integer,Description=\"Number of reads clipped here\">")
        w("##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes, for each ALT allele, in the same order as listed\">")
        w("##INFO=<ID=AF,Number=A,Type=Float,Description=\"Fraction of evidence in support of alternate allele(s)\">")
        w("##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Type of structural variant\">")
        w("##INFO=<ID=SVLEN,Number=.,Type=String,Description=\"Difference in length between REF and ALT alleles\">")
        w("##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the variant described in this record\">")
        w("##INFO=<ID=IMPRECISE,Number=0,Type=Flag,Description=\"Imprecise change from local reassembly (ALT contains Ns)\">")
        w("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">")
        w("##FORMAT=<ID=AD,Number=.,Type=String,Description=\"Allelic depths for the ref and alt alleles in the order listed\">")
        w("##FORMAT=<ID=DP,Number=1,Type=String,Description=\"Approximate read depth; some reads may have been filtered\">")
        w("##ALT=<ID=DUP,Description=\"Possible segmental duplication\">")
        w("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE")

    def writeRecord(self, region: GenomeRegionOut, index: int, embedded: bool = False, indelOkArg: bool = True):   # :74-176
        tab = "\t"
        indelOk = indelOkArg and index > 0
        locus = region.locus(index)
        pileUp = region.pileUpRegion[index]
        bc = pileUp.baseCall()
        bcString = bc.callString(indelOk)
        baseDP = i32(bc.baseSum)
        altBaseDP = i32(bc.altBaseSum)
        depth = i32(pileUp.depth)
        loc = locus
        if indelOk and not embedded and bc.isDeletion:
            loc -= 1
            rBase = region.refBase(loc)
            callType = "1/1" if bc.homoIndel else "0/1"
            p = pileUp.delPct
            rB, cB, refDP, altDP = rBase + bcString, rBase, 100 - p, p
        elif indelOk and not embedded and bc.isInsertion:
            loc -= 1
            rBase = region.refBase(loc)
            callType = "1/1" if bc.homoIndel else "0/1"
            p = pileUp.insPct
            rB, cB, refDP, altDP = rBase, rBase + bcString, 100 - p, p
        elif bc.homo:
            rBase = region.refBase(loc)
            if rBase == bc.base or bcString == "N":
                rB, cB, callType, refDP, altDP = rBase, bc.base, "0/0", baseDP, altBaseDP
            else:
                rB, cB, callType, refDP, altDP = rBase, bc.base, "1/1", altBaseDP, baseDP
        else:
            rBase = region.refBase(loc)
            if rBase == bc.base:
                rB, cB, callType, refDP, altDP = rBase, bc.altBase, "0/1", baseDP, altBaseDP
            else:
                rB, cB, callType, refDP, altDP = rBase, bc.base, "0/1", altBaseDP, baseDP
        filters: List[str] = []
        if depth < region.minDepth:
            filters.insert(0, "LowCov")
        if not self.out.diploid and callType == "0/1":
            filters.insert(0, "Amb")
        if embedded:
            filters.insert(0, "Del")
        if not filters:
            filters.insert(0, "PASS")
        cBaseVcf = "." if (cB == "N" or cB == rB) else cB
        flt = ";".join(filters)
        ac = {"0/0": 0, "0/1": 1, "1/1": 2}[callType]
        if i32(refDP + altDP) > 0 and cBaseVcf != ".":
            af = float(np.float32(altDP) / np.float32(i32(refDP + altDP)))         # Float division, widened to Double
        else:
            af = 0.0
        info = ("DP=" + str(pileUp.depth if not embedded else pileUp.count) +
                ";TD=" + str(pileUp.depth + pileUp.badPair) +
                ";BQ=" + str(pileUp.meanQual) +
                ";MQ=" + str(pileUp.meanMq) +
                ";QD=" + str(bc.q) +
                ";BC=" + str(pileUp.baseCount) +
                (";QE=" + str(pileUp.qualSum) if self.out.vcfQE else ";QP=" + pileUp.qualSum.toStringPct()) +
                ";PC=" + str(pileUp.physCov) +
                ";IC=" + str(pileUp.insertions) +
                ";DC=" + str(pileUp.deletions) +
                ";XC=" + str(pileUp.clips) +
                ";AC=" + str(ac) +
                ";AF=" + java_fmt2(af))
        line = (region.name + tab + str(loc) + tab + "." + tab + rB + tab + cBaseVcf + tab +
                ("." if (indelOk and bc.isDeletion) else str(bc.score)) + tab + flt + tab + info + tab + "GT" + tab + callType)
        self.lines.append(line)
        if indelOk and bc.indel and not embedded:
            self.writeRecord(region, index, bc.isDeletion and bc.homoIndel, False)

    def writeDup(self, region: GenomeRegionOut, dup: Tuple[int, int]):              # :193-201
        tab = "\t"
        loc = dup[0] - 1
        rBase = region.refBase(loc)
        line = region.name + tab + str(loc) + tab + "." + tab
        line += rBase + tab + "<DUP>" + tab + "." + tab + "PASS" + tab
        line += "SVTYPE=DUP;SVLEN=" + str(dup[1] + 1 - dup[0]) + ";END=" + str(dup[1]) + ";IMPRECISE"
        line += tab + "GT" + tab + "./."
        self.lines.append(line)


# ---- GenomeFile-level output (GenomeFile.scala:79-82, 135-162, 178-187) ----------------------
def pilonName(name: str) -> str:                                                    # :137-141
    if "|" not in name:
        sep = "_"
    elif name[-1] == "|":
        sep = ""
    else:
        sep = "|"
    return name + sep + "pilon"


def fastaElement(header: str, sequence: str) -> List[str]:                          # :79-82
    return [">" + header] + [sequence[i:i + 80] for i in range(0, len(sequence), 80)]


def writeContig(name: str, chunks: Sequence[GenomeRegionOut], vcf: Optional[Vcf] = None, changes: bool = True):
    """The body of `regions foreach` (:135-162) for one contig: returns (changes lines, fasta lines)."""
    newName = pilonName(name)
    offset = 0
    changeLines: List[str] = []
    for r in chunks:
        if vcf is not None:
            r.writeVcf(vcf)
        if changes:
            changeLines += r.writeChanges(newName, offset)
            offset += len(r.bases) - r.size
    bases = "".join(r.bases.decode("latin1") for r in chunks)
    return changeLines, fastaElement(newName, bases)


def coverageSummary(bamTypesBaseCounts: Sequence[Tuple[str, int]], genomeSize: int) -> List[str]:     # :178-187
    """bamTypesBaseCounts: (bamType, BamFile.baseCount) per BAM file; grouping order follows first appearance (the
    reference iterates a HashMap: order unspecified, compare as a set)."""
    out, total, seen = [], 0, []
    for t, _ in bamTypesBaseCounts:
        if t not in seen:
            seen.append(t)
    for t in seen:
        typeBaseCount = sum(c for tt, c in bamTypesBaseCounts if tt == t)
        out.append("Mean " + t + " coverage: " + str(roundDivL(typeBaseCount, genomeSize)))
        total += typeBaseCount
    out.append("Mean total coverage: " + str(roundDivL(total, genomeSize)))
    return out


__all__ = ["OutConfig", "GenomeRegionOut", "Vcf", "pilonName", "fastaElement", "writeContig", "coverageSummary", "regionString",
           "toIUPAC", "java_fmt2", "smooth", "pctI"]
