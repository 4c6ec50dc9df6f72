"""Host-side mirror of the reference's per-locus consumers and fix application over the C ABI (pb_out_*,
include/pilon_b200.h): what GenomeRegion.identifyAndFixIssues / writeChanges / writeVcf, Vcf.writeRecord, Tracks.makeTrack
and GenomeFile's output loop do with the engine's per-locus results for `--fix snps,indels [--changes] [--vcf] [--tracks]`
(reference GenomeRegion.scala:275-283,307-380,557-657; Vcf.scala:28-201; Tracks.scala:56-186; GenomeFile.scala:79-82,
122-187).  Same member names and argument meaning as the reference where a member exists there.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from .packing import ResultBuffers


class pb_output_config(C.Structure):
    _fields_ = [("fix_snps", C.c_int32), ("fix_indels", C.c_int32), ("iupac", C.c_int32), ("diploid", C.c_int32),
                ("vcf_qe", C.c_int32), ("longread", C.c_int32), ("reserved", C.c_int32 * 2)]


class pb_out_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("confirmed", "non_n", "snps", "amb", "ins", "dels", "ins_bases", "del_bases",
                                         "n_fixes", "fix_mismatches", "n_dups")]


TRACKS = {"Changes": 0, "Unconfirmed": 1, "Copy Number": 2, "Coverage": 3, "Bad Coverage": 4, "Pct Bad": 5,
          "Delta Coverage": 6, "Dip Coverage": 7, "Physical Coverage": 8, "Clipped Alignments": 9, "Weighted Qual": 10,
          "Weighted MQ": 11}

_bound = False


def _lib() -> C.CDLL:
    global _bound
    lib = capi.load_library()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.pb_out_last_error.restype = C.c_char_p
        lib.pb_out_create.argtypes = [C.POINTER(capi.pb_region_result), vp, i64, C.c_char_p, i32, i32,
                                      C.POINTER(pb_output_config), C.POINTER(vp)]
        lib.pb_out_destroy.argtypes = [vp]
        lib.pb_out_stats_get.argtypes = [vp, C.POINTER(pb_out_stats)]
        lib.pb_out_bases.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_out_copy_number.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_out_log.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_out_changes.argtypes = [vp, C.c_char_p, i64, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_out_vcf.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_out_wig.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(i64)]
        lib.pb_pilon_name.argtypes = [C.c_char_p, vp, i64, C.POINTER(i64)]
        lib.pb_fasta_element.argtypes = [C.c_char_p, vp, i64, vp, i64, C.POINTER(i64)]
        lib.pb_vcf_header.argtypes = [C.POINTER(pb_output_config), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                      C.POINTER(C.c_char_p), C.POINTER(i64), i32, vp, i64, C.POINTER(i64)]
        for n in ("pb_out_create", "pb_out_destroy", "pb_out_stats_get", "pb_out_bases", "pb_out_copy_number", "pb_out_log",
                  "pb_out_changes", "pb_out_vcf", "pb_out_wig", "pb_pilon_name", "pb_fasta_element", "pb_vcf_header"):
            getattr(lib, n).restype = C.c_int
        _bound = True
    return lib


def _check(rc: int):
    if rc != capi.PB_OK:
        msg = _lib().pb_out_last_error()
        raise capi.EngineError(rc, msg.decode() if msg else "")


class OutputConfig:
    def __init__(self, fixSnps=True, fixIndels=True, iupac=False, diploid=False, vcfQE=False, longread=False):
        self.fixSnps, self.fixIndels, self.iupac = fixSnps, fixIndels, iupac
        self.diploid, self.vcfQE, self.longread = diploid, vcfQE, longread

    def to_c(self) -> pb_output_config:
        return pb_output_config(fix_snps=int(self.fixSnps), fix_indels=int(self.fixIndels), iupac=int(self.iupac),
                                diploid=int(self.diploid), vcf_qe=int(self.vcfQE), longread=int(self.longread))


class RegionOutput:
    """One GenomeRegion after identifyAndFixIssues, built from the engine's result for that region."""

    def __init__(self, res: ResultBuffers, contigBases: bytes, name: str, start: int, stop: int,
                 config: Optional[OutputConfig] = None):
        self.lib = _lib()
        self.res, self.name, self.start, self.stop = res, name, start, stop
        self.size = stop + 1 - start
        self._contig = np.frombuffer(contigBases, np.uint8)
        self._h = C.c_void_p()
        cfg = (config or OutputConfig()).to_c()
        _check(self.lib.pb_out_create(C.byref(res.c), self._contig.ctypes.data, len(contigBases), name.encode(), start, stop,
                                      C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.pb_out_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _text(self, fn, *args) -> str:
        p, n = C.c_void_p(), C.c_int64()
        _check(fn(self._h, *args, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value).decode("latin1") if n.value else ""

    @property
    def stats(self) -> Dict[str, int]:
        st = pb_out_stats()
        _check(self.lib.pb_out_stats_get(self._h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in pb_out_stats._fields_}

    @property
    def bases(self) -> bytes:                                    # GenomeRegion.bases
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib.pb_out_bases(self._h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    @property
    def copyNumber(self) -> np.ndarray:                           # GenomeRegion.copyNumber
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib.pb_out_copy_number(self._h, C.byref(p), C.byref(n)))
        if not n.value:
            return np.zeros(self.size, np.int16)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int16)), shape=(n.value,)).copy()

    def log(self) -> List[str]:
        return self._text(self.lib.pb_out_log).splitlines()

    def writeChanges(self, newName: Optional[str] = None, offset: int = 0) -> List[str]:      # GenomeRegion.scala:646-657
        return self._text(self.lib.pb_out_changes, newName.encode() if newName is not None else None, offset).splitlines()

    def writeVcf(self, threads: int = 0) -> str:                                             # GenomeRegion.scala:623-643
        import os
        return self._text(self.lib.pb_out_vcf, threads or len(os.sched_getaffinity(0)))

    def wig(self, track: str) -> str:                                                        # Tracks.scala:169-186, per region
        return self._text(self.lib.pb_out_wig, TRACKS[track])


def _sized(fn, *args) -> str:
    n = C.c_int64()
    _check(fn(*args, None, 0, C.byref(n)))
    buf = C.create_string_buffer(max(1, n.value))
    _check(fn(*args, buf, n.value, C.byref(n)))
    return buf.raw[:n.value].decode("latin1")


def pilonName(name: str) -> str:                                  # GenomeFile.scala:137-141
    return _sized(_lib().pb_pilon_name, name.encode())


def fastaElement(header: str, bases: bytes) -> str:               # GenomeFile.scala:79-82
    arr = np.frombuffer(bases, np.uint8)
    return _sized(_lib().pb_fasta_element, header.encode(), arr.ctypes.data if len(bases) else None, len(bases))


def vcfHeader(date: str, version: str, commandArgs: str, reference: str, contigsWithSizes: Sequence[Tuple[str, int]],
              config: Optional[OutputConfig] = None) -> str:      # Vcf.scala:28-68
    cfg = (config or OutputConfig()).to_c()
    names = (C.c_char_p * max(1, len(contigsWithSizes)))(*[c.encode() for c, _ in contigsWithSizes])
    sizes = (C.c_int64 * max(1, len(contigsWithSizes)))(*[int(s) for _, s in contigsWithSizes])
    return _sized(_lib().pb_vcf_header, C.byref(cfg), date.encode(), version.encode(), commandArgs.encode(), reference.encode(),
                  names, sizes, len(contigsWithSizes))


def writeContig(name: str, chunks: Sequence[RegionOutput], vcf: bool = False, changes: bool = True):
    """The body of GenomeFile.processRegions' output loop for one contig (GenomeFile.scala:135-162):
    returns (changes lines, FASTA text, VCF record text)."""
    newName = pilonName(name)
    offset = 0
    changeLines: List[str] = []
    vcfText = []
    for r in chunks:
        if vcf:
            vcfText.append(r.writeVcf())
        if changes:
            changeLines += r.writeChanges(newName, offset)
            offset += len(r.bases) - r.size
    return changeLines, fastaElement(newName, b"".join(r.bases for r in chunks)), "".join(vcfText)


def coverageSummary(perBam: Sequence[Tuple[str, int]], genomeSize: int) -> List[str]:       # GenomeFile.scala:178-187
    """perBam: (bamType, accumulated baseCount of that BAM) -- pb_region_result.batch_base_count summed over regions."""
    from .engine import roundDiv
    out, total, seen = [], 0, []
    for t, _ in perBam:
        if t not in seen:
            seen.append(t)
    for t in seen:
        s = sum(c for tt, c in perBam if tt == t)
        out.append("Mean %s coverage: %d" % (t, roundDiv(s, genomeSize)))
        total += s
    out.append("Mean total coverage: %d" % roundDiv(total, genomeSize))
    return out
