"""BAM in, BAM out over the C ABI (pb_bam_*, include/pilon_b200.h).

`BamFile.process(region)` mirrors the reader half of the reference's BamFile.process (BamFile.scala:108-148): the records
overlapping the region +-10 kb that pass validateRead, packed for the engine -- without htsjdk.  `write_bam` / `write_fasta`
produce the synthetic inputs a real Pilon JVM can be run on (tools/run_real_pilon.sh).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from .packing import ReadBatch

_bound = False


def _lib() -> C.CDLL:
    global _bound
    lib = capi.load_library()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.pb_bam_last_error.restype = C.c_char_p
        lib.pb_bam_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
        lib.pb_bam_close.argtypes = [vp]
        lib.pb_bam_n_refs.argtypes = [vp, C.POINTER(i32)]
        lib.pb_bam_ref.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64)]
        lib.pb_bam_query_pack.argtypes = [vp, i32, i32, i32, C.c_int, C.c_int, vp, C.POINTER(i64), C.POINTER(i64)]
        lib.pb_bam_writer_open.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(i64), i32, C.c_char_p, C.POINTER(vp)]
        lib.pb_bam_writer_add_batch.argtypes = [vp, i32, C.POINTER(capi.pb_batch), vp]
        lib.pb_bam_writer_close.argtypes = [vp, C.c_char_p]
        lib.pb_fasta_write.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i64), i32]
        for n in ("pb_bam_open", "pb_bam_close", "pb_bam_n_refs", "pb_bam_ref", "pb_bam_query_pack", "pb_bam_writer_open",
                  "pb_bam_writer_add_batch", "pb_bam_writer_close", "pb_fasta_write"):
            getattr(lib, n).restype = C.c_int
        _bound = True
    return lib


def _check(rc: int):
    if rc != capi.PB_OK:
        msg = _lib().pb_bam_last_error()
        raise capi.EngineError(rc, msg.decode() if msg else "")


def batch_from_view(v: capi.pb_batch) -> ReadBatch:
    """Copies a pb_batch view (e.g. a packer's) into numpy arrays."""
    def arr(ptr, n, dt):
        if not n:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
    return ReadBatch(arr(v.pos, v.n_reads, np.int32), arr(v.tlen, v.n_reads, np.int32), arr(v.read_len, v.n_reads, np.int32),
                     arr(v.mapq, v.n_reads, np.uint8), arr(v.flags, v.n_reads, np.uint8),
                     arr(v.cigar_off, v.n_reads + 1, np.uint32) if v.n_reads else np.zeros(1, np.uint32),
                     arr(v.cigar, v.n_cigar, np.uint32), arr(v.seq_off, v.n_reads, np.uint32), arr(v.quals, v.n_seq, np.uint8),
                     arr(v.bases2, v.n_seq // 4, np.uint8), arr(v.exc_idx, v.n_exc, np.uint32), arr(v.exc_base, v.n_exc, np.uint8),
                     arr(v.exc_qual, v.n_exc, np.uint8))


class BamFile:
    """The slice of the reference's BamFile the pileup path uses (BamFile.scala:49-53,101-148)."""

    def __init__(self, path: str, bamType: str = "frags", index: Optional[str] = None, nonPf: bool = False, duplicates: bool = False):
        self.lib = _lib()
        self.path, self.bamType, self.nonPf, self.duplicates = path, bamType, nonPf, duplicates
        self._h = C.c_void_p()
        _check(self.lib.pb_bam_open(path.encode(), index.encode() if index else None, C.byref(self._h)))
        n = C.c_int32()
        _check(self.lib.pb_bam_n_refs(self._h, C.byref(n)))
        self.refs: List[Tuple[str, int]] = []
        for i in range(n.value):
            nm, ln = C.c_char_p(), C.c_int64()
            _check(self.lib.pb_bam_ref(self._h, i, C.byref(nm), C.byref(ln)))
            self.refs.append((nm.value.decode(), int(ln.value)))
        self.baseCount = 0                                       # BamFile.scala:146 accumulates what coverageSummary prints
        self._pk = C.c_void_p()
        capi.check(self.lib.pb_packer_create(C.byref(self._pk)))

    def close(self):
        if self._h:
            self.lib.pb_bam_close(self._h)
            self.lib.pb_packer_destroy(self._pk)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, contig: str, start: int, stop: int) -> Tuple[ReadBatch, int]:
        """Packed batch of the records of `contig` overlapping [start, stop] that pass validateRead; also the number rejected."""
        ref_id = [n for n, _ in self.refs].index(contig)
        capi.check(self.lib.pb_packer_reset(self._pk))
        n_ok, n_rej = C.c_int64(), C.c_int64()
        _check(self.lib.pb_bam_query_pack(self._h, ref_id, start, stop, int(self.nonPf), int(self.duplicates), self._pk,
                                          C.byref(n_ok), C.byref(n_rej)))
        v = capi.pb_batch()
        capi.check(self.lib.pb_packer_view(self._pk, C.byref(v)))
        return batch_from_view(v), int(n_rej.value)

    def process(self, contig: str, start: int, stop: int) -> ReadBatch:
        """reader.queryOverlapping(name, (start - 10000) max 0, (stop + 10000) min contig.length) + validateRead
        (BamFile.scala:117-126); the caller adds the batch to its PileUpRegion."""
        clen = dict(self.refs)[contig]
        return self.query(contig, max(start - 10000, 0), min(stop + 10000, clen))[0]

    @property
    def countsTowardFragCoverage(self) -> bool:                  # GenomeRegion.scala:291,296
        return self.bamType != "jumps"


def write_bam(path: str, refs: Sequence[Tuple[str, int]], batches: Sequence[Tuple[int, ReadBatch]], bai: Optional[str] = None,
              extra_flags: Optional[Sequence[Optional[np.ndarray]]] = None, program_line: Optional[str] = None):
    """Coordinate-sorted BAM (+ BAI) from (reference index, batch) pairs in file order."""
    lib = _lib()
    names = (C.c_char_p * max(1, len(refs)))(*[n.encode() for n, _ in refs])
    lens = (C.c_int64 * max(1, len(refs)))(*[int(l) for _, l in refs])
    w = C.c_void_p()
    _check(lib.pb_bam_writer_open(path.encode(), names, lens, len(refs), program_line.encode() if program_line else None, C.byref(w)))
    try:
        for k, (ref_id, rb) in enumerate(batches):
            cb = rb.to_c() if hasattr(rb, "to_c") else rb
            xf = extra_flags[k] if extra_flags is not None else None
            xf = np.ascontiguousarray(xf, np.uint16) if xf is not None else None
            _check(lib.pb_bam_writer_add_batch(w, ref_id, C.byref(cb), xf.ctypes.data if xf is not None else None))
    finally:
        _check(lib.pb_bam_writer_close(w, (bai or path + ".bai").encode()))


def write_fasta(path: str, contigs: Sequence[Tuple[str, bytes]]):
    lib = _lib()
    n = len(contigs)
    keep = [np.frombuffer(s, np.uint8) for _, s in contigs]
    names = (C.c_char_p * max(1, n))(*[c.encode() for c, _ in contigs])
    seqs = (C.c_void_p * max(1, n))(*[a.ctypes.data for a in keep])
    lens = (C.c_int64 * max(1, n))(*[len(s) for _, s in contigs])
    _check(lib.pb_fasta_write(path.encode(), names, seqs, lens, n))
