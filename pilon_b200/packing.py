"""Host-side read batching: SAMRecord-like objects -> the struct-of-arrays `pb_batch`
(include/pilon_b200.h).  This is the re-plumbed half of `BamFile.process`
(reference BamFile.scala:108-148): instead of calling `PileUpRegion.addRead` once per record,
the records of a region query are packed once and handed to the engine.

htsjdk semantics restated here (htsjdk 2.23.0, SAM spec): bases are the upper-case ASCII letters
"=ACMGRSVTWYHKDBN"; an all-0xFF quality array means "no qualities" (PB_F_HAS_QUALS clear);
getAlignmentStart is 1-based.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _capi as capi

def base_delta_encode(cb, contig, start: int, stop: int):
    """pb_base_delta_encode on a C batch; returns (idx uint32[n + 16], code uint8[n + 16]) numpy copies (16 spare entries
    keep the arrays non-empty and addressable)."""
    lib = capi.load_library()
    pi, pc, n = C.c_void_p(), C.c_void_p(), C.c_int64()
    buf = contig if isinstance(contig, (bytes, bytearray)) else None
    cptr = C.cast(C.c_char_p(bytes(contig)) if buf is not None else C.c_void_p(contig.ctypes.data), C.c_void_p)
    clen = len(contig) if buf is not None else int(contig.shape[0])
    capi.check(lib.pb_base_delta_encode(C.byref(cb), cptr, clen, start, stop, C.byref(pi), C.byref(pc), C.byref(n)))
    idx = np.zeros(n.value + 16, np.uint32)
    code = np.zeros(n.value + 16, np.uint8)
    if n.value:
        idx[:n.value] = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint32)), shape=(n.value,))
        code[:n.value] = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint8)), shape=(n.value,))
    lib.pb_free(pi)
    lib.pb_free(pc)
    return idx, code


def meta_encode(cb):
    """pb_meta_encode on a C batch: (codes uint64[n], cigar uint32[m + 16], esc int32[3 k + 16], pos0, seq_stride) as numpy
    copies, or None when the batch cannot be put that way (long reads, another seq_off layout ...)."""
    lib = capi.load_library()
    pc, pg, pe = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ng, ne, pos0, stride = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    rc = lib.pb_meta_encode(C.byref(cb), C.byref(pc), C.byref(pg), C.byref(ng), C.byref(pe), C.byref(ne), C.byref(pos0), C.byref(stride))
    if rc == capi.PB_ERR_UNSUPPORTED:
        return None
    capi.check(rc)
    n = int(cb.n_reads)
    codes = np.zeros(n + 2, np.uint64)
    cigar = np.zeros(ng.value + 16, np.uint32)
    esc = np.zeros(3 * ne.value + 16, np.int32)
    if n:
        codes[:n] = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint64)), shape=(n,))
    if ng.value:
        cigar[:ng.value] = np.ctypeslib.as_array(C.cast(pg, C.POINTER(C.c_uint32)), shape=(ng.value,))
    if ne.value:
        esc[:3 * ne.value] = np.ctypeslib.as_array(C.cast(pe, C.POINTER(C.c_int32)), shape=(3 * ne.value,))
    for p in (pc, pg, pe):
        lib.pb_free(p)
    return codes, cigar, esc, int(ng.value), int(ne.value), int(pos0.value), int(stride.value)


def meta_decode(codes, cigar, esc, n_esc, pos0, stride, n):
    """Reference decoder of the compact metadata (numpy, for tests): the eight plain arrays."""
    c = codes[:n]
    delta = (c & np.uint64(0xFFFF)).astype(np.int64)
    tl = ((c >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint16).view(np.int16).astype(np.int32)
    for k in range(n_esc):
        r, f, v = (int(x) for x in esc[3 * k:3 * k + 3])
        if f == 0:
            delta[r] = v
        else:
            tl[r] = v
    pos = (pos0 + np.cumsum(delta)).astype(np.int32)
    read_len = ((c >> np.uint64(32)) & np.uint64(0xFF)).astype(np.int32)
    mapq = ((c >> np.uint64(40)) & np.uint64(0xFF)).astype(np.uint8)
    flags = ((c >> np.uint64(48)) & np.uint64(0xFF)).astype(np.uint8)
    nl = (c >> np.uint64(56)).astype(np.int64)
    nops = np.where(nl == 0, 1, nl)
    cigar_off = np.zeros(n + 1, np.uint32)
    cigar_off[1:] = np.cumsum(nops)
    out = np.zeros(int(cigar_off[-1]), np.uint32)
    lo = np.concatenate([[0], np.cumsum(nl)])
    for r in range(n):
        if nl[r] == 0:
            out[cigar_off[r]] = np.uint32(read_len[r]) << np.uint32(4)
        else:
            out[cigar_off[r]:cigar_off[r + 1]] = cigar[lo[r]:lo[r + 1]]
    pad = (read_len + 3) & ~3
    seq_off = (np.arange(n, dtype=np.int64) * stride if stride > 0 else np.concatenate([[0], np.cumsum(pad)[:-1]]) if n else np.zeros(0)).astype(np.uint32)
    return pos, tl, read_len, mapq, flags, cigar_off, out, seq_off


_OPCODE = {op: i for i, op in enumerate(capi.CIGAR_OPS)}
_BASE_CODE = np.full(256, 255, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _BASE_CODE[_c] = _i


@dataclass
class ReadBatch:
    """Numpy struct-of-arrays with the exact layout of `pb_batch`."""
    pos: np.ndarray
    tlen: np.ndarray
    read_len: np.ndarray
    mapq: np.ndarray
    flags: np.ndarray
    cigar_off: np.ndarray
    cigar: np.ndarray
    seq_off: np.ndarray
    quals: np.ndarray
    bases2: np.ndarray
    exc_idx: np.ndarray
    exc_base: np.ndarray
    exc_qual: np.ndarray
    qual_codes: Optional[np.ndarray] = None    # pb_batch.qual_codes: packed 3- or 4-bit codes (compact H2D transport)
    qual_code_bits: int = 0                    # pb_batch.qual_code_bits
    qual_lut: Optional[np.ndarray] = None      # pb_batch.qual_lut: code -> quality byte
    base_delta_idx: Optional[np.ndarray] = None    # pb_batch.base_delta_idx / base_delta_code: bases that differ from
    base_delta_code: Optional[np.ndarray] = None   # the reference prediction (compact H2D transport of bases2)
    meta: Optional[tuple] = None                   # compact per-read metadata (meta_encode): pb_batch.meta_codes & co.

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    @property
    def aligned_bases(self) -> int:
        ops = self.cigar & 15
        return int((self.cigar >> 4)[(ops == 0) | (ops == 7) | (ops == 8)].sum())

    def to_c(self) -> capi.pb_batch:
        b = capi.pb_batch()
        b.n_reads, b.n_cigar = self.n_reads, int(self.cigar.shape[0])
        b.n_seq, b.n_exc = int(self.quals.shape[0]), int(self.exc_idx.shape[0])
        for name in ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off",
                     "quals", "bases2", "exc_idx", "exc_base", "exc_qual"):
            arr = getattr(self, name)
            assert arr.flags["C_CONTIGUOUS"]
            setattr(b, name, arr.ctypes.data)
        if self.qual_codes is not None:
            assert self.qual_codes.flags["C_CONTIGUOUS"]
            assert self.qual_codes.shape[0] >= (self.quals.shape[0] * self.qual_code_bits + 7) // 8
            b.qual_codes = self.qual_codes.ctypes.data
            b.qual_code_bits = self.qual_code_bits
            for i in range(16):
                b.qual_lut[i] = int(self.qual_lut[i])
        if self.base_delta_idx is not None:
            b.base_delta_idx = self.base_delta_idx.ctypes.data
            b.base_delta_code = self.base_delta_code.ctypes.data
            b.n_base_delta = int(self.base_delta_idx.shape[0]) - 16      # (the arrays carry 16 spare entries)
        if self.meta is not None:
            codes, cigar, esc, ng, ne, pos0, stride = self.meta
            b.meta_codes, b.meta_cigar, b.meta_esc = codes.ctypes.data, cigar.ctypes.data, esc.ctypes.data
            b.n_meta_cigar, b.n_meta_esc, b.meta_pos0, b.meta_seq_stride = ng, ne, pos0, stride
        b.mem = capi.PB_MEM_HOST
        b._keepalive = self
        return b

    def with_compact_meta(self) -> "ReadBatch":
        """Adds the compact transport of the per-read arrays (pb_meta_encode); returns self unchanged when the batch cannot
        be put that way."""
        m = meta_encode(self.to_c())
        return self if m is None else dataclasses.replace(self, meta=m)

    def with_base_deltas(self, contig: bytes, start: int, stop: int) -> "ReadBatch":
        """Adds the reference-delta transport of bases2 for the region [start, stop] (pb_base_delta_encode)."""
        idx, code = base_delta_encode(self.to_c(), contig, start, stop)
        return dataclasses.replace(self, base_delta_idx=idx, base_delta_code=code)

    def with_packed_quals(self) -> "ReadBatch":
        """Adds the packed quality transport when the batch uses at most 8 (3-bit codes) or 16 (4-bit codes) distinct
        quality bytes -- what pb_packer_view does natively; returns self unchanged otherwise."""
        vals = np.unique(np.concatenate([self.quals, np.zeros(1, np.uint8)]))
        if vals.shape[0] > 16:
            return self
        bits = 3 if vals.shape[0] <= 8 else 4
        lut = np.zeros(16, np.uint8)
        lut[:vals.shape[0]] = vals
        code = np.zeros(256, np.uint8)
        code[vals] = np.arange(vals.shape[0], dtype=np.uint8)
        c = code[self.quals].astype(np.uint32)
        n = c.shape[0]
        if bits == 4:
            packed = (c[0::2] | (c[1::2] << 4)).astype(np.uint8)
        else:
            c8 = np.zeros((n + 7) // 8 * 8, np.uint32)
            c8[:n] = c
            v = (c8.reshape(-1, 8) << (3 * np.arange(8, dtype=np.uint32))).sum(axis=1, dtype=np.uint32)
            packed = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
        buf = np.zeros(packed.shape[0] + 32, np.uint8)
        buf[:packed.shape[0]] = packed
        return dataclasses.replace(self, qual_codes=buf, qual_code_bits=bits, qual_lut=lut)

    def slice_reads(self, lo: int, hi: int) -> "ReadBatch":
        """Sub-batch of reads [lo, hi) (re-based offsets); used to split a region's reads."""
        co = self.cigar_off[lo:hi + 1].astype(np.int64)
        if hi > lo:
            s0 = int(self.seq_off[lo])
            s1 = int(self.seq_off[hi - 1]) + ((int(self.read_len[hi - 1]) + 3) & ~3)
        else:
            s0 = s1 = 0
        e0, e1 = np.searchsorted(self.exc_idx, [s0, s1])
        return ReadBatch(self.pos[lo:hi].copy(), self.tlen[lo:hi].copy(), self.read_len[lo:hi].copy(),
                         self.mapq[lo:hi].copy(), self.flags[lo:hi].copy(),
                         (co - co[0]).astype(np.uint32), self.cigar[co[0]:co[-1]].copy(),
                         (self.seq_off[lo:hi] - np.uint32(s0)).astype(np.uint32),
                         self.quals[s0:s1].copy(), self.bases2[s0 // 4:s1 // 4].copy(),
                         (self.exc_idx[e0:e1] - np.uint32(s0)).astype(np.uint32),
                         self.exc_base[e0:e1].copy(), self.exc_qual[e0:e1].copy())


def encode_cigar(cigar: Sequence) -> List[int]:
    return [(int(ln) << 4) | _OPCODE[op] for op, ln in cigar]


def record_flags(r) -> int:
    f = 0
    if r.paired:
        f |= capi.PB_F_PAIRED
    if r.proper:
        f |= capi.PB_F_PROPER
    if r.mate_same_ref:
        f |= capi.PB_F_MATE_SAME_REF
    if len(r.quals) > 0:
        f |= capi.PB_F_HAS_QUALS
    if getattr(r, "unmapped", False):
        f |= capi.PB_F_UNMAPPED
    if getattr(r, "reverse", False):
        f |= capi.PB_F_REVERSE
    return f


def pack_soa(pos, tlen, mapq, flags, read_len, cigar_off, cigar, ascii_off, seq, qual) -> ReadBatch:
    """Vectorised packer: reads given as struct-of-arrays with one ASCII byte per base (unpadded,
    read r occupying seq[ascii_off[r] : ascii_off[r]+read_len[r]]) and raw quality bytes (ignored
    for reads without PB_F_HAS_QUALS)."""
    pos = np.ascontiguousarray(pos, np.int32)
    n = pos.shape[0]
    read_len = np.ascontiguousarray(read_len, np.int32)
    flags = np.ascontiguousarray(flags, np.uint8)
    padded = (read_len.astype(np.int64) + 3) & ~3
    seq_off64 = np.zeros(n + 1, np.int64)
    np.cumsum(padded, out=seq_off64[1:])
    n_seq = int(seq_off64[-1])
    if n_seq >= 2 ** 32:
        raise ValueError("batch too large: split it (n_seq must be < 2^32)")
    seq = np.ascontiguousarray(seq, np.uint8)
    qual = np.ascontiguousarray(qual, np.uint8)
    ascii_off = np.ascontiguousarray(ascii_off, np.int64)
    # destination index of every source base
    total = int(read_len.sum())
    rid = np.repeat(np.arange(n, dtype=np.int64), read_len)
    within = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(read_len.astype(np.int64)) - read_len, read_len)
    src = ascii_off[rid] + within
    dst = seq_off64[rid] + within
    letters = seq[src]
    hasq = (flags[rid] & capi.PB_F_HAS_QUALS) != 0
    q = np.where(hasq, qual[src] if qual.shape[0] else np.zeros(total, np.uint8), 0).astype(np.uint8)
    code = _BASE_CODE[letters]
    exc = (code == 255) | (q >= 128)
    quals = np.zeros(n_seq, np.uint8)
    quals[dst] = np.where(exc, 0x80, q)
    code2 = np.where(code == 255, 0, code).astype(np.uint8)
    codes_full = np.zeros(n_seq, np.uint8)
    codes_full[dst] = code2
    c4 = codes_full.reshape(-1, 4)
    bases2 = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    return ReadBatch(pos, np.ascontiguousarray(tlen, np.int32), read_len,
                     np.ascontiguousarray(mapq, np.uint8), flags,
                     np.ascontiguousarray(cigar_off, np.uint32), np.ascontiguousarray(cigar, np.uint32),
                     seq_off64[:-1].astype(np.uint32), quals, bases2,
                     dst[exc].astype(np.uint32), letters[exc].copy(), q[exc].copy())


def pack_records(records: Iterable) -> ReadBatch:
    """Pack SAMRecord-like objects (attributes: pos, cigar [(op, len)], bases, quals, mapq, paired,
    proper, mate_same_ref, tlen, unmapped, reverse).  Records must already be sorted by pos."""
    recs = list(records)
    n = len(recs)
    pos = np.array([r.pos for r in recs], np.int32).reshape(n)
    tlen = np.array([r.tlen for r in recs], np.int32).reshape(n)
    mapq = np.array([r.mapq for r in recs], np.uint8).reshape(n)
    flags = np.array([record_flags(r) for r in recs], np.uint8).reshape(n)
    read_len = np.array([len(r.bases) for r in recs], np.int32).reshape(n)
    cig: List[int] = []
    cigar_off = np.zeros(n + 1, np.uint32)
    for i, r in enumerate(recs):
        cig.extend(encode_cigar(r.cigar))
        cigar_off[i + 1] = len(cig)
    ascii_off = np.zeros(n, np.int64)
    if n:
        ascii_off[1:] = np.cumsum(read_len[:-1].astype(np.int64))
    seq = np.frombuffer(b"".join(bytes(r.bases) for r in recs), np.uint8)
    qual = np.frombuffer(b"".join(bytes(r.quals) if len(r.quals) else bytes(len(r.bases)) for r in recs), np.uint8)
    return pack_soa(pos, tlen, mapq, flags, read_len, cigar_off, np.array(cig, np.uint32), ascii_off, seq, qual)


class ResultBuffers:
    """Caller-owned host arrays behind a `pb_region_result`."""

    def __init__(self, size: int, planes: Optional[Sequence[str]] = None, indels_cap: int = 0,
                 indel_bytes_cap: int = 0, pinned: bool = False, batch_cap: int = 256, calls_cap: int = 0):
        self.size = size
        self.arrays = {}
        self.c = capi.pb_region_result()
        want = set(planes) if planes is not None else {p[0] for p in capi.RESULT_PLANES}
        for name, dt, per in capi.RESULT_PLANES:
            if name in want:
                arr = _alloc(size * per, np.dtype(dt), pinned)
                self.arrays[name] = arr
                setattr(self.c, name, arr.ctypes.data)
        self.indels_cap, self.indel_bytes_cap = indels_cap, indel_bytes_cap
        if indels_cap:
            self._indels = (capi.pb_indel * indels_cap)()
            self.c.indels = C.addressof(self._indels)
            self.c.indels_cap = indels_cap
        if indel_bytes_cap:
            self.indel_bytes = np.zeros(indel_bytes_cap, np.uint8)
            self.c.indel_bytes = self.indel_bytes.ctypes.data
            self.c.indel_bytes_cap = indel_bytes_cap
        # per-BAM deltas (BamFile.scala:120-122,142-146)
        self.batch_read_count = np.zeros(batch_cap, np.int32)
        self.batch_base_count = np.zeros(batch_cap, np.int64)
        self.batch_coverage = np.zeros(batch_cap, np.int64)
        self.c.batch_read_count = self.batch_read_count.ctypes.data
        self.c.batch_base_count = self.batch_base_count.ctypes.data
        self.c.batch_coverage = self.batch_coverage.ctypes.data
        self.c.batch_cap = batch_cap
        # the call plane in sparse form (pb_region_result.calls): changed / ambiguous loci only
        self.calls_cap = calls_cap
        if calls_cap:
            self._calls = _alloc(calls_cap, np.dtype(capi.CALL_ENTRY_DTYPE), pinned)
            self.c.calls = self._calls.ctypes.data
            self.c.calls_cap = calls_cap

    def calls(self) -> np.ndarray:
        """The entries written (locus_index, flags, call); raises if the capacity was too small."""
        n = int(self.c.n_calls)
        if n > self.calls_cap:
            raise ValueError("calls_cap %d < %d changed / ambiguous loci" % (self.calls_cap, n))
        return self._calls[:n]

    def __getitem__(self, name: str) -> np.ndarray:
        a = self.arrays[name]
        per = {p[0]: p[2] for p in capi.RESULT_PLANES}[name]
        return a.reshape(self.size, per) if per > 1 else a

    def per_bam(self):
        """[(nReads, baseCount delta, meanCoverage)] per pb_region_add_batch call (BamFile.scala:142-147)."""
        n = min(int(self.c.n_batches), int(self.c.batch_cap))
        return [(int(self.batch_read_count[b]), int(self.batch_base_count[b]), int(self.batch_coverage[b])) for b in range(n)]

    def indels(self):
        out = []
        n = min(int(self.c.n_indels), self.indels_cap)
        for k in range(n):
            e = self._indels[k]
            s = bytes(self.indel_bytes[e.str_off:e.str_off + e.win_len]) if self.indel_bytes_cap else b""
            out.append(dict(locus_index=e.locus_index, kind=e.kind, list_len=e.list_len, win_count=e.win_count,
                            win_len=e.win_len, win_has_n=e.win_has_n, string=s))
        return out


def _alloc(n: int, dt: np.dtype, pinned: bool) -> np.ndarray:
    if pinned:
        import torch
        t = torch.empty(max(n, 1) * dt.itemsize, dtype=torch.uint8, pin_memory=True)
        arr = t.numpy().view(dt)[:n]   # the ndarray keeps the pinned tensor alive through .base
        arr.fill(0)
        return arr
    return np.zeros(n, dt)
