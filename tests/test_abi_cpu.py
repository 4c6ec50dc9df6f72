"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, and the host packer agrees with the numpy packer.  No compute calls (no GPU here)."""
import ctypes as C
import os
import random
import re

import numpy as np

from pilon_b200 import _capi as capi
from pilon_b200.packing import pack_records
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "pilon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libpilonb200.so does not export %s" % n
    assert lib.pb_abi_version() == capi.ABI_VERSION == 6


def test_struct_sizes_match_header():
    # sizes implied by the header's field lists (LP64)
    assert C.sizeof(capi.pb_config) == 40
    assert C.sizeof(capi.pb_batch) == 4 * 8 + 13 * 8 + 8 + 8 + 16 + 24 + 3 * 8 + 2 * 8 + 2 * 4
    assert C.sizeof(capi.pb_indel) == 32
    assert C.sizeof(capi.pb_region_result) == 4 * 8 + 4 * 4 + 2 * 8 + 18 * 8 + 4 * 8 + 5 * 8 + 3 * 8


def test_engine_rejects_bad_arguments_without_a_gpu():
    lib = capi.load_library()
    h = C.c_void_p()
    cfg = capi.pb_config(min_qual=-1, default_qual=10, flank=10, min_min_depth=5, min_depth=0.1)
    assert lib.pb_create(0, C.byref(cfg), C.byref(h)) == capi.PB_ERR_UNSUPPORTED
    assert b"min_qual" in lib.pb_last_error()


def test_c_packer_matches_numpy_packer():
    lib = capi.load_library()
    for seed in range(5):
        contig, start, stop, reads = H.random_case(seed)
        want = pack_records(reads)
        pk = C.c_void_p()
        assert lib.pb_packer_create(C.byref(pk)) == 0
        for r in reads:
            cig = np.array([(l << 4) | capi.CIGAR_OPS.index(op) for op, l in r.cigar], np.uint32)
            seq = np.frombuffer(r.bases, np.uint8)
            # BAM stores missing qualities as 0xFF bytes
            q = np.frombuffer(r.quals if len(r.quals) else b"\xff" * len(r.bases), np.uint8)
            flags = 0
            for bit, on in ((1, r.paired), (2, r.proper), (4, r.mate_same_ref), (16, r.unmapped), (32, r.reverse)):
                flags |= bit if on else 0
            assert lib.pb_packer_add(pk, r.pos, r.tlen, r.mapq, flags, cig.ctypes.data, len(cig),
                                     seq.ctypes.data, q.ctypes.data, len(seq)) == 0
        b = capi.pb_batch()
        assert lib.pb_packer_view(pk, C.byref(b)) == 0
        assert (b.n_reads, b.n_cigar, b.n_seq, b.n_exc) == (want.n_reads, len(want.cigar), len(want.quals), len(want.exc_idx))

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(max(n, 1),))[:n].copy()
        assert np.array_equal(arr(b.pos, b.n_reads, C.c_int32), want.pos)
        assert np.array_equal(arr(b.flags, b.n_reads, C.c_uint8), want.flags)
        assert np.array_equal(arr(b.cigar, b.n_cigar, C.c_uint32), want.cigar)
        assert np.array_equal(arr(b.seq_off, b.n_reads, C.c_uint32), want.seq_off)
        assert np.array_equal(arr(b.quals, b.n_seq, C.c_uint8), want.quals)
        assert np.array_equal(arr(b.bases2, b.n_seq // 4, C.c_uint8), want.bases2)
        assert np.array_equal(arr(b.exc_idx, b.n_exc, C.c_uint32), want.exc_idx)
        assert np.array_equal(arr(b.exc_base, b.n_exc, C.c_uint8), want.exc_base)
        assert np.array_equal(arr(b.exc_qual, b.n_exc, C.c_uint8), want.exc_qual)
        lib.pb_packer_destroy(pk)


def _decode_codes(b):
    bits = b.qual_code_bits
    nbytes = (b.n_seq * bits + 7) // 8
    raw = np.ctypeslib.as_array(C.cast(b.qual_codes, C.POINTER(C.c_uint8)), shape=(nbytes,))
    stream = np.unpackbits(raw, bitorder="little")[:b.n_seq * bits].reshape(b.n_seq, bits)
    codes = (stream * (1 << np.arange(bits, dtype=np.uint8))).sum(axis=1).astype(np.uint8)
    return np.array(list(b.qual_lut), np.uint8)[codes]


def test_packer_offers_packed_quality_transport_only_for_small_alphabets():
    """pb_batch.qual_codes / qual_code_bits / qual_lut (include/pilon_b200.h): offered iff the stored quality bytes take
    <= 16 values (3-bit codes up to 8 values, 4-bit up to 16), and then it decodes to exactly `quals`;
    packing.ReadBatch.with_packed_quals is the numpy twin."""
    lib = capi.load_library()
    rng = np.random.default_rng(5)
    for n_values, want_bits in ((4, 3), (6, 3), (7, 4), (14, 4), (16, 0), (40, 0)):   # + padding byte 0 and the 0x80 mark
        pk = C.c_void_p()
        assert lib.pb_packer_create(C.byref(pk)) == 0
        alphabet = rng.choice(np.arange(1, 94), n_values, replace=False).astype(np.uint8)
        for r in range(40):
            L = int(rng.integers(1, 90)) if r else 95
            seq = rng.choice(np.frombuffer(b"ACGTN", np.uint8), L, p=[.24, .24, .24, .24, .04]).astype(np.uint8)
            if r == 0:
                seq[-1] = ord("N")                                               # the 0x80 mark is certainly present
            q = rng.choice(alphabet, L).astype(np.uint8)
            if r == 0:
                q[:n_values] = alphabet                                          # every value is certainly present
            cig = np.array([(L << 4) | 0], np.uint32)
            assert lib.pb_packer_add(pk, 100 + r, 0, 60, 0, cig.ctypes.data, 1, seq.ctypes.data, q.ctypes.data, L) == 0
        b = capi.pb_batch()
        assert lib.pb_packer_view(pk, C.byref(b)) == 0
        q8 = np.ctypeslib.as_array(C.cast(b.quals, C.POINTER(C.c_uint8)), shape=(b.n_seq,)).copy()
        assert (b.qual_code_bits if b.qual_codes else 0) == want_bits
        if b.qual_codes:
            assert np.array_equal(_decode_codes(b), q8)
        lib.pb_packer_destroy(pk)

    from pilon_b200.packing import pack_records
    from tests import helpers as H
    for seed in (3, 4):
        contig, start, stop, reads = H.random_case(seed)
        rb = pack_records(reads).with_packed_quals()
        if rb.qual_codes is not None:
            assert np.array_equal(_decode_codes(rb.to_c()), rb.quals)


def test_base_delta_transport_round_trips():
    """pb_base_delta_encode (include/pilon_b200.h, pb_batch.base_delta_idx): the deltas applied to an independent Python
    restatement of the reference prediction give back bases2 exactly."""
    from pilon_b200.packing import pack_records
    from tests import helpers as H
    halo = 16384
    table = {ord("C"): 1, ord("G"): 2, ord("T"): 3}
    for seed in (0, 1, 5, 9, 14):
        contig, start, stop, reads = H.random_case(seed)
        rb = pack_records(reads)
        rd = rb.with_base_deltas(contig, start, stop)
        n = rd.base_delta_idx.shape[0] - 16
        assert np.all(np.diff(rd.base_delta_idx[:n].astype(np.int64)) > 0)
        lo, hi = max(1, start - halo), min(len(contig), stop + halo)
        codes = np.zeros(rb.quals.shape[0], np.uint8)
        for r in range(rb.n_reads):
            bi, L, done, locus = int(rb.seq_off[r]), int(rb.read_len[r]), 0, int(rb.pos[r])
            for k in range(int(rb.cigar_off[r]), int(rb.cigar_off[r + 1])):
                e = int(rb.cigar[k])
                op, ln = e & 15, e >> 4
                if op in (0, 7, 8):
                    for j in range(ln):
                        if done >= L:
                            break
                        l = locus + j
                        codes[bi + done] = table.get(contig[l - 1], 0) if lo <= l <= hi else 0
                        done += 1
                    locus += ln
                elif op in (1, 4):
                    done = min(L, done + ln)
                elif op in (2, 3):
                    locus += ln
        codes[rd.base_delta_idx[:n]] = rd.base_delta_code[:n]
        c4 = codes.reshape(-1, 4)
        assert np.array_equal((c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8), rb.bases2)


def test_compact_metadata_round_trips_through_an_independent_decoder():
    """pb_meta_encode (pb_batch.meta_codes): 8 bytes per read + listed CIGARs + escapes decode, by a numpy decoder written
    from the header's description, to exactly the eight plain arrays -- including reads with indels and clips, a gap of more
    than 65534 loci between neighbours and template lengths beyond int16."""
    import random
    from pilon_b200.packing import meta_decode, meta_encode, pack_records
    from tests import helpers as H
    contig, start, stop, reads = H.random_case(3, contig_len=2000, n_reads=500)
    reads = sorted(reads, key=lambda r: r.pos)
    rng = random.Random(3)
    for r in reads[::7]:
        r.tlen = rng.choice([-70000, 40000, 32767, -32768, -32767, 1 << 30])
    far = [r for r in reads[-20:]]
    for r in far:
        r.pos += 200000                                      # > 65534 past the previous read
    reads.sort(key=lambda r: r.pos)
    rb = pack_records(reads)
    m = meta_encode(rb.to_c())
    assert m is not None
    codes, cigar, esc, ng, ne, pos0, stride = m
    assert ne >= len(reads[::7]) // 2 and stride == 0 and ng > 0
    got = meta_decode(codes, cigar, esc, ne, pos0, stride, rb.n_reads)
    for name, g in zip(("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off"), got):
        assert np.array_equal(g, getattr(rb, name)), name
    # 8 B per read + the listed CIGARs instead of 22 B per read + every CIGAR
    assert codes[:rb.n_reads].nbytes + 4 * ng + 12 * ne < 0.7 * sum(getattr(rb, f).nbytes for f in ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off"))


def test_compact_metadata_refuses_what_it_cannot_hold():
    from oracle import pilon_oracle as po
    from pilon_b200.packing import meta_encode, pack_records
    long_read = po.Read(pos=5, cigar=[("M", 300)], bases=b"A" * 300, quals=bytes([30]) * 300, mapq=60)
    assert meta_encode(pack_records([long_read]).to_c()) is None
    a = po.Read(pos=50, cigar=[("M", 20)], bases=b"C" * 20, quals=bytes([30]) * 20, mapq=60)
    b = po.Read(pos=10, cigar=[("M", 20)], bases=b"C" * 20, quals=bytes([30]) * 20, mapq=60)
    assert meta_encode(pack_records([a, b]).to_c()) is None     # not sorted by pos
    assert meta_encode(pack_records([b, a]).to_c()) is not None
    assert meta_encode(pack_records([]).to_c()) is not None
