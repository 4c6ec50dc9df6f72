"""ctypes mirror of include/pilon_b200.h and the loader for libpilonb200.so.

The product path fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libpilonb200.so")

PB_OK = 0
ABI_VERSION = 6
PB_ERR_INVALID, PB_ERR_CUDA, PB_ERR_UNSORTED, PB_ERR_UNSUPPORTED, PB_ERR_OOM, PB_ERR_HASH = -1, -2, -3, -4, -5, -6

PB_F_PAIRED, PB_F_PROPER, PB_F_MATE_SAME_REF, PB_F_HAS_QUALS, PB_F_UNMAPPED, PB_F_REVERSE = 1, 2, 4, 8, 16, 32
PB_MEM_HOST, PB_MEM_DEVICE = 0, 1

PB_FL_CONFIRMED, PB_FL_CHANGED, PB_FL_AMBIGUOUS, PB_FL_DELETED, PB_FL_KIND_SHIFT = 1, 2, 4, 8, 4
PB_KIND_SNP, PB_KIND_INS, PB_KIND_DEL, PB_KIND_AMB = 0, 1, 2, 3

CIGAR_OPS = "MIDNSHP=X"


class pb_config(C.Structure):
    _fields_ = [("min_qual", C.c_int32), ("min_mq", C.c_int32), ("flank", C.c_int32),
                ("default_qual", C.c_int32), ("min_min_depth", C.c_int32), ("old_indel", C.c_int32),
                ("fix_amb", C.c_int32), ("reserved", C.c_int32), ("min_depth", C.c_double)]


class pb_batch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_cigar", C.c_int64), ("n_seq", C.c_int64), ("n_exc", C.c_int64),
                ("pos", C.c_void_p), ("tlen", C.c_void_p), ("read_len", C.c_void_p),
                ("mapq", C.c_void_p), ("flags", C.c_void_p), ("cigar_off", C.c_void_p),
                ("cigar", C.c_void_p), ("seq_off", C.c_void_p), ("quals", C.c_void_p),
                ("bases2", C.c_void_p), ("exc_idx", C.c_void_p), ("exc_base", C.c_void_p),
                ("exc_qual", C.c_void_p), ("mem", C.c_int32), ("qual_code_bits", C.c_int32),
                ("qual_codes", C.c_void_p), ("qual_lut", C.c_uint8 * 16),
                ("base_delta_idx", C.c_void_p), ("base_delta_code", C.c_void_p), ("n_base_delta", C.c_int64),
                ("meta_codes", C.c_void_p), ("meta_cigar", C.c_void_p), ("meta_esc", C.c_void_p),
                ("n_meta_cigar", C.c_int64), ("n_meta_esc", C.c_int64), ("meta_pos0", C.c_int32), ("meta_seq_stride", C.c_int32)]


class pb_indel(C.Structure):
    _fields_ = [("locus_index", C.c_int32), ("kind", C.c_int32), ("list_len", C.c_int32),
                ("win_count", C.c_int32), ("win_len", C.c_int32), ("win_has_n", C.c_int32),
                ("str_off", C.c_int64)]


class pb_region_result(C.Structure):
    _fields_ = [("size", C.c_int64), ("base_count", C.c_int64), ("coverage", C.c_int64),
                ("aligned_bases", C.c_int64), ("read_count", C.c_int32), ("min_depth", C.c_int32),
                ("unknown_ops", C.c_int32), ("dropped_oob", C.c_int32),
                ("n_indels", C.c_int64), ("n_indel_bytes", C.c_int64),
                ("base_count4", C.c_void_p), ("qual_sum4", C.c_void_p), ("mq_sum", C.c_void_p),
                ("q_sum", C.c_void_p), ("phys_cov", C.c_void_p), ("insert_size", C.c_void_p),
                ("bad_pair", C.c_void_p), ("deletions", C.c_void_p), ("del_qual", C.c_void_p),
                ("insertions", C.c_void_p), ("ins_qual", C.c_void_p), ("clips", C.c_void_p),
                ("coverage_arr", C.c_void_p), ("frag_coverage", C.c_void_p),
                ("weighted_qual", C.c_void_p), ("weighted_mq", C.c_void_p), ("flags", C.c_void_p),
                ("call", C.c_void_p),
                ("indels", C.c_void_p), ("indels_cap", C.c_int64),
                ("indel_bytes", C.c_void_p), ("indel_bytes_cap", C.c_int64),
                ("batch_read_count", C.c_void_p), ("batch_base_count", C.c_void_p), ("batch_coverage", C.c_void_p),
                ("batch_cap", C.c_int64), ("n_batches", C.c_int64),
                ("calls", C.c_void_p), ("calls_cap", C.c_int64), ("n_calls", C.c_int64)]


# pb_call_entry as a numpy record
CALL_ENTRY_DTYPE = [("locus_index", "<i4"), ("flags", "<u4"), ("call", "<u8")]


# name -> numpy dtype, elements per locus; order follows the struct
RESULT_PLANES = [("base_count4", "i4", 4), ("qual_sum4", "i8", 4), ("mq_sum", "i4", 1), ("q_sum", "i4", 1),
                 ("phys_cov", "i4", 1), ("insert_size", "i4", 1), ("bad_pair", "i4", 1), ("deletions", "i4", 1),
                 ("del_qual", "i4", 1), ("insertions", "i4", 1), ("ins_qual", "i4", 1), ("clips", "i4", 1),
                 ("coverage_arr", "i4", 1), ("frag_coverage", "i4", 1), ("weighted_qual", "i1", 1),
                 ("weighted_mq", "i1", 1), ("flags", "u1", 1), ("call", "u8", 1)]

_lib = None


class EngineLibraryMissing(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libpilonb200.so (built in-tree by __graft_entry__.build / pilon_b200/csrc/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineLibraryMissing(
            "%s not found: build it with `python -m pilon_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for the engine." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.pb_abi_version.restype = C.c_int
    lib.pb_last_error.restype = C.c_char_p
    lib.pb_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.pb_create.argtypes = [C.c_int, C.POINTER(pb_config), C.POINTER(vp)]
    lib.pb_destroy.argtypes = [vp]
    lib.pb_region_begin.argtypes = [vp, vp, i64, i32, i32]
    lib.pb_region_add_batch.argtypes = [vp, C.POINTER(pb_batch), C.c_int, C.c_int]
    lib.pb_region_finish.argtypes = [vp, C.POINTER(pb_region_result), vp]
    lib.pb_region_compute_timed.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                            C.POINTER(i64)]
    lib.pb_region_compute.argtypes = [vp]
    lib.pb_stream.argtypes = [vp, C.POINTER(vp)]
    lib.pb_packer_create.argtypes = [C.POINTER(vp)]
    lib.pb_packer_destroy.argtypes = [vp]
    lib.pb_packer_reset.argtypes = [vp]
    lib.pb_packer_add.argtypes = [vp, i32, i32, i32, C.c_uint32, vp, i32, vp, vp, i32]
    lib.pb_packer_add_many.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pb_packer_view.argtypes = [vp, C.POINTER(pb_batch)]
    lib.pb_base_delta_encode.argtypes = [C.POINTER(pb_batch), vp, i64, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    lib.pb_base_delta_encode.restype = C.c_int
    lib.pb_meta_encode.argtypes = [C.POINTER(pb_batch), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64),
                                   C.POINTER(i32), C.POINTER(i32)]
    lib.pb_meta_encode.restype = C.c_int
    lib.pb_free.argtypes = [vp]
    lib.pb_free.restype = None
    for name in ("pb_device_count", "pb_create", "pb_destroy", "pb_region_begin", "pb_region_add_batch",
                 "pb_region_finish", "pb_region_compute_timed", "pb_region_compute", "pb_stream", "pb_packer_create",
                 "pb_packer_destroy", "pb_packer_reset", "pb_packer_add", "pb_packer_add_many",
                 "pb_packer_view", "pb_packer_add_bam", "pb_base_delta_encode", "pb_meta_encode"):
        getattr(lib, name).restype = C.c_int
    if lib.pb_abi_version() != ABI_VERSION:
        raise RuntimeError("libpilonb200.so ABI version mismatch")
    _lib = lib
    return lib


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("pilon_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc: int):
    if rc != PB_OK:
        msg = load_library().pb_last_error()
        raise EngineError(rc, msg.decode() if msg else "")
