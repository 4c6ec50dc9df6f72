"""Synthetic workloads for BASELINE.json's five configs: deterministic genome +
coordinate-sorted packed read batches, generated region by region by csrc/pb_synth.cpp.

Chunking follows GenomeFile.contigRegions (reference GenomeFile.scala:67-74); each chunk's batch
holds the reads whose start lies within +-10 kb of the chunk (BamFile.scala:118-119).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Tuple

import numpy as np

from . import _capi as capi
from .packing import ReadBatch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "pb_synth.cpp")
LIB = os.path.join(HERE, "csrc", "libpilonsynth.so")

CHUNK_SIZE = 10_000_000      # Pilon.chunkSize (Pilon.scala:52)
HALO = 10_000                # BamFile.scala:118-119


class ps_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("contig_len", C.c_int64), ("chunk_size", C.c_int64),
                ("read_len", C.c_int32), ("snp_block", C.c_int32), ("indel_block", C.c_int32),
                ("n_period", C.c_int32), ("depth", C.c_double), ("ins_mean", C.c_double), ("ins_sd", C.c_double),
                ("het_frac", C.c_double), ("sub_err", C.c_double), ("indel_err", C.c_double),
                ("clip_frac", C.c_double), ("improper_frac", C.c_double), ("lowmq_frac", C.c_double),
                ("n_frac", C.c_double), ("lib_seed", C.c_uint64)]


_lib = None


def build(force: bool = False) -> str:
    hdr = os.path.join(os.path.dirname(HERE), "include", "pilon_b200.h")      # pb_batch's layout lives there
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC])
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.ps_contig.argtypes = [C.POINTER(ps_params), C.c_int64, C.c_int64, C.c_void_p]
        l.ps_contig.restype = None
        l.ps_generate.argtypes = [C.POINTER(ps_params), C.c_int64, C.c_int64]
        l.ps_generate.restype = C.c_void_p
        l.ps_view.argtypes = [C.c_void_p, C.POINTER(capi.pb_batch), C.POINTER(C.c_int64)]
        l.ps_view.restype = None
        l.ps_free.argtypes = [C.c_void_p]
        l.ps_free.restype = None
        _lib = l
    return _lib


def chunks_of(contig_len: int, chunk_size: int = CHUNK_SIZE) -> List[Tuple[int, int]]:
    """GenomeFile.contigRegions (GenomeFile.scala:67-74): equal chunks of at most chunk_size."""
    n = (contig_len + chunk_size - 1) // chunk_size
    cs = (contig_len + n - 1) // n
    return [(b, min(contig_len, b + cs - 1)) for b in range(1, contig_len + 1, cs)]


@dataclass
class Library:
    name: str            # "frags" | "jumps"
    depth: float
    ins_mean: float
    ins_sd: float
    lib_seed: int

    @property
    def counts_toward_frag_coverage(self) -> bool:   # GenomeRegion.scala:291,296
        return self.name != "jumps"


@dataclass
class Workload:
    name: str
    seed: int
    contig_lens: List[int]
    libraries: List[Library]
    het_frac: float = 0.0
    read_len: int = 150
    chunk_size: int = CHUNK_SIZE
    description: str = ""

    def regions(self) -> List[Tuple[int, int, int]]:
        """[(contig index, start, stop)] in the order Pilon processes them."""
        out = []
        for ci, n in enumerate(self.contig_lens):
            out += [(ci, a, b) for a, b in chunks_of(n, self.chunk_size)]
        return out

    def params(self, contig: int, libr: Library) -> ps_params:
        n = self.contig_lens[contig]
        nch = (n + self.chunk_size - 1) // self.chunk_size
        return ps_params(seed=(self.seed << 8) + contig + 1, contig_len=n, chunk_size=(n + nch - 1) // nch,
                         read_len=self.read_len, snp_block=1000, indel_block=5000, n_period=20000,
                         depth=libr.depth, ins_mean=libr.ins_mean, ins_sd=libr.ins_sd, het_frac=self.het_frac,
                         sub_err=0.002, indel_err=0.0001, clip_frac=0.01, improper_frac=0.01, lowmq_frac=0.05,
                         n_frac=0.001, lib_seed=libr.lib_seed)

    def contig_bases(self, contig: int, lo: int = 1, hi: Optional[int] = None) -> np.ndarray:
        p = self.params(contig, self.libraries[0])
        hi = hi or self.contig_lens[contig]
        out = np.empty(hi - lo + 1, np.uint8)
        lib().ps_contig(C.byref(p), lo, hi, out.ctypes.data)
        return out

    def region_batches(self, contig: int, start: int, stop: int) -> List["SynthBatch"]:
        n = self.contig_lens[contig]
        lo, hi = max(start - HALO, 1), min(stop + HALO, n)
        return [SynthBatch(self.params(contig, l), lo, hi, l.counts_toward_frag_coverage) for l in self.libraries]

    @property
    def total_loci(self) -> int:
        return sum(self.contig_lens)


class SynthBatch:
    """Owns one generated batch (host memory inside libpilonsynth) and views it as a pb_batch."""

    def __init__(self, p: ps_params, lo: int, hi: int, frag: bool):
        self._h = lib().ps_generate(C.byref(p), lo, hi)
        self.frag = frag
        self.c = capi.pb_batch()
        al = C.c_int64()
        lib().ps_view(self._h, C.byref(self.c), C.byref(al))
        self.aligned_bases = al.value
        self.n_reads = self.c.n_reads

    def to_c(self) -> capi.pb_batch:
        return self.c

    def nbytes(self) -> int:
        c = self.c
        return int(c.n_reads * 14 + (c.n_reads + 1) * 4 + c.n_cigar * 4 + c.n_reads * 4 + c.n_seq + c.n_seq // 4 + c.n_exc * 6)

    def as_read_batch(self) -> ReadBatch:
        c = self.c

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
        return ReadBatch(arr(c.pos, c.n_reads, np.int32), arr(c.tlen, c.n_reads, np.int32),
                         arr(c.read_len, c.n_reads, np.int32), arr(c.mapq, c.n_reads, np.uint8),
                         arr(c.flags, c.n_reads, np.uint8), arr(c.cigar_off, c.n_reads + 1, np.uint32),
                         arr(c.cigar, c.n_cigar, np.uint32), arr(c.seq_off, c.n_reads, np.uint32),
                         arr(c.quals, c.n_seq, np.uint8), arr(c.bases2, c.n_seq // 4, np.uint8),
                         arr(c.exc_idx, c.n_exc, np.uint32), arr(c.exc_base, c.n_exc, np.uint8),
                         arr(c.exc_qual, c.n_exc, np.uint8))

    def close(self):
        if self._h:
            lib().ps_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


FRAGS = lambda depth: Library("frags", depth, 400.0, 40.0, 1)      # noqa: E731
JUMPS = lambda depth: Library("jumps", depth, 3000.0, 300.0, 2)    # noqa: E731

_C2_CONTIGS = [int(x * 1_000_000) for x in
               (10, 8, 6, 5, 4, 3, 2.5, 2, 1.5, 1.2, 1, 0.9, 0.8, 0.7, 0.6, 0.6, 0.55, 0.55, 0.55, 0.55)]


def workload(name: str, scale: float = 1.0) -> Workload:
    """The five BASELINE.json configs; `scale` shrinks the genome (not the depth) for quick runs."""
    s = lambda n: max(20_000, int(n * scale))   # noqa: E731
    if name == "C1":
        d = int(os.environ.get("PB_SYNTH_DEPTH", "100"))      # kernel-selection experiments only (tools/ab_kernels.sh): says so in the name
        return Workload("C1", 1, [s(5_000_000)], [FRAGS(d)],
                        description="5 Mb bacterial-like, 100x 2x150" if d == 100 else "NOT A BASELINE CONFIG: C1 genome at %dx" % d)
    if name == "C2":
        return Workload("C2", 2, [s(n) for n in _C2_CONTIGS], [FRAGS(60), JUMPS(10)],
                        description="50 Mb fungal-scale, 20 contigs, 60x frags + 10x jumps")
    if name == "C3":
        return Workload("C3", 3, [s(64_000_000)], [FRAGS(30)], het_frac=0.5,
                        description="64 Mb chr20-like, 30x, het SNPs, --vcf")
    if name == "C4":
        return Workload("C4", 4, [s(125_000_000)] * 24, [FRAGS(30)], description="3 Gb human-scale, 24 contigs, 30x")
    if name == "C5":
        return Workload("C5", 5, [s(200_000)], [FRAGS(5000)], description="200 kb amplicon at 5000x")
    raise ValueError(name)
