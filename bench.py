#!/usr/bin/env python
"""bench.py -- aligned pileup bases/sec of the pileup + BaseCall hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--scale 1.0]
  python bench.py --impl reference ...      # the CPU restatement of the reference path on host cores

A "step" is one pass of the whole hot path (prep -> physCov scan -> indel grouping -> pileup+BaseCall
-> deletion spill) over every region of the workload.  `value` is measured with the packed read
batches already resident in HBM (CUDA events on the engine streams); `e2e` runs the same regions
through the public C ABI from pinned HOST buffers, host->device copies of the reads and the
device->host copy of the per-locus results inside the timed region.

Under torchrun each rank owns one GPU and its own copy-sized workload (regions are independent, no
data-path collective): scaling is weak, the job value is total bases / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "aligned pileup bases/sec"
UNIT = "bases/s"
# planes the Scala driver reads back for `--fix snps,indels --changes` (GenomeRegion.scala:247-253 +
# flags + call record for the change list); --vcf needs every counter
FIX_PLANES = ["coverage_arr", "bad_pair", "phys_cov", "insert_size", "weighted_qual", "weighted_mq", "clips",
              "frag_coverage", "flags", "call"]
# what the fix path itself consumes (pb_out_create: pass 2 + identifyAndFixIssues, GenomeRegion.scala:275-283, 307-380):
# flags, fragCoverage and the call records of the changed / ambiguous loci (pb_region_result.calls) -- the other
# per-locus arrays of :247-253 only feed --tracks and --vcf.  tests/test_output_cpu.py / test_output_gpu.py check that the
# FASTA, the change list and the log from this set equal those from every plane.
FIXMIN_PLANES = ["flags", "frag_coverage"]
CALLS_CAP = 1 << 20


def algorithmic_bytes(aligned: int, n_reads: int, n_cigar: int, loci: int) -> float:
    """per aligned base 1.25 B (2-bit base + quality) + per read 24 B + 4 B per CIGAR op
    + per locus 88 B counters out + 1 B reference in + 8 B call record out."""
    return 1.25 * aligned + 24.0 * n_reads + 4.0 * n_cigar + 97.0 * loci


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines, self.first = gpu_index, None, [], 0

    def mark(self):
        """Samples before this point (warm-up) are ignored."""
        self.first = len(self.lines)

    def count(self) -> int:
        return len(self.lines) - self.first

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
class Region:
    def __init__(self, wl, ci, start, stop, contig_np):
        self.ci, self.start, self.stop = ci, start, stop
        self.size = stop + 1 - start
        self.contig = contig_np                      # whole contig, uint8
        self.batches = wl.region_batches(ci, start, stop)
        self.aligned = sum(b.aligned_bases for b in self.batches)
        self.n_reads = sum(b.n_reads for b in self.batches)
        self.n_cigar = sum(b.c.n_cigar for b in self.batches)
        self.alg_bytes = algorithmic_bytes(self.aligned, self.n_reads, self.n_cigar, self.size)


def build_workload(name, scale, seed_shift, threads):
    from pilon_b200 import synth
    os.environ.setdefault("OMP_NUM_THREADS", str(max(1, threads)))
    wl = synth.workload(name, scale)
    wl.seed += 1000 * seed_shift
    contigs = {}
    regions = []
    for ci, a, b in wl.regions():
        if ci not in contigs:
            contigs[ci] = wl.contig_bases(ci)
        regions.append(Region(wl, ci, a, b, contigs[ci]))
    return wl, regions


_BATCH_FIELDS = [("pos", "n_reads", 4), ("tlen", "n_reads", 4), ("read_len", "n_reads", 4), ("mapq", "n_reads", 1),
                 ("flags", "n_reads", 1), ("cigar_off", "n_reads+1", 4), ("cigar", "n_cigar", 4),
                 ("seq_off", "n_reads", 4), ("quals", "n_seq", 1), ("bases2", "n_seq/4", 1),
                 ("exc_idx", "n_exc", 4), ("exc_base", "n_exc", 1), ("exc_qual", "n_exc", 1)]


_META_FIELDS = ("pos", "tlen", "read_len", "mapq", "flags", "cigar_off", "cigar", "seq_off")


def _field_bytes(c, count, width):
    n = {"n_reads": c.n_reads, "n_reads+1": c.n_reads + 1, "n_cigar": c.n_cigar, "n_seq": c.n_seq,
         "n_seq/4": c.n_seq // 4, "n_exc": c.n_exc}[count]
    return int(n) * width


def device_batch(torch, c, dev):
    """Copy one host pb_batch to the GPU; returns (pb_batch with PB_MEM_DEVICE, keepalive tensors)."""
    from pilon_b200 import _capi as capi
    d = capi.pb_batch()
    d.n_reads, d.n_cigar, d.n_seq, d.n_exc = c.n_reads, c.n_cigar, c.n_seq, c.n_exc
    keep = []
    for name, count, width in _BATCH_FIELDS:
        nb = _field_bytes(c, count, width)
        t = torch.empty(nb + 64, dtype=torch.uint8, device=dev)      # kernels fetch aligned 16-byte blocks
        if nb:
            host = np.ctypeslib.as_array(C.cast(getattr(c, name), C.POINTER(C.c_uint8)), shape=(nb,))
            t[:nb].copy_(torch.from_numpy(host))
        keep.append(t)
        setattr(d, name, t.data_ptr())
    d.mem = capi.PB_MEM_DEVICE
    return d, keep


def pin_batch(torch, c):
    """cudaHostRegister every array of a host batch that the engine uploads, so that H2D copies are true async
    DMA; returns the bytes uploaded per pass (4-bit quality codes replace the quality bytes when the batch has them)."""
    rt = torch.cuda.cudart()
    n = 0
    if c.meta_codes:                                       # compact per-read metadata instead of the eight plain arrays
        for nm, nb in (("meta_codes", int(c.n_reads) * 8), ("meta_cigar", int(c.n_meta_cigar) * 4), ("meta_esc", int(c.n_meta_esc) * 12)):
            if nb:
                _register(rt, getattr(c, nm), nb)
                n += nb
    for name, count, width in _BATCH_FIELDS:
        nb = _field_bytes(c, count, width)
        if c.meta_codes and name in _META_FIELDS:
            continue
        if name == "bases2" and c.base_delta_idx:
            for nm, w in (("base_delta_idx", 4), ("base_delta_code", 1)):
                if c.n_base_delta:
                    _register(rt, getattr(c, nm), int(c.n_base_delta) * w)
                    n += int(c.n_base_delta) * w
            continue
        if name == "quals" and c.qual_codes:
            nb = (int(c.n_seq) * int(c.qual_code_bits) + 7) // 8
            name = "qual_codes"
        if nb:
            _register(rt, getattr(c, name), nb)
            n += nb
    return n


def _register(rt, ptr, nbytes):
    err = rt.cudaHostRegister(ptr, nbytes, 0)
    if int(err) != 0:
        raise RuntimeError("cudaHostRegister(%d bytes) failed: %s" % (nbytes, err))


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C restatement of the reference path on host cores
# ---------------------------------------------------------------------------------------------
def oracle_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    from pilon_b200 import _capi as capi
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpilon_oracle.so"))
    lib.po_region_new.restype = C.c_void_p
    lib.po_region_new.argtypes = [C.POINTER(capi.pb_config), C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    lib.po_region_add_batch.argtypes = [C.c_void_p, C.POINTER(capi.pb_batch), C.c_int, C.c_int, C.c_void_p]
    lib.po_region_finish.argtypes = [C.c_void_p, C.POINTER(capi.pb_region_result)]
    lib.po_region_free.argtypes = [C.c_void_p]
    return lib


def oracle_region(lib, reg, planes=FIX_PLANES, indels=False):
    from pilon_b200.engine import EngineConfig
    from pilon_b200.packing import ResultBuffers
    cfg = EngineConfig().to_c()
    h = lib.po_region_new(C.byref(cfg), reg.contig.ctypes.data, len(reg.contig), reg.start, reg.stop)
    for b in reg.batches:
        assert lib.po_region_add_batch(h, C.byref(b.c), int(b.frag), 0, None) == 0
    res = ResultBuffers(reg.size, planes, *(_indel_caps(reg) if indels else (0, 0)))
    assert lib.po_region_finish(h, C.byref(res.c)) == 0
    lib.po_region_free(h)
    return res


def _indel_caps(reg):
    return max(1 << 16, reg.n_cigar // 8), 1 << 24


def parity_check(eng, lib_res, reg):
    """One region through the public C ABI from host buffers (the e2e path) against the C oracle's result for the same
    region: every scalar, every per-locus plane and every indel evidence entry, bit for bit.  Returns a list of
    mismatch descriptions (empty = identical)."""
    from pilon_b200 import _capi as capi
    from pilon_b200.packing import ResultBuffers
    res = ResultBuffers(reg.size, None, *_indel_caps(reg))
    eng.region_begin(reg.contig, reg.start, reg.stop)
    for b in reg.batches:
        eng.add_batch(b, b.frag)
    eng.finish(res)
    bad = []
    for f in ("size", "base_count", "coverage", "aligned_bases", "read_count", "min_depth", "unknown_ops", "dropped_oob",
              "n_indels"):       # (n_indel_bytes: the engine only materialises strict-majority strings, see below)
        if getattr(res.c, f) != getattr(lib_res.c, f):
            bad.append("scalar %s: %r != %r" % (f, getattr(res.c, f), getattr(lib_res.c, f)))
    for name, _, _ in capi.RESULT_PLANES:
        if not np.array_equal(res[name], lib_res[name]):
            bad.append("plane %s differs at %s" % (name, np.argwhere(res[name] != lib_res[name])[:3].tolist()))
    if res.per_bam() != lib_res.per_bam():
        bad.append("per-BAM deltas: %r != %r" % (res.per_bam(), lib_res.per_bam()))
    ia, ib = res.indels(), lib_res.indels()
    if len(ia) != len(ib):
        bad.append("indel entries: %d != %d" % (len(ia), len(ib)))
    for x, y in zip(ia, ib):
        kx, ky = (x["locus_index"], x["kind"], x["list_len"]), (y["locus_index"], y["kind"], y["list_len"])
        wx = (x["win_count"], x["win_len"], x["win_has_n"], x["string"]) if x["win_count"] >= 2 and 2 * x["win_count"] > x["list_len"] else None
        wy = (y["win_count"], y["win_len"], y["win_has_n"], y["string"]) if y["win_count"] >= 2 and 2 * y["win_count"] > y["list_len"] else None
        if kx != ky or wx != wy:        # the winner is only defined (and consumed, PileUp.scala:219-220) as a strict majority
            bad.append("indel entry %r != %r" % (x, y))
            break
    return bad


def time_oracle(regions, threads):
    lib = oracle_lib()
    t0 = time.perf_counter()
    if threads <= 1:
        for r in regions:
            oracle_region(lib, r)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda r: oracle_region(lib, r), regions))
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    wl, regions = build_workload(args.workload, args.scale, 0, ncpu)
    # bounded sample: the regions of at most 3 Mb (largest first), one per thread
    sample = sorted([r for r in regions if r.size <= 3_000_000] or regions[:1], key=lambda r: -r.size)
    threads = min(ncpu, len(sample))
    for _ in range(args.warmup):
        time_oracle(sample[:threads], threads)
    ts = [time_oracle(sample, threads) for _ in range(args.steps)]
    bases = sum(r.aligned for r in sample)
    t = sum(ts) / len(ts)
    val = bases / t
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": {"workload": "%s: %s" % (wl.name, wl.description), "scale": args.scale,
                      "sample": "%d regions <= 3 Mb, %d aligned bases per step" % (len(sample), bases)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": "%d regions of %s (<= 3 Mb each), %d aligned bases, region-parallel over %d threads; "
                                      "the reference itself is single-threaded (Pilon.scala:219-222)" % (
                                          len(sample), wl.name, bases, threads)},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pilon_b200 import build as pbuild
    pbuild.build()
    from pilon_b200.engine import Engine
    from pilon_b200.packing import ResultBuffers

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    placement = bind_rank(local, world) if world > 1 else None      # each rank on the cores next to its GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ncpu = len(os.sched_getaffinity(0))
    wl, regions = build_workload(args.workload, args.scale, rank, ncpu)
    total_aligned = sum(r.aligned for r in regions)
    total_loci = sum(r.size for r in regions)

    # ---- device-resident arm -------------------------------------------------------------
    engines, keep = [], []
    for r in regions:
        e = Engine(local)
        e.region_begin(r.contig, r.start, r.stop)
        for b in r.batches:
            d, k = device_batch(torch, b.c, dev)
            keep.append(k)
            e.add_batch(d, b.frag)
        engines.append(e)
    torch.cuda.synchronize()

    # (a) the kernels one region at a time: per-kernel numbers for the roofline (pileup kernel timed alone)
    def step_sequential():
        tot = pil = 0.0
        launches = 0
        for e in engines:
            a, p, n = e.compute_timed(1)
            tot += a; pil += p; launches += n
        return tot, pil, launches

    # (b) the job: every region's pass launched from a small pool of host threads onto the engines' own
    # streams, so that the short kernels of one region overlap the long kernels of another
    streams = [torch.cuda.ExternalStream(e.stream_ptr(), device=dev) for e in engines]
    order = sorted(range(len(engines)), key=lambda i: -regions[i].aligned)
    pool = ThreadPoolExecutor(args.host_threads)

    def step_concurrent():
        list(pool.map(lambda i: engines[i].compute(), order))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()              # started before the warm-up: nvidia-smi needs ~0.1 s to produce its first sample
    for _ in range(max(args.warmup, 3)):
        step_sequential()
        step_concurrent()
    barrier()
    sampler.mark()                   # only samples taken from here on (both timed passes) are reported
    seq_ms = pil_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        a, p, n = step_sequential()
        seq_ms += a; pil_ms += p; launches += n
    barrier()
    wall0 = time.perf_counter()
    ev_start = torch.cuda.Event(enable_timing=True)
    ev_start.record(streams[0])
    for _ in range(args.steps):
        step_concurrent()
    ends = []
    for st in streams:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(st)
        ends.append(ev)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    dev_ms = max(ev_start.elapsed_time(ev) for ev in ends)       # device clock: first launch -> last stream done
    if rank == 0 and sampler.proc:   # a timed region shorter than three 20 ms sampling periods: keep the same load running
        for _ in range(200):         # (untimed) until nvidia-smi has reported the clocks under it
            if sampler.count() >= 3:
                break
            step_concurrent()
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    pool.shutdown()

    # ---- end-to-end arm: host buffers -> C ABI -> host results -------------------------------
    for e in engines:
        e.close()
    engines, keep = [], []
    torch.cuda.empty_cache()
    if args.quals8:
        for r in regions:
            for b in r.batches:
                b.c.qual_codes = None
    q4 = all(bool(b.c.qual_codes) for r in regions for b in r.batches)
    if args.base_deltas:                                   # reference-delta transport of the 2-bit bases (pb_base_delta_encode)
        from pilon_b200.packing import base_delta_encode
        for r in regions:
            for b in r.batches:
                b.delta_idx, b.delta_code = base_delta_encode(b.c, r.contig, r.start, r.stop)
                b.c.base_delta_idx = b.delta_idx.ctypes.data
                b.c.base_delta_code = b.delta_code.ctypes.data
                b.c.n_base_delta = int(b.delta_idx.shape[0]) - 16
    meta_on = False
    if args.compact_meta:                                  # 8 bytes per read instead of 22 + 4 per CIGAR op (pb_meta_encode)
        from pilon_b200.packing import meta_encode
        encs = [(b, meta_encode(b.c)) for r in regions for b in r.batches]
        if all(m is not None for _, m in encs):
            meta_on = True
            for b, m in encs:
                b.meta = m
                codes, cigar, esc, ng, ne, pos0, stride = m
                b.c.meta_codes, b.c.meta_cigar, b.c.meta_esc = codes.ctypes.data, cigar.ctypes.data, esc.ctypes.data
                b.c.n_meta_cigar, b.c.n_meta_esc, b.c.meta_pos0, b.c.meta_seq_stride = ng, ne, pos0, stride
    qbits = max([int(b.c.qual_code_bits) for r in regions for b in r.batches] or [0]) if q4 else 8
    h2d = sum(pin_batch(torch, b.c) for r in regions for b in r.batches)
    halo = 16384                                          # PB_REF_HALO: the reference window a pass uploads
    h2d += sum(min(len(r.contig), r.stop + halo) - max(1, r.start - halo) + 1 for r in regions)
    for cbuf in {id(r.contig): r.contig for r in regions}.values():      # the contigs are pinned once, like the reads
        if os.environ.get("PB_BENCH_PIN_CONTIG") != "1":     # measured: pinning them costs ~10 % (profiles/README.md)
            break
        _register(torch.cuda.cudart(), np.frombuffer(cbuf, np.uint8).ctypes.data, len(cbuf))
    from pilon_b200 import _capi as capi
    n_workers = args.e2e_workers
    max_size = max(r.size for r in regions)
    e2e_depth = max(1, args.e2e_depth)                       # passes in flight per host thread (one engine + result set each)

    def plane_set(which):
        return {"fixmin": FIXMIN_PLANES, "fix": FIX_PLANES, "vcf": None}[which]

    def make_workers(which):
        pl = plane_set(which)
        return [[(Engine(local), ResultBuffers(max_size, pl, indels_cap=1 << 20, indel_bytes_cap=1 << 23, pinned=True,
                                                calls_cap=CALLS_CAP if which == "fixmin" else 0))
                 for _ in range(e2e_depth)] for _ in range(n_workers)]

    def d2h_bytes(which, n_calls):
        pl = plane_set(which)
        per_locus = sum(np.dtype(dt).itemsize * per for name, dt, per in capi.RESULT_PLANES if pl is None or name in pl)
        return per_locus * total_loci + 16 * n_calls
    workers = make_workers(args.planes)
    calls_seen = [0]

    trace = [[0.0, 0.0, 0.0] for _ in range(n_workers)] if os.environ.get("PB_E2E_TRACE") else None

    def e2e_submit(slot, k, r):
        """pb_region_begin + pb_region_add_batch: asynchronous, the uploads are queued on the engine's stream."""
        eng, res = workers[slot][k]
        res.c.size = r.size
        t_a = time.perf_counter()
        eng.region_begin(r.contig, r.start, r.stop)
        t_b = time.perf_counter()
        for b in r.batches:
            eng.add_batch(b, b.frag)
        if trace is not None:                              # host seconds inside begin / add_batch / finish, per host thread
            trace[slot][0] += t_b - t_a
            trace[slot][1] += time.perf_counter() - t_b

    def e2e_collect(slot, k):
        """pb_region_finish: the pass, the download into the pinned result planes, and the wait for both."""
        eng, res = workers[slot][k]
        t_c = time.perf_counter()
        eng.finish(res)
        if trace is not None:
            trace[slot][2] += time.perf_counter() - t_c
        if res.calls_cap:
            assert res.c.n_calls <= res.calls_cap, "CALLS_CAP too small for this region"
            calls_seen[0] += int(res.c.n_calls)              # (GIL-protected; only used for the byte count)
        return int(res.c.aligned_bases)

    def e2e_step():
        order = sorted(range(len(regions)), key=lambda i: -regions[i].aligned)
        lock = threading.Lock()
        done = [0]

        def work(slot):
            # each host thread keeps `e2e_depth` regions in flight: region i+1 is uploading while region i computes and downloads
            inflight, n, got = [], 0, 0
            while True:
                with lock:
                    i = order.pop(0) if order else None
                if i is None:
                    break
                if len(inflight) == e2e_depth:
                    got += e2e_collect(slot, inflight.pop(0))
                k = n % e2e_depth
                n += 1
                e2e_submit(slot, k, regions[i])
                inflight.append(k)
            for k in inflight:
                got += e2e_collect(slot, k)
            with lock:
                done[0] += got
        ts = [threading.Thread(target=work, args=(s,)) for s in range(n_workers)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return done[0]

    def prime():
        """Every engine takes the largest region once: which host thread gets which region varies from step to step, and an
        engine that met its largest region only in a timed step would grow its buffers there (cudaFree + cudaMalloc: a
        device-wide synchronisation; seen as 60 ms vs 100 ms steps before this was here)."""
        big = max(regions, key=lambda r: r.aligned)
        for slot in range(n_workers):
            for k in range(e2e_depth):
                e2e_submit(slot, k, big)
                e2e_collect(slot, k)
        calls_seen[0] = 0
    e2e_steps = max(1, args.steps)
    prime()
    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    got = 0
    for _ in range(e2e_steps):
        got += e2e_step()
    ev1.record()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert got == total_aligned * e2e_steps, (got, total_aligned)
    d2h = d2h_bytes(args.planes, calls_seen[0] // (e2e_steps + max(1, min(args.warmup, 3))))

    def timed_variant(n):
        prime()
        e2e_step()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            e2e_step()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / n
    # the same arm as round 1 ran it: every array of GenomeRegion.scala:247-253 + the call plane down (35 B per locus) and the
    # bases up as 2-bit codes -- what the minimal result set and the reference deltas bought
    e2e_classic = None
    if args.planes == "fixmin" and args.base_deltas and world == 1:      # (needs the reference deltas on: it re-pins bases2)
        for w in workers:
            for eng, _ in w:
                eng.close()
        workers = make_workers("fix")
        saved_d = [(b, b.c.base_delta_idx, b.c.base_delta_code, b.c.n_base_delta) for r in regions for b in r.batches]
        h2d_c = h2d
        saved_m = [(b, b.c.meta_codes) for b, _, _, _ in saved_d]
        for b, _, _, _ in saved_d:
            h2d_c += int(b.c.n_seq) // 4 - 5 * int(b.c.n_base_delta)
            b.c.base_delta_idx = None; b.c.base_delta_code = None; b.c.n_base_delta = 0
            _register(torch.cuda.cudart(), b.c.bases2, int(b.c.n_seq) // 4)
            if b.c.meta_codes:
                h2d_c -= int(b.c.n_reads) * 8 + int(b.c.n_meta_cigar) * 4 + int(b.c.n_meta_esc) * 12
                for name, count, width in _BATCH_FIELDS:
                    if name in _META_FIELDS and _field_bytes(b.c, count, width):
                        _register(torch.cuda.cudart(), getattr(b.c, name), _field_bytes(b.c, count, width))
                        h2d_c += _field_bytes(b.c, count, width)
                b.c.meta_codes = None
        nc = max(1, min(args.steps, 3))
        tc = timed_variant(nc)
        e2e_classic = {"value": total_aligned / tc, "unit": UNIT, "ms_per_step": 1e3 * tc, "h2d_bytes_per_step": h2d_c,
                       "d2h_bytes_per_step": d2h_bytes("fix", 0), "steps": nc,
                       "what": "round-1 transport: plain per-read arrays and 2-bit bases up, the ten per-locus arrays of the fix + tracks path down"}
        for b, di, dc, nd in saved_d:
            b.c.base_delta_idx = di; b.c.base_delta_code = dc; b.c.n_base_delta = nd
        for b, mc in saved_m:
            b.c.meta_codes = mc
        for w in workers:
            for eng, _ in w:
                eng.close()
        workers = make_workers(args.planes)
    # the same arm with one quality byte per base (what a BAM with unbinned qualities needs), a few steps
    e2e8 = None
    if q4 and not args.quals8 and world == 1:
        saved = [(b, b.c.qual_codes) for r in regions for b in r.batches]
        h2d8 = 0
        for b, _ in saved:
            b.c.qual_codes = None
            _register(torch.cuda.cudart(), b.c.quals, int(b.c.n_seq))
            h2d8 += int(b.c.n_seq) - (int(b.c.n_seq) * int(b.c.qual_code_bits) + 7) // 8
        n8 = max(1, min(args.steps, 3))
        t8 = timed_variant(n8)
        e2e8 = {"value": total_aligned / t8, "unit": UNIT, "ms_per_step": 1e3 * t8, "h2d_bytes_per_step": h2d + h2d8, "steps": n8}
        for b, qc in saved:
            b.c.qual_codes = qc
    if trace is not None and rank == 0:
        for slot, t in enumerate(trace):
            print("e2e worker %d: begin %.1f ms, add_batch %.1f ms, finish %.1f ms per step" %
                  ((slot,) + tuple(1e3 * x / (e2e_steps + max(1, min(args.warmup, 3))) for x in t)), file=sys.stderr)
    for w in workers:
        for eng, _ in w:
            eng.close()

    # ---- aggregate over ranks -----------------------------------------------------------------
    vals = torch.tensor([dev_ms / args.steps, pil_ms / args.steps, wall_ms / args.steps, e2e_s, seq_ms / args.steps], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(total_aligned), float(launches), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    step_ms, pileup_ms, wall_step_ms, e2e_sec, seq_step_ms = [float(x) for x in vals.tolist()]
    job_aligned, job_launches, job_h2d, job_d2h = [float(x) for x in sums.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        traffic = None
        try:     # dram bytes of the kernel from the committed ncu capture, scaled to this run's average launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["algorithmic_bytes"] * (sum(r.alg_bytes for r in regions) / len(regions))
        except Exception:
            pass
        alg = sum(r.alg_bytes for r in regions)                 # this rank's launches
        achieved = alg / (pileup_ms * 1e-3) / 1e9               # GB/s over the pileup kernel's launches of one step
        depth = total_aligned / total_loci
        cpu = parity = None
        if world == 1 and not args.no_cpu_baseline:
            sample, acc = [], 0            # ~2 G aligned bases = 10-20 s of one host core
            for r in sorted(regions, key=lambda r: -r.size):
                if r.size <= 6_000_000 and acc < 2.0e9:
                    sample.append(r); acc += r.aligned
            if not sample:                 # every chunk is larger than that (C3: seven 9.1 Mb chunks): the smallest one
                sample = [min(regions, key=lambda r: r.aligned)]
            # the oracle is timed per region; what it returns is then compared with the engine's result for the same region
            # through the same C-ABI calls the e2e arm makes (host buffers in, host planes out), outside the timed part
            lib = oracle_lib()
            peng = Engine(local)
            t, mismatches = 0.0, []
            for r in sample:
                t0 = time.perf_counter()
                ref = oracle_region(lib, r, None, True)
                t += time.perf_counter() - t0
                mismatches += ["%s contig %d %d-%d: %s" % (wl.name, r.ci, r.start, r.stop, m) for m in parity_check(peng, ref, r)]
                del ref
            peng.close()
            sb = sum(r.aligned for r in sample)
            cpu = {"value": sb / t, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "C restatement of the reference path (oracle/pilon_oracle.c), 1 thread like the reference, "
                             "%d regions of %s (%d loci, %d aligned bases), %.1f s" % (
                                 len(sample), wl.name, sum(r.size for r in sample), sb, t)}
            parity = {"regions": len(sample), "loci": sum(r.size for r in sample), "aligned_bases": sb,
                      "compared": "every scalar, all 18 per-locus planes, every indel evidence entry; engine through the C ABI "
                                  "from host buffers vs oracle/pilon_oracle.c on the same regions",
                      "ok": not mismatches}
            if mismatches:
                parity["mismatches"] = mismatches[:10]
        from_bam = from_bam_leg(args, wl, regions, local) if (args.from_bam and world == 1) else None
        # shared-memory reductions of the scatter kernels: one 32-bit RED lane-op per counted base (+ one per counted base of a
        # batch outside fragCoverage); peak = 32 lanes x 1 wavefront per clock per SM.  The engine's choice (pb_engine.cu,
        # compute()): mean depth > 1000 -> k_pileup7c, else >= 256 tiles of 2048 loci -> k_pileup7, else the gather kernel,
        # which accumulates in registers and issues no per-base atomics.
        def _depth(r):
            return sum(int(b.c.n_seq) for b in r.batches) // r.size
        deep = [r for r in regions if _depth(r) > 1000 and len(r.batches) <= 20]
        scatter = deep + [r for r in regions if _depth(r) <= 1000 and (r.size + 2047) // 2048 >= 256 and len(r.batches) <= 20]
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
        lane_ops = sum(b.aligned_bases * (1 if b.frag else 2) for r in scatter for b in r.batches) * (130.0 / 150.0)
        atomic = {"kernel": ("k_pileup7c (cluster scatter)" if deep and len(deep) == len(scatter) else "k_pileup7 (scatter)" if scatter
                             else "k_pileup5 (gather: no per-base atomics)"),
                  "regions": len(scatter),
                  "shared_red_lane_ops_per_s": (lane_ops / (pileup_ms * 1e-3)) if scatter else 0.0,
                  "peak_lane_ops_per_s": 148 * 32 * sm_hz,
                  "frac": (lane_ops / (pileup_ms * 1e-3)) / (148 * 32 * sm_hz) if scatter else 0.0,
                  "note": "counted bases ~ aligned bases x 130/150 (trusted flank); ncu: profiles/r2_pileup7_raw.csv, profiles/r2q/pileup7c_c5_raw.csv "
                          "(smsp__inst_executed_op_shared_atom, l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom)"}
        out = {"metric": METRIC, "value": job_aligned / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "int64", "data": "synthetic",
               "config": {"workload": "%s: %s" % (wl.name, wl.description), "scale": args.scale,
                          "regions_per_gpu": len(regions), "loci_per_gpu": total_loci, "reads_per_gpu": sum(r.n_reads for r in regions),
                          "aligned_bases_per_gpu": total_aligned, "mean_depth": depth, "l2": "inputs_exceed_l2 (%.1f GB per step)" % (h2d / 1e9),
                          "timing": "CUDA events: first launch of the timed steps -> last engine stream done, regions launched by %d host threads onto one stream per region; "
                                    "sequential_ms_per_step = sum of per-region event intervals with one region at a time (the pileup kernel's launches are timed in that pass)" % args.host_threads,
                          "e2e_planes": {"fixmin": "flags + frag_coverage + sparse call records (pb_region_result.calls) + indel evidence: what pb_out_create consumes for --fix snps,indels --changes",
                                         "fix": "the ten per-locus arrays of the fix + tracks path (35 B per locus)", "vcf": "every plane"}[args.planes],
                          "e2e_quals": ("%d-bit codes + table (the workload has <= %d distinct quality bytes), expanded on the device"
                                        % (qbits, 1 << qbits) if q4 else "1 byte per base"),
                          "e2e_reads": ("8 bytes per read + listed CIGARs + escapes (pb_meta_encode), plain arrays rebuilt on the device" if meta_on
                                        else "22 bytes per read + 4 per CIGAR op"),
                          "e2e_bases": ("deltas against the reference (5 B per differing base), bases2 rebuilt on the device"
                                        if args.base_deltas else "2 bits per base")},
               "wall_ms_per_step": wall_step_ms, "sequential_ms_per_step": seq_step_ms,
               "e2e": {"value": job_aligned / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": job_h2d, "d2h_bytes_per_step": job_d2h,
                       "ms_per_step": 1e3 * e2e_sec, "steps": e2e_steps, "streams_per_gpu": n_workers * e2e_depth,
                       "host_threads_per_gpu": n_workers, "passes_in_flight_per_thread": e2e_depth,
                       "quals8": e2e8, "classic": e2e_classic},
               "gpu_launches": int(job_launches),
               "roofline": {"bound": "hbm", "kernel": atomic["kernel"].split(" ")[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": sum(r.alg_bytes for r in regions) / len(regions),
                            "algorithmic_bytes_per_base": alg / total_aligned,
                            "pileup_ms_per_step": pileup_ms, "pileup_share_of_sequential_step": pileup_ms / seq_step_ms,
                            "atomic": atomic},
               "cpu_baseline": cpu, "parity": parity, "e2e_from_bam": from_bam, "host_placement": placement, "clocks": clocks}
        print(json.dumps(out))
        if parity is not None and not parity["ok"]:
            print("PARITY MISMATCH: " + "; ".join(parity["mismatches"]), file=sys.stderr)
            sys.exit(3)
    if world > 1:
        dist.destroy_process_group()



def from_bam_leg(args, wl, regions, local):
    """End to end from BAM BYTES on a bounded sample: the sample contigs' reads are written as coordinate-
    sorted BAM + BAI (untimed), then timed: native BGZF inflate + record decode + validateRead + packing (host threads, one
    BAM handle each) -> C ABI -> per-locus results on the host.  What the reference does here is htsjdk + addRead."""
    import tempfile
    from pilon_b200 import bamio, synth
    from pilon_b200.engine import Engine
    from pilon_b200.packing import ResultBuffers
    sample = [r for r in regions if r.size <= 3_000_000][:6] or regions[:1]
    tmp = tempfile.mkdtemp(prefix="pb_bam_")
    refs = [("contig%02d" % (r.ci + 1), len(r.contig)) for r in sample]
    paths = []
    for libr in wl.libraries:
        p = os.path.join(tmp, libr.name + ".bam")
        batches = [(k, synth.SynthBatch(wl.params(r.ci, libr), 1, len(r.contig), libr.counts_toward_frag_coverage)) for k, r in enumerate(sample)]
        bamio.write_bam(p, refs, [(k, sb.c) for k, sb in batches])
        paths.append((p, libr.name))
        del batches
    bam_bytes = sum(os.path.getsize(p) for p, _ in paths)
    nthreads = max(1, min(len(sample), len(os.sched_getaffinity(0)), 6))
    max_size = max(r.size for r in sample)
    slots = [(Engine(local), ResultBuffers(max_size, FIXMIN_PLANES, indels_cap=1 << 20, indel_bytes_cap=1 << 23, pinned=True, calls_cap=CALLS_CAP),
              [bamio.BamFile(p, t) for p, t in paths]) for _ in range(nthreads)]
    order = list(range(len(sample)))
    lock = threading.Lock()
    t_ingest = [0.0] * nthreads
    got = [0]

    def work(slot):
        eng, res, bams = slots[slot]
        while True:
            with lock:
                k = order.pop(0) if order else None
            if k is None:
                return
            r = sample[k]
            t0 = time.perf_counter()
            batches = [(b.process(refs[k][0], r.start, r.stop), b.countsTowardFragCoverage) for b in bams]
            t_ingest[slot] += time.perf_counter() - t0
            eng.region_begin(r.contig, r.start, r.stop)
            for rb, frag in batches:
                eng.add_batch(rb, frag)
            eng.finish(res)
            with lock:
                got[0] += int(res.c.aligned_bases)
    def one_pass():
        order[:] = list(range(len(sample)))
        ts = [threading.Thread(target=work, args=(s,)) for s in range(nthreads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    big = max(range(len(sample)), key=lambda k: sample[k].aligned)
    for slot in range(nthreads):                 # warm-up: every engine, packer and BAM handle takes the largest region once
        order[:] = [big]
        work(slot)
    rounds = 3
    t_ingest[:] = [0.0] * nthreads
    got[0] = 0
    t0 = time.perf_counter()
    for _ in range(rounds):
        one_pass()
    dt = time.perf_counter() - t0
    for eng, _, bams in slots:
        eng.close()
        [b.close() for b in bams]
    for p, _ in paths:
        os.remove(p); os.remove(p + ".bai")
    os.rmdir(tmp)
    return {"value": got[0] / dt, "unit": UNIT, "regions": len(sample), "passes": rounds, "aligned_bases": got[0] // rounds, "bam_bytes": bam_bytes,
            "host_threads": nthreads, "inflate_threads_per_query": 4, "seconds_per_pass": dt / rounds,
            "ingest_seconds_per_thread_per_pass": sum(t_ingest) / nthreads / rounds,
            "note": "BGZF inflate + BAM decode + packing are inside the timed region (zlib; a read-ahead window of blocks inflated by 4 threads per "
                    "query); the packed batches are uploaded from pageable memory; engines warmed with one untimed region each"}



# ---------------------------------------------------------------------------------------------
# host placement: one rank per GPU, each on the cores (and memory) next to its GPU
# ---------------------------------------------------------------------------------------------
def _parse_cpulist(txt):
    out = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def gpu_local_cpus(index):
    """Cores the kernel reports as local to GPU `index` (sysfs local_cpulist of its PCI function)."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, stderr=subprocess.DEVNULL).strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus                     # 00000000:1B:00.0 -> 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bus) as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def bind_rank(local, world):
    """Pin this rank (and the threads it starts later) to its share of the cores local to its GPU.  Returns a
    description for the JSON line.  Ranks whose GPUs report the same core list split it evenly."""
    allowed = sorted(os.sched_getaffinity(0))
    lists = [gpu_local_cpus(i) for i in range(world)]
    mine = lists[local] if local < len(lists) else None
    if not mine:
        mine, peers, k = allowed, world, local
    else:
        same = [i for i in range(world) if lists[i] == mine]
        peers, k = len(same), same.index(local)
    mine = [c for c in mine if c in allowed] or allowed
    share = mine[k * len(mine) // peers:(k + 1) * len(mine) // peers] or mine
    try:
        os.sched_setaffinity(0, share)
    except Exception:
        share = allowed
    os.environ["OMP_NUM_THREADS"] = str(max(1, len(share)))
    return {"cores": len(share), "first_core": share[0], "gpu_local_cores": len(mine), "ranks_sharing_them": peers}


# ---------------------------------------------------------------------------------------------
# sharded arm (BASELINE config 4): ONE genome, its chunks distributed over the ranks, streamed in waves
# ---------------------------------------------------------------------------------------------
class StreamRegion:
    """One chunk generated on its own: the reads of its +-10 kb window and the reference window the engine uploads.
    Nothing of the genome outside the window is ever held (GenomeFile.scala:67-74 chunking, BamFile.scala:118-119)."""

    def __init__(self, wl, index, chunk):
        ci, a, b = chunk
        self.index, self.ci, self.start, self.stop = index, ci, a, b
        self.size = b + 1 - a
        n = wl.contig_lens[ci]
        self.contig_len = n
        self.lo, self.hi = max(1, a - 16384), min(n, b + 16384)          # PB_REF_HALO on either side
        self.window = wl.contig_bases(ci, self.lo, self.hi)
        self.batches = wl.region_batches(ci, a, b)
        self.aligned = sum(x.aligned_bases for x in self.batches)
        self.n_reads = sum(x.n_reads for x in self.batches)
        self.n_cigar = sum(x.c.n_cigar for x in self.batches)

    def begin(self, eng):
        from pilon_b200 import _capi as capi
        # the engine reads contig[start - 16384 - 1 .. stop + 16384 - 1] only: hand it the window under the address
        # the whole contig would have had
        capi.check(eng.lib.pb_region_begin(eng._h, self.window.ctypes.data - (self.lo - 1), self.contig_len, self.start, self.stop))
        eng._keep = [self.window]


def plane_digest(res, planes):
    """Order-sensitive 64-bit digest of a region's result planes and scalars (cheap: two reductions per plane).
    The buffers may be larger than the region: only the first res.c.size loci are read."""
    from pilon_b200 import _capi as capi
    size = int(res.c.size)
    per = {p[0]: p[2] for p in capi.RESULT_PLANES}
    acc = [size, int(res.c.base_count), int(res.c.read_count), int(res.c.min_depth), int(res.c.n_indels)]
    w = (np.arange(size, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)
    for name in planes:
        x = res.arrays[name][:size * per[name]].astype(np.uint64)
        if per[name] > 1:
            x = x.reshape(size, per[name]).sum(axis=1, dtype=np.uint64)
        acc.append(int(x.sum(dtype=np.uint64)))
        acc.append(int((x * w).sum(dtype=np.uint64)))
    return acc


def run_sharded_arm(args):
    import zlib
    import torch
    import torch.distributed as dist
    from pilon_b200 import build as pbuild
    pbuild.build()
    from pilon_b200 import sharding, synth
    from pilon_b200.engine import Engine
    from pilon_b200.packing import ResultBuffers

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    placement = bind_rank(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = synth.workload(args.workload, args.scale)
    chunks = wl.regions()
    depth = sum(l.depth for l in wl.libraries)
    mine = sharding.assign(chunks, world, [depth] * len(chunks))[rank]
    from pilon_b200 import _capi as capi
    # (the sharded arm keeps the ten fix + tracks planes also for --planes fixmin: its chunk digests are taken over them)
    planes = FIX_PLANES if args.planes in ("fix", "fixmin") else [p[0] for p in capi.RESULT_PLANES]
    max_size = max(b + 1 - a for _, a, b in chunks)
    n_workers = args.e2e_workers
    workers = [(Engine(local), ResultBuffers(max_size, planes, indels_cap=1 << 20, indel_bytes_cap=1 << 23, pinned=True))
               for _ in range(n_workers)]
    rt = torch.cuda.cudart()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    totals = dict(aligned=0, loci=0, reads=0, h2d=0, d2h=0, launches=0)
    t_e2e = t_dev = t_pile = 0.0
    digests = {}
    per_locus = sum(np.dtype(dt).itemsize * per for name, dt, per in capi.RESULT_PLANES if name in planes)
    wave_n = max(1, args.wave)
    warm = True
    for w0 in range(0, len(mine), wave_n):
        regs = [StreamRegion(wl, i, chunks[i]) for i in mine[w0:w0 + wave_n]]       # generated now, dropped after the wave
        pinned = []
        for r in regs:
            for b in r.batches:
                totals["h2d"] += pin_batch(torch, b.c)
                pinned.append(b.c)
            totals["h2d"] += r.hi - r.lo + 1
        # ---- end to end: host buffers -> C ABI -> pinned host planes, `n_workers` chunks in flight ----
        def e2e_pass(keep_digests):
            order = list(range(len(regs)))
            lock = threading.Lock()

            def work(slot):
                eng, res = workers[slot]
                while True:
                    with lock:
                        k = order.pop(0) if order else None
                    if k is None:
                        return
                    r = regs[k]
                    r.begin(eng)
                    for b in r.batches:
                        eng.add_batch(b, b.frag)
                    eng.finish(res)
                    assert int(res.c.aligned_bases) == r.aligned
                    if keep_digests:
                        digests[r.index] = plane_digest(res, planes)
            ts = [threading.Thread(target=work, args=(s,)) for s in range(n_workers)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        if warm:                                   # first wave of the run: clocks, allocator pools, page faults
            e2e_pass(False)
            warm = False
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_pass(False)
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        e2e_pass(True)                             # untimed: the digests of this wave's results
        # ---- HBM-resident: batches uploaded (untimed), the pass timed with CUDA events on the engine's stream ----
        eng = workers[0][0]
        for r in regs:
            r.begin(eng)
            keep = []
            for b in r.batches:
                d, k = device_batch(torch, b.c, dev)
                keep.append(k)
                eng.add_batch(d, b.frag)
            eng.compute_timed(1)                                   # sizes every buffer
            a, p, n = eng.compute_timed(args.steps)
            t_dev += a / args.steps * 1e-3
            t_pile += p / args.steps * 1e-3
            totals["launches"] += n // args.steps
            res = workers[0][1]
            eng.finish(res)
            assert plane_digest(res, planes) == digests[r.index], "device-resident and host-fed results differ for chunk %d" % r.index
            del keep
        for r in regs:
            totals["aligned"] += r.aligned; totals["loci"] += r.size; totals["reads"] += r.n_reads
        for c in pinned:
            for name, count, width in _BATCH_FIELDS:
                nb = _field_bytes(c, count, width)
                if name == "quals" and c.qual_codes:
                    name = "qual_codes"
                if nb:
                    rt.cudaHostUnregister(getattr(c, name))
        del regs, pinned
    totals["d2h"] = per_locus * totals["loci"]
    clocks = sampler.stop() if rank == 0 else None
    for eng, _ in workers:
        eng.close()
    vals = torch.tensor([t_dev, t_pile, t_e2e], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(totals[k]) for k in ("aligned", "loci", "reads", "h2d", "d2h", "launches")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t_dev_max, t_pile_max, t_e2e_max = [float(x) for x in vals.tolist()]
    aligned, loci, reads, h2d, d2h, launches = [float(x) for x in sums.tolist()]
    ordered = sharding.gather_in_chunk_order(digests, world)     # host-side gather in chunk order (GenomeFile.scala:135-162)
    if rank == 0:
        assert len(ordered) == len(chunks)
        digest = "%08x" % (zlib.crc32(repr(ordered).encode()) & 0xFFFFFFFF)
        peak, peak_src = measured_peak_gbs()
        out = {"metric": METRIC, "value": aligned / t_dev_max, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": 1,
               "ms_per_step": 1e3 * t_dev_max, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
               "data": "synthetic",
               "config": {"workload": "%s: %s" % (wl.name, wl.description), "scale": args.scale, "chunks": len(chunks),
                          "chunks_per_gpu_max": max(len(x) for x in sharding.assign(chunks, world, [depth] * len(chunks))),
                          "loci": loci, "reads": reads, "aligned_bases": aligned, "mean_depth": aligned / loci,
                          "streaming": "chunks generated and dropped in waves of %d per rank; no rank ever holds more than a wave" % wave_n,
                          "sharding": "pilon_b200.sharding.assign (greedy by loci x depth), no data-path collective; per-chunk digests "
                                      "gathered on the host in chunk order", "l2": "inputs_exceed_l2",
                          "timing": "per rank: sum over its chunks of the device time of one pass (CUDA events on the engine's stream, "
                                    "batches resident); job = max over ranks.  e2e: wall time of each wave through the C ABI from pinned "
                                    "host buffers, %d chunks in flight, summed per rank, max over ranks" % n_workers,
                          "e2e_planes": "fix" if args.planes in ("fix", "fixmin") else args.planes},
               "digest": digest,
               "e2e": {"value": aligned / t_e2e_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": 1e3 * t_e2e_max},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "k_pileup", "achieved": algorithmic_bytes(aligned, reads, reads * 1.1, loci) / t_pile_max / world / 1e9,
                            "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
                            "frac": algorithmic_bytes(aligned, reads, reads * 1.1, loci) / t_pile_max / world / 1e9 / peak,
                            "note": "per GPU: job algorithmic bytes / ranks / max-over-ranks pileup-kernel time"},
               "cpu_baseline": None, "host_placement": placement, "clocks": clocks}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks the genome (not the depth); 1.0 = the BASELINE config")
    ap.add_argument("--planes", default="fixmin", choices=["fixmin", "fix", "vcf"],
                    help="results copied back in the e2e arm: fixmin = what the fix path consumes (flags, fragCoverage, sparse call "
                    "records), fix = the ten per-locus arrays of the fix + tracks path, vcf = every plane")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-workers", type=int, default=3, help="host threads per GPU in the e2e arm")
    ap.add_argument("--e2e-depth", type=int, default=2, help="regions in flight per host thread in the e2e arm (one engine "
                    "and one result set each): the next region uploads while the previous one computes and downloads")
    ap.add_argument("--no-compact-meta", dest="compact_meta", action="store_false", help="e2e arm: upload the eight plain per-read arrays "
                    "instead of 8-byte records (pb_meta_encode)")
    ap.add_argument("--no-base-deltas", dest="base_deltas", action="store_false", help="e2e arm: upload the bases as 2-bit codes "
                    "instead of their deltas against the reference (pb_base_delta_encode; 5 B per differing base)")
    ap.add_argument("--quals8", action="store_true", help="e2e arm: upload one quality byte per base even when the batch "
                    "offers the packed transport (pb_batch.qual_codes)")
    ap.add_argument("--host-threads", type=int, default=4, help="host threads feeding region passes to the GPU")
    ap.add_argument("--from-bam", action="store_true", help="also time a bounded sample end to end from BAM bytes (native BGZF/BAM ingest)")
    ap.add_argument("--sharded", action="store_true", help="ONE genome, its chunks distributed over the ranks and streamed in waves "
                    "(strong scaling; the default for --workload C4): see run_sharded_arm")
    ap.add_argument("--wave", type=int, default=4, help="sharded arm: chunks a rank generates, processes and drops at a time")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.sharded or args.workload == "C4":
        run_sharded_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
