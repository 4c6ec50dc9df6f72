"""End-to-end arm under different settings on one resident workload (what limits it: upload, download, host calls?).
usage: python tools/e2e_sweep.py [scale]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pilon_b200.engine import Engine
from pilon_b200.packing import ResultBuffers

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
wl, regions = bench.build_workload("C2", scale, 0, len(os.sched_getaffinity(0)))
h2d = sum(bench.pin_batch(torch, b.c) for r in regions for b in r.batches)
total = sum(r.aligned for r in regions)
max_size = max(r.size for r in regions)
print("workload ready: %.2f G bases, h2d %.2f GB" % (total / 1e9, h2d / 1e9), flush=True)

def run(n_workers, planes, steps=4, label=""):
    workers = [(Engine(0), ResultBuffers(max_size, planes, indels_cap=1 << 20, indel_bytes_cap=1 << 23, pinned=True)) for _ in range(n_workers)]
    def step():
        order = sorted(range(len(regions)), key=lambda i: -regions[i].aligned)
        lock = threading.Lock()
        def work(slot):
            eng, res = workers[slot]
            while True:
                with lock:
                    i = order.pop(0) if order else None
                if i is None: return
                r = regions[i]
                eng.region_begin(r.contig, r.start, r.stop)
                for b in r.batches: eng.add_batch(b, b.frag)
                eng.finish(res)
        ts = [threading.Thread(target=work, args=(s,)) for s in range(n_workers)]
        [t.start() for t in ts]; [t.join() for t in ts]
    step(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    for e, _ in workers: e.close()
    print("%-28s workers %d planes %-6s %6.1f ms/step  %5.1f G bases/s" % (label, n_workers, "none" if planes == [] else ("fix" if planes else "all"), 1e3 * dt, total / dt / 1e9), flush=True)

for nw in (1, 2, 3, 4, 6):
    run(nw, bench.FIX_PLANES)
run(3, [], label="no per-locus download")
run(3, ["flags"], label="flags only")
os.environ["PB_BENCH_DUMMY"] = "1"
# reference-delta transport of the bases (pb_base_delta_encode): encode once (untimed, like the packing), upload deltas
from pilon_b200.packing import base_delta_encode
t0 = time.perf_counter()
extra = 0
for r in regions:
    for b in r.batches:
        b.delta_idx, b.delta_code = base_delta_encode(b.c, r.contig, r.start, r.stop)
        b.c.base_delta_idx = b.delta_idx.ctypes.data; b.c.base_delta_code = b.delta_code.ctypes.data
        b.c.n_base_delta = int(b.delta_idx.shape[0]) - 16
        rt = torch.cuda.cudart()
        if b.c.n_base_delta:
            bench._register(rt, b.c.base_delta_idx, int(b.c.n_base_delta) * 4); bench._register(rt, b.c.base_delta_code, int(b.c.n_base_delta))
        extra += int(b.c.n_base_delta) * 5 - int(b.c.n_seq) // 4
print("delta encode %.1f s; upload changes by %.2f GB" % (time.perf_counter() - t0, extra / 1e9), flush=True)
run(3, bench.FIX_PLANES, label="base deltas")
run(3, [], label="base deltas, no download")
run(4, bench.FIX_PLANES, label="base deltas")
