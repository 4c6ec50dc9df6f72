"""Host time of one region pass (pb_region_compute: launches + the scalar read-back) vs its device time.
usage: PB_EXP=256 python tools/host_overhead.py   (256 = skip the pileup kernel: everything else)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from pilon_b200.engine import Engine
wl, regions = bench.build_workload("C2", 0.4, 0, 8)
dev = torch.device("cuda", 0)
engines, keep = [], []
for r in regions:
    e = Engine(0)
    e.region_begin(r.contig, r.start, r.stop)
    for b in r.batches:
        d, k = bench.device_batch(torch, b.c, dev); keep.append(k); e.add_batch(d, b.frag)
    engines.append(e)
torch.cuda.synchronize()
for _ in range(3):
    for e in engines: e.compute()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    for e in engines: e.compute()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
n = 5 * len(engines)
print("regions %d: host time per region pass %.1f us (returns before the device is done), + %.1f us to drain at the end"
      % (len(engines), 1e6 * (t1 - t0) / n, 1e6 * (t2 - t1)))
