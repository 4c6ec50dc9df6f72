import sys, os
sys.path.insert(0, '/root/repo')
import bench, torch
from pilon_b200.engine import Engine
wl, regions = bench.build_workload("C2", 0.1, 0, 8)
r = max(regions, key=lambda r: r.size)
dev = torch.device("cuda", 0)
e = Engine(0)
e.region_begin(r.contig, r.start, r.stop)
keep = []
for b in r.batches:
    d, k = bench.device_batch(torch, b.c, dev); keep.append(k); e.add_batch(d, b.frag)
print(e.compute_timed(1))
os.environ["PB_DEBUG_TILE"] = "2000"
print(e.compute_timed(1))
