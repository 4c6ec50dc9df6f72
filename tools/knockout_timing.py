"""Knock-out timing of the pileup kernel on one 2 Mb region.

usage: python tools/knockout_timing.py VER:EXP [VER:EXP ...]   (PB_PILEUP version : PB_EXP flag mask)
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from pilon_b200.engine import Engine
wl, regions = bench.build_workload("C2", 0.2, 0, 8)
r = max(regions, key=lambda r: r.size)
dev = torch.device("cuda", 0)
print("region loci", r.size, "aligned", r.aligned)
for ver in sys.argv[1:]:
    pv, exp = ver.split(":")
    os.environ["PB_PILEUP"] = pv; os.environ["PB_EXP"] = exp
    e = Engine(0)
    e.region_begin(r.contig, r.start, r.stop)
    keep = []
    for b in r.batches:
        d, k = bench.device_batch(torch, b.c, dev); keep.append(k); e.add_batch(d, b.frag)
    e.compute_timed(2)
    tot, pil, n = e.compute_timed(5)
    print("pileup v%s exp=%s: pileup %.3f ms  total %.3f ms  -> %.1f Gbases/s kernel-only" % (pv, exp, pil / 5, tot / 5, r.aligned / (pil / 5) / 1e6))
    e.close()
