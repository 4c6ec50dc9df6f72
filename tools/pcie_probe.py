"""Host <-> device copy bandwidth with pinned 1 GB buffers: one direction at a time, then both.
usage: python tools/pcie_probe.py [N]     N GPUs of the box copy at once (default 1): the aggregate shows what the host side
(memory bandwidth, root complexes) can feed when every rank of an 8-GPU job moves data at the same time."""
import sys, time
import torch
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = 1 << 30
H = [(torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()) for _ in range(ng)]
D = [(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % i), torch.empty(n, dtype=torch.uint8, device="cuda:%d" % i)) for i in range(ng)]
S = [(torch.cuda.Stream(device=i), torch.cuda.Stream(device=i)) for i in range(ng)]
def sync():
    for i in range(ng): torch.cuda.synchronize(i)
def run(h2d, d2h, reps=5):
    sync(); t = time.perf_counter()
    for _ in range(reps):
        for i in range(ng):
            if h2d:
                with torch.cuda.stream(S[i][0]): D[i][0].copy_(H[i][0], non_blocking=True)
            if d2h:
                with torch.cuda.stream(S[i][1]): H[i][1].copy_(D[i][1], non_blocking=True)
    sync(); return reps * n * ng / (time.perf_counter() - t) / 1e9
run(True, True, 1)
print("%d GPU(s): H2D alone %.1f GB/s aggregate" % (ng, run(True, False)))
print("%d GPU(s): D2H alone %.1f GB/s aggregate" % (ng, run(False, True)))
print("%d GPU(s): both directions: %.1f GB/s aggregate each way" % (ng, run(True, True)))
