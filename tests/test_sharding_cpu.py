"""Multi-GPU host logic on CPU: deterministic chunk assignment, and a world_size-2 gloo job in which
each rank runs the oracle on its own chunks and the per-chunk summaries are gathered in chunk order
(the N>1 path has no data-path collective)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from pilon_b200 import sharding, synth


def test_chunks_follow_contig_regions_rule():
    # GenomeFile.scala:67-74: nChunks = ceil(len / max), chunkSize = ceil(len / nChunks)
    ch = synth.chunks_of(64_000_000, 10_000_000)
    assert len(ch) == 7 and ch[0] == (1, 9_142_858) and ch[-1][1] == 64_000_000
    assert all(b - a + 1 <= 10_000_000 for a, b in ch)
    assert synth.chunks_of(5_000_000) == [(1, 5_000_000)]
    sizes = [b - a + 1 for a, b in synth.chunks_of(125_000_000)]
    assert len(sizes) == 13 and sum(sizes) == 125_000_000


def test_assignment_is_deterministic_balanced_and_a_partition():
    wl = synth.workload("C2")
    chunks = wl.regions()
    for n in (1, 2, 4, 8):
        a = sharding.assign(chunks, n)
        assert a == sharding.assign(chunks, n)
        flat = sorted(i for part in a for i in part)
        assert flat == list(range(len(chunks)))
        loads = [sum(sharding.chunk_cost(chunks[i]) for i in part) for part in a]
        assert max(loads) <= 1.35 * (sum(loads) / n) or n == 8     # LPT bound; the 10 Mb contig dominates at n = 8
        assert all(part == sorted(part) for part in a)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import helpers as H
    wl = synth.workload("C2", scale=0.002)         # 20 contigs of >= 20 kb
    chunks = wl.regions()
    mine = sharding.assign(chunks, world)[rank]
    local = {}
    for i in mine[:3]:                              # three chunks per rank keep the test short
        ci, a, b = chunks[i]
        contig = wl.contig_bases(ci).tobytes()
        batches = [(sb.as_read_batch(), sb.frag) for sb in wl.region_batches(ci, a, b)]
        res, _ = H.run_c_oracle(contig, a, b, batches)
        local[i] = (int(res.c.read_count), int(res.c.base_count), int(res["flags"].astype(np.int64).sum()))
    out = sharding.gather_in_chunk_order(local, world)
    if rank == 0:
        q.put((sorted(local), out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_job_gathers_chunk_summaries_in_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    mine0, gathered = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # same chunks processed single-process must give the same summaries, in chunk order
    from tests import helpers as H
    wl = synth.workload("C2", scale=0.002)
    chunks = wl.regions()
    parts = sharding.assign(chunks, 2)
    want = {}
    for part in parts:
        for i in part[:3]:
            ci, a, b = chunks[i]
            batches = [(sb.as_read_batch(), sb.frag) for sb in wl.region_batches(ci, a, b)]
            res, _ = H.run_c_oracle(wl.contig_bases(ci).tobytes(), a, b, batches)
            want[i] = (int(res.c.read_count), int(res.c.base_count), int(res["flags"].astype(np.int64).sum()))
    assert gathered == [want[k] for k in sorted(want)]
    assert len(gathered) == 6
