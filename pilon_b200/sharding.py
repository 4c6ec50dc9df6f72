"""Region -> GPU sharding.

Pilon's work units are the chunks produced by GenomeFile.contigRegions (reference
GenomeFile.scala:67-74); each chunk is processed without reference to any other
(GenomeFile.scala:101-120), so they are distributed over the GPUs of one box with no data-path
collective.  The only cross-rank traffic is the host-side gather of per-chunk summaries, in chunk
order, exactly where the reference concatenates its per-chunk outputs (GenomeFile.scala:135-162).

Chunks are never re-cut: chunk-level scalars (mean coverage -> minDepth, GenomeRegion.scala:216-224)
make the chunk boundaries part of the result.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

Chunk = Tuple[int, int, int]          # (contig index, start, stop), 1-based inclusive


def chunk_cost(chunk: Chunk, depth: float = 1.0) -> float:
    """Expected work of a chunk: loci x depth (aligned bases dominate the pass)."""
    _, a, b = chunk
    return (b - a + 1) * depth


def assign(chunks: Sequence[Chunk], world_size: int, depths: Sequence[float] = ()) -> List[List[int]]:
    """Greedy longest-processing-time assignment of chunk indices to ranks.

    Deterministic (ties broken by chunk index, then by rank) so that every rank computes the same
    table without communicating.  Each rank's list is returned in chunk order."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cost = [chunk_cost(c, depths[i] if i < len(depths) else 1.0) for i, c in enumerate(chunks)]
    order = sorted(range(len(chunks)), key=lambda i: (-cost[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += cost[i]
    return [sorted(x) for x in out]


def gather_in_chunk_order(local: Dict[int, object], world_size: int, group=None) -> List[object]:
    """All ranks contribute {chunk index: summary}; everyone gets the summaries back in chunk order.

    Uses torch.distributed object collectives on whatever backend the group has (gloo on CPU in the
    tests, the NCCL job's default group under torchrun); with world_size == 1 no process group is needed."""
    if world_size == 1:
        merged = dict(local)
    else:
        import torch.distributed as dist
        parts: List[Dict[int, object]] = [None] * world_size   # type: ignore[list-item]
        dist.all_gather_object(parts, local, group=group)
        merged = {}
        for p in parts:
            for k, v in p.items():
                if k in merged:
                    raise RuntimeError("chunk %d was processed by two ranks" % k)
                merged[k] = v
    return [merged[k] for k in sorted(merged)]
