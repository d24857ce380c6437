"""Genome-axis sharding for multi-GPU runs.

Sites are independent units: a site's result depends only on the reads spanning it, and its deletion
columns / depth on the site immediately to its left (R:src/exe/bam-readcount/bamreadcount.cpp:393,
R:src/lib/bamrc/IndelQueue.cpp:9).  So a region [beg, end) is cut into contiguous shards, one per rank;
shard [b, e) is simply the region (b, e) of the reference's own loop — it computes site b-1 as its halo
(R:…:269 vs :414) and fetches the reads overlapping [b-1, e).  No collective during compute; the only
exchange is the ordered gather of each rank's output to rank 0.

The `-d` max-count rule depends on file order from the start of the *unsharded* region
(V:htslib-1.10/sam.c:4491), so callers that use a small `-d` must apply admission before sharding
(`engine.admitted`) and run the shards with the default max_cnt.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def plan_shards(read_pos: np.ndarray, beg: int, end: int, world_size: int) -> List[Tuple[int, int]]:
    """Cut [beg, end) (0-based, end exclusive) into `world_size` contiguous shards with about equal
    numbers of read starts (the coverage proxy a BAI linear index gives)."""
    assert end > beg and world_size >= 1
    pos = np.asarray(read_pos, dtype=np.int64)
    inside = pos[(pos >= beg) & (pos < end)]
    cuts = [beg]
    for r in range(1, world_size):
        if inside.size:
            c = int(inside[min(inside.size - 1, (inside.size * r) // world_size)])
        else:
            c = beg + (end - beg) * r // world_size
        c = max(c, cuts[-1])
        cuts.append(min(c, end))
    cuts.append(end)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def shard_read_indices(batch, tid: int, b: int, e: int) -> np.ndarray:
    """Reads the index iterator yields for samfetch(b-1, e): every read a site of [b-1, e) can see."""
    return batch.fetch(tid, b - 1, e)


def gather_ordered(local_text: str, rank: int, world_size: int, group=None) -> str:
    """Ordered emit: ranks hold ascending site ranges, so concatenation in rank order is already sorted.
    Uses the process group's backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    if world_size == 1:
        return local_text
    import torch.distributed as dist
    out = [None] * world_size if rank == 0 else None
    dist.gather_object(local_text, out, dst=0, group=group)
    return "".join(out) if rank == 0 else ""
