"""CPU (gloo, world_size 2): the sharding + ordered-gather host logic.  Each rank computes its contiguous
shard (here with the CPU oracle as the per-shard engine — on the GPU box the CUDA engine takes its place,
see test_engine_parity.py::test_sharded_engine_equals_unsharded) and rank 0's concatenation must equal the
unsharded run byte for byte, including deletions whose anchor and emit sites fall on different shards."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import cases
from bam_readcount_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_region_text(case, flags, b, e):
    from oracle.oracle import Oracle
    name, clen, seq, wb = case["contigs"][0]
    o = Oracle(lib_names=case["lib_names"], **flags)
    sub = case["batch"].select(shard.shard_read_indices(case["batch"], 0, b, e))
    o.region(sub, tid=0, beg=b, end=e, contig=name, chrom_len=clen, ref_seq=seq, ref_win_beg=wb, site_list_mode=True)
    return o.text()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.synthetic_case(L=9000, depth=30, seed=21, regions=((0, 501, 8500),))
    flags = dict(per_lib=True, insertion_centric=True)
    beg, end = 500, 8500
    shards = shard.plan_shards(case["batch"].pos, beg, end, world)
    b, e = shards[rank]
    text = _oracle_region_text(case, flags, b, e)
    full = shard.gather_ordered(text, rank, world)
    if rank == 0:
        q.put((full, shards))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.synthetic_case(L=9000, depth=30, seed=21, regions=((0, 501, 8500),))
    want = _oracle_region_text(case, dict(per_lib=True, insertion_centric=True), 500, 8500)
    assert shards[0][1] == shards[1][0] and shards[0][0] == 500 and shards[1][1] == 8500
    assert full == want
    assert len(want.splitlines()) == 8000


def test_plan_shards_balances_reads_and_covers_region():
    rng = np.random.default_rng(0)
    pos = np.sort(np.concatenate([rng.integers(0, 1000, 9000), rng.integers(1000, 10000, 1000)]))
    sh = shard.plan_shards(pos, 0, 10000, 4)
    assert sh[0][0] == 0 and sh[-1][1] == 10000 and all(sh[i][1] == sh[i + 1][0] for i in range(3))
    counts = [int(((pos >= b) & (pos < e)).sum()) for b, e in sh]
    assert max(counts) - min(counts) <= 0.05 * len(pos)
