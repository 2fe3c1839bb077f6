"""N>1 host logic on CPU: world_size-2 gloo processes shard the clips, 'process' them and
reduce the statistics exactly like bench.py does under torchrun."""
import os
import socket
import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from deepconvsep_b200.sharding import shard_clips, reduce_stats, gather_stems


def test_shard_clips_partition_and_balance():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lengths = rng.integers(44100, 44100 * 300, size=37)
        parts = [shard_clips(lengths, world, r) for r in range(world)]
        allidx = sorted(i for p in parts for i in p)
        assert allidx == list(range(37))                       # a partition
        loads = [int(lengths[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= lengths.max()        # LPT bound
    assert shard_clips([5] * 256, 8, 3) == list(range(3, 256, 8))  # equal clips: 32 per GPU, round robin


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [44100 * (1 + (i % 5)) for i in range(11)]
    mine = shard_clips(lengths, world, rank)
    secs = sum(lengths[i] for i in mine) / 44100.0
    local = [np.full(3, i, dtype=np.float32) for i in mine]      # stand-in for stems
    tot, mx, chk = reduce_stats(secs, 10.0 * (rank + 1), float(sum(mine)))
    gathered = gather_stems((mine, local), world, rank)
    dist.barrier()
    if rank == 0:
        q.put((tot, mx, chk, [g[0] for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, mx, chk, parts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lengths = [1 + (i % 5) for i in range(11)]
    assert tot == sum(lengths) and mx == 20.0 and chk == sum(range(11))
    assert sorted(parts[0] + parts[1]) == list(range(11))
