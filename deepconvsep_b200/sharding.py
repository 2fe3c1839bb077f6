"""Multi-GPU plumbing: the path shards by clip (clips are independent; the reference's only
multi-clip driver is the notebook loop `os.system("python separate_dsd.py ...")`, one OS process
per file, examples/dsd100/separate_multiple.ipynb cell 3).  One process per GPU, static
longest-first assignment, no data-path collective; torch.distributed is used only for a barrier
and for reducing a handful of scalars (audio seconds, elapsed time, checksum)."""
import numpy as np


def shard_clips(lengths, world_size, rank):
    """Indices of the clips rank `rank` processes: longest-processing-time-first greedy
    assignment (balanced to within one clip length), deterministic on every rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    mine = []
    for i in order:
        r = int(np.argmin(load))
        load[r] += lengths[i]
        if r == rank:
            mine.append(int(i))
    return sorted(mine)


def reduce_stats(audio_seconds, elapsed_ms, checksum=0.0, group=None):
    """(sum of audio seconds, max of elapsed ms, sum of checksums) over all ranks; identity
    without an initialised process group."""
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return audio_seconds, elapsed_ms, checksum
    if not (dist.is_available() and dist.is_initialized()):
        return audio_seconds, elapsed_ms, checksum
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    s = torch.tensor([audio_seconds, checksum], dtype=torch.float64, device=dev)
    m = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return float(s[0]), float(m[0]), float(s[1])


def gather_stems(local, world_size, rank, group=None):
    """Optional final gather of per-rank results (list of numpy arrays) to rank 0."""
    import torch.distributed as dist
    out = [None] * world_size if rank == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    return out
