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


# ---- host placement -------------------------------------------------------------------------------
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def effective_cores():
    """Host threads this process may really use: the scheduler affinity mask clipped by the cgroup CPU
    quota (cgroup v2 `cpu.max`, v1 `cpu.cfs_quota_us`).  `os.cpu_count()` reports the machine, not the
    lease: sizing a worker pool with it oversubscribes a quota-limited container many times over."""
    import math
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, p = f.read().split()[:2]
            if q != "max":
                quota = float(q) / float(p)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                q = float(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                p = float(f.read())
            if q > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = min(n, max(1, int(math.floor(quota + 1e-9))))
    return max(1, n)


def gpu_numa_node(pci_bus_id):
    """NUMA node of a GPU from its PCI address ('0000:1b:00.0' / '00000000:1B:00.0'), or None."""
    import os
    s = pci_bus_id.strip().lower()
    if s.count(":") == 2 and len(s.split(":")[0]) == 8:
        s = s[4:]                      # nvidia-smi prints an 8-digit domain, sysfs uses 4
    try:
        with open(os.path.join("/sys/bus/pci/devices", s, "numa_node")) as f:
            node = int(f.read())
        return node if node >= 0 else None
    except (OSError, ValueError):
        return None


def bind_to_gpu_numa(device_index):
    """Pin this process (and the threads it starts later) to the CPUs of the NUMA node the GPU hangs off,
    BEFORE pinned host buffers are allocated: first-touch places them on that node, so the 0.6 GB of
    H2D + D2H per step of every rank stays on the GPU's own socket instead of crossing the
    inter-socket link (8 ranks x ~60 GB/s does not fit through it).  Returns a dict describing what was
    done (node None = nothing to bind to: single-node box or no sysfs)."""
    import os
    info = {"gpu": int(device_index), "numa_node": None, "cpus": None}
    bus = None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        if hasattr(p, "pci_bus_id") and hasattr(p, "pci_device_id"):
            bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    except Exception:
        bus = None
    if bus is None:
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
            bus = out[0].strip() if out else None
        except Exception:
            bus = None
    if not bus:
        return info
    node = gpu_numa_node(bus)
    info["pci"] = bus
    if node is None:
        return info
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        mine = cpus & allowed
        if mine:
            os.sched_setaffinity(0, mine)
            info["numa_node"], info["cpus"] = node, len(mine)
    except (OSError, ValueError, AttributeError):
        pass
    return info
