"""Multi-GPU layer: one process per GPU, particles sharded by contiguous POSITION_IN_STACK
range, zero communication while scoring, one sum-reduce of the half-volume accumulators before
merge3d's finalise.

This mirrors how the reference parallelises the same path with OS processes and files:
  * range split           src/pyp/system/local_run.py:507-516 (`increment = ceil(frames/cores)`)
  * parameter gather      Parameters.merge — vstack + sort by POSITION_IN_STACK
                          (src/pyp/inout/metadata/cistem_star_file.py:656-692)
  * volume sum-reduce     local_merge3d / merge3d over dump files
                          (src/pyp/refine/frealign/frealign.py:1878-1888, 2075-2093)
torch.distributed is plumbing only (NCCL on GPUs; gloo in the CPU tests).
"""
import math

import numpy as np

from ._lib import ROW_DTYPE


def split_ranges(frames, cores):
    """The reference's exact ranges: 1-based inclusive, `increment + 1` rows each."""
    if frames <= 0 or cores <= 0:
        return []
    increment = math.ceil(frames / cores)
    out = []
    for first in range(1, frames + 1, increment + 1):
        out.append((first, min(first + increment, frames)))
    return out


def shard_range(first, last, rank, world):
    """Contiguous shard of [first, last] (1-based inclusive) for `rank`; sizes differ by <= 1.
    Contiguity keeps POSITION_IN_STACK order so a plain concatenation is already sorted."""
    n = last - first + 1
    if n <= 0 or world <= 0:
        return (first, first - 1)
    base, extra = divmod(n, world)
    lo = first + rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0) - 1
    return (lo, hi)


def shard_entities(ids, rank, world):
    """CSP shards whole entities: all tilts of a particle (modes 1/2/5) or all particles of a tilt
    (modes 0/3/4/6) stay on one GPU, because the objective couples them (local_run.py:411-431 hands
    contiguous index ranges of ptlind_list / scanord_list to its processes).  Returns the sorted
    unique ids of this rank's contiguous slice."""
    u = np.unique(np.asarray(ids))
    lo, hi = shard_range(0, u.size - 1, rank, world)
    return u[lo:hi + 1]


def merge_entity_tables(base, parts, key):
    """Overlay refined extended-table entries on the input table, the way Parameters.merge updates
    its particle / tilt dictionaries (cistem_star_file.py:674-690): entries of later parts win."""
    out = np.array(base, copy=True)
    keys = key if isinstance(key, (tuple, list)) else (key,)
    index = {tuple(int(r[k]) for k in keys): i for i, r in enumerate(out)}
    for part in parts:
        for r in part:
            index_key = tuple(int(r[k]) for k in keys)
            if index_key in index:
                out[index[index_key]] = r
    return out


def gather_table(table, dst=0):
    """Gather variable-length structured tables (extended-table entries) on `dst` as a list per rank."""
    import torch

    d = _dist()
    rank, ws = world()
    table = np.ascontiguousarray(table)
    if ws == 1:
        return [table]
    item = table.dtype.itemsize
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
    d.all_gather(counts, torch.tensor([table.size], dtype=torch.int64))
    counts = [int(c.item()) for c in counts]
    mx = max(counts + [1])
    buf = torch.zeros(mx * item, dtype=torch.uint8)
    if table.size:
        buf[: table.size * item] = torch.from_numpy(table.view(np.uint8).reshape(-1).copy())
    bufs = [torch.zeros_like(buf) for _ in range(ws)] if rank == dst else None
    d.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    return [b.numpy()[: c * item].view(table.dtype).copy() for b, c in zip(bufs, counts)]


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: run (and allocate pinned host memory) on the CPUs of the GPU's own NUMA
    node, so that host->device copies do not cross the socket interconnect.  Uses NVML's ideal CPU
    affinity of the device; silently does nothing when NVML or sched_setaffinity is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def _dist():
    import torch.distributed as dist

    return dist


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d.is_available() and d.is_initialized() else (0, 1)


def gather_rows(rows, dst=0):
    """Gather variable-length row tables on `dst`, sorted by POSITION_IN_STACK (Parameters.merge)."""
    import torch

    d = _dist()
    rank, ws = world()
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    if ws == 1:
        return rows[np.argsort(rows["position_in_stack"], kind="stable")]
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
    d.all_gather(counts, torch.tensor([rows.size], dtype=torch.int64))
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    buf = torch.zeros(mx * ROW_DTYPE.itemsize, dtype=torch.uint8)
    if rows.size:
        buf[: rows.size * ROW_DTYPE.itemsize] = torch.from_numpy(rows.view(np.uint8).reshape(-1).copy())
    bufs = [torch.zeros_like(buf) for _ in range(ws)] if rank == dst else None
    d.gather(buf, bufs, dst=dst)
    if rank != dst:
        return None
    parts = [b.numpy()[: c * ROW_DTYPE.itemsize].view(ROW_DTYPE) for b, c in zip(bufs, counts)]
    allrows = np.concatenate(parts) if parts else np.zeros(0, ROW_DTYPE)
    return allrows[np.argsort(allrows["position_in_stack"], kind="stable")]


def reduce_sum(tensor, dst=0):
    """Sum-reduce an accumulator tensor onto `dst` (ncclReduce over NVLink on GPUs)."""
    d = _dist()
    if world()[1] > 1:
        d.reduce(tensor, dst=dst, op=d.ReduceOp.SUM)
    return tensor


def allreduce_noise_curve(ring_sums, n_images):
    """Partition-invariant whitening: every rank contributes its per-ring |F|^2 sums and image
    count; all ranks end with the same mean curve."""
    import torch

    d = _dist()
    t = torch.as_tensor(np.concatenate([np.asarray(ring_sums, dtype=np.float64), [float(n_images)]]))
    if world()[1] > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    t = t.numpy()
    return (t[:-1] / max(t[-1], 1.0)).astype(np.float32)


def allreduce_parameter_statistics(rows):
    """`<name>_stat.cistem` over the rows of ALL ranks (a 2 x 32 float allreduce): every rank
    contributes per-column (count, sum, sum of squares) of its shard in float64; all ranks get the same two
    rows (means, variances) as `tables.parameter_statistics` would give on the merged table."""
    import torch

    from . import tables

    names = rows.dtype.names
    m = np.zeros(2 * len(names) + 1, dtype=np.float64)
    for k, name in enumerate(names):
        col = rows[name].astype(np.float64)
        m[k], m[len(names) + k] = col.sum(), (col * col).sum()
    m[-1] = rows.size
    t = torch.as_tensor(m)
    if world()[1] > 1:
        _dist().all_reduce(t, op=_dist().ReduceOp.SUM)
    m = t.numpy()
    return tables.statistics_from_moments(rows.dtype, m[-1], m[:len(names)], m[len(names):-1])


def device_tensor(ptr, nfloats, device):
    """torch float32 view of an engine-owned device buffer (for NCCL on the accumulators)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=device)
