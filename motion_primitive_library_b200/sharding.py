"""Multi-GPU: the path shards over independent units (frontier batches, or whole start/goal queries
— BASELINE.json config 5), one process per GPU, every rank holding a replica of the map.  There is
no data-path collective: the map is broadcast once at set-up (broadcast_array) and ranks exchange
only the per-query results and the counters at the end (torch.distributed: NCCL on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

import numpy as np


def bind_to_gpu_numa(cuda_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the GPU's PCI function ->
    numa_node -> cpulist), BEFORE any pinned host buffer is allocated: cudaHostAlloc takes its pages from
    the allocating thread's node (default local policy), and a rank whose result stream crosses the
    socket interconnect loses PCIe bandwidth once several ranks stream at the same time (round 1: 8 ranks
    reached 0.54 of linear end to end).  Returns what was done; never raises (a box without the sysfs
    entries, or a cgroup that forbids the affinity, keeps the default)."""
    import os

    info = {"bound": False}
    try:
        import torch

        props = torch.cuda.get_device_properties(cuda_index)
        if hasattr(props, "pci_bus_id"):
            path = (f"/sys/bus/pci/devices/{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:"
                    f"{getattr(props, 'pci_device_id', 0):02x}.0/numa_node")
        else:  # older torch: ask NVML (same ordinal unless CUDA_VISIBLE_DEVICES reorders)
            import pynvml

            pynvml.nvmlInit()
            bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(cuda_index)).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
            path = f"/sys/bus/pci/devices/{bus_id.lower()[-12:]}/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if use:
            os.sched_setaffinity(0, use)
            info.update(bound=True, cpus=len(use))
    except Exception as e:  # noqa: BLE001 - best effort by design
        info["error"] = str(e)[:120]
    return info


def shard_slice(n_items: int, rank: int, world: int) -> slice:
    """Contiguous, balanced partition: rank r owns [n*r/world, n*(r+1)/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return slice(n_items * rank // world, n_items * (rank + 1) // world)


def broadcast_array(arr, src: int = 0, group=None, device=None) -> np.ndarray:
    """Set-up collective of the replicated design: the map (and potential / search-region arrays) exist
    on rank `src` and every rank needs a copy before it uploads its replica (one
    broadcast per map, 128 MiB at 512^3; NCCL over NVLink on GPUs, gloo in the CPU tests).
    `arr` is the numpy array on rank src and ignored (may be None) elsewhere; returns the array on
    every rank.  Without torch.distributed (world 1) it returns `arr` unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.ascontiguousarray(arr)
    rank = dist.get_rank(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    # header: dtype string (<= 16 bytes), ndim, shape (<= 8 dims)
    hdr = torch.zeros(16 + 9, dtype=torch.int64, device=dev)
    if rank == src:
        a = np.ascontiguousarray(arr)
        name = a.dtype.str.encode()
        if len(name) > 16 or a.ndim > 8:
            raise ValueError("unsupported array")
        h = np.zeros(25, dtype=np.int64)
        h[: len(name)] = np.frombuffer(name, dtype=np.uint8)
        h[16] = a.ndim
        h[17 : 17 + a.ndim] = a.shape
        hdr.copy_(torch.from_numpy(h))
    dist.broadcast(hdr, src=src, group=group)
    h = hdr.cpu().numpy()
    dtype = np.dtype(bytes(h[:16][h[:16] > 0].astype(np.uint8)).decode())
    shape = tuple(int(x) for x in h[17 : 17 + int(h[16])])
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if rank == src:
        payload = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    else:
        payload = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if nbytes:
        dist.broadcast(payload, src=src, group=group)
    return payload.cpu().numpy().view(dtype).reshape(shape)


def run_sharded(items: np.ndarray, fn, group=None, device=None):
    """Run fn(my_items) -> (structured result array of len(my_items), counters dict) on this rank's
    slice and return (all results in the original order, counters reduced over ranks).
    Counter reduction: keys ending in '_max' are max-reduced (e.g. device seconds), the rest summed.
    Works without torch.distributed (world 1)."""
    import torch
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    sl = shard_slice(len(items), rank, world)
    mine, counters = fn(items[sl])
    mine = np.ascontiguousarray(mine)
    if len(mine) != sl.stop - sl.start:
        raise ValueError("fn must return one result per item")
    if not on or world == 1:
        return mine, dict(counters)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    # results: pad each rank's bytes to the largest shard, all_gather, strip
    per = mine.dtype.itemsize
    longest = max(shard_slice(len(items), r, world).stop - shard_slice(len(items), r, world).start for r in range(world))
    buf = np.zeros(longest * per, dtype=np.uint8)
    buf[: mine.nbytes] = mine.view(np.uint8).reshape(-1)
    send = torch.from_numpy(buf).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = []
    for r in range(world):
        s = shard_slice(len(items), r, world)
        nbytes = (s.stop - s.start) * per
        parts.append(recv[r].cpu().numpy()[:nbytes].view(mine.dtype))
    allres = np.concatenate(parts) if parts else mine
    keys = sorted(counters)
    sums = torch.tensor([float(counters[k]) for k in keys if not k.endswith("_max")], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(counters[k]) for k in keys if k.endswith("_max")], dtype=torch.float64, device=dev)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith("_max"):
            out[k] = float(maxs[mi])
            mi += 1
        else:
            out[k] = float(sums[si])
            si += 1
    return allres, out
