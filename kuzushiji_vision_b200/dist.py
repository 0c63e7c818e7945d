"""Image-sharded multi-GPU harness for the hot path.

Assignment, decode and NMS never look across images, so a global batch is cut into contiguous
per-rank shards (what the reference's ``DistributedSampler`` + ``batch // world_size`` do:
data/build.py:140, engine/trainer.py:286) and the data path has NO collective. The only exchange is
a small all-reduce / all-gather of per-rank counters (``target_scores_sum``, ``n_fg``, pages, kept
boxes) through ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Initialise the default process group when launched by torchrun; no-op for world_size 1."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def bind_to_gpu_numa(device_index):
    """Pin this process (and, through first-touch, the pinned host buffers it allocates afterwards) to the
    CPUs of the NUMA node the GPU hangs off. With one process per GPU feeding page batches from host memory
    (the end-to-end path: 566 MB of head maps per step and GPU), buffers on the far socket halve the H2D rate
    when all eight GPUs pull at once. Returns the node, or None when the topology is not exposed
    (containers without /sys PCI info) - then nothing is changed."""
    try:
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id  # torch >= 2.3
    except Exception:
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = device_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[device_index])
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception:
            return None
    try:
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:  # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n_images, rank, world):
    """Contiguous shard [lo, hi) of ``n_images`` for ``rank``; the first ``n % world`` ranks get one more."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_counters(counters: torch.Tensor):
    """Sum a small counter vector over ranks in place (SUM all-reduce); identity for world_size 1."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def all_gather_summaries(summary: torch.Tensor):
    """Gather a fixed-size per-rank summary vector -> (world, n)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(summary) for _ in range(dist.get_world_size())]
        dist.all_gather(out, summary)
        return torch.stack(out)
    return summary.unsqueeze(0)


def max_over_ranks(value: float, device="cpu"):
    """Max of a per-rank scalar (used for device-timed step durations)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
