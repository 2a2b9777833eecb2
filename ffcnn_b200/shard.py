"""Batched-image frontend over several GPUs: one process per GPU (torchrun), frames are independent so they are
sharded contiguously across ranks with NO data-path collective; the only thing on the wire is one broadcast of the
packed weight buffer (NET.weight_buf, 356 576 floats for yolo-fastest-1.1) from rank 0 at load time
(SURVEY 8e).  torch.distributed is plumbing here: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: GPU g of G gets frames [g*N/G, (g+1)*N/G)."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError((n_frames, rank, world))
    return rank * n_frames // world, (rank + 1) * n_frames // world


class _CudaArrayView:
    """Wraps a raw device pointer so torch can view it (``__cuda_array_interface__`` v2)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def broadcast_weights(net, dist, device=None) -> int:
    """Broadcast rank 0's packed weights into every rank's device buffer, then rebuild the kernel-side layouts.

    GPU path: the NCCL broadcast writes straight into the library's device buffer (no host round trip).
    Host path (gloo tests, net not attached): broadcasts NET.weight_buf on the host. Returns bytes broadcast."""
    import torch
    if getattr(net, "attached", False):
        ptr, n = net.packed_weights_device()
        t = torch.as_tensor(_CudaArrayView(ptr, n), device=device if device is not None else "cuda")
        dist.broadcast(t, src=0)
        torch.cuda.current_stream().synchronize()
        net.commit_weights()
        return n * 4
    n = net.net.weight_size
    host = np.ctypeslib.as_array(net.net.weight_buf, shape=(n,))
    t = torch.from_numpy(host)
    dist.broadcast(t, src=0)
    return n * 4


def gather_box_counts(counts: list[int], dist) -> list[int]:
    """Frame-ordered concatenation of per-rank box counts (boxes themselves stay on their rank's host)."""
    import torch
    world = dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(counts)], dtype=torch.int64))
    mx = int(max(int(s) for s in sizes))
    pad = torch.full((mx,), -1, dtype=torch.int64)
    pad[:len(counts)] = torch.tensor(counts, dtype=torch.int64)
    bufs = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out: list[int] = []
    for r in range(world):
        out += [int(v) for v in bufs[r][:int(sizes[r])]]
    return out


def bind_near_gpu(device_index: int, uuid: str | None = None) -> list[int] | None:
    """Pin this process to the CPUs NVML reports as local to the GPU, so pinned host frames are allocated on the GPU's NUMA
    node and the H2D copies of the e2e path do not cross the socket interconnect (what `numactl --cpunodebind` does for a
    per-GPU process).  Best effort: returns the CPU list applied, or None when NVML has no answer, the set is not a proper
    subset of the CPUs this process may use, or it has fewer than 4 CPUs.  Call before allocating pinned memory; restore
    with os.sched_setaffinity(0, previous) before starting CPU-side work that should use every core."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByUUID(uuid) if uuid else nv.nvmlDeviceGetHandleByIndex(device_index)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(device_index)
        allowed = os.sched_getaffinity(0)
        words = nv.nvmlDeviceGetCpuAffinity(h, (max(allowed) + 64) // 64)
        near = {b + 64 * w for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        if len(near) < 4 or near == allowed:
            return None
        os.sched_setaffinity(0, near)
        return sorted(near)
    except Exception:
        return None
