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


def weighted_shards(n_frames: int, rates, quantum: int = 8, max_per_rank: int | None = None) -> list[tuple[int, int]]:
    """Contiguous shards [lo, hi) per rank, sized in proportion to each rank's measured ingest rate (frames per unit time).

    The end-to-end path of a multi-GPU box is bounded by the host->device copies, and the GPUs of one box do not all get
    the same share of the host's memory path (profiles/r2h_e2e_8gpu.txt: GPUs 0-3 of the 8 x B200 box get 23 GB/s each when
    all eight copy, GPUs 4-7 get 36).  Equal shards then finish at the pace of the slowest rank; shards proportional to the
    measured rates finish together.  Frames stay contiguous and nothing goes on the wire.  Sizes are multiples of `quantum`
    (except that the last rank absorbs n_frames % quantum), never exceed `max_per_rank`, and always sum to n_frames."""
    world = len(rates)
    if world < 1 or n_frames < 0 or quantum < 1 or any(not (r > 0) for r in rates):
        raise ValueError((n_frames, list(rates), quantum))
    cap = n_frames if max_per_rank is None else max_per_rank
    if cap * world < n_frames:
        raise ValueError("max_per_rank too small for n_frames")
    units, tail = divmod(n_frames, quantum)
    cap_u = [(cap - (tail if i == world - 1 else 0)) // quantum for i in range(world)]   # the last rank also carries the tail
    total = float(sum(rates))
    want = [units * r / total for r in rates]
    size = [min(int(w), cap_u[i]) for i, w in enumerate(want)]
    left = units - sum(size)
    # largest remainders first; a rank at its cap passes its turn
    order = sorted(range(world), key=lambda i: (-(want[i] - int(want[i])), i))
    while left > 0:
        progressed = False
        for i in order:
            if left > 0 and size[i] < cap_u[i]:
                size[i] += 1; left -= 1; progressed = True
        if not progressed:
            raise ValueError("cannot place every frame under max_per_rank")
    out, lo = [], 0
    for i in range(world):
        n = size[i] * quantum + (tail if i == world - 1 else 0)
        out.append((lo, lo + n)); lo += n
    assert lo == n_frames
    return out


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
