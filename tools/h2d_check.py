"""Developer diagnostic: pinned-host -> device copy bandwidth per rank when all ranks copy at once (torchrun)."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(78643200, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
for mode in ("alone" if world == 1 else "all ranks at once", "one rank at a time"):
    for r in range(world if mode == "one rank at a time" else 1):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        if mode == "one rank at a time" and r != rank: 
            if world > 1: dist.barrier()
            continue
        t0 = time.time()
        for _ in range(20): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); dt = time.time() - t0
        print("rank %d %s: %.1f GB/s" % (rank, mode, 20 * h.numel() / dt / 1e9), flush=True)
        if mode == "one rank at a time" and world > 1: dist.barrier()
