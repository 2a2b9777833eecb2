"""Developer diagnostic: pinned-host -> device copy bandwidth per rank (torchrun): every rank at once, halves of the box,
one rank at a time; plain and write-combined pinned memory.  One line per (mode, rank)."""
import os, sys, time, ctypes, torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 78643200
h = torch.empty(N, dtype=torch.uint8).pin_memory(); d = torch.empty(N, dtype=torch.uint8, device="cuda")
import ffcnn_b200 as fb
wc = fb.lib().ffb_host_alloc_pinned_wc(N)
cudart = ctypes.CDLL("libcudart.so.12") if False else None

def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()

def run(active, src_ptr=None):
    barrier()
    t0 = time.time()
    if active:
        for _ in range(20):
            if src_ptr is None: d.copy_(h, non_blocking=True)
            else: fb.lib().ffb_copy_h2d(d.data_ptr(), src_ptr, N)
        torch.cuda.synchronize()
    dt = time.time() - t0
    barrier()
    return 20 * N / dt / 1e9 if active else 0.0

modes = [("all ranks at once", lambda r: True)]
if world >= 8:
    modes += [("ranks 0-3 only", lambda r: r < 4), ("ranks 4-7 only", lambda r: r >= 4), ("even ranks", lambda r: r % 2 == 0)]
for r0 in range(world):
    modes.append(("rank %d alone" % r0, lambda r, r0=r0: r == r0))
for name, pred in modes:
    bw = run(pred(rank))
    if pred(rank): print("%-20s rank %d: %6.1f GB/s" % (name, rank, bw), flush=True)
bw = run(True, wc)
print("%-20s rank %d: %6.1f GB/s" % ("all, write-combined", rank, bw), flush=True)
if world > 1: dist.destroy_process_group()
