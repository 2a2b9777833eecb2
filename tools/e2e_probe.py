"""Developer diagnostic (one GPU): where does the pipelined end-to-end step go?  Times (CUDA events) the resident step, the
ffb_submit_u8 / ffb_collect loop, one batch's pinned-host -> device copy alone, and the same copy while forward passes run."""
import os, sys, time, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import ffcnn_b200 as fb
from ffcnn_b200 import synth
B, W, H = 256, 320, 320; PITCH = 960; NB = 4
cfg, wts = fb.default_model()
torch.cuda.set_device(0)
net = fb.Net(cfg, wts, W, H, device=0, max_batch=B)
stream = torch.cuda.ExternalStream(net.stream_handle()) if hasattr(net, "stream_handle") else torch.cuda.current_stream()
host = torch.empty((NB, B, H, PITCH), dtype=torch.uint8).pin_memory()
base = synth.frames_u8(16, W, H)
hv = host.numpy()
for b in range(NB):
    for f in range(B): hv[b, f] = base[(b * 5 + f) % 16]
dev = host.cuda()
K = int(os.environ.get("K", 60))

def resident(k):
    for i in range(k):
        net.input_u8(dev[i % NB].data_ptr(), B, W, H, PITCH, on_device=True); net.forward(); net.detect_enqueue()
def timed(fn, k):
    torch.cuda.synchronize(); t0 = time.time(); fn(k); torch.cuda.synchronize(); return 1e3 * (time.time() - t0) / k
resident(5); net.detect_finish()
print("resident            %.4f ms/step" % timed(resident, K)); net.detect_finish()

def e2e(k):
    D = int(os.environ.get("DEPTH", 2))             # batches queued behind the collected one
    for j in range(min(D, k)): net.submit_u8(host[j % NB].data_ptr(), B, W, H, PITCH)
    for i in range(k):
        if i + D < k: net.submit_u8(host[(i + D) % NB].data_ptr(), B, W, H, PITCH)
        net.collect()
e2e(5)
print("e2e                 %.4f ms/step" % timed(e2e, K))

cs = torch.cuda.Stream()
dst = torch.empty((B, H, PITCH), dtype=torch.uint8, device="cuda")
def copies(k):
    with torch.cuda.stream(cs):
        for i in range(k): dst.copy_(host[i % NB], non_blocking=True)
copies(3)
ms = timed(copies, 20); print("H2D alone           %.4f ms/copy  %.1f GB/s" % (ms, dst.numel() / ms / 1e6))
def both(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for i in range(k): dst.copy_(host[i % NB], non_blocking=True)
        e1.record(cs)
    resident(k)
    torch.cuda.synchronize()
    both.copy_ms = e0.elapsed_time(e1) / k
ms = timed(both, K); net.detect_finish()
print("H2D under forward   %.4f ms/copy  %.1f GB/s ; forward+copy loop %.4f ms/step" % (both.copy_ms, dst.numel() / both.copy_ms / 1e6, ms))
net.close()

# ---- the same two loops on picture-derived frames (every frame yields candidates: the host decode + NMS has work) ----
if os.environ.get("PICTURE", "1") == "1":
    net = fb.Net(cfg, wts, W, H, device=0, max_batch=B)
    raw = np.fromfile(os.path.join(fb.ASSETS, "test.bmp"), np.uint8)
    bw, bh = int(raw[18:22].view("<u4")[0]), int(raw[22:26].view("<u4")[0]); bp = (bw * 3 + 3) & ~3
    img = np.ascontiguousarray(raw[54:54 + bp * bh].reshape(bh, bp)[::-1])
    pic = synth.shifted_frames_from(img, bw, bh, B, W, H).reshape(B, H, PITCH)
    for b in range(NB): hv[b] = pic if b % 2 == 0 else pic[::-1]
    dev = host.cuda()
    resident(5); net.detect_finish()
    print("picture: resident (filter enqueued, never read back)   %.4f ms/step" % timed(resident, K)); net.detect_finish()
    def resident_read(k):
        for i in range(k):
            net.input_u8(dev[i % NB].data_ptr(), B, W, H, PITCH, on_device=True); net.forward(); net.detect_enqueue(); net.detect_finish()
    print("picture: resident + blocking detect_finish each step  %.4f ms/step (boxes %d)" % (timed(resident_read, K), sum(len(net.boxes(f)) for f in range(B))))
    e2e(5)
    print("picture: e2e                                           %.4f ms/step" % timed(e2e, K))
    t0 = time.time(); net.detect_finish(); print("picture: detect_finish alone (host decode + NMS, D2H)   %.4f ms" % (1e3 * (time.time() - t0)))
    net.close()
