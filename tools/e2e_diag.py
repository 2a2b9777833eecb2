import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import numpy as np, torch
import ffcnn_b200 as fb
from ffcnn_b200 import synth
B = 256
h = torch.empty(78643200, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(20): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("H2D alone: %.1f GB/s (%.3f ms per 78.6 MB batch)" % (20 * h.numel() / dt / 1e9, dt / 20 * 1e3), flush=True)
cfg, wts = fb.default_model()
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); net.set_stream(st.cuda_stream)
for NB in (2, 4):
    host = torch.empty((NB, B, 320, 960), dtype=torch.uint8).pin_memory()
    host.numpy()[:] = np.concatenate([synth.frames_u8(8)] * (B // 8), axis=0)[None]
    def e2e(K):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        net.submit_u8(host[0].data_ptr(), B, 320, 320, 960)
        for i in range(K):
            if i + 1 < K: net.submit_u8(host[(i + 1) % NB].data_ptr(), B, 320, 320, 960)
            net.collect()
        e1.record(st); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    e2e(5)
    for rep in range(3):
        print("NB=%d e2e submit/collect %.3f ms/step" % (NB, e2e(100)), flush=True)
    # copy concurrent with compute, both timed
    dev = host.cuda()
    cs = torch.cuda.Stream(); tmp = torch.empty_like(dev[0])
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); c0.record(cs)
    for i in range(50):
        with torch.cuda.stream(cs): tmp.copy_(host[i % NB], non_blocking=True)
        net.input_u8(dev[i % NB].data_ptr(), B, 320, 320, 960, on_device=True); net.forward(); net.detect_enqueue()
    e1.record(st); c1.record(cs); torch.cuda.synchronize(); net.detect_finish()
    print("NB=%d concurrent: compute %.3f ms/step, copies %.3f ms/batch" % (NB, e0.elapsed_time(e1) / 50, c0.elapsed_time(c1) / 50), flush=True)
    del host, dev
