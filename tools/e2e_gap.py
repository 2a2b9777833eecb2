"""Developer diagnostic: where does the gap between the resident step and the e2e step come from?"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import numpy as np, torch
import ffcnn_b200 as fb
from ffcnn_b200 import synth
B = 256
cfg, wts = fb.default_model()
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); net.set_stream(st.cuda_stream)
host = torch.empty((2, B, 320, 960), dtype=torch.uint8).pin_memory()
host.numpy()[:] = np.concatenate([synth.frames_u8(8)] * (B // 8), axis=0)[None]
dev = host.cuda()
def resident(K, bg_copy=False, detect=True):
    cs = torch.cuda.Stream(); tmp = torch.empty_like(dev[0])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(K):
        if bg_copy:
            with torch.cuda.stream(cs): tmp.copy_(host[i & 1], non_blocking=True)
        net.input_u8(dev[i & 1].data_ptr(), B, 320, 320, 960, on_device=True); net.forward()
        if detect: net.detect_enqueue()
    e1.record(st); torch.cuda.synchronize()
    if detect: net.detect_finish()
    return e0.elapsed_time(e1) / K
def e2e(K):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    net.submit_u8(host[0].data_ptr(), B, 320, 320, 960)
    for i in range(K):
        if i + 1 < K: net.submit_u8(host[(i + 1) & 1].data_ptr(), B, 320, 320, 960)
        net.collect()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for _ in range(3): resident(5)
print("resident, no detect       %.3f ms" % resident(50, detect=False))
print("resident + detect_enqueue %.3f ms" % resident(50))
print("resident + background H2D %.3f ms" % resident(50, bg_copy=True))
e2e(5)
print("e2e submit/collect        %.3f ms" % e2e(50))
net.set_option("fuse_input", 0)
print("resident, fuse_input=0    %.3f ms" % resident(50, detect=False))
