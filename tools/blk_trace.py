"""Developer tool: per-stage timeline of CTA 0 (warps 0 and 7) of a fused block kernel launch (needs a -DFFB_BLK_TRACE build:
make -C ffcnn_b200/csrc B=build_trace OUT=../libffcnn_b200_trace.so EXTRA="-DFFB_TC_TRACE -DFFB_BLK_TRACE" ../libffcnn_b200_trace.so).
usage: FFCNN_LIB=.../libffcnn_b200_trace.so FFCNN_BLK_TRACE_SHAPE=96,1 blk_trace.py [batch]      (shape = expanded channels, stride)
events: 1 tile top, 2 after tile barrier, 3 x landed, 10 weights landed, 11 after chunk barrier, 12 expand accumulators ready (tcgen05),
13 stage A done, 14 after barrier, 15 stage B done, 20 epilogue start"""
import os, sys, ctypes as C
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import numpy as np, torch
import ffcnn_b200 as fb
from ffcnn_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = fb.lib(); L.ffb_blk_set_trace.argtypes = [C.c_void_p]
cfg, wts = fb.default_model()
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
net.set_option("graph", 0)
fr = np.concatenate([synth.frames_u8(8)] * (B // 8), axis=0)
d = fb.DeviceBuffer(fr.nbytes).upload(fr)
for _ in range(2):
    net.input_u8(d.ptr, B, 320, 320, 960, on_device=True); net.forward(); net.sync()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
L.ffb_blk_set_trace(buf.data_ptr())
net.input_u8(d.ptr, B, 320, 320, 960, on_device=True); net.forward(); net.sync()
L.ffb_blk_set_trace(None)
t = buf.cpu().numpy().reshape(2, 256, 2)
t0 = min(int(t[w, 0, 1]) for w in range(2) if t[w, 0, 1] > 0)
names = {1: "tile", 2: "tile-bar", 3: "x", 10: "w", 11: "bar", 12: "dfull", 13: "A", 14: "bar", 15: "B", 20: "epi"}
for w, wn in ((0, "warp0"), (1, "warp7")):
    row = []
    for k in range(250):
        ev, ts = int(t[w, k, 0]), int(t[w, k, 1])
        if ev == 0: break
        if ev == 1: row.append("\n   ")
        row.append("%s@%d" % (names.get(ev, str(ev)), ts - t0))
    print(wn, " ".join(row))
