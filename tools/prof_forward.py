"""Plain-launch forward passes at batch 256 for ncu (set FFCNN_GRAPH=0). usage: prof_forward.py [batch] [passes]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import ffcnn_b200 as fb
from ffcnn_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg, wts = fb.default_model()
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
net.set_option("graph", 0)
fr = synth.frames_u8(8)
big = np.concatenate([fr] * (B // 8), axis=0)
d = fb.DeviceBuffer(big.nbytes).upload(big)
for _ in range(passes):
    net.input_u8(d.ptr, B, 320, 320, 960, on_device=True)
    net.forward()
    net.detect_enqueue()
net.sync()
print("done", net.launches_per_forward())
