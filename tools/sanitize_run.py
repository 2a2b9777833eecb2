"""Tiny fused-plan run for compute-sanitizer (memcheck / racecheck / synccheck): two frames, every fused block shape."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import ffcnn_b200 as fb
from ffcnn_b200 import synth
cfg, wts = fb.default_model()
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=2)
net.set_option("fuse_block", 2)          # every supported chain fused
net.set_option("graph", 0)
fr = synth.frames_u8(2)
net.detect_batch_u8(fr, 2, 320, 320, 960)
print("blocks", net.get_option("blocks"), "launches", net.launches_per_forward(), "boxes", [len(net.boxes(f)) for f in range(2)])
net.close()
