"""Developer tool (GPU box): sweep the fused-block tile shape (TH, TW, GC) for one block shape and print the block's time.
usage: blk_sweep.py OH CEXP first_layer "TH,TW,GC" ...      (e.g. blk_sweep.py 40 96 38 8,8,3 10,20,2)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
oh, cexp, layer = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
import ffcnn_b200 as fb
from ffcnn_b200 import synth
cfg, wts = fb.default_model()
B = 256
fr = synth.frames_u8(8)
big = np.concatenate([fr] * (B // 8), axis=0)
d = fb.DeviceBuffer(big.nbytes).upload(big)
for cand in sys.argv[4:]:
    os.environ["FFCNN_BLK_TILE_%d_%d" % (oh, cexp)] = cand
    os.environ["FFCNN_FUSE_BLOCK"] = "2"
    try:
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
        net.input_u8(d.ptr, B, 320, 320, 960, on_device=True)
        net.forward(); net.sync()
        by, fl, name = net.layer_cost(layer)
        if not name.startswith("block"):
            print("%-10s not fused (no kernel instance / does not fit)" % cand, flush=True)
        else:
            lt = net.layer_times(reps=10)
            print("%-10s L%d %.4f ms" % (cand, layer, lt[layer]), flush=True)
        net.close()
    except Exception as ex:
        print(cand, "failed:", ex, flush=True)
