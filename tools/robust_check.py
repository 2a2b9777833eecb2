"""robustness: odd batch sizes and a non-square geometry through the fused plan vs the layer-by-layer plan"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import ref
cfg, wts = fb.default_model()
img, w, h = ref.load_bmp(os.path.join(fb.ASSETS, "test.bmp"))
for (nw, nh, n) in ((0, 0, 1), (0, 0, 3), (0, 0, 7), (0, 0, 33), (640, 448, 2), (416, 256, 5), (352, 352, 3)):
    W, H = (nw or 320), (nh or 320)
    frames = synth.shifted_frames_from(img, w, h, n, W, H) if hasattr(synth, "shifted_frames_from") and synth.shifted_frames_from.__code__.co_argcount >= 6 else None
    if frames is None:
        # build frames of the net size by nearest resize of test.bmp + shifts
        ys = (np.arange(H) * h // H); xs = (np.arange(W) * w // W)
        base = img[:, :w * 3].reshape(h, w, 3)[ys][:, xs]
        pitch = (W * 3 + 3) & ~3
        frames = np.zeros((n, H, pitch), np.uint8)
        for f in range(n): frames[f, :, :W * 3] = np.roll(base, (f * 3, f * 5), (0, 1)).reshape(H, W * 3)
    pitch = frames.shape[2]
    res = []
    for fuse in (1, 0):
        net = fb.Net(cfg, wts, nw, nh, device=0, max_batch=n)
        net.set_option("fuse_block", fuse); net.set_option("fuse_tail", fuse)
        net.detect_batch_u8(frames, n, W, H, pitch)
        net.detect_batch_u8(frames, n, W, H, pitch)      # graph replay
        res.append([net.boxes(f) for f in range(n)]); blocks = net.get_option("blocks")
        net.close()
    worst = 0.0; cnt = 0
    for a, b in zip(*res):
        assert len(a) == len(b), (len(a), len(b)); cnt += len(a)
        for x, y in zip(a, b):
            assert int(x["type"]) == int(y["type"])
            worst = max(worst, max(abs(float(x[k]) - float(y[k])) for k in ("x1", "y1", "x2", "y2")))
    print("net %dx%d batch %d: %d boxes, fused vs unfused max |d| = %.2e px" % (W, H, n, cnt, worst), flush=True)
