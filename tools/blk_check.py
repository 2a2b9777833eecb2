"""Developer diagnostic (GPU box): fused-block path (fuse_block=1, keep_all=2) vs the oracle on test.bmp and synthetic frames,
then step time at batch 256 with and without block fusion.  Writes gpurun_out/blk_check.txt."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import oracle as orc, ref

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
out = open(os.path.join(REPO, "gpurun_out", "blk_check.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); out.write(s + "\n"); out.flush()

cfg, wts = fb.default_model()
bmp = os.path.join(fb.ASSETS, "test.bmp")
layers = orc.load_net(cfg, wts, 0, 0)
img, w, h = ref.load_bmp(bmp)
if os.environ.get("SKIP_PARITY") != "1":
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=4)
    net.set_option("keep_all", 2)
    P("blocks fused:", net.get_option("blocks"))
    net.net_input(img, w, h)
    x = net.input_tensor().copy()
    got = net.net_forward()
    oo, oraw, ofin = orc.forward(layers, x, net.net.s1, net.net.s2, v6_quirk=True)
    worst = 0
    for i, L in enumerate(layers):
        if oo[i] is None: continue
        a = net.layer_output(i, 0)
        if a is None or a.size == 0: continue
        rel = float(np.abs(a - oo[i]).max() / max(1e-30, np.abs(oo[i]).max()))
        worst = max(worst, rel)
        _, _, kn = net.layer_cost(i)
        if rel > 2e-5 or i in (3, 8, 11, 16, 24, 37, 42, 60, 65, 83, 88, 108, 120, 129): P("layer", i, orc.TYPE_NAMES[L.type], kn, oo[i].shape, "rel", "%.3e" % rel)
    P("worst rel over materialised layers", "%.3e" % worst)
    P("boxes gpu", got); P("boxes orc", ofin)
    if len(got) == len(ofin) and len(got):
        P("max box abs diff", max(abs(float(g[k]) - float(e[k])) for g, e in zip(got, ofin) for k in ("x1", "y1", "x2", "y2")),
          "score diff", max(abs(float(g["score"]) - float(e["score"])) for g, e in zip(got, ofin)))
    fr = synth.frames_u8(4)
    for rep in range(2):
        net.input_u8(fr, 4, 320, 320, 960); net.forward(); net.detect()
    for f in range(4):
        x0, s1, s2 = orc.net_input(fr[f], 320, 320, 320, 320)
        o2, r2, f2 = orc.forward(layers, x0, s1, s2, v6_quirk=True)
        rels = [float(np.abs(net.layer_output(i, f) - o2[i]).max() / np.abs(o2[i]).max()) for i in (3, 57, 108, 120, 129)]
        P("synthetic frame", f, "rel L3/L57/L108/L120/L129", ["%.2e" % r for r in rels], "raw", len(net.boxes(f, raw=True)), len(r2), "final", len(net.boxes(f)), len(f2))
    net.close()

B = int(os.environ.get("BATCH", "256"))
fr = synth.frames_u8(8)
big = np.concatenate([fr] * (B // 8), axis=0)
d = fb.DeviceBuffer(big.nbytes).upload(big)
for fuse in (1, 0):
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
    net.set_option("fuse_block", fuse)
    net.input_u8(d.ptr, B, 320, 320, 960, on_device=True)
    net.forward(); net.sync()
    for _ in range(3): net.forward()
    net.sync()
    t0 = time.time(); K = 30
    for _ in range(K): net.forward()
    net.sync()
    dt = (time.time() - t0) / K
    P("fuse_block=%d batch %d: %.3f ms/step  %.0f frames/s  launches %d" % (fuse, B, dt * 1e3, B / dt, net.launches_per_forward()))
    if fuse:
        lt = net.layer_times(reps=5)
        for i in range(net.layer_num):
            by, fl, name = net.layer_cost(i)
            if lt[i] > 0 and name.startswith("block"):
                P("  L%-3d %-18s %8.4f ms  %7.1f GB/s unfused-algorithmic  %6.2f TFLOP/s" % (i, name, lt[i], by * B / (lt[i] * 1e-3) / 1e9, fl * B / (lt[i] * 1e-3) / 1e12))
    net.close()
