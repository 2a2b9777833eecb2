"""Developer diagnostic (GPU box): per-layer parity vs the oracle + quick timings. Writes gpurun_out/gpu_check.txt."""
import os, sys, time, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import oracle as orc, ref

out = open(os.path.join(REPO, "gpurun_out", "gpu_check.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()

cfg, wts = fb.default_model()
bmp = os.path.join(fb.ASSETS, "test.bmp")
layers = orc.load_net(cfg, wts, 0, 0)
img, w, h = ref.load_bmp(bmp)
pw_mode = int(os.environ.get("PW_MODE", "0"))
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=4)
net.set_option("pw_mode", pw_mode)
net.set_option("keep_all", 1)
net.net_input(img, w, h)
x = net.input_tensor().copy()
got = net.net_forward()
oo, oraw, ofin = orc.forward(layers, x, net.net.s1, net.net.s2, v6_quirk=True)
worst = 0
for i, L in enumerate(layers):
    if oo[i] is None: continue
    a = net.layer_output(i, 0)
    rel = float(np.abs(a - oo[i]).max() / max(1e-30, np.abs(oo[i]).max()))
    worst = max(worst, rel)
    _, _, kn = net.layer_cost(i)
    if rel > 2e-5 or i < 3 or i in (116, 120, 129): P("layer", i, orc.TYPE_NAMES[L.type], kn, oo[i].shape, "rel", "%.3e" % rel)
P("worst rel over layers", "%.3e" % worst)
P("boxes gpu", got); P("boxes orc", ofin)
P("raw count gpu", len(net.boxes(0, raw=True)), "orc", len(oraw))
if len(got) == len(ofin) and len(got):
    P("max box abs diff", max(abs(float(g[k]) - float(e[k])) for g, e in zip(got, ofin) for k in ("x1", "y1", "x2", "y2")),
      "score diff", max(abs(float(g["score"]) - float(e["score"])) for g, e in zip(got, ofin)))

# batch of 4 synthetic + graph path twice
fr = synth.frames_u8(4)
for rep in range(2):
    net.input_u8(fr, 4, 320, 320, 960); net.forward(); net.detect()
for f in range(4):
    x0, s1, s2 = orc.net_input(fr[f], 320, 320, 320, 320)
    o2, r2, f2 = orc.forward(layers, x0, s1, s2, v6_quirk=True)
    rels = [float(np.abs(net.layer_output(i, f) - o2[i]).max() / np.abs(o2[i]).max()) for i in (0, 57, 120, 129)]
    P("synthetic frame", f, "rel L0/L57/L120/L129", ["%.2e" % r for r in rels], "raw", len(net.boxes(f, raw=True)), len(r2), "final", len(net.boxes(f)), len(f2))
net.close()

# timing at batch 256 (device-resident frames)
B = int(os.environ.get("BATCH", "256"))
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
net.set_option("pw_mode", pw_mode)
fr = synth.frames_u8(8)
big = np.concatenate([fr] * (B // 8), axis=0)
d = fb.DeviceBuffer(big.nbytes).upload(big)
net.input_u8(d.ptr, B, 320, 320, 960, on_device=True)
net.forward(); net.sync()
for rep in range(3):
    net.forward()
net.sync()
t = time.time(); K = 10
for rep in range(K): net.forward()
net.sync(); dt = (time.time() - t) / K
P("batch", B, "forward ms", "%.3f" % (dt * 1e3), "frames/s", "%.0f" % (B / dt), "launches", net.launches_per_forward(), "arena MB", net.get_option("arena_mb"))
ms = net.layer_times(reps=10)
tot_b = 0; rows = []
for i in range(net.layer_num):
    b, fl, kn = net.layer_cost(i)
    if ms[i] > 0:
        gbs = b * B / (ms[i] * 1e-3) / 1e9
        rows.append((i, kn, ms[i], gbs, fl * B / (ms[i] * 1e-3) / 1e12))
P("sum of per-layer ms", "%.3f" % float(ms.sum()))
for r in rows: P("L%-3d %-14s %8.4f ms %8.1f GB/s %6.2f TFLOP/s" % r)
agg = {}
for i, kn, m, g, tf in rows: agg[kn] = agg.get(kn, 0) + m
P(json.dumps({k: round(float(v), 4) for k, v in agg.items()}))
t = time.time(); net.detect(); P("detect ms", (time.time() - t) * 1e3)
