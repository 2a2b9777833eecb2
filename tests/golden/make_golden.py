"""Regenerates the committed golden vectors by RUNNING the compiled, unmodified reference (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py

The reference repo ships no golden vectors or tests of its own, so parity is pinned on outputs of the
reference executed here.  Files written (all small):
  testbmp_320.npz      test.bmp through net_load(cfg,w,0,0): heads, per-layer checksums, raw + final boxes (v6, v6_O2, v0)
  testbmp_640x448.npz  test.bmp through the stock main() geometry (net sized to the bmp): boxes + head checksums
  synth_320.npz        seeded synthetic u8 frames (ffcnn_b200.synth): heads of frame 0, per-layer checksums, box counts
  groupconv_cases.npz  the operator seam (conv.h:4-7) on seeded small cases: v6_O2 and v0 outputs
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import oracle as orc, ref          # noqa: E402
from ffcnn_b200 import synth                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
A = orc.ASSETS
CFG, WTS, BMP = A + "/yolo-fastest-1.1.cfg", A + "/yolo-fastest-1.1.weights", A + "/test.bmp"


def checksums(outs):
    s = np.zeros(len(outs)); m = np.zeros(len(outs))
    for i, a in enumerate(outs):
        if a is not None:
            s[i] = a.astype(np.float64).sum(); m[i] = np.abs(a).max()
    return s, m


def run_net(variant, bgr, w, h, iw, ih):
    rn = ref.RefNet(CFG, WTS, iw, ih, variant)
    rn.input_bgr(bgr, w, h)
    x = rn.input_tensor().copy()
    outs, raw, fin = rn.forward_dump()
    rn.close()
    return x, outs, raw, fin


def head_ids(outs_info):
    return [i - 1 for i, t in enumerate(outs_info) if t == orc.YOLO]


def main():
    img, w, h = ref.load_bmp(BMP)
    layers = orc.load_net(CFG, WTS, 0, 0)
    heads = [i - 1 for i, L in enumerate(layers) if L.type == orc.YOLO]

    # ---- test.bmp @ 320x320
    d = {}
    for v in ("v6", "v6_O2", "v0"):
        x, outs, raw, fin = run_net(v, img, w, h, 0, 0)
        s, m = checksums(outs)
        d[f"{v}_sum"], d[f"{v}_maxabs"], d[f"{v}_raw"], d[f"{v}_final"] = s, m, raw, fin
        if v != "v6":                                          # -Ofast heads differ only by re-association noise
            for hid in heads:
                d[f"{v}_head{hid}"] = outs[hid]
        if v == "v6_O2":
            for lid in (0, 8, 10, 57, 108, 114, 116, 124):     # a few full tensors of small/interesting layers
                if outs[lid].size <= 120 * 20 * 20 or lid in (116,):
                    d[f"{v}_layer{lid}"] = outs[lid]
    d["input_checksum"] = np.array([x.astype(np.float64).sum()])
    np.savez_compressed(OUT + "/testbmp_320.npz", **d)

    # ---- test.bmp @ bmp size (stock main(): 640x424 -> 640x448)
    d = {}
    for v in ("v6", "v6_O2", "v0"):
        x, outs, raw, fin = run_net(v, img, w, h, w, h)
        s, m = checksums(outs)
        d[f"{v}_sum"], d[f"{v}_maxabs"], d[f"{v}_raw"], d[f"{v}_final"] = s, m, raw, fin
    d["net_wh"] = np.array([x.shape[2], x.shape[1]])
    np.savez_compressed(OUT + "/testbmp_640x448.npz", **d)

    # ---- seeded synthetic frames (S1) and picture-derived frames (S2)
    d = {}
    fr = synth.frames_u8(4)
    for f in range(4):
        for v in ("v6_O2", "v0"):
            x, outs, raw, fin = run_net(v, fr[f], 320, 320, 0, 0)
            s, m = checksums(outs)
            d[f"s1_f{f}_{v}_sum"], d[f"s1_f{f}_{v}_maxabs"] = s, m
            d[f"s1_f{f}_{v}_raw"], d[f"s1_f{f}_{v}_final"] = raw, fin
            if f == 0 and v == "v6_O2":
                for hid in heads:
                    d[f"s1_f0_{v}_head{hid}"] = outs[hid]
    s2 = synth.shifted_frames_from(img, w, h, 20)
    for f in (0, 3, 7, 19):
        x, outs, raw, fin = run_net("v6_O2", s2[f], 320, 320, 0, 0)
        d[f"s2_f{f}_raw"], d[f"s2_f{f}_final"] = raw, fin
        s, m = checksums(outs)
        d[f"s2_f{f}_sum"] = s
    np.savez_compressed(OUT + "/synth_320.npz", **d)

    # ---- operator seam cases: (iw, ih, ic, groups, pad(eff), stride, fs, fn, act)
    cases = [
        (12, 10, 8, 1, 0, 1, 1, 12, 2),     # 1x1 fast path (conv-v6.c:481)
        (7, 5, 4, 1, 0, 1, 1, 7, 0),        # 1x1, oc not a multiple of 4 (tail loop conv-v6.c:80-90)
        (9, 7, 16, 1, 0, 1, 1, 255 % 16 + 3, 1),
        (11, 9, 6, 6, 1, 1, 3, 6, 2),       # dw3x3 s1 (conv-v6.c:487)
        (2, 5, 3, 3, 1, 1, 3, 3, 2),        # dw3x3 s1, ow == 2 special case (conv-v6.c:154-161)
        (1, 4, 2, 2, 1, 1, 3, 2, 0),        # dw3x3 s1, ow == 1
        (12, 10, 5, 5, 1, 2, 3, 5, 2),      # dw3x3 s2 even (conv-v6.c:493)
        (11, 9, 4, 4, 1, 2, 3, 4, 0),       # dw3x3 s2 odd sizes
        (10, 10, 6, 6, 2, 1, 5, 6, 2),      # dw5x5 (conv-v6.c:499) -- the row oh-2 quirk
        (7, 6, 3, 3, 2, 1, 5, 3, 0),
        (20, 20, 4, 4, 2, 1, 5, 4, 2),
        (16, 12, 3, 1, 1, 2, 3, 8, 2),      # generic im2row path (stem geometry)
        (9, 8, 6, 2, 1, 1, 3, 4, 2),        # generic, grouped with 3 in-ch per group
        (8, 8, 4, 1, 0, 2, 2, 5, 1),        # generic, even kernel, no pad, stride 2, relu
        (6, 6, 8, 1, 0, 2, 1, 4, 0),        # 1x1 stride 2 -> generic
        (10, 7, 4, 4, 1, 1, 3, 8, 2),       # channel multiplier 2: gc_ic == 1 but oc != ic (v6 fast path mis-handles; v0 is truth)
    ]
    rng = np.random.default_rng(20261017)
    d = {"cases": np.array(cases, np.int32)}
    for n, (iw, ih, ic, g, pad, st, fs, fn, act) in enumerate(cases):
        k = fs * fs * (ic // g)
        row = ((k + 3) & ~3) + 4
        x = rng.standard_normal((ic, ih, iw)).astype(np.float32)
        f = np.zeros((fn, row), np.float32)
        f[:, :k] = (rng.standard_normal((fn, k)) / np.sqrt(k)).astype(np.float32)
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn).astype(np.float32)
        f[:, row - 3] = rng.uniform(-0.5, 0.5, fn).astype(np.float32)
        d[f"x{n}"], d[f"f{n}"] = x, f
        d[f"v0_{n}"] = ref.groupconv(x, f, iw, ih, ic, g, pad, st, fs, fn, act, "v0")
        if fn == ic or g == 1 or ic // g > 1:      # v6 fast paths assume oc == ic for depthwise
            d[f"v6_{n}"] = ref.groupconv(x, f, iw, ih, ic, g, pad, st, fs, fn, act, "v6_O2")
    np.savez_compressed(OUT + "/groupconv_cases.npz", **d)
    for fn_ in os.listdir(OUT):
        if fn_.endswith(".npz"):
            print(fn_, os.path.getsize(os.path.join(OUT, fn_)))


if __name__ == "__main__":
    main()
