"""Golden vectors for the second darknet graph (ffcnn_b200/tinygraph.py), produced by RUNNING the compiled, unmodified
reference (oracle/_ref) on the generated cfg / weights.  Build container only:  python tests/golden/make_tinygraph.py
Writes tests/golden/tinygraph.npz: every layer output of frame 0 (conv-v0 -O2 and conv-v6 -O2), raw + final boxes of 3 frames."""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import oracle as orc, ref          # noqa: E402
from ffcnn_b200 import tinygraph as tg          # noqa: E402


def main():
    d = {}
    with tempfile.TemporaryDirectory() as tmp:
        cfg, wts = tg.write(tmp)
        fr = tg.frames(3)
        for v in ("v0", "v6_O2"):
            for f in range(3):
                rn = ref.RefNet(cfg, wts, 0, 0, v)
                rn.input_bgr(fr[f], tg.W, tg.H)
                outs, raw, fin = rn.forward_dump()
                rn.close()
                d[f"{v}_f{f}_raw"], d[f"{v}_f{f}_final"] = raw, fin
                if f == 0:
                    for i, a in enumerate(outs):
                        if a is not None:
                            d[f"{v}_L{i}"] = a
    for i in range(64):
        if f"v0_L{i}" in d:                                   # no 5x5 depthwise here: v6 == v0 bit for bit, keep one copy
            assert np.array_equal(d[f"v0_L{i}"].view(np.uint32), d[f"v6_O2_L{i}"].view(np.uint32)), i
            del d[f"v6_O2_L{i}"]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tinygraph.npz"), **d)
    print("tinygraph.npz:", len(d), "arrays; boxes per frame (raw/final):", [(len(d[f"v0_f{f}_raw"]), len(d[f"v0_f{f}_final"])) for f in range(3)])


if __name__ == "__main__":
    main()
