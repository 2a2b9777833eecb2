"""GPU (B200): random darknet graphs (tests/cfg_fuzz.py -- every layer type, ragged maps, grouped / dense / depthwise convs
of sizes 1/3/5, both pool spellings, upsample, 1-4-input routes, shortcuts, 1-2 yolo heads) through the loader and the
engine.  The CPU half (host loader and oracle against the compiled reference, bit-exact) is tests/test_host.py and
tests/test_oracle.py; this is the engine half: every layer of the layer-by-layer plan against the oracle, and the default
fused plan's candidates against the oracle's.  Seeds are fixed; tools/graph_fuzz_gpu.py runs the same check on more."""
import os

import numpy as np
import pytest

import ffcnn_b200 as fb
from oracle import oracle as orc
import cfg_fuzz
from conftest import note_feat

pytestmark = pytest.mark.gpu
FEAT_TOL = 2e-5


def check_graph(cfg, wts, seed):
    rng = np.random.default_rng(seed)
    layers = orc.load_net(cfg, wts, 0, 0)
    W, H = layers[0].w, layers[0].h
    n = 3
    pitch = (3 * W + 3) & ~3
    frames = rng.integers(0, 256, (n, H, pitch), dtype=np.uint8)
    want = []
    for f in range(n):
        x, s1, s2 = orc.net_input(frames[f], W, H, W, H)
        want.append(orc.forward(layers, x, s1, s2, True))
    kernels = set()
    for keep in (1, 0):                                     # 1: layer-by-layer plan, every tensor readable; 0: the default fused plan
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=n)
        net.set_option("keep_all", keep)
        for _ in range(2):                                  # second pass replays the CUDA graph
            net.detect_batch_u8(frames, n, W, H, pitch)
        for f in range(n):
            outs, raw, fin = want[f]
            if keep:
                for i, o in enumerate(outs):
                    if o is None or o.size == 0:
                        continue
                    got = net.layer_output(i, f)
                    err = float(np.abs(got - o).max() / max(1e-30, float(np.abs(o).max())))
                    note_feat(err)
                    assert err < FEAT_TOL, (i, layers[i].type, f, err, net.layer_cost(i)[2])
                    kernels.add(net.layer_cost(i)[2])
            graw = net.boxes(f, raw=True)
            # a candidate whose confidence sits within rounding of the threshold may flip on random weights; far off is a failure
            assert abs(len(graw) - len(raw)) <= max(2, len(raw) // 200), (keep, f, len(graw), len(raw))
            if len(graw) == len(raw):
                assert [int(t) for t in graw["type"]] == [int(t) for t in raw["type"]]
        net.close()
    return kernels


@pytest.mark.parametrize("case", range(12))
def test_random_graph_against_oracle(case, tmp_path):
    rng = np.random.default_rng(20261017)
    for _ in range(case + 1):                               # case k = the k-th graph of one fixed stream
        text, convs, (W, H) = cfg_fuzz.gen(rng)
        wbytes = cfg_fuzz.weights(rng, convs)
    cfg, wts = str(tmp_path / "g.cfg"), str(tmp_path / "g.weights")
    with open(cfg, "w", newline="") as f:
        f.write(text)
    with open(wts, "wb") as f:
        f.write(wbytes)
    kernels = check_graph(cfg, wts, 1000 + case)
    assert kernels
