"""CPU: the oracle (oracle/ffcnn_oracle.c + oracle/oracle.py) against the committed golden vectors, which were
produced by RUNNING the unmodified reference (tests/golden/make_golden.py), and -- where oracle/_ref is present --
against the compiled reference live.  Integer/byte work and exact-math fp32 are compared bit-exactly."""
import os

import numpy as np
import pytest

from oracle import oracle as orc, ref
from ffcnn_b200 import synth
from conftest import boxes_close


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_groupconv_cases_bit_exact(golden):
    g = golden["groupconv_cases"]
    for n, (iw, ih, ic, grp, pad, st, fs, fn, act) in enumerate(g["cases"]):
        x, f = g[f"x{n}"], g[f"f{n}"]
        exact = orc.conv_raw(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, v6_quirk=False)
        assert np.array_equal(bits(exact), bits(g[f"v0_{n}"])), f"case {n}: oracle(exact) != conv-v0"
        if f"v6_{n}" in g.files:
            quirk = orc.conv_raw(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, v6_quirk=True)
            assert np.array_equal(bits(quirk), bits(g[f"v6_{n}"])), f"case {n}: oracle(v6) != conv-v6"


def test_v6_quirk_is_row_oh_minus_2(golden):
    """conv-v6's 5x5 depthwise path ignores kernel row 0 on output row oh-2 only (conv-v6.c:422-441)."""
    g = golden["groupconv_cases"]
    n = 8                                            # 10x10 dw5x5 case
    diff = np.abs(g[f"v6_{n}"] - g[f"v0_{n}"]).max(axis=(0, 2))
    assert diff[8] > 1e-3 and np.all(np.delete(diff, 8) < 1e-5)


def test_network_matches_reference_goldens(assets, golden, oracle_layers):
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    x, s1, s2 = orc.net_input(img, w, h, 320, 320)
    g = golden["testbmp_320"]
    assert abs(float(x.astype(np.float64).sum()) - float(g["input_checksum"][0])) < 1e-6
    for variant, quirk in (("v6_O2", True), ("v0", False)):
        outs, raw, fin = orc.forward(oracle_layers, x, s1, s2, v6_quirk=quirk)
        for i, o in enumerate(outs):
            if o is None:
                continue
            assert float(o.astype(np.float64).sum()) == pytest.approx(float(g[f"{variant}_sum"][i]), rel=1e-12, abs=1e-9), (variant, i)
            assert float(np.abs(o).max()) == float(g[f"{variant}_maxabs"][i])
        for hid in (120, 129):
            assert np.array_equal(bits(outs[hid]), bits(g[f"{variant}_head{hid}"]))
        assert raw.tobytes() == g[f"{variant}_raw"].tobytes()
        assert fin.tobytes() == g[f"{variant}_final"].tobytes()
    # -Ofast build of the named oracle: same boxes up to its own re-association noise (<= 9.1e-5 px)
    boxes_close(fin, g["v6_final"], px=2e-4, score=1e-6)


def test_known_answers_from_survey(golden):
    """boxes of test.bmp at 320x320 and at the stock 640x448 geometry."""
    f320 = golden["testbmp_320"]["v6_O2_final"]
    assert [int(t) for t in f320["type"]] == [0, 18, 16]
    assert float(f320["score"][0]) == pytest.approx(0.983754694, abs=1e-8)
    assert float(f320["x1"][0]) == pytest.approx(195.218903, abs=1e-5)
    f640 = golden["testbmp_640x448"]["v6_O2_final"]
    assert [int(t) for t in f640["type"]] == [0, 18, 16]
    assert float(f640["x1"][0]) == pytest.approx(188.843079, abs=1e-5)
    assert len(golden["testbmp_640x448"]["v6_O2_raw"]) == 31
    assert list(golden["testbmp_640x448"]["net_wh"]) == [640, 448]


def test_synthetic_frames_match_goldens(golden, oracle_layers):
    g = golden["synth_320"]
    fr = synth.frames_u8(2)
    for f in range(2):
        x, s1, s2 = orc.net_input(fr[f], 320, 320, 320, 320)
        outs, raw, fin = orc.forward(oracle_layers, x, s1, s2, v6_quirk=True)
        for i, o in enumerate(outs):
            if o is not None:
                assert float(np.abs(o).max()) == float(g[f"s1_f{f}_v6_O2_maxabs"][i]), (f, i)
        assert raw.tobytes() == g[f"s1_f{f}_v6_O2_raw"].tobytes()
        if f == 0:
            for hid in (120, 129):
                assert np.array_equal(bits(outs[hid]), bits(g[f"s1_f0_v6_O2_head{hid}"]))


def test_picture_frames_have_boxes(assets, golden, oracle_layers):
    _, _, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    s2f = synth.shifted_frames_from(img, w, h, 8)
    g = golden["synth_320"]
    for f in (0, 3, 7):
        x, s1, s2 = orc.net_input(s2f[f], 320, 320, 320, 320)
        _, raw, fin = orc.forward(oracle_layers, x, s1, s2, v6_quirk=True)
        assert raw.tobytes() == g[f"s2_f{f}_raw"].tobytes()
        assert fin.tobytes() == g[f"s2_f{f}_final"].tobytes()
        assert len(fin) >= 1


def test_pools_and_resample_edge_cases():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 7, 6)).astype(np.float32)
    # maxpool9 on a 7x6 map == global max over the clamped window; max is idempotent: mp5 == mp3 o mp3 (stride 1)
    m3, m5, m9 = orc.maxpool(x, 3, 1), orc.maxpool(x, 5, 1), orc.maxpool(x, 9, 1)
    assert np.array_equal(orc.maxpool(m3, 3, 1), m5)
    assert np.array_equal(orc.maxpool(m5, 5, 1), m9)
    assert np.array_equal(orc.maxpool(x, 1, 1), x)
    u = orc.upsample(x, 2)
    assert u.shape == (3, 14, 12) and np.array_equal(u[:, ::2, ::2], x) and np.array_equal(u[:, 1::2, 1::2], x)
    a = rng.standard_normal(100).astype(np.float32)
    assert np.array_equal(orc.shortcut(a, -a, 0), np.zeros(100, np.float32))
    assert np.array_equal(orc.shortcut(a, a, 2), np.where(a > 0, a + a, np.float32(0.1) * (a + a)).astype(np.float32))


def test_nms_and_decode_edge_cases():
    boxes = np.zeros(4, orc.BOX_DTYPE)
    assert orc.nms(boxes, 0, 1, 1) == 0                                   # empty list
    boxes[0] = (1, 0.9, 0, 0, 10, 10); boxes[1] = (1, 0.8, 1, 1, 9, 9)    # contained box of same class -> min-area ratio 1
    boxes[2] = (2, 0.7, 1, 1, 9, 9); boxes[3] = (1, 0.6, 20, 20, 30, 30)  # other class / disjoint survive
    n = orc.nms(boxes, 4, 2, 1)
    assert n == 3 and [int(t) for t in boxes["type"][:3]] == [1, 2, 1]
    assert float(boxes["x2"][0]) == 20.0 and float(boxes["score"][3]) == 0.0   # rescaled by s1/s2 = 2, tail zeroed


@pytest.mark.skipif(not ref.available("v0"), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_bit_exact(assets):
    """The restatement against the compiled reference on fresh seeded inputs (not only the committed fixtures)."""
    cfg, wts, _ = assets
    rng = np.random.default_rng(99)
    for (iw, ih, ic, grp, pad, st, fs, fn, act) in [(13, 11, 8, 8, 2, 1, 5, 8, 2), (17, 9, 12, 1, 0, 1, 1, 20, 2), (14, 14, 8, 2, 1, 2, 3, 6, 1)]:
        k = fs * fs * (ic // grp); row = ((k + 3) & ~3) + 4
        x = rng.standard_normal((ic, ih, iw)).astype(np.float32)
        f = np.zeros((fn, row), np.float32); f[:, :k] = rng.standard_normal((fn, k)); f[:, row - 4] = 1.25; f[:, row - 3] = -0.5
        assert np.array_equal(bits(orc.conv_raw(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, False)), bits(ref.groupconv(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, "v0")))
        assert np.array_equal(bits(orc.conv_raw(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, True)), bits(ref.groupconv(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, "v6_O2")))
    fr = synth.frames_u8(1, seed0=0xABC)[0]
    rn = ref.RefNet(cfg, wts, 0, 0, "v6_O2")
    rn.input_bgr(fr, 320, 320)
    x = rn.input_tensor().copy()
    routs, rraw, rfin = rn.forward_dump(want={0, 57, 116, 129})
    layers = orc.load_net(cfg, wts, 0, 0)
    xo, s1, s2 = orc.net_input(fr, 320, 320, 320, 320)
    assert np.array_equal(bits(x), bits(xo))
    outs, raw, fin = orc.forward(layers, xo, s1, s2, True)
    for i in (0, 57, 116, 129):
        assert np.array_equal(bits(outs[i]), bits(routs[i])), i
    assert raw.tobytes() == rraw.tobytes() and fin.tobytes() == rfin.tobytes()


@pytest.mark.skipif(not ref.available("v6_O2"), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_is_immune_to_heap_garbage(assets):
    """The reference's im2row leaves the pad lane of its scratch rows unwritten (conv-v6.c:9-42: 27 taps in rows of 28 for the stem) and
    multiplies it by the filter's zero pad -- NaN * 0 = NaN when malloc hands back memory that held NaNs.  oracle/ref.py makes
    allocations start as zero bytes (glibc M_PERTURB), the state the reference's own fresh process sees: with the heap deliberately
    salted with NaNs the stem output must still be finite and bit-equal to the restatement."""
    cfg, wts, _ = assets
    ref.lib("v6_O2")                                                  # M_PERTURB is set when the library is first loaded
    for _ in range(8):                                               # leave freed chunks full of NaNs in several size classes
        junk = [np.full(n, np.nan, np.float32) for n in (28 * 160, 28 * 320, 4096, 65536, 3 * 322 * 322)]
        del junk
    fr = synth.frames_u8(1, seed0=0x5EED)[0]
    rn = ref.RefNet(cfg, wts, 0, 0, "v6_O2")
    rn.input_bgr(fr, 320, 320)
    routs, _, _ = rn.forward_dump(want={0})
    assert np.isfinite(routs[0]).all()
    layers = orc.load_net(cfg, wts, 0, 0)
    xo, s1, s2 = orc.net_input(fr, 320, 320, 320, 320)
    outs, _, _ = orc.forward(layers, xo, s1, s2, True)
    assert np.array_equal(bits(outs[0]), bits(routs[0]))


@pytest.mark.skipif(not ref.available("v0"), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_random_shapes_bit_exact():
    """160 seeded random operator shapes (ragged maps from 4x4 up, 1x1 / 3x3 s1,s2 / 5x5, dense, grouped, depthwise, all three
    activations) through the restatement and the compiled reference: bit-exact against conv-v0 and against conv-v6.
    Shapes stay inside what the reference itself computes without undefined behaviour:
      * "same" padding (pad = fs/2): im2row walks the INPUT columns (conv-v6.c:16), other paddings overrun its scratch rows;
      * fs*fs*ic/groups a multiple of 4 on the generic path: im2row never writes the pad lanes (conv-v6.c:9-24);
      * depthwise (ic/groups == 1) with one filter per channel, the only form conv-v6's fast paths handle (conv-v6.c:486-503);
      * 5x5 depthwise maps of at least 4x4 (conv-v6.c:291-465 reads out of bounds below that)."""
    rng = np.random.default_rng(2026)
    done = 0
    while done < 160:
        fs = int(rng.choice([1, 3, 5]))
        st = int(rng.choice([1, 2])) if fs == 3 else 1
        pad = fs // 2
        ic = int(rng.choice([4, 8, 12, 16, 24]))
        if fs > 1 and rng.random() < 0.5:
            grp, fn = ic, ic                                                        # depthwise
        else:
            grp = int(rng.choice([1, 1, 2, 4]))
            fn = grp * int(rng.integers(1, 6))
            if grp > 1 and (fs * fs * (ic // grp)) % 4:
                continue
            if ic // grp == 1:
                continue
        iw, ih, act = int(rng.integers(4, 24)), int(rng.integers(4, 24)), int(rng.choice([0, 1, 2]))
        k = fs * fs * (ic // grp); row = ((k + 3) & ~3) + 4
        x = rng.standard_normal((ic, ih, iw)).astype(np.float32)
        f = np.zeros((fn, row), np.float32); f[:, :k] = rng.standard_normal((fn, k))
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
        shape = (iw, ih, ic, grp, pad, st, fs, fn, act)
        assert np.array_equal(bits(orc.conv_raw(x, f, *shape, False)), bits(ref.groupconv(x, f, *shape, "v0"))), shape
        assert np.array_equal(bits(orc.conv_raw(x, f, *shape, True)), bits(ref.groupconv(x, f, *shape, "v6_O2"))), shape
        done += 1
    # the conv-v6 5x5 quirk on the smallest maps it is defined for (4 wide / 4 high)
    for (iw, ih) in ((4, 4), (4, 9), (9, 4), (5, 5)):
        x = rng.standard_normal((4, ih, iw)).astype(np.float32)
        f = np.zeros((4, 32), np.float32); f[:, :25] = rng.standard_normal((4, 25)); f[:, 28] = 1.0
        a, b = orc.conv_raw(x, f, iw, ih, 4, 4, 2, 1, 5, 4, 0, True), ref.groupconv(x, f, iw, ih, 4, 4, 2, 1, 5, 4, 0, "v6_O2")
        assert np.array_equal(bits(a), bits(b)), (iw, ih)
        exact = orc.conv_raw(x, f, iw, ih, 4, 4, 2, 1, 5, 4, 0, False)
        assert [int(r) for r in np.nonzero(np.abs(a - exact).max(axis=(0, 2)) > 1e-4)[0]] == [ih - 2]


def test_second_graph_matches_reference_goldens(tmp_path):
    """Widening case: a yolov3-tiny-like graph (dense 3x3, stride-2 max pools, avgpool, relu, grouped conv,
    relative + absolute routes, two heads) -- the restatement must reproduce the compiled reference bit for bit."""
    from ffcnn_b200 import tinygraph as tg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tinygraph.npz"))
    cfg, wts = tg.write(str(tmp_path))
    layers = orc.load_net(cfg, wts, 0, 0)
    assert [L.type for L in layers].count(orc.YOLO) == 2 and [L.type for L in layers].count(orc.AVGPOOL) == 1
    fr = tg.frames(3)
    for f in range(3):
        x, s1, s2 = orc.net_input(fr[f], tg.W, tg.H, tg.W, tg.H)
        outs, raw, fin = orc.forward(layers, x, s1, s2, True)
        if f == 0:
            for i, o in enumerate(outs):
                if o is not None:
                    assert np.array_equal(bits(o), bits(g[f"v0_L{i}"])), i
        assert raw.tobytes() == g[f"v0_f{f}_raw"].tobytes() and fin.tobytes() == g[f"v0_f{f}_final"].tobytes()
        assert fin.tobytes() == g[f"v6_O2_f{f}_final"].tobytes() and len(fin) > 50


@pytest.mark.skipif(not ref.available("v0"), reason="oracle/_ref not built (needs /root/reference)")
def test_random_graphs_forward_bit_exact_against_live_reference(tmp_path):
    """25 random darknet graphs (tests/cfg_fuzz.py: every layer type, ragged maps, grouped / depthwise / unpadded convs,
    1-4 input routes, one or two yolo heads) on random pictures of random size: every layer output, the pre-NMS candidates
    and the final boxes of the restatement equal the compiled reference's (conv-v0, exact math) bit for bit."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cfg_fuzz
    rng = np.random.default_rng(77)
    boxes = 0
    for case in range(25):
        text, convs, _ = cfg_fuzz.gen(rng)
        cfg, wts = str(tmp_path / ("g%d.cfg" % case)), str(tmp_path / ("g%d.weights" % case))
        with open(cfg, "w", newline="") as f:
            f.write(text)
        with open(wts, "wb") as f:
            f.write(cfg_fuzz.weights(rng, convs))
        r = ref.RefNet(cfg, wts, 0, 0, "v0")
        layers = orc.load_net(cfg, wts, 0, 0)
        w, h = int(rng.integers(20, 200)), int(rng.integers(20, 200))
        img = rng.integers(0, 256, (h, (3 * w + 3) & ~3), dtype=np.uint8)
        r.input_bgr(img, w, h)
        routs, rraw, rfin = r.forward_dump()
        x, s1, s2 = orc.net_input(img, w, h, r.W, r.H)
        outs, raw, fin = orc.forward(layers, x, s1, s2, False)
        for i, (a, b) in enumerate(zip(outs, routs)):
            assert (a is None) == (b is None), (case, i)
            if a is not None:
                assert a.shape == b.shape and np.array_equal(bits(a), bits(b)), (case, i, layers[i].type)
        assert raw[:4096].tobytes() == rraw.tobytes(), case            # the harness returns at most 4096 candidates
        assert fin.tobytes() == rfin.tobytes(), case
        boxes += len(fin)
        r.close()
    assert boxes > 1000
