"""CPU: the host-C half of libffcnn_b200.so (cfg/weights loader, net_input, yolo decode, NMS, BMP, net_dump) against
the oracle, and the shape of the C-ABI itself: the library loads without a GPU, exports every symbol the headers
declare, and refuses to compute (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import oracle as orc, ref

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = fb.lib()
    declared = set()
    for hdr in ("ffcnn.h", "conv.h", "bmpfile.h", "ffcnn_b200.h"):
        text = open(os.path.join(REPO, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b((?:ffb|net|bmp)_\w+|groupconv)\s*\(", text))
    assert len(declared) >= 45
    missing = sorted(s for s in declared if not hasattr(L, s))
    assert not missing, missing
    assert declared <= set(fb.EXPORTS) | {"ffb_conv"}, sorted(declared - set(fb.EXPORTS))


def test_struct_layout_matches_reference_abi():
    # LP64 layout of ffcnn.h:16-46 -- callers read NET fields directly (ffcnn.c:583-586)
    # numbers printed by a C program compiled against /root/reference/ffcnn.h (gcc, x86-64)
    assert (C.sizeof(fb.LAYER), C.sizeof(fb.BBOX), C.sizeof(fb.NET)) == (120, 24, 104)
    assert (fb.NET.bbox_list.offset, fb.NET.bbox_num.offset, fb.NET.s1.offset, fb.NET.weight_buf.offset,
            fb.NET.cnntempbuf.offset, fb.NET.timeused.offset) == (16, 24, 32, 48, 56, 68)
    assert (fb.LAYER.data.offset, fb.LAYER.w.offset, fb.LAYER.depend_list.offset, fb.LAYER.anchor_list.offset,
            fb.LAYER.scale_x_y.offset) == (8, 24, 64, 88, 116)
    # and our own header compiled by gcc agrees with the ctypes mirror
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "ffcnn.h"\nint main(){printf("%zu %zu %zu %zu %zu", sizeof(LAYER), sizeof(BBOX), '
           'sizeof(NET), offsetof(NET, weight_buf), offsetof(LAYER, anchor_list)); return 0;}')
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(d, "a.c"), "-o", os.path.join(d, "a")], check=True)
        assert subprocess.run([os.path.join(d, "a")], capture_output=True, text=True).stdout == "120 24 104 48 88"


def test_parse_matches_oracle_loader(assets, oracle_layers):
    cfg, wts, _ = assets
    net = fb.Net(cfg, wts, 0, 0, device=None)
    assert net.layer_num == len(oracle_layers) == 131
    assert net.net.weight_size == 356576 and net.net.bbox_max == 320 * 320 * 3 * 4 // 24
    for i, L in enumerate(oracle_layers):
        a, b = net.layer(i), net.layer(i + 1)
        assert (a.type, a.w, a.h, a.c) == (L.type, L.w, L.h, L.c), i
        if L.type != orc.YOLO:
            assert (b.w, b.h, b.c) == (L.ow, L.oh, L.oc), i
        if L.type == orc.CONV:
            assert (a.fn, a.fs, a.stride, a.groups, a.pad, a.batchnorm, a.activation) == (L.fn, L.fs, L.stride, L.groups, L.pad, L.batchnorm, L.activation)
        if L.type in (orc.SHORTCUT, orc.ROUTE):
            assert list(a.depend_list)[:a.depend_num] == L.deps
        if L.type == orc.YOLO:
            assert a.class_num == 80 and [tuple(p) for p in a.anchor_list] == L.anchors
            assert a.ignore_thres == np.float32(0.45) and a.scale_x_y == 1.0
    mine = net.packed_weights()
    want = np.concatenate([L.filt.reshape(-1) for L in oracle_layers if L.type == orc.CONV])
    assert np.array_equal(mine.view(np.uint32), want.view(np.uint32))           # BN fold bit-exact (ffcnn.c:230-231)
    net.close()


def test_parse_input_override_and_errors(assets, tmp_path):
    cfg, wts, _ = assets
    net = fb.Net(cfg, wts, 640, 424, device=None)           # rounded up to x32 (ffcnn.c:133-134)
    assert net.input_whc == (640, 448, 3) and (net.layer(130).w, net.layer(130).h) == (40, 28)
    net.close()
    with pytest.raises(fb.FfcnnError):
        fb.Net(str(tmp_path / "missing.cfg"), wts, device=None)
    zero = fb.Net(cfg, str(tmp_path / "missing.weights"), device=None)        # reference: zero weights, no error
    assert zero.net.weight_size == 356576 and not zero.packed_weights().any()
    zero.close()
    # minimal hand-written cfg: unknown sections skipped, defaults (stride/groups 0 -> 1), pad flag semantics
    p = tmp_path / "tiny.cfg"
    p.write_text("[net]\nwidth=8\nheight=8\nchannels=4\n[foo]\nx=1\n[conv]\nfilters=8\nsize=3\npad=1\nactivation=relu\n"
                 "[maxpool]\nsize=2\nstride=2\n[avgpool]\nsize=3\n[upsample]\nstride=2\n[route]\nlayers=-1,-4\n")
    t = fb.Net(str(p), None, device=None)
    assert t.layer_num == 5
    l0 = t.layer(0)
    assert (l0.fn, l0.fs, l0.pad, l0.stride, l0.groups, l0.activation) == (8, 3, 1, 1, 1, 1)
    assert (t.layer(2).w, t.layer(2).h) == (4, 4) and t.layer(2).type == 1
    assert (t.layer(5).c, t.layer(5).w) == (16, 8)
    t.close()


def test_net_input_host_matches_oracle(assets):
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    for iw, ih in ((0, 0), (640, 424), (96, 224)):
        net = fb.Net(cfg, None, iw, ih, device=None)
        W, H, _ = net.input_whc
        net.net_input(img, w, h)
        want, s1, s2 = orc.net_input(img, w, h, W, H)
        assert (net.net.s1, net.net.s2) == (s1, s2)
        assert np.array_equal(net.input_tensor().view(np.uint32), want.view(np.uint32))
        net.close()
    net = fb.Net(cfg, None, 0, 0, device=None)
    fr = synth.frames_u8(1)[0]
    net.net_input(fr, 320, 320, mean=(10, 20, 30), norm=(0.5, 0.25, 2.0))
    want, _, _ = orc.net_input(fr, 320, 320, 320, 320, mean=(10, 20, 30), norm=(0.5, 0.25, 2.0))
    assert np.array_equal(net.input_tensor(), want)
    net.close()


def test_host_decode_and_nms_match_reference_goldens(assets, golden):
    """ffb_decode_head_chw + ffb_nms on the reference's own head tensors reproduce its raw and final boxes bit-exactly."""
    cfg, wts, _ = assets
    L = fb.lib()
    L.ffb_decode_head_chw.argtypes = [C.POINTER(fb.LAYER), C.POINTER(C.c_float)] + [C.c_int] * 4 + [C.POINTER(fb.BBOX), C.c_int, C.c_int]
    L.ffb_decode_head_chw.restype = C.c_int
    L.ffb_nms.argtypes = [C.POINTER(fb.BBOX), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ffb_nms.restype = C.c_int
    net = fb.Net(cfg, wts, 0, 0, device=None)
    g = golden["testbmp_320"]
    for variant in ("v6_O2", "v0"):
        boxes = (fb.BBOX * 4096)()
        n = 0
        for yolo_layer, hid in ((121, 120), (130, 129)):
            head = np.ascontiguousarray(g[f"{variant}_head{hid}"])
            n = L.ffb_decode_head_chw(C.byref(net.layer(yolo_layer)), head.ctypes.data_as(C.POINTER(C.c_float)), head.shape[2], head.shape[1], 320, 320, boxes, n, 4096)
        raw = np.frombuffer(C.string_at(boxes, n * 24), fb.BOX_DTYPE)
        assert raw.tobytes() == g[f"{variant}_raw"].tobytes()
        m = L.ffb_nms(boxes, n, 0.5, 1, 640, 320)
        fin = np.frombuffer(C.string_at(boxes, m * 24), fb.BOX_DTYPE)
        assert fin.tobytes() == g[f"{variant}_final"].tobytes()
    assert L.ffb_nms(boxes, 0, 0.5, 1, 1, 1) == 0
    net.close()


def test_no_cpu_fallback(assets):
    if fb.device_count() > 0:
        pytest.skip("a GPU is present")
    cfg, wts, _ = assets
    assert not fb.lib().net_load(cfg.encode(), wts.encode(), 0, 0)             # NULL, message on stderr
    net = fb.Net(cfg, wts, device=None)
    with pytest.raises(fb.FfcnnError, match="no CUDA device"):
        net.attach(0, 1)
    with pytest.raises(fb.FfcnnError):
        net.forward()
    with pytest.raises(fb.FfcnnError):
        fb.groupconv(np.zeros((4, 4, 4), np.float32), np.zeros((4, 8), np.float32), 4, 4, 4, 1, 0, 1, 1, 4, 0)
    net.close()


def test_bmp_roundtrip(assets, tmp_path):
    _, _, bmp = assets
    L = fb.lib()

    class BMP(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("stride", C.c_int), ("cdepth", C.c_int), ("pdata", C.c_void_p)]
    for f in (L.bmp_load, L.bmp_save):
        f.argtypes, f.restype = [C.POINTER(BMP), C.c_char_p], C.c_int
    L.bmp_free.argtypes = [C.POINTER(BMP)]
    L.bmp_rectangle.argtypes = [C.POINTER(BMP)] + [C.c_int] * 7
    L.bmp_getpixel.argtypes = [C.POINTER(BMP), C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    b = BMP()
    assert L.bmp_load(C.byref(b), bmp.encode()) == 0
    want, w, h = ref.load_bmp(bmp)
    assert (b.width, b.height, b.stride, b.cdepth) == (w, h, want.shape[1], 24)
    assert C.string_at(b.pdata, b.stride * b.height) == want.tobytes()
    L.bmp_rectangle(C.byref(b), 5, 6, 50, 40, 0, 255, 0)
    r, g_, bl = C.c_int(), C.c_int(), C.c_int()
    L.bmp_getpixel(C.byref(b), 5, 20, C.byref(r), C.byref(g_), C.byref(bl))
    assert (r.value, g_.value, bl.value) == (0, 255, 0)
    out = str(tmp_path / "o.bmp")
    assert L.bmp_save(C.byref(b), out.encode()) == 0
    b2 = BMP()
    assert L.bmp_load(C.byref(b2), out.encode()) == 0
    assert C.string_at(b2.pdata, b2.stride * b2.height) == C.string_at(b.pdata, b.stride * b.height)
    assert L.bmp_load(C.byref(b2), b"/nonexistent.bmp") == -1
    L.bmp_free(C.byref(b)); L.bmp_free(C.byref(b2))


@pytest.mark.skipif(not ref.available("v6_O2"), reason="oracle/_ref not built")
def test_net_dump_prints_the_reference_table(assets):
    cfg, wts, _ = assets
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import ffcnn_b200 as fb; from oracle import ref\n"
            "which = sys.argv[1]\n"
            "if which == 'mine':\n    n = fb.Net(%r, None, 0, 0, device=None); fb.lib().net_dump(n.p)\n"
            "else:\n    r = ref.RefNet(%r, %r, 0, 0, 'v6_O2'); r.L.net_dump.argtypes=[__import__('ctypes').c_void_p]; r.L.net_dump(r.net)\n") % (REPO, cfg, cfg, wts)
    outs = [subprocess.run([sys.executable, "-c", code, w], capture_output=True, text=True, check=True).stdout for w in ("mine", "ref")]
    assert outs[0] == outs[1] and outs[0].count("\n") == 132


def test_parse_second_graph_matches_oracle_loader(tmp_path):
    """The host C loader on the yolov3-tiny-like widening graph (ffcnn_b200/tinygraph.py): pools, avgpool, relu, grouped conv,
    absolute + relative routes, 2-class heads -- same layer table and bit-identical packed weights as the oracle's loader
    (itself bit-exact against the compiled reference on this graph, tests/test_oracle.py)."""
    from ffcnn_b200 import tinygraph as tg
    cfg, wts = tg.write(str(tmp_path))
    layers = orc.load_net(cfg, wts, 0, 0)
    net = fb.Net(cfg, wts, 0, 0, device=None)
    assert net.layer_num == len(layers) == 17
    for i, L in enumerate(layers):
        a, b = net.layer(i), net.layer(i + 1)
        assert (a.type, a.w, a.h, a.c) == (L.type, L.w, L.h, L.c), i
        if L.type != orc.YOLO:
            assert (b.w, b.h, b.c) == (L.ow, L.oh, L.oc), i
        if L.type == orc.CONV:
            assert (a.fn, a.fs, a.stride, a.groups, a.pad, a.batchnorm, a.activation) == (L.fn, L.fs, L.stride, L.groups, L.pad, L.batchnorm, L.activation), i
        if L.type in (orc.MAXPOOL, orc.AVGPOOL, orc.UPSAMPLE):
            assert a.stride == L.stride and (L.type == orc.UPSAMPLE or a.fs == L.fs), i
        if L.type in (orc.SHORTCUT, orc.ROUTE):
            assert list(a.depend_list)[:a.depend_num] == L.deps, i
        if L.type == orc.YOLO:
            assert a.class_num == 2 and [tuple(p) for p in a.anchor_list] == L.anchors and a.ignore_thres == np.float32(0.3)
    want = np.concatenate([L.filt.reshape(-1) for L in layers if L.type == orc.CONV])
    assert np.array_equal(net.packed_weights().view(np.uint32), want.view(np.uint32))
    net.close()


def _cli():
    fb.build()
    exe = os.path.join(REPO, "ffcnn_b200", "ffcnn_cli")
    assert os.path.exists(exe)
    return exe


def test_cli_argument_and_error_behaviour(assets, tmp_path):
    """tools/ffcnn_cli.c = the reference's test driver (ffcnn.c:552-593): same banner, same message and status for an
    unreadable picture; without a GPU it reports the refused net_load instead of dereferencing NULL like the reference."""
    cfg, wts, bmp = assets
    exe = _cli()
    r = subprocess.run([exe, "1", "/nonexistent.bmp"], capture_output=True, text=True, cwd=tmp_path)
    assert r.stdout.splitlines() == ["file_bmp    : /nonexistent.bmp", "file_cfg    : yolo-fastest-1.1.cfg",
                                     "file_weights: yolo-fastest-1.1.weights", "failed to load bmp file: /nonexistent.bmp !"]
    assert r.returncode == 255                                                       # main() returns -1
    r = subprocess.run([exe, "--batch", cfg], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "usage" in r.stderr
    if fb.device_count() <= 0:
        for args in (["1", bmp, cfg, wts], ["--batch", cfg, wts, bmp]):
            r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=tmp_path)
            assert r.returncode == 2 and "no CPU fallback" in r.stderr
        assert not os.path.exists(tmp_path / "out.bmp")


@pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "oracle", "_ref", "ffcnn_ref_bench")), reason="oracle/_ref not built")
def test_bench_reference_arm_prints_the_contract_line(assets):
    """`bench.py --impl reference`: the compiled reference (conv-v6) on the host cores, one JSON line with the keys the
    driver reads; under torchrun only rank 0 works."""
    import json
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "frames/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"].startswith("frames/sec yolo-fastest-1.1 320x320")
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=60, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(not ref.available("v6_O2"), reason="oracle/_ref not built")
def test_net_input_and_size_rounding_against_live_reference(assets):
    """Random picture sizes (1x1 up), net sizes (including ones net_load rounds up to a multiple of 32, ffcnn.c:133-134), means
    and norms: the host C net_input, the oracle and the compiled reference give the same input tensor bit for bit and the
    same s1/s2 box rescale pair (ffcnn.c:259-289)."""
    cfg, wts, _ = assets
    rng = np.random.default_rng(7)
    for case in range(16):
        nw, nh = int(rng.choice([0, 32, 64, 96, 160, 320, 416])), int(rng.choice([0, 32, 64, 128, 224, 320]))
        if nw and rng.random() < 0.3:
            nw += int(rng.integers(1, 31))
        w, h = (int(rng.integers(1, 500)), int(rng.integers(1, 400))) if case else (1, 1)
        img = rng.integers(0, 256, (h, (3 * w + 3) & ~3), dtype=np.uint8)
        mean, norm = tuple(float(v) for v in rng.uniform(0, 128, 3)), tuple(float(v) for v in rng.uniform(0.001, 0.02, 3))
        r = ref.RefNet(cfg, wts, nw, nh, "v6_O2")
        r.input_bgr(img, w, h, mean, norm)
        want, hd = r.input_tensor().copy(), r.head()
        rs, (RW, RH) = (hd.s1, hd.s2), (r.W, r.H)
        r.close()
        n = fb.Net(cfg, None, nw, nh, device=None)
        assert tuple(n.input_whc[:2]) == (RW, RH), (nw, nh)
        n.net_input(img, w, h, mean=mean, norm=norm)
        assert (n.net.s1, n.net.s2) == rs, (w, h, RW, RH)
        assert np.array_equal(n.input_tensor().view(np.uint32), want.view(np.uint32)), (w, h, RW, RH)
        n.close()
        xo, s1, s2 = orc.net_input(img, w, h, RW, RH, mean=mean, norm=norm)
        assert (s1, s2) == rs and np.array_equal(xo.view(np.uint32), want.view(np.uint32))


@pytest.mark.skipif(not ref.available("v6_O2"), reason="oracle/_ref not built")
def test_loader_against_live_reference_on_random_graphs(tmp_path):
    """40 random darknet graphs (tests/cfg_fuzz.py) with seeded weights, some truncated, some with an input-size override:
    ffb_net_parse builds the same layer table as the compiled reference's net_load (type, geometry in and out, kernel,
    stride, pad, groups, activation -- ffcnn.c:123-211) and the same packed, BN-folded weight buffer bit for bit
    (ffcnn.c:213-236, short reads included); dependency lists and yolo parameters agree with the oracle's loader."""
    import cfg_fuzz
    rng = np.random.default_rng(41)
    kinds = set()
    for case in range(40):
        text, convs, _ = cfg_fuzz.gen(rng)
        cfg, wts = str(tmp_path / ("g%d.cfg" % case)), str(tmp_path / ("g%d.weights" % case))
        with open(cfg, "w", newline="") as f:
            f.write(text)
        cut = None if rng.random() < 0.8 else float(rng.uniform(0.2, 0.95))
        with open(wts, "wb") as f:
            f.write(cfg_fuzz.weights(rng, convs, cut))
        iw, ih = (0, 0) if rng.random() < 0.7 else (int(rng.integers(33, 200)), int(rng.integers(33, 200)))
        r = ref.RefNet(cfg, wts, iw, ih, "v6_O2")
        n = fb.Net(cfg, wts, iw, ih, device=None)
        assert n.layer_num == r.n, case
        for i in range(r.n):
            a, b = n.layer(i), n.layer(i + 1)
            assert [a.type, a.w, a.h, a.c, b.w, b.h, b.c, a.fs, a.stride, a.pad, a.groups, a.activation] == r.info[i], (case, i)
            kinds.add(a.type)
        assert np.array_equal(n.packed_weights().view(np.uint32), r.packed_weights().view(np.uint32)), (case, cut)
        for i, L in enumerate(orc.load_net(cfg, wts, iw, ih)):
            a = n.layer(i)
            if L.type in (orc.SHORTCUT, orc.ROUTE):
                assert list(a.depend_list)[:a.depend_num] == L.deps, (case, i)
            if L.type == orc.YOLO:
                assert a.class_num == L.classes and [tuple(p) for p in a.anchor_list] == L.anchors, (case, i)
                assert a.ignore_thres == np.float32(L.ignore_thresh) and a.scale_x_y == np.float32(L.scale_x_y), (case, i)
        n.close(); r.close()
    assert kinds == set(range(8))                       # every layer type of ffcnn.h:4-14 was exercised


def test_host_decode_and_nms_on_random_graphs(tmp_path):
    """The host C decode + NMS (host_decode.c, what ffb_detect_finish runs on the GPU's candidates) on the head tensors of
    random graphs -- 1-4 classes, random masks / anchors / thresholds / scale_x_y, one or two heads, thousands of candidates:
    same raw candidates and same final boxes, bit for bit, as the oracle's decode (ffcnn.c:438-474, 291-335), which
    tests/test_oracle.py pins against the compiled reference on the same generator."""
    import cfg_fuzz
    L = fb.lib()
    L.ffb_decode_head_chw.argtypes = [C.POINTER(fb.LAYER), C.POINTER(C.c_float)] + [C.c_int] * 4 + [C.POINTER(fb.BBOX), C.c_int, C.c_int]
    L.ffb_decode_head_chw.restype = C.c_int
    L.ffb_nms.argtypes = [C.POINTER(fb.BBOX), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ffb_nms.restype = C.c_int
    rng = np.random.default_rng(303)
    heads = total = 0
    for case in range(30):
        text, convs, _ = cfg_fuzz.gen(rng)
        if "[yolo]" not in text:
            continue
        cfg, wts = str(tmp_path / ("g%d.cfg" % case)), str(tmp_path / ("g%d.weights" % case))
        with open(cfg, "w", newline="") as f:
            f.write(text)
        with open(wts, "wb") as f:
            f.write(cfg_fuzz.weights(rng, convs))
        layers = orc.load_net(cfg, wts, 0, 0)
        W, H = layers[0].w, layers[0].h
        w, h = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        img = rng.integers(0, 256, (h, (3 * w + 3) & ~3), dtype=np.uint8)
        x, s1, s2 = orc.net_input(img, w, h, W, H)
        outs, raw, fin = orc.forward(layers, x, s1, s2, False)
        net = fb.Net(cfg, wts, 0, 0, device=None)
        cap = W * H * 3 * 4 // 24                                             # bbox_max of the reference, ffcnn.c:243
        boxes = (fb.BBOX * cap)()
        n = 0
        for i, Ly in enumerate(layers):
            if Ly.type == orc.YOLO:
                head = np.ascontiguousarray(outs[i - 1])
                n = L.ffb_decode_head_chw(C.byref(net.layer(i)), head.ctypes.data_as(C.POINTER(C.c_float)), head.shape[2], head.shape[1], W, H, boxes, n, cap)
                heads += 1
        assert n == len(raw), (case, n, len(raw))
        assert C.string_at(boxes, n * 24) == raw.tobytes(), case
        m = L.ffb_nms(boxes, n, 0.5, 1, s1, s2)
        assert m == len(fin) and C.string_at(boxes, m * 24) == fin.tobytes(), case
        total += n
        net.close()
    assert heads >= 10 and total > 5000


def test_public_headers_compile_standalone_as_c_and_cpp():
    """include/*.h is the drop-in boundary: each header must compile on its own, as strict C99 (the reference is C) and as C++."""
    inc = os.path.join(REPO, "include")
    for h in sorted(os.listdir(inc)):
        for cc, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
            r = subprocess.run([cc, std, "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", lang, "-"],
                               input='#include "%s"\n' % h, capture_output=True, text=True)
            assert r.returncode == 0, (h, cc, r.stderr[:400])


@pytest.mark.skipif(not os.path.exists("/root/reference/ffcnn.c"), reason="reference sources not present (GPU box)")
def test_reference_driver_source_links_against_the_library(tmp_path):
    """INTEGRATION.md, way A: the reference's own test driver -- main() and its static tick helper, cut out of ffcnn.c where it
    lies, not copied into the repo -- compiles against include/ffcnn.h + include/bmpfile.h and links against
    libffcnn_b200.so with no source change; its bmp error path runs (net_load itself needs a GPU)."""
    fb.build()
    src = open("/root/reference/ffcnn.c", errors="replace").read().splitlines()
    keep, on = [], False
    for line in src:
        if line.startswith("#ifdef WIN32") or line.startswith("#if _TEST_"):
            on = True
        if on:
            keep.append(line)
        if on and line.startswith("#endif"):
            on = False
    main_c = tmp_path / "main_only.c"
    main_c.write_text("\n".join(keep) + "\n")
    exe = str(tmp_path / "ffcnn_ref_driver")
    r = subprocess.run(["gcc", "-O2", "-D_TEST_=1", "-I", os.path.join(REPO, "include"), "-include", "stdio.h", "-include", "stdlib.h",
                        "-include", "stdint.h", "-include", "ffcnn.h", str(main_c), "-L", os.path.join(REPO, "ffcnn_b200"), "-lffcnn_b200",
                        "-Wl,-rpath," + os.path.join(REPO, "ffcnn_b200"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    r = subprocess.run([exe, "1", "/nonexistent.bmp"], capture_output=True, text=True, cwd=tmp_path)
    assert "failed to load bmp file: /nonexistent.bmp !" in r.stdout


@pytest.mark.skipif(not ref.available("v6_O2"), reason="oracle/_ref not built")
def test_net_dump_on_random_graphs_prints_what_the_reference_prints(tmp_path):
    """net_dump + net_profile (ffcnn.c:522-550) on 8 random graphs (pool / upsample / route / shortcut / yolo rows, unknown
    activations): byte-identical stdout to the compiled reference."""
    import cfg_fuzz
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import ffcnn_b200 as fb; from oracle import ref; import ctypes\n"
            "which, cfg, wts = sys.argv[1:4]\n"
            "if which == 'mine':\n    n = fb.Net(cfg, wts, 0, 0, device=None); fb.lib().net_dump(n.p); fb.lib().net_profile(n.p)\n"
            "else:\n    r = ref.RefNet(cfg, wts, 0, 0, 'v6_O2'); r.L.net_dump.argtypes = [ctypes.c_void_p]; r.L.net_dump(r.net)\n"
            "    r.L.net_profile.argtypes = [ctypes.c_void_p]; r.L.net_profile(r.net)\n") % REPO
    rng = np.random.default_rng(91)
    for case in range(8):
        text, convs, _ = cfg_fuzz.gen(rng)
        cfg, wts = str(tmp_path / ("g%d.cfg" % case)), str(tmp_path / ("g%d.weights" % case))
        with open(cfg, "w", newline="") as f:
            f.write(text)
        with open(wts, "wb") as f:
            f.write(cfg_fuzz.weights(rng, convs))
        outs = [subprocess.run([sys.executable, "-c", code, w, cfg, wts], capture_output=True, text=True, check=True).stdout for w in ("mine", "ref")]
        assert outs[0] == outs[1] and outs[0].count("\n") > 10, case


@pytest.mark.skipif(not ref.available("v0"), reason="oracle/_ref not built")
def test_bmp_helpers_against_live_reference(tmp_path):
    """bmp_load / bmp_rectangle (corners partly or wholly outside the picture, as detection boxes can be) / bmp_getpixel /
    bmp_save on 40 random 24-bit pictures from 1x1 up: same BMP struct, same pixel bytes, same saved file as the reference's
    bmpfile.c (bmpfile.c:42-156).  Out-of-picture bmp_getpixel is left out: the reference reads outside its buffer there."""
    class BMP(C.Structure):
        _fields_ = [("width", C.c_int), ("height", C.c_int), ("stride", C.c_int), ("cdepth", C.c_int), ("pdata", C.c_void_p)]

    def bind(L):
        for f in (L.bmp_load, L.bmp_save):
            f.argtypes, f.restype = [C.POINTER(BMP), C.c_char_p], C.c_int
        L.bmp_free.argtypes = [C.POINTER(BMP)]
        L.bmp_rectangle.argtypes = [C.POINTER(BMP)] + [C.c_int] * 7
        L.bmp_getpixel.argtypes = [C.POINTER(BMP), C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
        return L
    M, R = bind(fb.lib()), bind(ref.lib("v0"))
    rng = np.random.default_rng(3)
    u32 = lambda v: np.frombuffer(np.uint32(v).tobytes(), np.uint8)
    for case in range(40):
        w, h = int(rng.integers(1, 70)), int(rng.integers(1, 50))
        pitch = (3 * w + 3) & ~3
        hdr = np.zeros(54, np.uint8)
        hdr[0:2] = (66, 77); hdr[2:6] = u32(54 + pitch * h); hdr[10:14] = u32(54); hdr[14:18] = u32(40)
        hdr[18:22] = u32(w); hdr[22:26] = u32(h); hdr[26:28] = (1, 0); hdr[28:30] = (24, 0); hdr[34:38] = u32(pitch * h)
        path = str(tmp_path / "in.bmp")
        with open(path, "wb") as f:
            f.write(hdr.tobytes() + rng.integers(0, 256, (h, pitch), dtype=np.uint8).tobytes())
        a, b = BMP(), BMP()
        assert M.bmp_load(C.byref(a), path.encode()) == R.bmp_load(C.byref(b), path.encode()) == 0
        assert (a.width, a.height, a.stride, a.cdepth) == (b.width, b.height, b.stride, b.cdepth) == (w, h, pitch, 24)
        pix = lambda q: C.string_at(q.pdata, q.stride * q.height)
        assert pix(a) == pix(b)
        for _ in range(6):
            x1, x2 = sorted(int(v) for v in rng.integers(-10, w + 10, 2))
            y1, y2 = sorted(int(v) for v in rng.integers(-10, h + 10, 2))
            col = [int(v) for v in rng.integers(0, 256, 3)]
            M.bmp_rectangle(C.byref(a), x1, y1, x2, y2, *col); R.bmp_rectangle(C.byref(b), x1, y1, x2, y2, *col)
            assert pix(a) == pix(b), (w, h, x1, y1, x2, y2)
        for _ in range(8):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            va, vb = [C.c_int(-7) for _ in range(3)], [C.c_int(-7) for _ in range(3)]
            M.bmp_getpixel(C.byref(a), x, y, *[C.byref(v) for v in va]); R.bmp_getpixel(C.byref(b), x, y, *[C.byref(v) for v in vb])
            assert [v.value for v in va] == [v.value for v in vb]
        pa, pb = str(tmp_path / "a.bmp"), str(tmp_path / "b.bmp")
        assert M.bmp_save(C.byref(a), pa.encode()) == R.bmp_save(C.byref(b), pb.encode()) == 0
        assert open(pa, "rb").read() == open(pb, "rb").read()
        M.bmp_free(C.byref(a)); R.bmp_free(C.byref(b))
