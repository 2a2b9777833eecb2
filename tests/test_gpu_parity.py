"""GPU (B200): the CUDA path behind the C-ABI against the oracle and the committed reference goldens.

Tolerances (written here as the contract):
  * feature maps:  max|gpu - oracle| <= 2e-5 * max|oracle layer|   (the reference differs from itself by 8.8e-6 between
                   its -O2 and -Ofast builds; fp32 FFMA in a different summation order lands around 1e-6)
  * boxes:         same candidate set (class, cell); scores within 5e-6; coordinates in source-image pixels:
                     BOX_TOL    = 1e-4 px  BASELINE.json's contract -- asserted for the configuration it is quoted on (the default
                                           plan on the 320x320 net: test.bmp, the seeded frames, batch 256) and for the strict
                                           fp32 mode (pw_mode = 1);  measured there: <= 9.2e-5 px
                     BOX_TOL_TC = 1.5e-4   SURVEY 8(d)'s figure for 3xTF32 ("fp32 mode <= 1e-4 px, 3xTF32 ~ 1.5e-4 px"): plans that put MORE
                                           layers on the tensor cores than the default (fuse_block = 2, pw_mode = 2); measured 1.22e-4
                     both scale with net width / 320 on larger nets (coordinates and their fp32 ulps grow with the frame: at
                     640x448 the default plan measures 2.14e-4 px = 7 ulp at y = 345, limit 3e-4).
                   For scale: 1 fp32 ulp at x ~ 600 is 6.1e-5 px, and the reference differs from ITSELF by 9.1e-5 px between
                   its -O2 and -Ofast builds.  3xTF32 carries 22 significand bits per operand (hi + lo of
                   11 each) against fp32's 24, so it sits ~2x above that floor; no summation order can do better.
                   The worst deviation every run measures is printed in the pytest summary (conftest.MEASURED).
  * integer/byte:  net_input's u8 -> fp32 conversion is bit-exact.
"""
import os
import sys

import numpy as np
import pytest

import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import oracle as orc, ref
from conftest import boxes_close, note_feat, REPO

pytestmark = pytest.mark.gpu

FEAT_TOL = 2e-5
BOX_TOL = 1e-4
BOX_TOL_TC = 1.5e-4
SCORE_TOL = 5e-6
PW_MODES = [int(m) for m in os.environ.get("FFCNN_TEST_PW_MODES", "0,1").split(",")]


def rel_err(a, b):
    e = float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))
    note_feat(e)
    return e


@pytest.fixture(scope="module")
def net4(assets):
    cfg, wts, _ = assets
    n = fb.Net(cfg, wts, 0, 0, device=0, max_batch=4)
    n.set_option("keep_all", 1)
    yield n
    n.close()


def test_extension_is_native_and_on_gpu():
    assert fb.device_count() >= 1
    assert os.path.exists(fb.LIB_PATH)


def test_groupconv_seam_against_reference_goldens(golden):
    """conv.h:4-7 through the GPU on every committed case: v6 semantics by default, exact math with FFCNN_DW5_EXACT."""
    g = golden["groupconv_cases"]
    for n, (iw, ih, ic, grp, pad, st, fs, fn, act) in enumerate(g["cases"]):
        x, f = g[f"x{n}"], g[f"f{n}"]
        want = g[f"v6_{n}"] if f"v6_{n}" in g.files else g[f"v0_{n}"]
        got = fb.groupconv(x, f, iw, ih, ic, grp, pad, st, fs, fn, act)
        assert got.shape == want.shape and rel_err(got, want) < FEAT_TOL, (n, rel_err(got, want))
    os.environ["FFCNN_DW5_EXACT"] = "1"
    try:
        for n in (8, 9, 10):
            iw, ih, ic, grp, pad, st, fs, fn, act = g["cases"][n]
            got = fb.groupconv(g[f"x{n}"], g[f"f{n}"], iw, ih, ic, grp, pad, st, fs, fn, act)
            assert rel_err(got, g[f"v0_{n}"]) < FEAT_TOL, n
    finally:
        del os.environ["FFCNN_DW5_EXACT"]


@pytest.mark.parametrize("pw_mode", PW_MODES)
def test_every_layer_of_testbmp_against_oracle(assets, oracle_layers, net4, pw_mode):
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    net4.set_option("pw_mode", pw_mode)
    net4.net_input(img, w, h)
    x = net4.input_tensor().copy()
    got = net4.net_forward()
    outs, raw, fin = orc.forward(oracle_layers, x, net4.net.s1, net4.net.s2, v6_quirk=True)
    worst = 0.0
    for i, o in enumerate(outs):
        if o is None:
            continue
        e = rel_err(net4.layer_output(i, 0), o)
        worst = max(worst, e)
        assert e < FEAT_TOL, (i, e)
    graw = net4.boxes(0, raw=True)
    assert len(graw) == len(raw) and [int(t) for t in graw["type"]] == [int(t) for t in raw["type"]]
    boxes_close(got, fin, px=BOX_TOL, score=SCORE_TOL)
    net4.set_option("pw_mode", 0)


def test_reference_api_flow_and_goldens(assets, golden):
    """net_load -> net_input -> net_forward -> bbox_list exactly as ffcnn.c's main() drives it, both geometries."""
    cfg, wts, bmp = assets
    L = fb.lib()
    img, w, h = ref.load_bmp(bmp)
    mean, norm = (fb.C.c_float * 3)(0, 0, 0), (fb.C.c_float * 3)(1 / 255., 1 / 255., 1 / 255.)
    # the contract (1e-4 px) on the 320x320 net it is quoted on; the 640x448 geometry of the reference's main() gets the
    # 3xTF32 figure scaled by the net width (1.5e-4 * 2; measured there: 2.14e-4 px = 7 ulp at y = 345)
    for (iw, ih, key, tol) in ((0, 0, "testbmp_320", BOX_TOL), (w, h, "testbmp_640x448", BOX_TOL_TC * 2)):
        p = L.net_load(cfg.encode(), wts.encode(), iw, ih)
        assert p
        for _ in range(2):                                                    # second pass replays the CUDA graph
            L.net_input(p, img.ctypes.data, w, h, mean, norm)
            L.net_forward(p)
        net = p.contents
        got = np.frombuffer(fb.C.string_at(net.bbox_list, net.bbox_num * 24), fb.BOX_DTYPE)
        boxes_close(got, golden[key]["v6_O2_final"], px=tol, score=SCORE_TOL)
        boxes_close(got, golden[key]["v6_final"], px=tol + 1.5e-4, score=SCORE_TOL)     # the -Ofast build (own noise 9e-5 px)
        L.net_free(p)


def test_heads_against_committed_goldens(assets, golden, net4):
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    net4.net_input(img, w, h)
    net4.net_forward()
    for hid in (120, 129):
        assert rel_err(net4.layer_output(hid, 0), golden["testbmp_320"][f"v6_O2_head{hid}"]) < FEAT_TOL


def test_batched_u8_path_against_goldens(assets, golden, oracle_layers, net4):
    """ffb_input_u8 (GPU net_input) + batch of 4 distinct frames: input bit-exact, every frame's layers match its golden."""
    g = golden["synth_320"]
    fr = synth.frames_u8(4)
    net4.input_u8(fr, 4, 320, 320, 960)
    net4.forward(); net4.detect()
    for f in range(4):
        want_in, _, _ = orc.net_input(fr[f], 320, 320, 320, 320)
        assert np.array_equal(net4.layer_output(-1, f).view(np.uint32), want_in.view(np.uint32))     # byte work: bit-exact
        for i in (0, 10, 57, 108, 114, 116, 124, 129):
            o = net4.layer_output(i, f)
            assert abs(float(np.abs(o).max()) / float(g[f"s1_f{f}_v6_O2_maxabs"][i]) - 1) < 1e-4, (f, i)
            assert abs(float(o.astype(np.float64).sum()) - float(g[f"s1_f{f}_v6_O2_sum"][i])) <= 2e-5 * float(g[f"s1_f{f}_v6_O2_maxabs"][i]) * o.size
        assert len(net4.boxes(f, raw=True)) == len(g[f"s1_f{f}_v6_O2_raw"])
    for hid in (120, 129):
        assert rel_err(net4.layer_output(hid, 0), g[f"s1_f0_v6_O2_head{hid}"]) < FEAT_TOL


def test_picture_frames_boxes_match_reference(assets, golden):
    """Set S2 (frames derived from test.bmp, shifted): decode + NMS are exercised; boxes equal the reference's."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    s2f = synth.shifted_frames_from(img, w, h, 20)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=20)
    net.detect_batch_u8(s2f, 20, 320, 320, 960)
    g = golden["synth_320"]
    for f in (0, 3, 7, 19):
        want_raw, want = g[f"s2_f{f}_raw"], g[f"s2_f{f}_final"]
        raw = net.boxes(f, raw=True)
        assert len(raw) == len(want_raw) and [int(t) for t in raw["type"]] == [int(t) for t in want_raw["type"]]
        boxes_close(net.boxes(f), want, px=BOX_TOL, score=SCORE_TOL)
    net.close()


def test_pipelined_submit_collect_equals_blocking_call(assets):
    """ffb_submit_u8 / ffb_collect (copy of batch i+1 overlapped with batch i) returns exactly the blocking call's boxes."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    batches = [np.ascontiguousarray(synth.shifted_frames_from(img, w, h, 12)[k:k + 6]) for k in (0, 3, 6)]
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=6)
    want = []
    for b in batches:
        net.detect_batch_u8(b, 6, 320, 320, 960)
        want.append([net.boxes(f).tobytes() for f in range(6)])
    got = []
    net.submit_u8(batches[0], 6, 320, 320, 960)
    for i in range(3):
        if i + 1 < 3:
            net.submit_u8(batches[i + 1], 6, 320, 320, 960)
        net.collect()
        got.append([net.boxes(f).tobytes() for f in range(6)])
    assert got == want and any(len(x) for x in got[0])
    with pytest.raises(fb.FfcnnError):
        net.collect()                                   # nothing in flight
    # three batches in flight (the pipeline's depth): all are queued on the GPU before the host blocks on the first; a fourth is refused
    for b in batches:
        net.submit_u8(b, 6, 320, 320, 960)
    with pytest.raises(fb.FfcnnError):
        net.submit_u8(batches[0], 6, 320, 320, 960)
    got3 = []
    for i in range(3):
        net.collect()
        got3.append([net.boxes(f).tobytes() for f in range(6)])
    assert got3 == want
    # steady state with two batches queued behind the collected one, batch sizes changing on the way
    sizes = [6, 4, 6, 2, 5, 6, 6]
    net.submit_u8(batches[0], sizes[0], 320, 320, 960); net.submit_u8(batches[1], sizes[1], 320, 320, 960)
    for i in range(len(sizes)):
        if i + 2 < len(sizes):
            net.submit_u8(batches[(i + 2) % 3], sizes[i + 2], 320, 320, 960)
        net.collect()
        assert [net.boxes(f).tobytes() for f in range(sizes[i])] == want[i % 3][:sizes[i]]
    net.close()


def test_resized_input_matches_host_net_input(assets):
    """GPU net_input with a real resize (640x424 bmp -> 320x212 corner) equals the host/reference arithmetic bit for bit."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=2)
    net.set_option("keep_all", 1)
    two = np.stack([img, img[::-1].copy()])
    net.input_u8(two, 2, w, h, img.shape[1])
    net.forward()
    for f in range(2):
        want, s1, s2 = orc.net_input(two[f], w, h, 320, 320)
        assert np.array_equal(net.layer_output(-1, f).view(np.uint32), want.view(np.uint32))
    assert (net.net.s1, net.net.s2) == (640, 320)
    net.close()


def test_full_batch_256_properties(assets, golden):
    """BASELINE size (batch 256): frames are independent, so every copy of a frame must give identical bits wherever
    it sits in the batch, and each distinct frame must still match its golden checksum; graph replay is idempotent."""
    cfg, wts, _ = assets
    B = 256
    base = synth.frames_u8(4)
    frames = np.concatenate([base] * (B // 4), axis=0)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=B)
    net.set_option("keep_all", 1)
    d = fb.DeviceBuffer(frames.nbytes).upload(frames)
    net.input_u8(d.ptr, B, 320, 320, 960, on_device=True)
    net.forward(); net.forward(); net.forward()
    net.detect()
    g = golden["synth_320"]
    first = {f: net.layer_output(129, f) for f in range(4)}
    for f in range(4):
        assert abs(float(first[f].astype(np.float64).sum()) - float(g[f"s1_f{f}_v6_O2_sum"][129])) <= 2e-5 * float(g[f"s1_f{f}_v6_O2_maxabs"][129]) * first[f].size
    for pos in (4, 127, 128, 252, 255):
        assert np.array_equal(net.layer_output(129, pos).view(np.uint32), first[pos % 4].view(np.uint32)), pos
        assert np.array_equal(net.layer_output(120, pos).view(np.uint32), net.layer_output(120, pos % 4).view(np.uint32)), pos
    assert net.launches_per_forward() > 0
    net.close(); d.free()


def test_liveness_arena_gives_same_heads_as_keep_all(assets):
    """The reused-buffer plan (ffcnn.c:511-517's free-when-unreferenced, done statically) must not change results:
    keep_all=2 runs the same (fused) kernels with one private buffer per tensor, so the boxes must be bit-identical."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    s2f = synth.shifted_frames_from(img, w, h, 8)
    res = []
    for keep in (2, 0):
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=8)
        net.set_option("keep_all", keep)
        net.detect_batch_u8(s2f, 8, 320, 320, 960)
        res.append([net.boxes(f).tobytes() for f in range(8)])
        if keep == 0:
            assert net.get_option("arena_mb") < 60
        net.close()
    assert res[0] == res[1]


def test_forked_tail_is_found_and_changes_nothing(assets):
    """fork_tail: the conv chain of the first yolo head (L116-L121) runs on a second stream beside the layers that follow it, inside
    the captured graph and in eager mode.  Same kernels, same operands: boxes and heads must be bit-identical with the fork off,
    over several replays of the graph (a buffer recycled too early between the two chains would show up as a difference)."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    s2f = synth.shifted_frames_from(img, w, h, 16)
    res = []
    for fork in (1, 0):
        for graph in (1, 0):
            net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=16)
            net.set_option("fork_tail", fork)
            net.set_option("graph", graph)
            out = []
            for rep in range(3):
                net.detect_batch_u8(s2f, 16, 320, 320, 960)
                out.append([net.boxes(f).tobytes() for f in range(16)] + [net.layer_output(120, f).tobytes() for f in (0, 15)] + [net.layer_output(129, f).tobytes() for f in (0, 15)])
            assert out[0] == out[1] == out[2]
            assert (net.get_option("side_branch") == 116 * 1000 + 121) == bool(fork)
            res.append(out[0])
            net.close()
    assert res[0] == res[1] == res[2] == res[3]


def test_stem_block_fusion_is_bit_identical(assets):
    """fuse_stem: the stem (net_input + 3x3 s2 conv on the u8 frames) and the first 8->8->4 block run as one kernel (stem_block.cuh)
    whenever the frames are read by the stem directly.  It performs the two kernels' arithmetic in their order, so the block's output
    (layer 3), the heads and the boxes must be bit-identical with the fusion off -- on seeded frames, on picture frames, on a batch
    that is not a multiple of anything, and on a net whose width is not a multiple of the 32-pixel tile."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    for (nw, nh, n) in ((320, 320, 11), (416, 256, 3)):
        fr = synth.frames_u8(n, nw, nh)
        pitch = fr.shape[-1] if fr.ndim == 3 else (3 * nw + 3) & ~3
        res = []
        for fuse in (1, 0):
            net = fb.Net(cfg, wts, nw, nh, device=0, max_batch=n)
            net.set_option("fuse_stem", fuse)
            net.set_option("keep_all", 2)
            net.detect_batch_u8(fr, n, nw, nh, pitch)
            assert net.get_option("stem_block") == fuse
            res.append([net.layer_output(3, f).tobytes() for f in range(n)] + [net.layer_output(129, f).tobytes() for f in (0, n - 1)] + [net.boxes(f).tobytes() for f in range(n)])
            net.close()
        assert res[0] == res[1]
    s2f = synth.shifted_frames_from(img, w, h, 8)
    out = []
    for fuse in (1, 0):
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=8)
        net.set_option("fuse_stem", fuse)
        net.detect_batch_u8(s2f, 8, 320, 320, 960)
        out.append([net.boxes(f).tobytes() for f in range(8)])
        net.close()
    assert out[0] == out[1] and any(len(b) for b in out[0])


def test_tiled_stride2_block_is_bit_identical_to_the_sliced_kernel(assets):
    """block_s2.cuh: the 4->24->8 stride-2 block (L9-L11) as one shared-memory-tiled kernel performs the arithmetic of the three
    8-channel slices of k_block_reg_s2 in their order; the block's output (layer 11), the heads and the boxes must be bit-identical
    (child process: the choice is read from FFCNN_S2_TILE when the plan is made)."""
    import subprocess
    code = r"""
import sys, hashlib
sys.path.insert(0, %r)
import numpy as np
import ffcnn_b200 as fb
from ffcnn_b200 import synth
cfg, wts = fb.default_model()
h = hashlib.sha256()
for (nw, nh, n) in ((320, 320, 7), (416, 256, 3)):
    fr = synth.frames_u8(n, nw, nh)
    net = fb.Net(cfg, wts, nw, nh, device=0, max_batch=n)
    net.set_option("keep_all", 2)
    net.detect_batch_u8(fr, n, nw, nh, fr.shape[-1])
    for f in range(n):
        h.update(net.layer_output(11, f).tobytes()); h.update(net.layer_output(129, f).tobytes()); h.update(net.boxes(f).tobytes())
    net.close()
print("DIGEST", h.hexdigest())
""" % REPO
    digests = []
    for tile in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, FFCNN_S2_TILE=tile, FFCNN_BLK_VERBOSE="1"), timeout=280)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert ("shared-memory tiles" in r.stderr) == (tile == "1")
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]


def test_row_ring_register_blocks_are_bit_identical_to_register_prefetch(assets):
    """block_reg.cuh "row ring": the register-resident block kernels fetch x rows through a per-lane cp.async ring in shared memory
    (default) instead of register loads two rows ahead (FFCNN_REG_RING=0).  Only the way x arrives changes, so every block output
    (layers 3, 8, 11 -- the stem fusion is switched off so that the 8->8->4 block runs on the register kernel too), the heads and the
    boxes must be bit-identical; frames whose strips end on ragged rows / columns are included (child process: the choice is read
    from FFCNN_REG_RING when the plan is made)."""
    import subprocess
    code = r"""
import sys, hashlib
sys.path.insert(0, %r)
import numpy as np
import ffcnn_b200 as fb
from ffcnn_b200 import synth
cfg, wts = fb.default_model()
h = hashlib.sha256()
for (nw, nh, n) in ((320, 320, 7), (416, 256, 3), (352, 288, 2)):
    fr = synth.frames_u8(n, nw, nh)
    net = fb.Net(cfg, wts, nw, nh, device=0, max_batch=n)
    net.set_option("keep_all", 2)
    net.set_option("fuse_stem", 0)
    net.detect_batch_u8(fr, n, nw, nh, fr.shape[-1])
    for f in range(n):
        for l in (3, 8, 11, 120, 129): h.update(net.layer_output(l, f).tobytes())
        h.update(net.boxes(f).tobytes())
    net.close()
print("DIGEST", h.hexdigest())
""" % REPO
    digests = []
    for ring in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, FFCNN_REG_RING=ring, FFCNN_BLK_VERBOSE="1"), timeout=280)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert ("cp.async row ring" in r.stderr) == (ring == "1"), r.stderr[-1500:]
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]


def test_threaded_host_decode_returns_the_same_boxes(assets):
    """ffb_detect_finish splits the exact decode + NMS of a large candidate set (thousands of candidates: picture-derived frames) at
    frame boundaries over a few host threads.  The boxes, raw boxes and their order must be those of the single-threaded decode
    (FFCNN_DECODE_THREADS=1), on a batch that is not a multiple of the thread count and through the pipelined calls too."""
    import subprocess
    code = r"""
import sys, hashlib
sys.path.insert(0, %r)
import numpy as np
import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import ref
cfg, wts = fb.default_model()
img, w, h = ref.load_bmp(fb.ASSETS + "/test.bmp")
N = 203
fr = np.ascontiguousarray(synth.shifted_frames_from(img, w, h, N))
hsh = hashlib.sha256(); total = 0
net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=N)
net.detect_batch_u8(fr, N, 320, 320, 960)
total = (net.last_d2h_bytes() - 4) // 36                     # candidates the GPU filter passed to the host
for f in range(N):
    hsh.update(net.boxes(f).tobytes()); hsh.update(net.boxes(f, raw=True).tobytes())
net.submit_u8(fr, N, 320, 320, 960); net.submit_u8(fr[::-1].copy(), N, 320, 320, 960)
for _ in range(2):
    net.collect()
    for f in range(N): hsh.update(net.boxes(f).tobytes())
net.close()
print("DIGEST", hsh.hexdigest(), total)
""" % REPO
    out = []
    for t in ("1", "4", "3"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, FFCNN_DECODE_THREADS=t), timeout=280)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        out.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert out[0] == out[1] == out[2]
    assert int(out[0].split()[2]) >= 2048                    # enough candidates for the threaded path to have run


def test_dw5_exact_mode_matches_conv_v0(assets, oracle_layers):
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=1)
    net.set_option("keep_all", 1)
    net.set_option("dw5_exact", 1)
    net.net_input(img, w, h)
    x = net.input_tensor().copy()
    net.net_forward()
    outs, _, _ = orc.forward(oracle_layers, x, 640, 320, v6_quirk=False)
    for i in (116, 118, 125, 127, 129):
        assert rel_err(net.layer_output(i, 0), outs[i]) < FEAT_TOL, i
    quirk, _, _ = orc.forward(oracle_layers, x, 640, 320, v6_quirk=True)
    assert float(np.abs(net.layer_output(116, 0) - quirk[116]).max() / np.abs(quirk[116]).max()) > 1e-3     # and it really differs from the v6 default
    net.close()


@pytest.mark.parametrize("pw_mode", PW_MODES)
def test_microbench_shapes_against_oracle(pw_mode):
    """BASELINE configs 3 and 4 at a batch the oracle finishes in seconds: dw3x3 160x160x96 and 1x1 40x40 192->192."""
    rng = np.random.default_rng(7)
    for (ih, iw, ic, grp, pad, st, fs, fn, act, n) in ((160, 160, 96, 96, 1, 1, 3, 96, 2, 2), (40, 40, 192, 1, 0, 1, 1, 192, 2, 3),
                                                       (20, 20, 120, 1, 0, 1, 1, 255, 0, 2), (40, 40, 96, 96, 1, 2, 3, 96, 2, 2)):
        k = fs * fs * (ic // grp); row = ((k + 3) & ~3) + 4
        f = np.zeros((fn, row), np.float32)
        f[:, :k] = rng.standard_normal((fn, k)) / np.sqrt(k)
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
        x = rng.standard_normal((n, ic, ih, iw)).astype(np.float32)
        op = fb.ConvOp(f, ic, grp, pad, st, fs, fn, act, pw_mode=pw_mode)
        y = op(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
        for b in range(n):
            want = orc.conv_raw(x[b], f, iw, ih, ic, grp, pad, st, fs, fn, act, v6_quirk=True)
            assert rel_err(y[b].transpose(2, 0, 1), want) < FEAT_TOL, (op.kernel, b)
        op.close()


def test_ragged_and_odd_shapes_through_generic_kernel():
    """Shapes no specialised kernel takes (channels not a multiple of 4, grouped with >1 channel per group, even kernels)."""
    rng = np.random.default_rng(11)
    for (iw, ih, ic, grp, pad, st, fs, fn, act) in ((9, 7, 6, 2, 1, 1, 3, 4, 2), (5, 5, 3, 1, 0, 1, 1, 5, 1), (8, 6, 4, 1, 0, 2, 2, 5, 0), (1, 1, 8, 1, 0, 1, 1, 8, 2)):
        k = fs * fs * (ic // grp); row = ((k + 3) & ~3) + 4
        f = np.zeros((fn, row), np.float32); f[:, :k] = rng.standard_normal((fn, k)); f[:, row - 4] = 0.75; f[:, row - 3] = 0.1
        x = rng.standard_normal((ic, ih, iw)).astype(np.float32)
        got = fb.groupconv(x, f, iw, ih, ic, grp, pad, st, fs, fn, act)
        assert rel_err(got, orc.conv_raw(x, f, iw, ih, ic, grp, pad, st, fs, fn, act, False)) < FEAT_TOL


def test_dw5_quirk_on_the_smallest_maps():
    """conv-v6's dropped kernel row on output row oh-2 (conv-v6.c:422-441) holds down to 4x4 maps (a 128-pixel net side);
    below that the reference reads out of bounds and the GPU path computes the exact convolution."""
    rng = np.random.default_rng(12)
    for (iw, ih, ic, quirk) in ((4, 4, 8, True), (4, 9, 4, True), (9, 4, 12, True), (5, 5, 4, True), (3, 6, 4, False), (6, 3, 4, False)):
        f = np.zeros((ic, 32), np.float32); f[:, :25] = rng.standard_normal((ic, 25)); f[:, 28] = rng.uniform(0.5, 1.5, ic); f[:, 29] = 0.1
        x = rng.standard_normal((ic, ih, iw)).astype(np.float32)
        got = fb.groupconv(x, f, iw, ih, ic, ic, 2, 1, 5, ic, 2)
        want = orc.conv_raw(x, f, iw, ih, ic, ic, 2, 1, 5, ic, 2, quirk)
        assert rel_err(got, want) < FEAT_TOL, (iw, ih, rel_err(got, want))
        if quirk:
            exact = orc.conv_raw(x, f, iw, ih, ic, ic, 2, 1, 5, ic, 2, False)
            assert float(np.abs(got - exact).max() / np.abs(exact).max()) > 1e-3


@pytest.mark.gpu
def test_warp_specialised_block_kernel_opt_in():
    """k_block_ws (block_ws.cuh: fixed stage-A / stage-B warp roles, opt-in because it measured slower) must stay correct: the fused-block
    parity test in a child process with FFCNN_BLK_WS=1 -- the planner reads the variable once per process -- and the kernel really selected."""
    import subprocess
    env = dict(os.environ, FFCNN_BLK_WS="1", FFCNN_BLK_VERBOSE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-s", "-k", "test_fused_blocks_against_oracle"],
                       capture_output=True, text=True, env=env, timeout=280, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "2 passed" in r.stdout and "warp-specialised" in (r.stdout + r.stderr)


@pytest.mark.parametrize("fuse_block", [1, 2])
def test_fused_blocks_against_oracle(assets, oracle_layers, fuse_block):
    """Block fusion (1x1 expand -> 3x3 depthwise -> 1x1 project [+ shortcut] as one kernel, block_mma.cu): every tensor the
    fused plan still materialises (block outputs, tail layers, heads) against the oracle, boxes against the oracle's, and
    the fused plan must give the same boxes on a batch as the layer-by-layer plan.  fuse_block 1 = default policy,
    2 = every supported block (all 24 chains of the graph, both strides, every channel configuration)."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=8)
    net.set_option("fuse_block", fuse_block)
    net.set_option("keep_all", 2)                       # fused plan, but no buffer reuse: surviving tensors stay readable
    nblocks = net.get_option("blocks")
    assert nblocks == 24 if fuse_block == 2 else 0 < nblocks <= 24      # the default policy fuses every chain since round 2s
    net.net_input(img, w, h)
    x = net.input_tensor().copy()
    got = net.net_forward()
    outs, raw, fin = orc.forward(oracle_layers, x, net.net.s1, net.net.s2, v6_quirk=True)
    checked = 0
    for i, o in enumerate(outs):
        _, _, kn = net.layer_cost(i)
        if o is None or kn in ("in_block", "in_spp"):
            continue
        # a projection conv followed by dropout + shortcut shares its buffer with the shortcut: only the shortcut's values live there
        nxt = [oracle_layers[j].type for j in range(i + 1, min(i + 3, len(oracle_layers)))]
        if orc.SHORTCUT in nxt and oracle_layers[i].type in (orc.CONV, orc.DROPOUT):
            continue
        a = net.layer_output(i, 0)
        if a is None:
            continue
        assert rel_err(a, o) < FEAT_TOL, (i, kn, rel_err(a, o))
        checked += 1
    assert checked >= 15
    graw = net.boxes(0, raw=True)
    assert len(graw) == len(raw) and [int(t) for t in graw["type"]] == [int(t) for t in raw["type"]]
    boxes_close(got, fin, px=BOX_TOL if fuse_block == 1 else BOX_TOL_TC, score=SCORE_TOL)
    s2f = synth.shifted_frames_from(img, w, h, 8)
    net.set_option("keep_all", 0)
    net.detect_batch_u8(s2f, 8, 320, 320, 960)
    fused = [net.boxes(f) for f in range(8)]
    net.set_option("fuse_block", 0)
    assert net.get_option("blocks") == 0
    net.detect_batch_u8(s2f, 8, 320, 320, 960)
    for f in range(8):
        boxes_close(fused[f], net.boxes(f), px=BOX_TOL + BOX_TOL_TC, score=2 * SCORE_TOL)       # two GPU plans, each within its tolerance of the oracle
    net.close()


def test_fused_plan_odd_batches_and_geometries(assets):
    """Ragged cases for the fused plan: batch sizes that do not fill the tile grid, non-square nets (640x448 = the reference
    main()'s geometry, 416x256) and a net whose deepest maps are 11x11 (odd width: those chains must fall back to the
    per-layer kernels).  The fused plan must give the layer-by-layer plan's boxes (both are within tolerance of the oracle)."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    pix = img[:, :w * 3].reshape(h, w, 3)
    for (nw, nh, n) in ((0, 0, 1), (0, 0, 7), (640, 448, 2), (416, 256, 3), (352, 352, 2)):
        W, H = nw or 320, nh or 320
        base = pix[np.arange(H) * h // H][:, np.arange(W) * w // W]
        pitch = (W * 3 + 3) & ~3
        frames = np.zeros((n, H, pitch), np.uint8)
        for f in range(n):
            frames[f, :, :W * 3] = np.roll(base, (f * 3, f * 5), (0, 1)).reshape(H, W * 3)
        res = []
        for fuse in (1, 0):
            net = fb.Net(cfg, wts, nw, nh, device=0, max_batch=n)
            net.set_option("fuse_block", fuse); net.set_option("fuse_tail", fuse)
            for _ in range(2):                                              # second pass replays the CUDA graph
                net.detect_batch_u8(frames, n, W, H, pitch)
            res.append([net.boxes(f) for f in range(n)])
            assert (net.get_option("blocks") > 0) == bool(fuse)
            net.close()
        assert sum(len(b) for b in res[0]) > 0
        for a, b in zip(*res):
            boxes_close(a, b, px=(BOX_TOL + BOX_TOL_TC) * max(1, W // 320), score=2 * SCORE_TOL)    # plan vs plan (each within its tolerance of the oracle, scaled with the net width)


def test_second_darknet_graph_against_oracle(tmp_path):
    """Widening: the yolov3-tiny-like graph of ffcnn_b200/tinygraph.py through the same loader and engine --
    dense 3x3 and grouped convs (generic kernel), stride-2 max pools, avgpool, relu, upsample into a route, a tcgen05
    pointwise layer, 2-class yolo heads -- every layer and the boxes of three frames against the oracle (which is bit-exact
    against the compiled reference on this graph)."""
    from ffcnn_b200 import tinygraph as tg
    cfg, wts = tg.write(str(tmp_path))
    layers = orc.load_net(cfg, wts, 0, 0)
    fr = tg.frames(3)
    for keep in (1, 0):
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=3)
        net.set_option("keep_all", keep)
        net.detect_batch_u8(fr, 3, tg.W, tg.H, fr.shape[2])
        for f in range(3):
            x, s1, s2 = orc.net_input(fr[f], tg.W, tg.H, tg.W, tg.H)
            outs, raw, fin = orc.forward(layers, x, s1, s2, True)
            if keep:
                for i, o in enumerate(outs):
                    if o is not None:
                        assert rel_err(net.layer_output(i, f), o) < FEAT_TOL, (f, i, rel_err(net.layer_output(i, f), o))
            graw = net.boxes(f, raw=True)
            assert len(graw) == len(raw) and [int(t) for t in graw["type"]] == [int(t) for t in raw["type"]]
            # random weights give boxes hundreds to thousands of pixels wide: w = exp(tw) * anchor turns a logit error d into
            # w * d pixels, so the pixel tolerance grows with the box (4e-5 absolute on the logit = the feature-map tolerance
            # 2e-5 * max|head| at |head| = 2; measured 2.0e-5 * size on a 2180 px wide box)
            got = net.boxes(f)
            assert len(got) == len(fin)
            for g, e in zip(got, fin):
                # scores: 0.25 * logit error at the steepest point of the sigmoid (random weights put scores mid-range)
                assert int(g["type"]) == int(e["type"]) and abs(float(g["score"]) - float(e["score"])) <= 4 * SCORE_TOL
                tol = BOX_TOL + 4e-5 * max(float(e["x2"]) - float(e["x1"]), float(e["y2"]) - float(e["y1"]))
                assert max(abs(float(g[k]) - float(e[k])) for k in ("x1", "y1", "x2", "y2")) <= tol, (g, e, tol)
        net.close()


def test_cli_prints_the_reference_lines_and_draws_boxes(assets, golden, tmp_path):
    """tools/ffcnn_cli.c on test.bmp: the stdout of the reference's `./ffcnn 2 test.bmp cfg weights` (ffcnn.c:552-593) --
    banner, net_dump table, timing, net_profile, one line per box -- and out.bmp with the boxes drawn; the batched mode
    gives the same lines per frame."""
    import subprocess
    cfg, wts, bmp = assets
    fb.build()
    exe = os.path.join(REPO, "ffcnn_b200", "ffcnn_cli")
    want = ["score: %.2f, category: %2d, rect: (%3d %3d %3d %3d)" % (b["score"], b["type"], int(b["x1"]), int(b["y1"]), int(b["x2"]), int(b["y2"]))
            for b in golden["testbmp_640x448"]["v6_O2_final"]]
    r = subprocess.run([exe, "2", bmp, cfg, wts], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "file_bmp    : " + bmp and "2 times inference:" in r.stdout
    assert lines[-len(want):] == want
    import re
    kinds = "conv|avgpool|maxpool|upsample|dropout|shortcut|route|yolo"
    assert sum(bool(re.match(r"\s*\d+\s+(%s)\b" % kinds, l)) for l in lines) == 131                # net_dump: one row per layer
    assert sum(bool(re.match(r"\s*(%s):\s+\d+ ms$" % kinds, l)) for l in lines) == 8                # net_profile: one row per layer type
    src, w, h = ref.load_bmp(bmp)
    out, w2, h2 = ref.load_bmp(str(tmp_path / "out.bmp"))
    assert (w2, h2) == (w, h)
    b = golden["testbmp_640x448"]["v6_O2_final"][0]
    x1, y1, y2 = int(b["x1"]), int(b["y1"]), int(b["y2"])
    ym = min(max((y1 + y2) // 2, 0), h - 1)
    assert tuple(out[ym, 3 * x1:3 * x1 + 3]) == (0, 255, 0)                          # B,G,R of the left edge
    assert (out != src).any() and (out != src).mean() < 0.05
    r = subprocess.run([exe, "--batch", cfg, wts, bmp, bmp, bmp, "--out", "det"], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    for f in range(3):
        at = lines.index("frame %d: %s" % (f, bmp))
        assert lines[at + 1:at + 1 + len(want)] == want
        assert np.array_equal(ref.load_bmp(str(tmp_path / ("det%d.bmp" % f)))[0], out)


# ------------------------------------------------------------------------------------------------------------------
# round 2: the configuration bench.py times (default fused plan, batch 256), every rank's weight path, overflow


def _batch256_mixed(img, w, h):
    """256 frames with known frames at the positions a persistent tile scheduler treats differently (first, last, the
    wave boundary around 127/128): set S2 (picture-derived, golden boxes) and set S1 (seeded random, golden heads)."""
    s2 = synth.shifted_frames_from(img, w, h, 20).reshape(20, 320, 960)
    s1 = synth.frames_u8(4)
    frames = np.empty((256, 320, 960), np.uint8)
    for p in range(256):
        frames[p] = s2[p % 16] if p % 3 else s1[p % 4]
    known = {0: ("s2", 0), 1: ("s2", 3), 127: ("s2", 7), 128: ("s2", 19), 254: ("s1", 0), 255: ("s1", 1), 64: ("s1", 2), 200: ("s2", 3)}
    for p, (kind, f) in known.items():
        frames[p] = s2[f] if kind == "s2" else s1[f]
    return frames, known


@pytest.mark.parametrize("keep_all", [0, 2])
def test_batch_256_default_fused_plan_against_oracle_and_goldens(assets, golden, oracle_layers, keep_all):
    """The benchmarked configuration itself: default options (fused blocks, fused tail, buffer reuse, CUDA-graph replay) at
    batch 256.  Heads L120 / L129 of frames {0, 1, 64, 127, 128, 200, 254, 255} against the oracle, boxes of the picture
    frames against the reference goldens (tests/golden/synth_320.npz), candidate counts of the random frames."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    frames, known = _batch256_mixed(img, w, h)
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=256)
    net.set_option("keep_all", keep_all)
    assert net.get_option("fuse_block") == 1 and net.get_option("fuse_tail") == 1 and net.get_option("graph") == 1
    d = fb.DeviceBuffer(frames.nbytes).upload(frames)
    for _ in range(3):                                                  # eager pass, capture, replay
        net.input_u8(d.ptr, 256, 320, 320, 960, on_device=True)
        net.forward()
    net.detect()
    assert net.get_option("blocks") > 0 and net.get_option("input_fused") == 1
    g = golden["synth_320"]
    cache = {}
    for p, (kind, f) in sorted(known.items()):
        if (kind, f) not in cache:
            x, s1, s2 = orc.net_input(frames[p], 320, 320, 320, 320)
            cache[(kind, f)] = orc.forward(oracle_layers, x, s1, s2, v6_quirk=True)
        outs, raw, fin = cache[(kind, f)]
        for hid in (120, 129):
            e = rel_err(net.layer_output(hid, p), outs[hid])
            assert e < FEAT_TOL, (p, hid, e)
        graw = net.boxes(p, raw=True)
        if kind == "s2":
            want_raw, want = g[f"s2_f{f}_raw"], g[f"s2_f{f}_final"]
            assert len(graw) == len(want_raw) and [int(t) for t in graw["type"]] == [int(t) for t in want_raw["type"]], p
            boxes_close(net.boxes(p), want, px=BOX_TOL, score=SCORE_TOL)
        else:
            assert len(graw) == len(g[f"s1_f{f}_v6_O2_raw"]) == len(raw), p
    # every copy of a frame gives identical bits wherever it sits in the batch
    ref_pos = {}
    for p in range(0, 256, 7):
        key = frames[p].tobytes()
        if key in ref_pos:
            assert np.array_equal(net.layer_output(129, p).view(np.uint32), net.layer_output(129, ref_pos[key]).view(np.uint32)), p
        else:
            ref_pos[key] = p
    net.close(); d.free()


def test_commit_weights_path_of_nonzero_ranks(assets):
    """What every rank > 0 does in the multi-GPU frontend: a net parsed WITHOUT a weights file (zero weights), its device
    copy of the packed buffer filled from outside (stands in for the NCCL broadcast), ffb_commit_weights -- the heads must
    be bit-equal to a normally loaded net's, in the default fused plan and in the layer-by-layer plan."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    fr = np.ascontiguousarray(synth.shifted_frames_from(img, w, h, 6).reshape(6, 320, 960))
    a = fb.Net(cfg, wts, 0, 0, device=0, max_batch=6)
    b = fb.Net(cfg, None, 0, 0, device=0, max_batch=6)
    assert not b.packed_weights().any()
    b.detect_batch_u8(fr, 6, 320, 320, 960)                             # zero weights: runs, finds nothing or garbage -- and must not stick
    ptr, n = b.packed_weights_device()
    assert n == a.net.weight_size
    packed = a.packed_weights()
    fb._check(fb.lib().ffb_copy_h2d(ptr, packed.ctypes.data, packed.nbytes), "ffb_copy_h2d")
    b.commit_weights()
    assert np.array_equal(b.packed_weights().view(np.uint32), packed.view(np.uint32))     # host copy follows the device copy
    for keep in (0, 1):
        res = []
        for net in (a, b):
            net.set_option("keep_all", keep)
            net.detect_batch_u8(fr, 6, 320, 320, 960)
            res.append(([net.layer_output(hid, f).tobytes() for hid in (120, 129) for f in (0, 5)], [net.boxes(f).tobytes() for f in range(6)]))
        assert res[0] == res[1], keep
        assert any(len(x) for x in res[0][1])
    a.close(); b.close()


def test_candidate_overflow_is_reported_not_truncated(assets):
    """ffcnn.c:438-474 keeps every candidate above the threshold (up to bbox_max per frame).  When the device candidate
    buffer is too small the library must say so (FFB_E_OVERFLOW + ffb_last_error) and ffb_detect must retry with a grown
    buffer -- never return a silently truncated box set."""
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    fr = np.ascontiguousarray(synth.shifted_frames_from(img, w, h, 8).reshape(8, 320, 960))
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=8)
    net.detect_batch_u8(fr, 8, 320, 320, 960)
    want = [net.boxes(f).tobytes() for f in range(8)]
    assert sum(len(net.boxes(f, raw=True)) for f in range(8)) > 16
    net.set_option("cand_cap", 5)
    net.input_u8(fr, 8, 320, 320, 960); net.forward()
    assert net.detect_enqueue() == 2
    rc = fb.lib().ffb_detect_finish(net.p)
    assert rc == -2 and b"overflow" in fb.lib().ffb_last_error()
    assert all(len(net.boxes(f)) == 0 for f in range(8))                # nothing half-decoded is left behind
    net.set_option("cand_cap", 5)
    net.detect_batch_u8(fr, 8, 320, 320, 960)                           # the blocking call retries with the grown buffer
    assert [net.boxes(f).tobytes() for f in range(8)] == want
    net.close()


def test_large_pointwise_layers_do_not_break_the_forward():
    """1x1 convs whose weights fit neither pw_tc's resident plan nor the FFMA kernel's shared memory
    (yolov3's 1024 -> 256, 512 -> 1024 ...) must still run -- through the implicit-GEMM tcgen05 kernel or the generic one."""
    rng = np.random.default_rng(21)
    for (ic, fn, hw, n) in ((1024, 256, 6, 2), (512, 1024, 5, 1), (256, 2304, 4, 1)):
        row = ic + 4
        f = np.zeros((fn, row), np.float32)
        f[:, :ic] = rng.standard_normal((fn, ic)) / np.sqrt(ic)
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
        x = rng.standard_normal((n, ic, hw, hw)).astype(np.float32)
        op = fb.ConvOp(f, ic, 1, 0, 1, 1, fn, 2)
        y = op(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
        for b in range(n):
            want = orc.conv_raw(x[b], f, hw, hw, ic, 1, 0, 1, 1, fn, 2, v6_quirk=True)
            assert rel_err(y[b].transpose(2, 0, 1), want) < FEAT_TOL, (op.kernel, ic, fn)
        op.close()


def test_dense_convs_on_the_implicit_gemm_tcgen05_kernel():
    """dense k x k convs (the im2row / im2col + GEMM paths of conv-v6.c:9-42 and conv-v2.c:7-87) as an implicit GEMM
    on tcgen05 with TMA-fetched taps (conv_tc.cu): strides 1 and 2 (TMA element strides), 1x1 / 3x3 / 5x5, channel counts that are
    not multiples of 32 or 4, ragged maps smaller than a tile, several frames per tile, 255 filters, batch > 1."""
    rng = np.random.default_rng(31)
    cases = [  # iw, ih, ic, pad, stride, fs, fn, act, n
        (26, 26, 64, 1, 1, 3, 128, 2, 2), (27, 19, 32, 1, 2, 3, 64, 2, 3), (13, 13, 128, 1, 1, 3, 255, 0, 2), (40, 24, 16, 2, 1, 5, 24, 1, 1),
        (33, 17, 3, 1, 2, 3, 16, 2, 2), (9, 9, 40, 0, 1, 3, 48, 2, 5), (12, 12, 20, 1, 1, 3, 136, 0, 1), (64, 48, 8, 1, 2, 3, 32, 2, 1),
    ]
    for (iw, ih, ic, pad, st, fs, fn, act, n) in cases:
        k = fs * fs * ic; row = ((k + 3) & ~3) + 4
        f = np.zeros((fn, row), np.float32)
        f[:, :k] = rng.standard_normal((fn, k)) / np.sqrt(k)
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
        x = rng.standard_normal((n, ic, ih, iw)).astype(np.float32)
        op = fb.ConvOp(f, ic, 1, pad, st, fs, fn, act)
        assert op.kernel == ("igemm_tcgen05_3xtf32" if k >= 32 else "conv_generic"), (op.kernel, iw, ih, ic, fs, fn)    # tiny contractions stay on the generic kernel
        y = op(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
        for b in range(n):
            want = orc.conv_raw(x[b], f, iw, ih, ic, 1, pad, st, fs, fn, act, v6_quirk=True)
            e = rel_err(y[b].transpose(2, 0, 1), want)
            assert e < FEAT_TOL, (iw, ih, ic, pad, st, fs, fn, b, e)
        op.close()
