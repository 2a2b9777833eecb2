"""GPU (B200): the UNMODIFIED reference runtime -- /root/reference/ffcnn.c + bmpfile.c, compiled where they lie by
oracle/Makefile -- linked against libffcnn_b200.so's `groupconv` in place of conv-vN.c (build.sh:48 with the operator
swapped).  This is the reference's own plugin seam end to end: its net_load / net_input / net_forward
loop, malloc/free per layer, maxpool / upsample / shortcut / route / yolo decode / nms all run as shipped; only the 84
groupconv calls per frame (ffcnn.c:374-379) execute on the GPU, host CHW in, host CHW out.

The binaries are built in the build container (the reference sources are not shipped) and travel in oracle/_ref/."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc, ref
from conftest import boxes_close, note_feat, REPO

pytestmark = pytest.mark.gpu

FEAT_TOL = 2e-5
BOX_TOL = 1e-4
SCORE_TOL = 5e-6
EXE = os.path.join(REPO, "oracle", "_ref", "ffcnn_ref_gpuconv")


def test_reference_main_with_gpu_groupconv_prints_the_reference_boxes(assets, golden, tmp_path):
    """`./ffcnn 2 test.bmp cfg weights` (ffcnn.c:552-593) with our operator: the three box lines the CPU reference prints."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ffcnn_ref_gpuconv not built (needs /root/reference at build time)")
    cfg, wts, bmp = assets
    r = subprocess.run([EXE, "2", bmp, cfg, wts], capture_output=True, text=True, cwd=tmp_path, timeout=240)
    assert r.returncode == 0, r.stderr[-500:]
    want = ["score: %.2f, category: %2d, rect: (%3d %3d %3d %3d)" % (b["score"], b["type"], int(b["x1"]), int(b["y1"]), int(b["x2"]), int(b["y2"]))
            for b in golden["testbmp_640x448"]["v6_O2_final"]]
    lines = r.stdout.splitlines()
    assert lines[-len(want):] == want, lines[-6:]
    assert os.path.exists(tmp_path / "out.bmp")


def test_reference_layer_loop_with_gpu_groupconv_every_layer(assets, golden, oracle_layers):
    """Same link, driven through the harness: every layer output of the reference's loop (its own CPU pool / route /
    shortcut code fed by GPU conv results) against the oracle, raw candidates and final boxes against the -O2 goldens."""
    if not ref.available("gpuconv"):
        pytest.skip("oracle/_ref/libffcnn_ref_gpuconv.so not built")
    cfg, wts, bmp = assets
    img, w, h = ref.load_bmp(bmp)
    rn = ref.RefNet(cfg, wts, 0, 0, "gpuconv")
    rn.input_bgr(img, w, h)
    x = rn.input_tensor().copy()
    outs, raw, fin = rn.forward_dump()
    hd = rn.head()
    want_outs, want_raw, want_fin = orc.forward(oracle_layers, x, hd.s1, hd.s2, v6_quirk=True)
    for i, (a, b) in enumerate(zip(outs, want_outs)):
        if a is None or b is None:
            continue
        e = float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))
        note_feat(e)
        assert e < FEAT_TOL, (i, e)
    g = golden["testbmp_320"]
    assert [int(t) for t in raw["type"]] == [int(t) for t in g["v6_O2_raw"]["type"]]
    boxes_close(fin, g["v6_O2_final"], px=BOX_TOL, score=SCORE_TOL)
    boxes_close(fin, want_fin, px=BOX_TOL, score=SCORE_TOL)
    rn.close()
