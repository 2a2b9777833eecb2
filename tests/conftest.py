import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def assets():
    import ffcnn_b200 as fb
    cfg, wts = fb.default_model()
    bmp = os.path.join(fb.ASSETS, "test.bmp")
    for p in (cfg, wts, bmp):
        if not os.path.exists(p):
            pytest.skip(f"model asset missing: {p} (run `make -C oracle` in the build container)")
    return cfg, wts, bmp


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in
            ("testbmp_320", "testbmp_640x448", "synth_320", "groupconv_cases")}


@pytest.fixture(scope="session")
def oracle_layers(assets):
    from oracle import oracle as orc
    cfg, wts, _ = assets
    return orc.load_net(cfg, wts, 0, 0)


MEASURED = {"box_px": 0.0, "box_px_large_nets": 0.0, "score": 0.0, "boxes": 0, "feat_rel": 0.0}      # worst deviations seen by this session's comparisons


def boxes_close(got, want, px=1e-4, score=1e-6):
    """Same count, same classes, coordinates within px pixels, scores within `score`.  The worst deviation seen is
    recorded (printed in the terminal summary and written to gpurun_out/parity_measured.json on the GPU box)."""
    assert len(got) == len(want), (len(got), len(want))
    for g, e in zip(got, want):
        assert int(g["type"]) == int(e["type"])
        ds = abs(float(g["score"]) - float(e["score"]))
        dp = max(abs(float(g[k]) - float(e[k])) for k in ("x1", "y1", "x2", "y2"))
        if px <= 1.0001e-4:                              # comparisons against the named (-O2) oracle at the 320x320 geometry
            MEASURED["box_px"] = max(MEASURED["box_px"], dp); MEASURED["score"] = max(MEASURED["score"], ds); MEASURED["boxes"] += 1
        elif px <= 3.0001e-4:
            MEASURED["box_px_large_nets"] = max(MEASURED["box_px_large_nets"], dp)
        assert ds <= score, (g, e)
        assert dp <= px, (dp, g, e)


def note_feat(err):
    MEASURED["feat_rel"] = max(MEASURED["feat_rel"], float(err))


def pytest_terminal_summary(terminalreporter):
    if MEASURED["boxes"] or MEASURED["feat_rel"]:
        terminalreporter.write_line("parity measured this session: max box error %.3g px over %d boxes at 320x320 (%.3g px on larger nets / plan-vs-plan), "
                                    "max score error %.3g, max feature-map error %.3g of the layer max"
                                    % (MEASURED["box_px"], MEASURED["boxes"], MEASURED["box_px_large_nets"], MEASURED["score"], MEASURED["feat_rel"]))
        out = os.path.join(REPO, "gpurun_out")
        if os.path.isdir(out):
            import json
            with open(os.path.join(out, "parity_measured.json"), "w") as f:
                json.dump(MEASURED, f)
