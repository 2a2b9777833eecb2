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


def boxes_close(got, want, px=1e-4, score=1e-6):
    """Same count, same classes, coordinates within px pixels, scores within `score`."""
    assert len(got) == len(want), (len(got), len(want))
    for g, e in zip(got, want):
        assert int(g["type"]) == int(e["type"])
        assert abs(float(g["score"]) - float(e["score"])) <= score, (g, e)
        for k in ("x1", "y1", "x2", "y2"):
            assert abs(float(g[k]) - float(e[k])) <= px, (k, g, e)
