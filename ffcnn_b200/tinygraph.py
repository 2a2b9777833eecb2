"""A second darknet graph for widening tests ("other darknet graphs through the same loader").

yolov3-tiny-like, small enough for the CPU oracle: dense 3x3 convs, stride-2 max pools, a clamped-window avgpool,
relu / leaky / linear activations, a grouped conv (4 groups x 8 channels), 1x1 convs, upsample, routes written both
relative and absolute, two yolo heads with 2 classes.  The cfg text and the seeded darknet .weights image are generated
here so tests, the golden generator and the GPU box all see the same bytes; nothing is read from /root/reference.
"""
from __future__ import annotations

import os

import numpy as np

CFG = """[net]
width=96
height=64
channels=3

[convolutional]
batch_normalize=1
filters=16
size=3
stride=1
pad=1
activation=leaky

[maxpool]
size=2
stride=2

[convolutional]
batch_normalize=1
filters=32
size=3
stride=1
pad=1
activation=relu

[maxpool]
size=2
stride=2

[convolutional]
batch_normalize=1
filters=32
size=3
stride=1
pad=1
groups=4
activation=leaky

[convolutional]
batch_normalize=1
filters=64
size=1
stride=1
pad=1
activation=leaky

[avgpool]
size=3
stride=1

[convolutional]
batch_normalize=1
filters=64
size=3
stride=2
pad=1
activation=leaky

[convolutional]
filters=21
size=1
stride=1
pad=1
activation=linear

[yolo]
mask = 3,4,5
anchors = 4,6, 8,10, 12,18, 20,24, 36,30, 50,44
classes=2
ignore_thresh = .3

[route]
layers = -3

[convolutional]
batch_normalize=1
filters=32
size=1
stride=1
pad=1
activation=leaky

[upsample]
stride=2

[route]
layers = -1, 5

[convolutional]
batch_normalize=1
filters=48
size=3
stride=1
pad=1
activation=leaky

[convolutional]
filters=21
size=1
stride=1
pad=1
activation=linear

[yolo]
mask = 0,1,2
anchors = 4,6, 8,10, 12,18, 20,24, 36,30, 50,44
classes=2
ignore_thresh = .3
"""

# (filters, size, in_channels / groups, batch_normalize) of every conv layer in cfg order
_CONVS = [(16, 3, 3, 1), (32, 3, 16, 1), (32, 3, 8, 1), (64, 1, 32, 1), (64, 3, 64, 1), (21, 1, 64, 0), (32, 1, 64, 1), (48, 3, 96, 1), (21, 1, 48, 0)]
W, H = 96, 64


def weights_bytes(seed: int = 20261017) -> bytes:
    """darknet .weights image (readme.txt:77-97): 20-byte header, then per conv: biases, [scales, means, variances], weights."""
    rng = np.random.default_rng(seed)
    parts = [np.array([0, 2, 5], "<i4").tobytes(), np.array([12345], "<u8").tobytes()]
    for fn, k, cpg, bn in _CONVS:
        parts.append(rng.uniform(-0.3, 0.3, fn).astype("<f4").tobytes())
        if bn:
            parts.append(rng.uniform(0.6, 1.4, fn).astype("<f4").tobytes())
            parts.append(rng.uniform(-0.2, 0.2, fn).astype("<f4").tobytes())
            parts.append(rng.uniform(0.3, 1.2, fn).astype("<f4").tobytes())
        parts.append((rng.standard_normal(fn * cpg * k * k) * (1.6 / np.sqrt(cpg * k * k))).astype("<f4").tobytes())
    return b"".join(parts)


def write(dirpath: str) -> tuple[str, str]:
    """Write tinygraph.cfg / tinygraph.weights into dirpath; returns their paths."""
    cfg, wts = os.path.join(dirpath, "tinygraph.cfg"), os.path.join(dirpath, "tinygraph.weights")
    with open(cfg, "w") as f:
        f.write(CFG)
    with open(wts, "wb") as f:
        f.write(weights_bytes())
    return cfg, wts


def frames(n: int) -> np.ndarray:
    """[n, H, pitch] seeded u8 BGR frames of the net's size."""
    from . import synth
    return synth.frames_u8(n, W, H, seed0=0x71A9)
