"""Seeded synthetic inputs shared by tests, bench.py and the CPU baseline (oracle/ref_bench.c uses the same stream).

SURVEY 8(d) config 2: frame f of set S1 = splitmix64(seed = 0xFFC0 + f) bytes, 320x320x3 u8 BGR, pitch 960.
"""
from __future__ import annotations

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, n: int) -> np.ndarray:
    """First n outputs of splitmix64 seeded with `seed` (vectorised; wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        s = np.uint64(seed) + _GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = (s ^ (s >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def frame_bytes(nbytes: int, seed: int) -> np.ndarray:
    n = (nbytes + 7) // 8
    return splitmix64(seed, n).view(np.uint8)[:nbytes].copy()


def frames_u8(n: int, w: int = 320, h: int = 320, seed0: int = 0xFFC0) -> np.ndarray:
    """[n, h, pitch] u8 BGR rows top-down, pitch = ALIGN(3w, 4) -- the layout net_input takes (ffcnn.c:274)."""
    pitch = (3 * w + 3) & ~3
    out = np.empty((n, h, pitch), np.uint8)
    for f in range(n):
        out[f] = frame_bytes(h * pitch, seed0 + f).reshape(h, pitch)
    return out


def shifted_frames_from(img: np.ndarray, w: int, h: int, n: int, W: int = 320, H: int = 320) -> np.ndarray:
    """Set S2 of SURVEY 8(d): frames derived from a real picture so decode/NMS see boxes.

    img: [h, pitch] BGR bytes.  Frame f = nearest-resize of img to fit WxH (as net_input does), then rolled by
    (f mod 16) pixels in x and y; returned as [n, H, 3W] u8 BGR (pitch = 3W, W multiple of 4)."""
    if w * H > h * W:
        sw, sh, s1, s2 = W, W * h // w, w, W
    else:
        sh, sw, s1, s2 = H, H * w // h, h, H
    ys = (np.arange(sh) * s1 // s2)[:, None]
    xs = (np.arange(sw) * s1 // s2)[None, :]
    base = np.zeros((H, W, 3), np.uint8)
    px = img[:, :3 * w].reshape(h, w, 3)
    base[:sh, :sw] = px[ys, xs]
    out = np.empty((n, H, 3 * W), np.uint8)
    for f in range(n):
        s = f % 16
        out[f] = np.roll(base, (s, s), axis=(0, 1)).reshape(H, 3 * W)
    return out
