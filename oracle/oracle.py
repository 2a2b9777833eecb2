"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Python glue around the plain-C restatement in ``oracle/ffcnn_oracle.c``:

* an independent restatement of the reference's darknet cfg parser
  (ffcnn.c:50-84,128-208: sections by ``[name]``, keys by *first substring match*,
  effective pad = ``pad ? size/2 : 0``, inputw/h rounded up to a multiple of 32),
* the ``.weights`` reader + filter packing + BN fold (ffcnn.c:107-112,211-239),
* the layer-by-layer forward loop (ffcnn.c:476-520) over single-image CHW tensors,
  calling the C functions for every arithmetic step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
ASSETS = os.path.join(REPO, "baseline", "_ref")

CONV, AVGPOOL, MAXPOOL, UPSAMPLE, DROPOUT, SHORTCUT, ROUTE, YOLO = range(8)   # ffcnn.h:4-14
TYPE_NAMES = ["conv", "avgpool", "maxpool", "upsample", "dropout", "shortcut", "route", "yolo"]


class OrcBox(C.Structure):
    _fields_ = [("type", C.c_int), ("score", C.c_float), ("x1", C.c_float), ("y1", C.c_float),
                ("x2", C.c_float), ("y2", C.c_float)]


BOX_DTYPE = np.dtype([("type", "<i4"), ("score", "<f4"), ("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4")])

_lib = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "ffcnn_oracle.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/ffcnn.c"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        fp, ip, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
        L.orc_groupconv.argtypes = [fp, fp, fp] + [C.c_int] * 13
        L.orc_maxpool.argtypes = [fp, fp] + [C.c_int] * 5
        L.orc_avgpool.argtypes = [fp, fp] + [C.c_int] * 5
        L.orc_upsample.argtypes = [fp, fp] + [C.c_int] * 4
        L.orc_shortcut.argtypes = [fp, fp, fp, C.c_int, C.c_int]
        L.orc_net_input.argtypes = [u8p, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_int, ip, ip]
        L.orc_fold_bn.argtypes = [fp, fp, fp, fp, C.c_int]
        L.orc_yolo_decode.argtypes = [fp, C.c_int, C.c_int, C.c_int, ip, C.c_float, C.c_float, C.c_int, C.c_int,
                                      C.POINTER(OrcBox), C.c_int, C.c_int]
        L.orc_yolo_decode.restype = C.c_int
        L.orc_nms.argtypes = [C.POINTER(OrcBox), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_nms.restype = C.c_int
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ----------------------------------------------------------------------------- cfg

@dataclass
class Layer:
    type: int
    w: int = 0          # input geometry of this layer
    h: int = 0
    c: int = 0
    ow: int = 0         # output geometry
    oh: int = 0
    oc: int = 0
    fn: int = 0
    fs: int = 0
    stride: int = 1
    groups: int = 1
    pad: int = 0
    batchnorm: int = 0
    activation: int = 0
    deps: list = field(default_factory=list)
    classes: int = 0
    anchors: list = field(default_factory=list)   # 3 (w, h) pairs
    ignore_thresh: float = 0.0
    scale_x_y: float = 1.0
    filt: np.ndarray | None = None                # packed [fn, ALIGN4(k*k*c/g)+4]


def _atoi(s: str) -> int:
    s = s.lstrip(" \t\n\r\f\v")
    i, sign = 0, 1
    if i < len(s) and s[i] in "+-":
        sign = -1 if s[i] == "-" else 1
        i += 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    return sign * int(s[i:j]) if j > i else 0


def _atof(s: str) -> float:
    s = s.strip()
    n = len(s)
    while n > 0:
        try:
            return float(s[:n])
        except ValueError:
            n -= 1
    return 0.0


def _param(sec: str, key: str) -> str:
    """ffcnn.c:64-84 -- first substring hit of key in the section, skip '=' and ' ', read to EOL."""
    p = sec.find(key)
    if p < 0:
        return ""
    p += len(key)
    while p < len(sec) and sec[p] in "= ":
        p += 1
    e = sec.find("\n", p)
    return sec[p:] if e < 0 else sec[p:e]


def _activation(s: str) -> int:
    for i, name in enumerate(("linear", "relu", "leaky")):
        if s.startswith(name):
            return i
    return -1


def _align(x: int, n: int) -> int:
    return (x + n - 1) & ~(n - 1)


def parse_cfg(text: str, inputw: int = 0, inputh: int = 0) -> list[Layer]:
    layers: list[Layer] = []
    cur_w = cur_h = cur_c = 0
    pos = text.find("[")
    while pos >= 0:
        nxt = text.find("[", pos + 1)
        # the reference clips the section one character before the next '[' (ffcnn.c:129)
        sec = text[pos:] if nxt < 0 else text[pos:nxt - 1]
        head = text[pos:]
        idx = len(layers)
        L = None
        if head.startswith("[net]"):
            cur_w = _align(inputw, 32) if inputw else _atoi(_param(sec, "width"))
            cur_h = _align(inputh, 32) if inputh else _atoi(_param(sec, "height"))
            cur_c = _atoi(_param(sec, "channels"))
        elif head.startswith("[conv]") or head.startswith("[convolutional]"):
            L = Layer(CONV, cur_w, cur_h, cur_c)
            L.fn = _atoi(_param(sec, "filters"))
            L.fs = _atoi(_param(sec, "size"))
            L.stride = _atoi(_param(sec, "stride")) or 1
            L.groups = _atoi(_param(sec, "groups")) or 1
            L.pad = L.fs // 2 if _atoi(_param(sec, "pad")) else 0
            L.batchnorm = 1 if _atoi(_param(sec, "batch_normalize")) else 0
            L.activation = _activation(_param(sec, "activation"))
            L.oc = L.fn
            L.ow = (L.w - L.fs + 2 * L.pad) // L.stride + 1
            L.oh = (L.h - L.fs + 2 * L.pad) // L.stride + 1
        elif any(head.startswith(t) for t in ("[avg]", "[avgpool]", "[max]", "[maxpool]")):
            L = Layer(AVGPOOL if head.startswith("[avg") else MAXPOOL, cur_w, cur_h, cur_c)
            L.fs = _atoi(_param(sec, "size"))
            L.stride = _atoi(_param(sec, "stride")) or 1
            L.oc, L.ow, L.oh = L.c, L.w // L.stride, L.h // L.stride
        elif head.startswith("[upsample]"):
            L = Layer(UPSAMPLE, cur_w, cur_h, cur_c)
            L.stride = _atoi(_param(sec, "stride")) or 1
            L.oc, L.ow, L.oh = L.c, L.w * L.stride, L.h * L.stride
        elif head.startswith("[dropout]"):
            L = Layer(DROPOUT, cur_w, cur_h, cur_c)
            L.oc, L.ow, L.oh = L.c, L.w, L.h
        elif head.startswith("[shortcut]"):
            L = Layer(SHORTCUT, cur_w, cur_h, cur_c)
            L.deps = [_atoi(_param(sec, "from")) + idx]
            L.activation = _activation(_param(sec, "activation"))
            L.oc, L.ow, L.oh = L.c, L.w, L.h
        elif head.startswith("[route]"):
            L = Layer(ROUTE, cur_w, cur_h, cur_c)
            for tok in [t for t in _param(sec, "layers").split(",") if t != ""][:4]:
                d = _atoi(tok)
                d = d if d > 0 else idx + d
                L.deps.append(d)
                L.oc += layers[d].oc
                L.ow, L.oh = layers[d].ow, layers[d].oh
        elif head.startswith("[yolo]"):
            L = Layer(YOLO, cur_w, cur_h, cur_c)
            L.classes = _atoi(_param(sec, "classes"))
            sxy = _param(sec, "scale_x_y")
            L.scale_x_y = 1.0 if sxy == "" else float(np.float32(_atof(sxy)))
            L.ignore_thresh = float(np.float32(_atof(_param(sec, "ignore_thresh"))))
            masks = [_atoi(t) for t in _param(sec, "mask").split(",") if t != ""][:9]
            nums = [_atoi(t) for t in _param(sec, "anchors").split(",") if t != ""]
            pairs = [(nums[2 * i], nums[2 * i + 1]) for i in range(min(9, len(nums) // 2))]
            L.anchors = [pairs[masks[i]] for i in range(3)]
            # a yolo layer has no output tensor; the reference leaves its olayer geometry at 0
        if L is not None:
            layers.append(L)
            cur_w, cur_h, cur_c = L.ow, L.oh, L.oc
        pos = nxt
    return layers


def filter_row(L: Layer) -> int:
    return _align(L.fs * L.fs * (L.c // L.groups), 4) + 4


def load_weights(path: str, layers: list[Layer]) -> None:
    """ffcnn.c:211-239 -- 20-byte header, then per conv: bias, [scale, mean, var], filters."""
    try:
        blob = np.fromfile(path, dtype=np.uint8)
    except OSError:
        blob = None                                   # reference: missing file -> all-zero weights, no error
    off = 20
    for L in layers:
        if L.type != CONV:
            continue
        row, k = filter_row(L), L.fs * L.fs * (L.c // L.groups)
        L.filt = np.zeros((L.fn, row), np.float32)
        if blob is None:
            continue

        def take(n):
            nonlocal off
            a = np.zeros(n, np.float32)
            avail = max(0, min(n, (len(blob) - off) // 4))
            if avail:
                a[:avail] = blob[off:off + 4 * avail].view(np.float32)
            off += 4 * n
            return a

        bias = take(L.fn)
        scale = np.ones(L.fn, np.float32)
        mean = np.zeros(L.fn, np.float32)
        var = np.zeros(L.fn, np.float32)
        if L.batchnorm:
            scale, mean, var = take(L.fn), take(L.fn), take(L.fn)
            lib().orc_fold_bn(_fp(scale), _fp(bias), _fp(mean), _fp(var), L.fn)
        L.filt[:, :k] = take(L.fn * k).reshape(L.fn, k)
        L.filt[:, row - 4], L.filt[:, row - 3], L.filt[:, row - 2], L.filt[:, row - 1] = scale, bias, mean, var


def load_net(cfg_path: str, weights_path: str, inputw: int = 0, inputh: int = 0) -> list[Layer]:
    with open(cfg_path, "rb") as f:
        text = f.read().decode("latin-1")
    layers = parse_cfg(text, inputw, inputh)
    load_weights(weights_path, layers)
    return layers


# ----------------------------------------------------------------------------- ops

def groupconv(x: np.ndarray, filt: np.ndarray, L: Layer, v6_quirk: bool = True) -> np.ndarray:
    out = np.empty((L.oc, L.oh, L.ow), np.float32)
    x = np.ascontiguousarray(x, np.float32)
    filt = np.ascontiguousarray(filt, np.float32)
    lib().orc_groupconv(_fp(x), _fp(filt), _fp(out), L.w, L.h, L.c, L.groups, L.pad, L.stride,
                        L.fs, L.fn, L.ow, L.oh, L.oc, L.activation, 1 if v6_quirk else 0)
    return out


def conv_raw(x, filt, iw, ih, ic, ig, pad, stride, fs, fn, act, v6_quirk=True):
    ow, oh = (iw - fs + 2 * pad) // stride + 1, (ih - fs + 2 * pad) // stride + 1
    out = np.empty((fn, oh, ow), np.float32)
    lib().orc_groupconv(_fp(np.ascontiguousarray(x, np.float32)), _fp(np.ascontiguousarray(filt, np.float32)), _fp(out),
                        iw, ih, ic, ig, pad, stride, fs, fn, ow, oh, fn, act, 1 if v6_quirk else 0)
    return out


def maxpool(x, fs, stride):
    c, h, w = x.shape
    out = np.empty((c, -(-h // stride), -(-w // stride)), np.float32)
    lib().orc_maxpool(_fp(np.ascontiguousarray(x)), _fp(out), w, h, c, fs, stride)
    return out[:, :h // stride, :w // stride] if stride == 1 else out


def avgpool(x, fs, stride):
    c, h, w = x.shape
    out = np.empty((c, -(-h // stride), -(-w // stride)), np.float32)
    lib().orc_avgpool(_fp(np.ascontiguousarray(x)), _fp(out), w, h, c, fs, stride)
    return out


def upsample(x, stride):
    c, h, w = x.shape
    out = np.empty((c, h * stride, w * stride), np.float32)
    lib().orc_upsample(_fp(np.ascontiguousarray(x)), _fp(out), w, h, c, stride)
    return out


def shortcut(a, b, act):
    out = np.empty_like(a)
    lib().orc_shortcut(_fp(np.ascontiguousarray(a)), _fp(np.ascontiguousarray(b)), _fp(out), a.size, act)
    return out


def net_input(bgr: np.ndarray, w: int, h: int, W: int, H: int, mean=(0, 0, 0), norm=(1 / 255., 1 / 255., 1 / 255.)):
    """bgr: raw bytes, rows top-down, pitch ALIGN(3w,4). Returns (CHW fp32 [3,H,W], s1, s2)."""
    out = np.zeros((3, H, W), np.float32)
    m, n = np.asarray(mean, np.float32), np.asarray(norm, np.float32)
    s1, s2 = C.c_int(0), C.c_int(0)
    buf = np.ascontiguousarray(bgr, np.uint8).reshape(-1)
    lib().orc_net_input(buf.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, _fp(m), _fp(n), _fp(out), W, H,
                        C.byref(s1), C.byref(s2))
    return out, s1.value, s2.value


def yolo_decode(head: np.ndarray, L: Layer, netw: int, neth: int, boxes: np.ndarray, n0: int) -> int:
    anch = (C.c_int * 6)(*[v for p in L.anchors for v in p])
    c, gh, gw = head.shape
    return lib().orc_yolo_decode(_fp(np.ascontiguousarray(head)), gw, gh, L.classes, anch,
                                 L.ignore_thresh, L.scale_x_y, netw, neth,
                                 boxes.ctypes.data_as(C.POINTER(OrcBox)), len(boxes), n0)


def nms(boxes: np.ndarray, n: int, s1: int, s2: int, thr: float = 0.5, min_mode: int = 1) -> int:
    return lib().orc_nms(boxes.ctypes.data_as(C.POINTER(OrcBox)), n, thr, min_mode, s1, s2)


def forward(layers: list[Layer], x: np.ndarray, s1: int = 1, s2: int = 1, v6_quirk: bool = True, keep: bool = True):
    """Layer-by-layer forward (ffcnn.c:476-520). Returns (outs, raw_boxes, final_boxes).

    outs[i] = CHW output of layer i (None for yolo layers)."""
    outs: list[np.ndarray | None] = []
    netw, neth = layers[0].w, layers[0].h
    boxes = np.zeros(netw * neth * layers[0].c * 4 // 24, BOX_DTYPE)     # bbox_max, ffcnn.c:243
    nb = 0
    cur = np.ascontiguousarray(x, np.float32)
    for i, L in enumerate(layers):
        if L.type == CONV:
            y = groupconv(cur, L.filt, L, v6_quirk)
        elif L.type == MAXPOOL:
            y = maxpool(cur, L.fs, L.stride)
        elif L.type == AVGPOOL:
            y = avgpool(cur, L.fs, L.stride)
        elif L.type == UPSAMPLE:
            y = upsample(cur, L.stride)
        elif L.type == DROPOUT:
            y = cur
        elif L.type == SHORTCUT:
            y = shortcut(cur, outs[L.deps[0]], L.activation)
        elif L.type == ROUTE:
            y = np.ascontiguousarray(np.concatenate([outs[d] for d in L.deps], axis=0))
        elif L.type == YOLO:
            nb = yolo_decode(cur, L, netw, neth, boxes, nb)
            y = None
        else:
            raise ValueError(L.type)
        outs.append(y)
        cur = y
    raw = boxes[:nb].copy()
    nfinal = nms(boxes, nb, s1, s2)
    return outs, raw, boxes[:nfinal].copy()
