"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes access to the UNMODIFIED reference compiled into ``oracle/_ref/`` by
``oracle/Makefile`` (sources stay under /root/reference).  Variants:

* ``v6``    conv-v6.c with build.sh's -Ofast/-flto flags  -> BASELINE.json's named oracle
* ``v6_O2`` conv-v6.c at -O2                              -> v6 semantics without re-association
* ``v0``    conv-v0.c at -O2                              -> exact-math oracle (== v1..v5)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def available(variant: str = "v6") -> bool:
    return os.path.exists(os.path.join(REFDIR, f"libffcnn_ref_{variant}.so"))


_libs: dict = {}


def lib(variant: str = "v6"):
    if variant not in _libs:
        orc.build()
        # The reference reads memory it never wrote: im2row fills fs*fs*c lanes of each ALIGN(fs*fs*c, 4)-float scratch row and the dot
        # product runs over all of them (conv-v6.c:9-42; the stem has 27 taps in rows of 28), relying on the filter's pad lane being 0.
        # The scratch comes from malloc: in a fresh process it is zero, inside a long-lived Python process it can hold a NaN, and
        # NaN * 0 poisons column 0 of layer 0 (seen as a test that failed one run in ten).  glibc's M_PERTURB = 255 makes every
        # allocation start as zero bytes -- the state the reference's own binary sees -- without touching its sources or build.
        try:
            C.CDLL(None).mallopt(C.c_int(-6), C.c_int(255))
        except Exception:
            pass
        L = C.CDLL(os.path.join(REFDIR, f"libffcnn_ref_{variant}.so"), mode=os.RTLD_LOCAL)
        fp = C.POINTER(C.c_float)
        L.net_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.net_load.restype = C.c_void_p
        L.net_free.argtypes = [C.c_void_p]
        L.net_input.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, fp, fp]
        L.net_forward.argtypes = [C.c_void_p]
        L.refh_layer_num.argtypes = [C.c_void_p]
        L.refh_layer_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.refh_input_ptr.argtypes = [C.c_void_p]
        L.refh_input_ptr.restype = fp
        L.refh_weight_buf.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.refh_weight_buf.restype = fp
        L.refh_scale.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.refh_forward_dump.argtypes = [C.c_void_p, C.POINTER(fp), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.refh_forward_dump.restype = C.c_int
        L.refh_groupconv.argtypes = [fp, fp, fp] + [C.c_int] * 12
        _libs[variant] = L
    return _libs[variant]


def groupconv(x, filt, iw, ih, ic, ig, pad, stride, fs, fn, act, variant="v6"):
    """The reference operator seam (conv.h:4-7) on CHW input / packed filter rows."""
    ow, oh = (iw - fs + 2 * pad) // stride + 1, (ih - fs + 2 * pad) // stride + 1
    x = np.ascontiguousarray(x, np.float32)
    filt = np.ascontiguousarray(filt, np.float32)
    out = np.zeros((fn, oh, ow), np.float32)
    fp = C.POINTER(C.c_float)
    lib(variant).refh_groupconv(x.ctypes.data_as(fp), filt.ctypes.data_as(fp), out.ctypes.data_as(fp),
                                iw, ih, ic, ig, pad, stride, fs, fn, ow, oh, fn, act)
    return out


class RefNet:
    """The reference NET driven through its own public API (ffcnn.h:48-52)."""

    class _NetHead(C.Structure):       # leading fields of NET, ffcnn.h:34-46
        _fields_ = [("layer_list", C.c_void_p), ("layer_num", C.c_int), ("bbox_list", C.c_void_p),
                    ("bbox_num", C.c_int), ("bbox_max", C.c_int), ("s1", C.c_int), ("s2", C.c_int)]

    def __init__(self, cfg: str, weights: str, inputw: int = 0, inputh: int = 0, variant: str = "v6"):
        self.L = lib(variant)
        self.net = self.L.net_load(cfg.encode(), weights.encode(), inputw, inputh)
        if not self.net:
            raise RuntimeError("reference net_load failed")
        self.n = self.L.refh_layer_num(self.net)
        info = (C.c_int * 12)()
        self.info = []
        for i in range(self.n):
            self.L.refh_layer_info(self.net, i, info)
            self.info.append(list(info))
        self.W, self.H, self.Cin = self.info[0][1], self.info[0][2], self.info[0][3]

    def close(self):
        if self.net:
            self.L.net_free(self.net)
            self.net = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def head(self):
        return self._NetHead.from_address(self.net)

    def input_bgr(self, bgr: np.ndarray, w: int, h: int, mean=(0, 0, 0), norm=(1 / 255., 1 / 255., 1 / 255.)):
        m = (C.c_float * 3)(*mean)
        n = (C.c_float * 3)(*norm)
        buf = np.ascontiguousarray(bgr, np.uint8)
        self.L.net_input(self.net, buf.ctypes.data, w, h, m, n)

    def input_tensor(self) -> np.ndarray:
        p = self.L.refh_input_ptr(self.net)
        return np.ctypeslib.as_array(p, shape=(self.Cin, self.H, self.W))

    def set_input_tensor(self, x: np.ndarray, s1: int = 1, s2: int = 1):
        self.input_tensor()[...] = x
        hd = self.head()
        hd.bbox_num, hd.s1, hd.s2 = 0, s1, s2

    def packed_weights(self) -> np.ndarray:
        n = C.c_int(0)
        p = self.L.refh_weight_buf(self.net, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def boxes(self) -> np.ndarray:
        hd = self.head()
        if hd.bbox_num == 0:
            return np.zeros(0, orc.BOX_DTYPE)
        raw = C.string_at(hd.bbox_list, hd.bbox_num * 24)
        return np.frombuffer(raw, orc.BOX_DTYPE).copy()

    def forward(self) -> np.ndarray:
        self.L.net_forward(self.net)
        return self.boxes()

    def forward_dump(self, want=None):
        """Returns (outs, raw_boxes, final_boxes); outs[i] is CHW or None. want: iterable of layer ids (None = all)."""
        fp = C.POINTER(C.c_float)
        outs, ptrs = [], (fp * self.n)()
        for i, inf in enumerate(self.info):
            ow, oh, oc = inf[4], inf[5], inf[6]
            if inf[0] == orc.YOLO or (want is not None and i not in want):
                outs.append(None)
                ptrs[i] = None
            else:
                a = np.zeros((oc, oh, ow), np.float32)
                outs.append(a)
                ptrs[i] = a.ctypes.data_as(fp)
        cap = 4096
        raw = np.zeros(cap, orc.BOX_DTYPE)
        nraw = C.c_int(0)
        rc = self.L.refh_forward_dump(self.net, ptrs, raw.ctypes.data, cap, C.byref(nraw))
        if rc != 0:
            raise RuntimeError("refh_forward_dump failed")
        return outs, raw[:min(cap, nraw.value)].copy(), self.boxes()


def load_bmp(path: str):
    """24-bit BMP -> (bytes top-down with pitch ALIGN(3w,4), w, h) -- what bmp_load hands to net_input (bmpfile.c:42-69)."""
    raw = np.fromfile(path, np.uint8)
    w = int(raw[18:22].view("<u4")[0])
    h = int(raw[22:26].view("<u4")[0])
    pitch = (w * 3 + 3) & ~3
    body = raw[54:54 + pitch * h].reshape(h, pitch)
    return np.ascontiguousarray(body[::-1]), w, h
