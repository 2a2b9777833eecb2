"""ffcnn_b200 -- Python mirror of the C-ABI of ``libffcnn_b200.so`` (ctypes, no torch types).

The product is the shared library (host C + hand-written sm_100a CUDA, see ``ffcnn_b200/csrc`` and
``include/*.h``); this module only binds it for ``tests/`` and ``bench.py``.  Names follow the
reference's API (``net_load / net_input / net_forward / net_free``, ``groupconv``; ffcnn.h:48-52,
conv.h:4-7) plus the additive batched entry points of ``include/ffcnn_b200.h``.

There is no CPU fallback here either: everything that computes raises ``FfcnnError`` when the
library reports no CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.environ.get("FFCNN_LIB") or os.path.join(HERE, "libffcnn_b200.so")     # FFCNN_LIB: a developer build (e.g. `make tc`)
ASSETS = os.path.join(REPO, "baseline", "_ref")

BOX_DTYPE = np.dtype([("type", "<i4"), ("score", "<f4"), ("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4")])

LAYER_TYPES = ["conv", "avgpool", "maxpool", "upsample", "dropout", "shortcut", "route", "yolo"]


class FfcnnError(RuntimeError):
    pass


class LAYER(C.Structure):                      # include/ffcnn.h == reference ffcnn.h:16-27
    _fields_ = [("type", C.c_int), ("refcnt", C.c_int), ("data", C.POINTER(C.c_float)), ("filter", C.POINTER(C.c_float)),
                ("w", C.c_int), ("h", C.c_int), ("c", C.c_int), ("pad", C.c_int), ("stride", C.c_int), ("fn", C.c_int),
                ("fs", C.c_int), ("groups", C.c_int), ("batchnorm", C.c_int), ("activation", C.c_int),
                ("depend_list", C.c_int * 4), ("depend_num", C.c_int), ("class_num", C.c_int),
                ("anchor_list", (C.c_int * 2) * 3), ("ignore_thres", C.c_float), ("scale_x_y", C.c_float)]


class BBOX(C.Structure):
    _fields_ = [("type", C.c_int), ("score", C.c_float), ("x1", C.c_float), ("y1", C.c_float), ("x2", C.c_float), ("y2", C.c_float)]


class NET(C.Structure):
    _fields_ = [("layer_list", C.POINTER(LAYER)), ("layer_num", C.c_int), ("bbox_list", C.POINTER(BBOX)),
                ("bbox_num", C.c_int), ("bbox_max", C.c_int), ("s1", C.c_int), ("s2", C.c_int),
                ("weight_size", C.c_int), ("weight_buf", C.POINTER(C.c_float)), ("cnntempbuf", C.POINTER(C.c_float)),
                ("cnnbufsize", C.c_int), ("timeused", C.c_int * 8)]


def build(verbose: bool = False) -> str:
    """Compile libffcnn_b200.so in-tree (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    r = subprocess.run(["make", "-j4", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise FfcnnError("building libffcnn_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library (building it first if it is missing). Fails loudly if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
    fp, vp, ip = C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int)
    NP = C.POINTER(NET)
    sig = {
        "ffb_last_error": (C.c_char_p, []),
        "ffb_device_count": (C.c_int, []),
        "ffb_net_parse": (NP, [C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
        "ffb_net_attach": (C.c_int, [NP, C.c_int, C.c_int]),
        "ffb_packed_weights_device": (vp, [NP, C.POINTER(C.c_size_t)]),
        "ffb_commit_weights": (C.c_int, [NP]),
        "ffb_set_option": (C.c_int, [NP, C.c_char_p, C.c_int]),
        "ffb_get_option": (C.c_int, [NP, C.c_char_p]),
        "ffb_set_stream": (C.c_int, [NP, vp]),
        "ffb_get_stream": (vp, [NP]),
        "ffb_sync": (C.c_int, [NP]),
        "ffb_input_u8": (C.c_int, [NP, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, C.c_int]),
        "ffb_input_chw": (C.c_int, [NP, fp, C.c_int, C.c_int, C.c_int]),
        "ffb_forward": (C.c_int, [NP]),
        "ffb_detect": (C.c_int, [NP]),
        "ffb_detect_enqueue": (C.c_int, [NP]),
        "ffb_detect_finish": (C.c_int, [NP]),
        "ffb_last_d2h_bytes": (C.c_long, [NP]),
        "ffb_boxes": (C.c_int, [NP, C.c_int, C.POINTER(C.POINTER(BBOX))]),
        "ffb_raw_boxes": (C.c_int, [NP, C.c_int, C.POINTER(C.POINTER(BBOX))]),
        "ffb_detect_batch_u8": (C.c_int, [NP, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]),
        "ffb_submit_u8": (C.c_int, [NP, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]),
        "ffb_collect": (C.c_int, [NP]),
        "ffb_layer_output": (C.c_long, [NP, C.c_int, C.c_int, fp, C.c_long]),
        "ffb_layer_times": (C.c_int, [NP, fp, C.c_int, C.c_int, C.c_int]),
        "ffb_layer_cost": (C.c_int, [NP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p, C.c_int]),
        "ffb_launches_per_forward": (C.c_int, [NP]),
        "ffb_measure_tf32_peak": (C.c_int, [C.POINTER(C.c_double)]),
        "ffb_multi_create": (vp, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, ip, C.c_int, C.c_int]),
        "ffb_multi_destroy": (None, [vp]),
        "ffb_multi_devices": (C.c_int, [vp]),
        "ffb_multi_net": (NP, [vp, C.c_int]),
        "ffb_multi_broadcast_bytes": (C.c_long, [vp]),
        "ffb_multi_detect_u8": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]),
        "ffb_multi_submit_u8": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]),
        "ffb_multi_collect": (C.c_int, [vp]),
        "ffb_multi_boxes": (C.c_int, [vp, C.c_int, C.POINTER(C.POINTER(BBOX))]),
        "ffb_conv_create": (vp, [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "ffb_conv_destroy": (None, [vp]),
        "ffb_conv_run": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
        "ffb_conv_kernel_name": (C.c_char_p, [vp]),
        "ffb_dev_alloc": (vp, [C.c_size_t]),
        "ffb_dev_free": (None, [vp]),
        "ffb_copy_h2d": (C.c_int, [vp, vp, C.c_size_t]),
        "ffb_copy_d2h": (C.c_int, [vp, vp, C.c_size_t]),
        "ffb_host_alloc_pinned": (vp, [C.c_size_t]),
        "ffb_host_alloc_pinned_wc": (vp, [C.c_size_t]),
        "ffb_host_free_pinned": (None, [vp]),
        "ffb_chw_to_nhwc": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "ffb_nhwc_to_chw": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "net_load": (NP, [C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
        "net_free": (None, [NP]),
        "net_input": (None, [NP, vp, C.c_int, C.c_int, fp, fp]),
        "net_forward": (None, [NP]),
        "net_dump": (None, [NP]),
        "net_profile": (None, [NP]),
        "groupconv": (None, [fp, fp, fp] + [C.c_int] * 12 + [C.POINTER(fp), ip]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


EXPORTS = ["ffb_last_error", "ffb_device_count", "ffb_net_parse", "ffb_net_attach", "ffb_packed_weights_device",
           "ffb_commit_weights", "ffb_set_option", "ffb_get_option", "ffb_set_stream", "ffb_get_stream", "ffb_sync",
           "ffb_input_u8", "ffb_input_chw", "ffb_forward", "ffb_detect", "ffb_detect_enqueue", "ffb_detect_finish",
           "ffb_last_d2h_bytes", "ffb_boxes", "ffb_raw_boxes",
           "ffb_detect_batch_u8", "ffb_submit_u8", "ffb_collect", "ffb_layer_output", "ffb_layer_times", "ffb_layer_cost", "ffb_launches_per_forward", "ffb_measure_tf32_peak",
           "ffb_multi_create", "ffb_multi_destroy", "ffb_multi_devices", "ffb_multi_net", "ffb_multi_broadcast_bytes",
           "ffb_multi_detect_u8", "ffb_multi_submit_u8", "ffb_multi_collect", "ffb_multi_boxes",
           "ffb_conv_create", "ffb_conv_destroy", "ffb_conv_run", "ffb_conv_kernel_name", "ffb_dev_alloc", "ffb_dev_free",
           "ffb_copy_h2d", "ffb_copy_d2h", "ffb_host_alloc_pinned", "ffb_host_alloc_pinned_wc", "ffb_host_free_pinned", "ffb_chw_to_nhwc",
           "ffb_nhwc_to_chw", "net_load", "net_free", "net_input", "net_forward", "net_dump", "net_profile", "groupconv",
           "bmp_load", "bmp_save", "bmp_free", "bmp_setpixel", "bmp_getpixel", "bmp_rectangle"]


def _err() -> str:
    return (lib().ffb_last_error() or b"").decode(errors="replace")


def _check(rc: int, what: str):
    if rc is None or rc < 0:
        raise FfcnnError(f"{what}: {_err()}")
    return rc


def device_count() -> int:
    return lib().ffb_device_count()


def measure_tf32_peak() -> float:
    """Measured dense tcgen05 kind::tf32 TFLOP/s of the current device (roofline denominator)."""
    v = C.c_double(0)
    _check(lib().ffb_measure_tf32_peak(C.byref(v)), "ffb_measure_tf32_peak")
    return v.value


def default_model() -> tuple[str, str]:
    return os.path.join(ASSETS, "yolo-fastest-1.1.cfg"), os.path.join(ASSETS, "yolo-fastest-1.1.weights")


def _boxes_to_np(ptr, n: int) -> np.ndarray:
    if n <= 0:
        return np.zeros(0, BOX_DTYPE)
    return np.frombuffer(C.string_at(ptr, n * 24), BOX_DTYPE).copy()


class Net:
    """The reference NET behind its own API, plus the batched extension.

    ``Net(cfg, weights, w, h)`` == ``net_load`` (ffcnn.h:48) when ``device`` is given (default 0);
    ``device=None`` stops after the host half (cfg/weights parsing) so host logic is testable without a GPU.
    """

    def __init__(self, cfg: str, weights: str | None, inputw: int = 0, inputh: int = 0, device: int | None = 0, max_batch: int = 1):
        L = lib()
        self._L = L
        self.p = L.ffb_net_parse(cfg.encode(), weights.encode() if weights else None, inputw, inputh)
        if not self.p:
            raise FfcnnError(f"ffb_net_parse: {_err()}")
        self.attached = False
        if device is not None:
            self.attach(device, max_batch)

    # ---- lifecycle
    def attach(self, device: int = 0, max_batch: int = 1):
        _check(self._L.ffb_net_attach(self.p, device, max_batch), "ffb_net_attach")
        self.attached = True

    def close(self):
        if getattr(self, "p", None):
            self._L.net_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- geometry / host state
    @property
    def net(self) -> NET:
        return self.p.contents

    @property
    def layer_num(self) -> int:
        return self.net.layer_num

    def layer(self, i: int) -> LAYER:
        return self.net.layer_list[i]

    @property
    def input_whc(self):
        l0 = self.layer(0)
        return l0.w, l0.h, l0.c

    def packed_weights(self) -> np.ndarray:
        n = self.net.weight_size
        return np.ctypeslib.as_array(self.net.weight_buf, shape=(n,)).copy()

    def out_shape(self, i: int):
        o = self.layer(i + 1)
        return o.c, o.h, o.w

    # ---- reference API (single frame)
    def net_input(self, bgr: np.ndarray, w: int, h: int, mean=(0, 0, 0), norm=(1 / 255., 1 / 255., 1 / 255.)):
        buf = np.ascontiguousarray(bgr, np.uint8)
        self._L.net_input(self.p, buf.ctypes.data, w, h, (C.c_float * 3)(*mean), (C.c_float * 3)(*norm))

    def input_tensor(self) -> np.ndarray:
        w, h, c = self.input_whc
        return np.ctypeslib.as_array(self.layer(0).data, shape=(c, h, w))

    def net_forward(self) -> np.ndarray:
        if not self.attached:
            raise FfcnnError("net_forward: no GPU engine attached (no CPU fallback)")
        self._L.net_forward(self.p)
        return _boxes_to_np(self.net.bbox_list, self.net.bbox_num)

    # ---- batched extension
    def set_option(self, name: str, value: int):
        _check(self._L.ffb_set_option(self.p, name.encode(), int(value)), f"ffb_set_option({name})")

    def get_option(self, name: str) -> int:
        return self._L.ffb_get_option(self.p, name.encode())

    def set_stream(self, cuda_stream: int | None):
        _check(self._L.ffb_set_stream(self.p, C.c_void_p(cuda_stream) if cuda_stream else None), "ffb_set_stream")

    def sync(self):
        _check(self._L.ffb_sync(self.p), "ffb_sync")

    def input_u8(self, frames, n: int, w: int, h: int, pitch: int, on_device: bool = False, mean=None, norm=None):
        """frames: numpy array (host) or an integer device/host pointer."""
        ptr = frames if isinstance(frames, int) else np.ascontiguousarray(frames, np.uint8).ctypes.data
        m = (C.c_float * 3)(*mean) if mean is not None else None
        nn = (C.c_float * 3)(*norm) if norm is not None else None
        _check(self._L.ffb_input_u8(self.p, ptr, n, w, h, pitch, m, nn, 1 if on_device else 0), "ffb_input_u8")

    def input_chw(self, x: np.ndarray, s1: int = 1, s2: int = 1):
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 3:
            x = x[None]
        self._keep = x
        _check(self._L.ffb_input_chw(self.p, x.ctypes.data_as(C.POINTER(C.c_float)), x.shape[0], s1, s2), "ffb_input_chw")

    def forward(self):
        _check(self._L.ffb_forward(self.p), "ffb_forward")

    def detect(self):
        _check(self._L.ffb_detect(self.p), "ffb_detect")

    def detect_enqueue(self) -> int:
        return _check(self._L.ffb_detect_enqueue(self.p), "ffb_detect_enqueue")

    def detect_finish(self):
        _check(self._L.ffb_detect_finish(self.p), "ffb_detect_finish")

    def last_d2h_bytes(self) -> int:
        return self._L.ffb_last_d2h_bytes(self.p)

    def boxes(self, frame: int = 0, raw: bool = False) -> np.ndarray:
        ptr = C.POINTER(BBOX)()
        n = _check((self._L.ffb_raw_boxes if raw else self._L.ffb_boxes)(self.p, frame, C.byref(ptr)), "ffb_boxes")
        return _boxes_to_np(ptr, n)

    def detect_batch_u8(self, frames, n: int, w: int, h: int, pitch: int):
        ptr = frames if isinstance(frames, int) else np.ascontiguousarray(frames, np.uint8).ctypes.data
        _check(self._L.ffb_detect_batch_u8(self.p, ptr, n, w, h, pitch, None, None), "ffb_detect_batch_u8")

    def submit_u8(self, frames, n: int, w: int, h: int, pitch: int):
        ptr = frames if isinstance(frames, int) else np.ascontiguousarray(frames, np.uint8).ctypes.data
        _check(self._L.ffb_submit_u8(self.p, ptr, n, w, h, pitch, None, None), "ffb_submit_u8")

    def collect(self):
        _check(self._L.ffb_collect(self.p), "ffb_collect")

    def layer_output(self, layer: int, frame: int = 0) -> np.ndarray | None:
        n = _check(self._L.ffb_layer_output(self.p, layer, frame, None, 0), "ffb_layer_output")
        if n == 0:
            return None
        c, h, w = self.out_shape(layer) if layer >= 0 else (self.input_whc[2], self.input_whc[1], self.input_whc[0])
        out = np.empty((c, h, w), np.float32)
        _check(self._L.ffb_layer_output(self.p, layer, frame, out.ctypes.data_as(C.POINTER(C.c_float)), out.size), "ffb_layer_output")
        return out

    def layer_times(self, reps: int = 5, flush_l2: bool = False) -> np.ndarray:
        ms = np.zeros(self.layer_num, np.float32)
        _check(self._L.ffb_layer_times(self.p, ms.ctypes.data_as(C.POINTER(C.c_float)), self.layer_num, reps, 1 if flush_l2 else 0), "ffb_layer_times")
        return ms

    def layer_cost(self, i: int):
        b, f = C.c_double(0), C.c_double(0)
        name = C.create_string_buffer(64)
        _check(self._L.ffb_layer_cost(self.p, i, C.byref(b), C.byref(f), name, 64), "ffb_layer_cost")
        return b.value, f.value, name.value.decode()

    def launches_per_forward(self) -> int:
        return self._L.ffb_launches_per_forward(self.p)

    def packed_weights_device(self):
        n = C.c_size_t(0)
        p = self._L.ffb_packed_weights_device(self.p, C.byref(n))
        if not p:
            raise FfcnnError(f"ffb_packed_weights_device: {_err()}")
        return p, n.value

    def commit_weights(self):
        _check(self._L.ffb_commit_weights(self.p), "ffb_commit_weights")


class MultiNet:
    """ffb_multi_*: the C multi-GPU frontend (one NET, host thread and CUDA graph per device; NCCL weight broadcast at load)."""

    def __init__(self, cfg: str, weights: str | None, inputw: int = 0, inputh: int = 0, devices=None, max_batch_per_device: int = 1):
        L = lib()
        self._L = L
        arr = (C.c_int * len(devices))(*devices) if devices else None
        self.h = L.ffb_multi_create(cfg.encode(), weights.encode() if weights else None, inputw, inputh, arr, len(devices) if devices else 0, max_batch_per_device)
        if not self.h:
            raise FfcnnError(f"ffb_multi_create: {_err()}")

    @property
    def devices(self) -> int:
        return self._L.ffb_multi_devices(self.h)

    @property
    def broadcast_bytes(self) -> int:
        return self._L.ffb_multi_broadcast_bytes(self.h)

    def set_option(self, name: str, value: int):
        for g in range(self.devices):
            _check(self._L.ffb_set_option(self._L.ffb_multi_net(self.h, g), name.encode(), int(value)), f"ffb_set_option({name})")

    def detect_u8(self, frames, n: int, w: int, h: int, pitch: int):
        ptr = frames if isinstance(frames, int) else np.ascontiguousarray(frames, np.uint8).ctypes.data
        _check(self._L.ffb_multi_detect_u8(self.h, ptr, n, w, h, pitch, None, None), "ffb_multi_detect_u8")

    def submit_u8(self, frames, n: int, w: int, h: int, pitch: int):
        ptr = frames if isinstance(frames, int) else np.ascontiguousarray(frames, np.uint8).ctypes.data
        _check(self._L.ffb_multi_submit_u8(self.h, ptr, n, w, h, pitch, None, None), "ffb_multi_submit_u8")

    def collect(self):
        _check(self._L.ffb_multi_collect(self.h), "ffb_multi_collect")

    def boxes(self, frame: int) -> np.ndarray:
        ptr = C.POINTER(BBOX)()
        n = _check(self._L.ffb_multi_boxes(self.h, frame, C.byref(ptr)), "ffb_multi_boxes")
        return _boxes_to_np(ptr, n)

    def close(self):
        if getattr(self, "h", None):
            self._L.ffb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def groupconv(x: np.ndarray, filt: np.ndarray, iw, ih, ic, ig, pad, stride, fs, fn, act) -> np.ndarray:
    """The operator seam (conv.h:4-7): host CHW in, host CHW out, computed on the GPU."""
    if device_count() <= 0:
        raise FfcnnError("groupconv: no CUDA device (no CPU fallback)")
    ow, oh = (iw - fs + 2 * pad) // stride + 1, (ih - fs + 2 * pad) // stride + 1
    x = np.ascontiguousarray(x, np.float32)
    filt = np.ascontiguousarray(filt, np.float32)
    out = np.full((fn, oh, ow), np.nan, np.float32)
    fp = C.POINTER(C.c_float)
    buf, bufsize = fp(), C.c_int(0)
    lib().groupconv(x.ctypes.data_as(fp), filt.ctypes.data_as(fp), out.ctypes.data_as(fp), iw, ih, ic, ig, pad, stride,
                    fs, fn, ow, oh, fn, act, C.byref(buf), C.byref(bufsize))
    return out


class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.nbytes = nbytes
        self.ptr = lib().ffb_dev_alloc(nbytes)
        if not self.ptr:
            raise FfcnnError(f"ffb_dev_alloc: {_err()}")

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        _check(lib().ffb_copy_h2d(self.ptr, a.ctypes.data, a.nbytes), "ffb_copy_h2d")
        return self

    def download(self, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype)
        _check(lib().ffb_copy_d2h(out.ctypes.data, self.ptr, out.nbytes), "ffb_copy_d2h")
        return out

    def free(self):
        if self.ptr:
            lib().ffb_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ConvOp:
    """One convolution layer on device NHWC tensors (configs 3 and 4 of BASELINE.json, op-level parity)."""

    def __init__(self, filt: np.ndarray, ic, groups, pad, stride, fs, fn, act, dw5_exact=False, pw_mode=0, dw_mode=0):
        filt = np.ascontiguousarray(filt, np.float32)
        self.geom = (ic, groups, pad, stride, fs, fn, act)
        self.h = lib().ffb_conv_create(filt.ctypes.data_as(C.POINTER(C.c_float)), ic, groups, pad, stride, fs, fn, act,
                                       (1 if dw5_exact else 0) | (pw_mode << 8) | (dw_mode << 16))
        if not self.h:
            raise FfcnnError(f"ffb_conv_create: {_err()}")

    @property
    def kernel(self) -> str:
        return lib().ffb_conv_kernel_name(self.h).decode()

    def run_ptr(self, in_ptr: int, out_ptr: int, n: int, ih: int, iw: int, stream: int | None = None):
        """Raw device pointers (NHWC fp32, channel pitch = ALIGN(c, 4)); asynchronous on `stream`."""
        _check(lib().ffb_conv_run(self.h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), n, ih, iw, C.c_void_p(stream) if stream else None), "ffb_conv_run")

    def run(self, d_in: DeviceBuffer, d_out: DeviceBuffer, n: int, ih: int, iw: int, stream: int | None = None):
        _check(lib().ffb_conv_run(self.h, d_in.ptr, d_out.ptr, n, ih, iw, C.c_void_p(stream) if stream else None), "ffb_conv_run")

    def __call__(self, x_nhwc: np.ndarray) -> np.ndarray:
        """x: [n, h, w, ic] host -> [n, oh, ow, fn] host (channel dims padded to a multiple of 4 on device)."""
        ic, groups, pad, stride, fs, fn, act = self.geom
        n, ih, iw, _ = x_nhwc.shape
        oh, ow = (ih - fs + 2 * pad) // stride + 1, (iw - fs + 2 * pad) // stride + 1
        ldi, ldo = (ic + 3) & ~3, (fn + 3) & ~3
        xin = np.zeros((n, ih, iw, ldi), np.float32)
        xin[..., :ic] = x_nhwc
        d_in = DeviceBuffer(xin.nbytes).upload(xin)
        d_out = DeviceBuffer(n * oh * ow * ldo * 4)
        self.run(d_in, d_out, n, ih, iw)
        out = d_out.download((n, oh, ow, ldo))
        d_in.free(); d_out.free()
        return out[..., :fn]

    def close(self):
        if self.h:
            lib().ffb_conv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
