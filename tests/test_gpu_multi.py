"""GPU (B200): the C multi-GPU frontend ffb_multi_* (include/ffcnn_b200.h) -- one NET + host thread + CUDA graph
per device, contiguous frame shards, weights handed to every device but the first from outside its own file read.

On a one-GPU box the frontend is exercised with the SAME device listed twice (two replicas, device-to-device weight
hand-off, two worker threads sharing one GPU): threads, sharding, ffb_commit_weights and frame-ordered results are all the
real code.  With two or more GPUs the NCCL broadcast (dlopen'ed libnccl.so.2) is exercised too."""
import numpy as np
import pytest

import ffcnn_b200 as fb
from ffcnn_b200 import synth
from oracle import ref

pytestmark = pytest.mark.gpu


def _frames(bmp, n):
    img, w, h = ref.load_bmp(bmp)
    return np.ascontiguousarray(synth.shifted_frames_from(img, w, h, n).reshape(n, 320, 960))


def _single(cfg, wts, fr):
    net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=len(fr))
    net.detect_batch_u8(fr, len(fr), 320, 320, 960)
    want = [net.boxes(f).tobytes() for f in range(len(fr))]
    net.close()
    return want


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_multi_frontend_replicas_on_one_gpu(assets, devices):
    cfg, wts, bmp = assets
    fr = _frames(bmp, 11)                                   # 11 frames over 1/2/3 workers: ragged shards
    want = _single(cfg, wts, fr)
    m = fb.MultiNet(cfg, wts, 0, 0, devices=devices, max_batch_per_device=11)
    assert m.devices == len(devices) and m.broadcast_bytes == 0
    m.detect_u8(fr, 11, 320, 320, 960)
    assert [m.boxes(f).tobytes() for f in range(11)] == want and any(len(x) for x in want)
    # pipelined: two batches in flight, different sizes (the second smaller than the number of workers at 3 devices)
    m.submit_u8(fr, 11, 320, 320, 960)
    m.submit_u8(fr[3:5], 2, 320, 320, 960)
    m.collect()
    assert [m.boxes(f).tobytes() for f in range(11)] == want
    m.collect()
    assert [m.boxes(f).tobytes() for f in range(2)] == want[3:5]
    with pytest.raises(fb.FfcnnError):
        m.collect()
    with pytest.raises(fb.FfcnnError):
        m.boxes(2)
    m.close()


def test_multi_frontend_nccl_broadcast_across_gpus(assets):
    n = fb.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cfg, wts, bmp = assets
    fr = _frames(bmp, 16)
    want = _single(cfg, wts, fr)
    m = fb.MultiNet(cfg, wts, 0, 0, devices=list(range(min(n, 8))), max_batch_per_device=16)
    assert m.broadcast_bytes == 356576 * 4                  # ffcnn.c:150 weight_size of yolo-fastest-1.1, once, and nothing else on the wire
    for _ in range(2):
        m.detect_u8(fr, 16, 320, 320, 960)
        assert [m.boxes(f).tobytes() for f in range(16)] == want
    m.close()
