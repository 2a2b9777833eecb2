"""Operator microbenchmarks (BASELINE configs 3 and 4 and any other shape): device-resident NHWC tensors, CUDA-event timing.
usage: op_bench.py [--lib path.so] shape...   shape = ih,iw,ic,groups,pad,stride,fs,fn,batch[,pw_mode[,dw_mode]]"""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
args = sys.argv[1:]
if args and args[0] == "--lib":
    import ffcnn_b200 as fb
    fb.LIB_PATH = os.path.abspath(args[1]); args = args[2:]
import torch
import ffcnn_b200 as fb

PEAK = 6550.7
try:
    PEAK = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def bench(ih, iw, ic, groups, pad, stride, fs, fn, n, pw_mode=0, dw_mode=0, reps=20):
    rng = np.random.default_rng(1)
    k = fs * fs * (ic // groups); row = ((k + 3) & ~3) + 4
    f = np.zeros((fn, row), np.float32); f[:, :k] = rng.standard_normal((fn, k)) / np.sqrt(k)
    f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
    op = fb.ConvOp(f, ic, groups, pad, stride, fs, fn, 2, pw_mode=pw_mode, dw_mode=dw_mode)
    oh, ow = (ih - fs + 2 * pad) // stride + 1, (iw - fs + 2 * pad) // stride + 1
    ldi, ldo = (ic + 3) & ~3, (fn + 3) & ~3
    x = torch.randn((n, ih, iw, ldi), device="cuda", dtype=torch.float32)
    y = torch.empty((n, oh, ow, ldo), device="cuda", dtype=torch.float32)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    class P:  # duck-typed device buffer
        def __init__(s, t): s.ptr = t.data_ptr()
    for _ in range(3): op.run(P(x), P(y), n, ih, iw, st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): op.run(P(x), P(y), n, ih, iw, st.cuda_stream)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = 4.0 * (n * ih * iw * ic + n * oh * ow * fn + fn * (k + 2))
    fl = 2.0 * k * fn * n * oh * ow
    print("%-20s in %dx%dx%d -> %d  k%d s%d g%d batch %d: %.4f ms  %.1f GB/s (%.1f%% of %.0f)  %.1f TFLOP/s" %
          (op.kernel, ih, iw, ic, fn, fs, stride, groups, n, ms, by / ms / 1e6, 100 * by / ms / 1e6 / PEAK, PEAK, fl / ms / 1e9), flush=True)
    op.close()
    return ms

if __name__ == "__main__":
    for a in args:
        v = [int(t) for t in a.split(",")]
        bench(*v)
