"""Developer tool: per-tile timeline of CTA 0 of the tcgen05 pointwise kernel (needs a -DFFB_TC_TRACE build:
make -C ffcnn_b200/csrc clean && make -C ffcnn_b200/csrc EXTRA=-DFFB_TC_TRACE). usage: tc_trace.py ih,iw,K,N,batch"""
import os, sys, ctypes as C
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
import ffcnn_b200 as fb
ih, iw, K, N, n = [int(t) for t in sys.argv[1].split(",")]
L = fb.lib()
L.ffb_tc_set_trace.argtypes = [C.c_void_p]
buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
rng = np.random.default_rng(0)
f = np.zeros((N, K + 4), np.float32); f[:, :K] = rng.standard_normal((N, K)); f[:, K] = 1
op = fb.ConvOp(f, K, 1, 0, 1, 1, N, 2, pw_mode=2)
x = torch.randn((n, ih, iw, K), device="cuda"); y = torch.empty((n, ih, iw, (N + 3) & ~3), device="cuda")
class P:
    def __init__(s, t): s.ptr = t.data_ptr()
for _ in range(2): op.run(P(x), P(y), n, ih, iw)
torch.cuda.synchronize()
L.ffb_tc_set_trace(buf.data_ptr())
op.run(P(x), P(y), n, ih, iw); torch.cuda.synchronize()
L.ffb_tc_set_trace(None)
t = buf.cpu().numpy().reshape(6, 64, 4)
t0 = t[t > 0].min()
names = ["producer: empty-ok", "mma: A-ready, acc-free, committed", "split: enter, landed, done", "epilogue: enter, acc-ready, done"]
print("kernel entry %d, setup done %d, all roles done %d (cycles, CTA 0)" % tuple(int(v - t0) if v else -1 for v in t[5, 1, :3]))
for it in list(range(0, 3)) + list(range(3, 8)):
    row = []
    for r in range(4):
        row.append(" ".join("%6d" % (v - t0) if v else "     -" for v in t[r, it, :3]))
    inner = " ".join("%6d" % (v - t0) if v else "     -" for v in list(t[4, it, :4]) + [t[5, it + 2, 3] if it + 2 < 64 else 0])
    print("it %2d | P %s | M %s | S %s | E %s | Einner(chunk0: store-slot free, tmem loaded; chunk1: free, loaded; stores issued) %s" % (it, row[0][:6], row[1], row[2], row[3], inner))
d = np.diff(t[3, 3:40, 2]); print("cycles per tile (epilogue-done to epilogue-done), median:", np.median(d[d > 0]))
