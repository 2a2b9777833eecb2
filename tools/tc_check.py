"""Developer diagnostic (GPU box): tcgen05 pointwise kernel vs fp64 on a range of (K, N, M) shapes, both math modes."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import ffcnn_b200 as fb

def tf32_trunc(x): return (x.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
def tf32_rna(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x1000
    return (u.astype(np.uint32) & np.uint32(0xffffe000)).view(np.float32)

shapes = [(32, 32, 1, 16, 16), (32, 16, 1, 16, 16), (96, 16, 2, 40, 40), (16, 96, 2, 40, 40), (48, 16, 1, 40, 40), (24, 136, 2, 20, 20), (136, 24, 2, 20, 20),
          (224, 48, 3, 10, 10), (48, 224, 3, 10, 10), (192, 96, 3, 10, 10), (96, 255, 3, 10, 10), (120, 120, 2, 20, 20), (120, 255, 2, 20, 20),
          (192, 192, 3, 40, 40), (8, 32, 1, 20, 20), (32, 8, 1, 20, 20), (96, 96, 5, 13, 7)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
rng = np.random.default_rng(5)
for mode in (3, 2):
    for (K, N, n, h, w) in shapes:
        row = K + 4
        f = np.zeros((N, row), np.float32); f[:, :K] = rng.standard_normal((N, K)) / np.sqrt(K)
        f[:, K] = rng.uniform(0.5, 1.5, N); f[:, K + 1] = rng.uniform(-0.5, 0.5, N)
        x = rng.standard_normal((n, h, w, K)).astype(np.float32)
        try:
            op = fb.ConvOp(f, K, 1, 0, 1, 1, N, 2, pw_mode=mode)
            t = time.time(); y = op(x); dt = time.time() - t
        except fb.FfcnnError as e:
            print("mode", mode, (K, N), "ERROR", e); continue
        W = f[:, :K]
        def ref(xx, ww):
            acc = xx.reshape(-1, K).astype(np.float64) @ ww.astype(np.float64).T
            v = acc * f[:, K].astype(np.float64) + f[:, K + 1]
            return np.where(v > 0, v, 0.1 * v).reshape(n, h, w, N)
        exact = ref(x, W)
        e_exact = np.abs(y - exact).max() / np.abs(exact).max()
        msg = "mode %d K=%3d N=%3d M=%5d %-18s rel_err %.2e" % (mode, K, N, n * h * w, op.kernel, e_exact)
        if mode == 3:
            et = np.abs(y - ref(tf32_trunc(x), tf32_trunc(W))).max() / np.abs(exact).max()
            er = np.abs(y - ref(tf32_rna(x), tf32_rna(W))).max() / np.abs(exact).max()
            msg += "  vs trunc-model %.2e  vs rna-model %.2e" % (et, er)
        print(msg, flush=True)
        op.close()
