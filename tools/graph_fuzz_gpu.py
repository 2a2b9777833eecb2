"""Developer diagnostic (GPU box): random darknet graphs (tests/cfg_fuzz.py) through the loader and the engine, every layer
against the oracle.  The CPU half of this (host loader and oracle against the compiled reference, bit-exact) is already a
test; this is the GPU half and has NOT been run yet (round 1 ended without GPU minutes) -- run it first in round 2:

    gpurun --timeout 300 -- 'timeout 240 python tools/graph_fuzz_gpu.py 1 40 > gpurun_out/graph_fuzz.txt 2>&1; tail -30 gpurun_out/graph_fuzz.txt'

usage: graph_fuzz_gpu.py [seed [cases]]     exit status = number of failing graphs (capped at 100)
Every graph runs in its own subprocess under a timeout, so a kernel that faults or hangs on an odd geometry costs one case,
not the run."""
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
FEAT_TOL = 2e-5


def one(cfg, wts, seed):
    import numpy as np
    import ffcnn_b200 as fb
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    layers = orc.load_net(cfg, wts, 0, 0)
    W, H = layers[0].w, layers[0].h
    n = 3
    pitch = (3 * W + 3) & ~3
    frames = rng.integers(0, 256, (n, H, pitch), dtype=np.uint8)
    worst = (0.0, -1, "")
    for keep in (1, 0):                                     # 1: layer-by-layer plan, every tensor readable; 0: the default fused plan
        net = fb.Net(cfg, wts, 0, 0, device=0, max_batch=n)
        net.set_option("keep_all", keep)
        for _ in range(2):                                  # second pass replays the CUDA graph
            net.detect_batch_u8(frames, n, W, H, pitch)
        for f in range(n):
            x, s1, s2 = orc.net_input(frames[f], W, H, W, H)
            outs, raw, fin = orc.forward(layers, x, s1, s2, True)
            if keep:
                for i, o in enumerate(outs):
                    if o is None or o.size == 0:
                        continue
                    got = net.layer_output(i, f)
                    err = float(np.abs(got - o).max() / max(1e-30, np.abs(o).max()))
                    if not (err < FEAT_TOL):
                        print("  FAIL layer %d (type %d) frame %d: rel err %.3e" % (i, layers[i].type, f, err))
                        return 1
                    if err > worst[0]:
                        worst = (err, i, "keep_all")
            graw = net.boxes(f, raw=True)
            if len(graw) != len(raw) or [int(t) for t in graw["type"]] != [int(t) for t in raw["type"]]:
                # a candidate whose confidence sits within rounding of the threshold may flip; report, do not fail, unless far off
                print("  note: frame %d keep %d: %d candidates vs %d in the oracle" % (f, keep, len(graw), len(raw)))
                if abs(len(graw) - len(raw)) > max(2, len(raw) // 200):
                    return 1
        net.close()
    print("  ok  worst rel err %.2e at layer %d" % (worst[0], worst[1]))
    return 0


def main():
    import numpy as np
    import cfg_fuzz
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        sys.exit(one(sys.argv[2], sys.argv[3], int(sys.argv[4])))
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rng = np.random.default_rng(seed)
    d = tempfile.mkdtemp()
    bad = 0
    for case in range(cases):
        text, convs, (W, H) = cfg_fuzz.gen(rng)
        cfg, wts = os.path.join(d, "g%d.cfg" % case), os.path.join(d, "g%d.weights" % case)
        with open(cfg, "w", newline="") as f:
            f.write(text)
        with open(wts, "wb") as f:
            f.write(cfg_fuzz.weights(rng, convs))
        kinds = [l[1:-1] for l in text.replace("\r", "").split("\n") if l.startswith("[")]
        print("graph %d: %dx%d, %d sections" % (case, W, H, len(kinds)), flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", cfg, wts, str(seed * 1000 + case)],
                               capture_output=True, text=True, timeout=60)
            print(r.stdout.rstrip())
            if r.returncode != 0:
                bad += 1
                print("  rc %d %s" % (r.returncode, r.stderr[-400:].strip()))
                keep = os.path.join(REPO, "gpurun_out", "graph_fuzz_bad_%d_%d.cfg" % (seed, case))
                os.makedirs(os.path.dirname(keep), exist_ok=True)
                with open(keep, "w", newline="") as f:
                    f.write(text)
        except subprocess.TimeoutExpired:
            bad += 1
            print("  TIMEOUT")
    print("graphs %d, failing %d" % (cases, bad))
    sys.exit(min(bad, 100))


if __name__ == "__main__":
    main()
