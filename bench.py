#!/usr/bin/env python
"""bench.py -- frames/s of yolo-fastest-1.1 at 320x320, batch 256 per GPU (BASELINE.json's metric), on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      times the reference's own CPU path (oracle/_ref)

One "step" = one pass of the hot path over one batch of 256 synthetic frames per GPU: net_input fused into the stem
kernel, the 131-layer loop as 43 kernel launches (CUDA graph replay; 23 expand->depthwise->project chains and the SPP
block run as one fused kernel each) and the yolo candidate-filter kernels.  `value` is measured with the u8 frames
already resident in HBM (4 distinct batches rotated, 314 MB > L2); `e2e` goes through the public C-ABI calls
ffb_submit_u8 / ffb_collect with pinned HOST frames in and decoded boxes out (H2D + kernels + D2H + host decode/NMS
inside the timed region).  `roofline` charges a fused kernel the summed algorithmic bytes of the layers it replaces
(SURVEY 8d defines bytes per layer), so a fraction above 1 means faster than those layers could run unfused.  Multi-GPU: frames are independent -> contiguous shards per rank, no data-path collective;
the only traffic is one NCCL broadcast of the packed weights at load (weak scaling, 256 frames per GPU).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

BATCH = 256
NET_W = NET_H = 320
PITCH = 960
ALG_BYTES_PER_FRAME = 60.11e6          # unfused fp32 activations + weights, every layer, per frame
METRIC = "frames/sec yolo-fastest-1.1 320x320 batch256"
WORKLOAD = "yolo-fastest-1.1.cfg 320x320 batch=%d fp32 per GPU, all layers (net_input + 131-layer forward + yolo filter)"


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md: clocks.sm, clocks.max.sm and the
    clocks_event_reasons nvidia-smi prints).  Read in-process through NVML -- the library nvidia-smi itself is a client of --
    every 100 ms: a forked `nvidia-smi` per sample, and also one background `nvidia-smi -lms` process, held driver locks
    long enough to slow the host-driven e2e loop by 8-16 % (profiles/r1n_notes.txt).  BENCH_CLOCKS=smi uses the recipe's
    background `nvidia-smi -lms 200` process instead, BENCH_CLOCKS=off samples once before and once after."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.stop_flag = index, uuid, threading.Event()
        self.mode = os.environ.get("BENCH_CLOCKS", "nvml")
        self.sm, self.mask, self.max_mhz, self.proc, self.log, self.nv, self.h = [], 0, None, None, None, None, None

    def begin(self):
        if self.mode == "smi":
            import tempfile
            self.log = tempfile.TemporaryFile(mode="w+")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                                         stdout=self.log, stderr=subprocess.DEVNULL)
            return
        import pynvml as nv
        nv.nvmlInit()
        self.nv = nv
        try:
            self.h = nv.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid else nv.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        self.sample()
        if self.mode != "off":
            self.start()

    def sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def run(self):
        while not self.stop_flag.wait(0.1):
            try:
                self.sample()
            except Exception:
                pass

    def summary(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.log.seek(0)
            rows = [[c.strip() for c in line.split(",")] for line in self.log.read().splitlines() if line.count(",") >= 7]
            self.log.close()
            self.sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
            mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
            self.max_mhz = mx[0] if mx else None
            reasons = sorted({name for r in rows for (name, _), v in zip(self.REASONS, r[4:8]) if v.lower().startswith("active")})
        else:
            self.stop_flag.set()
            if self.is_alive():
                self.join(timeout=2)
            if self.nv is not None:
                try:
                    self.sample()
                except Exception:
                    pass
            reasons = sorted(name for name, bit in self.REASONS if self.mask & bit)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "source": "nvidia-smi -lms 200" if self.proc is not None else "nvml, in-process, every 100 ms" if self.mode != "off" else "nvml, before and after"}


FFMA_PEAK_TFLOPS = 74.4                # 148 SMs x 128 lanes x 2 flop x 1.965 GHz (nominal)


def parity_check(net, fb, synth, np):
    """In-run correctness (every rank, after the weight broadcast, default plan): the 4 seeded random frames (set S1) and 4
    picture-derived frames (set S2) whose reference outputs are committed under tests/golden/synth_320.npz (written by the
    compiled, unmodified reference -- tests/golden/make_golden.py).  S1: head checksums and candidate counts; S2: candidate
    classes and final boxes within the contract (1e-4 px, 5e-6 score).  Returns a dict with "ok"."""
    g = np.load(os.path.join(REPO, "tests", "golden", "synth_320.npz"))
    raw = np.fromfile(os.path.join(fb.ASSETS, "test.bmp"), np.uint8)
    bw, bh = int(raw[18:22].view("<u4")[0]), int(raw[22:26].view("<u4")[0])
    bp = (bw * 3 + 3) & ~3
    img = np.ascontiguousarray(raw[54:54 + bp * bh].reshape(bh, bp)[::-1])
    s2 = synth.shifted_frames_from(img, bw, bh, 20, NET_W, NET_H).reshape(20, NET_H, PITCH)
    s2_ids = (0, 3, 7, 19)
    frames = np.concatenate([synth.frames_u8(4, NET_W, NET_H), s2[list(s2_ids)]], axis=0)
    res = {"ok": True, "frames": 8, "max_box_px": 0.0, "max_score": 0.0, "max_head_rel": 0.0, "fail": []}
    for _ in range(2):                                   # eager pass + graph replay
        net.detect_batch_u8(frames, 8, NET_W, NET_H, PITCH)
    for f in range(4):
        for hid in (120, 129):
            o = net.layer_output(hid, f)
            tol = 2e-5 * float(g["s1_f%d_v6_O2_maxabs" % f][hid]) * o.size
            d = abs(float(o.astype(np.float64).sum()) - float(g["s1_f%d_v6_O2_sum" % f][hid]))
            res["max_head_rel"] = max(res["max_head_rel"], d / (float(g["s1_f%d_v6_O2_maxabs" % f][hid]) * o.size))
            if not d <= tol:
                res["fail"].append("s1 frame %d head %d checksum off by %.3g" % (f, hid, d))
        if len(net.boxes(f, raw=True)) != len(g["s1_f%d_v6_O2_raw" % f]):
            res["fail"].append("s1 frame %d candidate count" % f)
    for k, f in enumerate(s2_ids):
        want_raw, want = g["s2_f%d_raw" % f], g["s2_f%d_final" % f]
        graw, got = net.boxes(4 + k, raw=True), net.boxes(4 + k)
        if len(graw) != len(want_raw) or [int(t) for t in graw["type"]] != [int(t) for t in want_raw["type"]] or len(got) != len(want):
            res["fail"].append("s2 frame %d candidate set" % f)
            continue
        for a, b in zip(got, want):
            dp = max(abs(float(a[c]) - float(b[c])) for c in ("x1", "y1", "x2", "y2"))
            ds = abs(float(a["score"]) - float(b["score"]))
            res["max_box_px"] = max(res["max_box_px"], dp); res["max_score"] = max(res["max_score"], ds)
            if int(a["type"]) != int(b["type"]) or dp > 1e-4 or ds > 5e-6:
                res["fail"].append("s2 frame %d box off by %.3g px / %.3g score" % (f, dp, ds))
    res["ok"] = not res["fail"]
    res["fail"] = res["fail"][:4]
    return res


def launch_table(net, lt, B, peak_gbs, tf32_peak):
    """One row per kernel launch slot of the plan: the layers it covers, its unfused algorithmic bytes, the bytes
    a fused kernel MUST move (input of its first layer + output of its last + the weights, each once), its pointwise (tensor
    pipe) and stencil/stem (FFMA) flops, and the floor time  max(must-move bytes / HBM, 3 x pointwise flops / measured
    tcgen05 kind::tf32 peak [FFMA peak when the kernel computes them with FFMAs], FFMA flops / 74.4 TFLOP/s)."""
    L = net.layer_num
    names = [net.layer_cost(i)[2] for i in range(L)]
    rows = []
    for i in range(L):
        if not lt[i] > 0:
            continue
        cover = [i]
        if names[i].startswith("block_"):
            k = i + 1
            while k < L and names[k] == "in_block":
                cover.append(k); k += 1
        elif names[i] == "spp_fused":
            cover = [k for k in range(L) if names[k] == "in_spp"] + [i]
        alg = sum(net.layer_cost(k)[0] for k in ([i] if len(cover) > 1 else cover))      # a fused slot already reports the sum
        fl_pw = fl_ffma = wbytes = 0.0
        for k in cover:
            a = net.layer(k)
            if a.type != 0:
                continue
            b = net.layer(k + 1)
            taps = a.fs * a.fs * (a.c // a.groups)
            wbytes += 4.0 * a.fn * (taps + 2)
            fl = 2.0 * taps * b.w * b.h * b.c
            if a.fs == 1 and a.groups == 1:
                fl_pw += fl
            else:
                fl_ffma += fl
        first, last = net.layer(min(cover)), net.layer(max(cover) + 1)
        if names[i] == "spp_fused":
            first = net.layer(min(cover))
        must = 4.0 * (first.w * first.h * first.c + last.w * last.h * last.c) + wbytes
        if len(cover) == 1:
            must = alg
        tensor = names[i].startswith("block_mma") or "tcgen05" in names[i]
        t_hbm = must * B / (peak_gbs * 1e9)
        t_pw = (3.0 * fl_pw * B / (tf32_peak * 1e12)) if tensor else fl_pw * B / (FFMA_PEAK_TFLOPS * 1e12)
        t_ffma = fl_ffma * B / (FFMA_PEAK_TFLOPS * 1e12)
        floor_ms = 1e3 * max(t_hbm, t_pw, t_ffma)
        bound = "hbm" if t_hbm >= max(t_pw, t_ffma) else "tensor" if t_pw >= t_ffma and tensor else "ffma"
        rows.append({"slot": i, "kernel": names[i], "layers": len(cover), "ms": float(lt[i]), "alg_bytes": alg * B, "must_bytes": must * B,
                     "pw_flops": fl_pw * B, "ffma_flops": fl_ffma * B, "floor_ms": floor_ms, "bound": bound})
    return rows


def microbench(fb, torch, np, peak_gbs, tf32_peak, stream):
    """BASELINE configs 3 and 4 at their full sizes, CUDA events on the launching stream, 3 warm-up + 10 timed
    launches each; inputs far larger than L2 (10 GB / 1.26 GB)."""
    out = {}
    rng = np.random.default_rng(5)

    def packed(fn, k):
        row = ((k + 3) & ~3) + 4
        f = np.zeros((fn, row), np.float32)
        f[:, :k] = rng.standard_normal((fn, k)) / np.sqrt(k)
        f[:, row - 4] = rng.uniform(0.5, 1.5, fn); f[:, row - 3] = rng.uniform(-0.5, 0.5, fn)
        return f

    def timed(op, x, y, n, h, w, reps=10):
        for _ in range(3):
            op.run_ptr(x.data_ptr(), y.data_ptr(), n, h, w, stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            op.run_ptr(x.data_ptr(), y.data_ptr(), n, h, w, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # config 3: 3x3 depthwise, 160x160x96, batch 1024, leaky
    n, h, w, c = 1024, 160, 160, 96
    x = torch.empty((n, h, w, c), dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    op = fb.ConvOp(packed(c, 9), c, c, 1, 1, 3, c, 2)
    ms = timed(op, x, y, n, h, w)
    alg = 2.0 * n * h * w * c * 4 + c * 11 * 4
    out["dw3x3_160x160x96_b1024"] = {"ms": ms, "kernel": op.kernel, "GBps": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak_gbs,
                                    "alg_bytes": alg, "bound": "hbm"}
    op.close(); del x, y
    torch.cuda.empty_cache()
    # config 4: 1x1 pointwise 40x40, 192 -> 192, batch 1024, leaky
    n, h, w, c = 1024, 40, 40, 192
    x = torch.empty((n, h, w, c), dtype=torch.float32, device="cuda").normal_()
    y = torch.empty_like(x)
    op = fb.ConvOp(packed(c, c), c, 1, 0, 1, 1, c, 2)
    ms = timed(op, x, y, n, h, w)
    M = n * h * w
    alg = M * (c + c) * 4.0 + c * (c + 2) * 4.0
    fl = 2.0 * M * c * c
    passes = 1 if "1xtf32" in op.kernel else 3
    out["pw_192x192_40x40_b1024"] = {"ms": ms, "kernel": op.kernel, "mode": "%dxTF32" % passes, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak_gbs,
                                     "tensor_frac": passes * fl / (ms * 1e-3) / 1e12 / tf32_peak, "tflops_useful": fl / (ms * 1e-3) / 1e12,
                                     "alg_bytes": alg, "flops": fl, "bound": "hbm" if alg / (peak_gbs * 1e9) > passes * fl / (tf32_peak * 1e12) else "tensor"}
    op.close(); del x, y
    torch.cuda.empty_cache()
    return out


def run_reference(procs: int, frames_per_proc: int):
    """The reference's own CPU implementation of the path: oracle/_ref/ffcnn_ref_bench (unmodified ffcnn.c + conv-v6.c,
    build.sh flags), P independent single-threaded processes.  Returns (fps, description)."""
    import ffcnn_b200 as fb
    cfg, wts = fb.default_model()
    last = None
    for exe in ("ffcnn_ref_bench", "ffcnn_ref_bench_v3"):        # -march=native build first, portable x86-64-v3 if it cannot run here
        path = os.path.join(REPO, "oracle", "_ref", exe)
        if not os.path.exists(path):
            continue
        r = subprocess.run([path, cfg, wts, str(procs), str(frames_per_proc)], capture_output=True, text=True)
        last = r
        if r.returncode == 0 and "fps=" in r.stdout:
            fps = float(r.stdout.strip().split("fps=")[1])
            return fps, f"{exe}: {procs} procs x {frames_per_proc} frames (conv-v6, build.sh flags, 320x320 synthetic u8 frames)"
    raise RuntimeError("oracle/_ref/ffcnn_ref_bench unavailable or failed: " + (last.stderr[-300:] if last else "not built"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also print the per-layer roofline table to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, args.warmup)
    K = max(1, args.steps)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        per_proc = 12
        for _ in range(max(1, min(W, 2))):
            run_reference(cores, 2)
        t0 = time.time()
        vals = [run_reference(cores, per_proc) for _ in range(K)]
        wall = time.time() - t0
        fps = sum(v[0] for v in vals) / len(vals)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * cores * per_proc / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD % args.batch, "cpu_step": "%d single-threaded processes x %d frames" % (cores, per_proc)},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": vals[0][1]},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": wall}
        print(json.dumps(line))
        return

    # stdout carries exactly one JSON line: whatever libraries print while the job runs (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    import ffcnn_b200 as fb
    from ffcnn_b200 import synth, shard

    if not torch.cuda.is_available() or fb.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    cpus_before = os.sched_getaffinity(0)
    near = None
    if os.environ.get("FFCNN_BENCH_NUMA", "0") == "1":      # opt-in until measured at 8 GPUs: allocate pinned frames next to the GPU
        near = shard.bind_near_gpu(local)
        print("bench.py rank %d: bound to GPU-local CPUs %s" % (rank, near), file=sys.stderr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    cfg, wts = fb.default_model()
    # N > 1: the end-to-end leg may give a rank up to 1.5 x its equal shard (rate-proportional shards, see run_e2e below)
    balance = world > 1 and os.environ.get("FFCNN_E2E_BALANCE", "1") == "1"
    BE = (B * 3 // 2 + 7) // 8 * 8 if balance else B
    net = fb.Net(cfg, wts if rank == 0 else None, 0, 0, device=local, max_batch=BE)
    stream = torch.cuda.Stream()            # a real (non-legacy) stream shared by torch's events and the library's launches
    torch.cuda.set_stream(stream)
    net.set_stream(stream.cuda_stream)
    bcast_bytes = 0
    if world > 1:
        bcast_bytes = shard.broadcast_weights(net, dist, device=torch.device("cuda", local))

    pc = parity_check(net, fb, synth, np)
    pc_all = torch.tensor([1.0 if pc["ok"] else 0.0, pc["max_box_px"]], dtype=torch.float64, device="cuda")
    pc_min, pc_max = pc_all.clone(), pc_all.clone()
    if world > 1:
        dist.all_reduce(pc_min, op=dist.ReduceOp.MIN); dist.all_reduce(pc_max, op=dist.ReduceOp.MAX)
    if not pc["ok"]:
        print("bench.py rank %d: PARITY CHECK FAILED: %s" % (rank, pc["fail"]), file=sys.stderr)
    if float(pc_min[0]) < 1.0:                                  # some rank computes wrong results: no throughput number for wrong answers
        if rank == 0:
            os.dup2(saved_stdout, 1)
            print(json.dumps({"metric": METRIC, "value": None, "unit": "frames/s", "n_gpus": world, "parity_check": {"ok": False, "rank0": pc}}), flush=True)
        raise SystemExit(3)
    pc["ranks_ok"] = world
    pc["max_box_px_all_ranks"] = float(pc_max[1])

    # synthetic frames: 4 distinct resident batches per rank (seeded per global frame index), rotated across steps
    NB = 4
    lo, _ = shard.shard_range(B * world, rank, world)
    if os.environ.get("BENCH_PINNED", "") == "wc":          # developer knob: write-combined pinned frames (profiles/r2h_e2e_8gpu.txt)
        wc_ptr = fb.lib().ffb_host_alloc_pinned_wc(NB * BE * NET_H * PITCH)
        if not wc_ptr:
            raise SystemExit("bench.py: write-combined pinned allocation failed")
        hv = np.ctypeslib.as_array((fb.C.c_uint8 * (NB * BE * NET_H * PITCH)).from_address(wc_ptr)).reshape(NB, BE, NET_H, PITCH)

        class _HostView:                                  # host[i].data_ptr() as the pinned torch tensor offers it
            def __init__(self, base, stride): self.base, self.stride = base, stride
            def __getitem__(self, i):
                p = self.base + i * self.stride
                return type("P", (), {"data_ptr": staticmethod(lambda p=p: p)})
        host = _HostView(wc_ptr, BE * NET_H * PITCH)
        host_t = None
    else:
        host = torch.empty((NB, BE, NET_H, PITCH), dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        host_t = host
    base = synth.frames_u8(16, NET_W, NET_H, seed0=0xFFC0 + 16 * rank)
    for b in range(NB):
        for f in range(BE):
            hv[b, f] = base[(b * 5 + f) % 16]
            hv[b, f, f % NET_H, :8] = (lo + f + b) & 0xFF          # every frame distinct
    dev = torch.from_numpy(np.ascontiguousarray(hv[:, :B])).cuda(non_blocking=False)      # the resident leg: equal shards of B frames
    frame_bytes = B * NET_H * PITCH

    def step_resident(i):
        net.input_u8(dev[i % NB].data_ptr(), B, NET_W, NET_H, PITCH, on_device=True)
        net.forward()
        net.detect_enqueue()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step_resident(i)
    net.detect_finish()
    launches_per_step = (0 if net.get_option("input_fused") == 1 else 1) + net.launches_per_forward() + 2
    sampler = ClockSampler(local, "GPU-" + str(torch.cuda.get_device_properties(local).uuid) if hasattr(torch.cuda.get_device_properties(local), "uuid") else None)
    if rank == 0:
        try:
            sampler.begin()
        except Exception as ex:
            print("bench.py: clock sampler unavailable: %s" % ex, file=sys.stderr)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(K):
        step_resident(i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    net.detect_finish()

    # end to end through the public calls: pinned host frames -> decoded boxes on the host, every batch's H2D copy and
    # D2H read inside the timed region (ffb_submit_u8 / ffb_collect: the copy of batch i+1 overlaps the work on batch i)
    def run_e2e(steps, n=B):
        moved = 0
        net.submit_u8(host[0].data_ptr(), n, NET_W, NET_H, PITCH)
        if steps > 1:
            net.submit_u8(host[1 % NB].data_ptr(), n, NET_W, NET_H, PITCH)
        for i in range(steps):                  # two batches stay queued behind the one being collected (three in flight at most)
            if i + 2 < steps:
                net.submit_u8(host[(i + 2) % NB].data_ptr(), n, NET_W, NET_H, PITCH)
            net.collect()
            moved += net.last_d2h_bytes()
        return moved

    def timed_e2e(steps, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.time()
        e0.record(stream)
        moved = run_e2e(steps, n)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), time.time() - t_wall, moved

    run_e2e(W)                      # W untimed warm-up steps of the same pipelined path (copy stream, staging slots, graph)
    KE = K
    ms_e2e, wall_e2e, d2h = timed_e2e(KE, B)                    # equal shards: every rank B frames per step
    n_e2e, shards, equal_e2e = B, None, None
    if balance:
        # Rate-proportional shards (ffcnn_b200/shard.py::weighted_shards): on a box whose GPUs do not get equal shares of the host's
        # memory path, equal shards run at the pace of the slowest copy.  Each rank's measured end-to-end rate with everybody
        # running (the equal-shard loop just timed) sizes its contiguous shard of the same world x B frames; frames stay where
        # they are, nothing goes on the wire.  Both results are reported: `e2e` = proportional shards, `e2e.equal_shards`.
        rate = torch.tensor([B * KE / ms_e2e], dtype=torch.float64, device="cuda")
        rates = [torch.zeros_like(rate) for _ in range(world)]
        dist.all_gather(rates, rate)
        shards = shard.weighted_shards(B * world, [float(r[0]) for r in rates], quantum=8, max_per_rank=BE)
        n_e2e = shards[rank][1] - shards[rank][0]
        t_eq = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_eq, op=dist.ReduceOp.MAX)
        equal_e2e = {"value": world * B * KE / (float(t_eq[0]) * 1e-3), "unit": "frames/s", "ms_per_step": float(t_eq[0]) / KE,
                     "ms_per_step_rank0": ms_e2e / KE}
        run_e2e(max(4, W // 2), n_e2e)                          # graph + staging for the new batch size (at least one pass over every pipeline slot)
        # one refinement: the rates move when the shares do (a rank that copies less leaves bandwidth to its neighbours)
        ms_c, _, _ = timed_e2e(max(5, KE // 3), n_e2e)
        rate = torch.tensor([n_e2e * max(5, KE // 3) / ms_c], dtype=torch.float64, device="cuda")
        dist.all_gather(rates, rate)
        shards = shard.weighted_shards(B * world, [float(r[0]) for r in rates], quantum=8, max_per_rank=BE)
        if shards[rank][1] - shards[rank][0] != n_e2e:
            n_e2e = shards[rank][1] - shards[rank][0]
        run_e2e(max(4, W // 2), n_e2e)                          # every rank takes the same path whether or not its own share moved
        ms_e2e, wall_e2e, d2h = timed_e2e(KE, n_e2e)
    nboxes = sum(len(net.boxes(f)) for f in range(n_e2e))
    clocks = sampler.summary() if rank == 0 else None

    pic_result = None
    if world == 1 and os.environ.get("BENCH_PICTURE", "1") == "1":
        # the same end-to-end loop on set S2 of SURVEY 8(d): frames derived from test.bmp, so every frame yields candidates
        # and the host decode + NMS inside ffb_collect has real work (the seeded random frames above produce none).
        # Reported beside the headline e2e, never instead of it; a failure here must not lose the line.
        try:
            raw = np.fromfile(os.path.join(fb.ASSETS, "test.bmp"), np.uint8)
            bw, bh = int(raw[18:22].view("<u4")[0]), int(raw[22:26].view("<u4")[0])
            bp = (bw * 3 + 3) & ~3
            img = np.ascontiguousarray(raw[54:54 + bp * bh].reshape(bh, bp)[::-1])
            pic = synth.shifted_frames_from(img, bw, bh, B, NET_W, NET_H)
            host_pic = torch.empty((2, B, NET_H, PITCH), dtype=torch.uint8).pin_memory()
            hp = host_pic.numpy()
            hp[0] = pic.reshape(B, NET_H, PITCH); hp[1] = pic[::-1].reshape(B, NET_H, PITCH)

            def run_pic(steps):
                moved = 0
                net.submit_u8(host_pic[0].data_ptr(), B, NET_W, NET_H, PITCH)
                if steps > 1:
                    net.submit_u8(host_pic[1].data_ptr(), B, NET_W, NET_H, PITCH)
                for i in range(steps):
                    if i + 2 < steps:
                        net.submit_u8(host_pic[(i + 2) % 2].data_ptr(), B, NET_W, NET_H, PITCH)
                    net.collect()
                    moved += net.last_d2h_bytes()
                return moved

            run_pic(W)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            moved = run_pic(K)
            p1.record(stream)
            torch.cuda.synchronize()
            ms_pic = p0.elapsed_time(p1)
            pic_result = {"value": B * K / (ms_pic * 1e-3), "unit": "frames/s", "ms_per_step": ms_pic / K,
                          "d2h_bytes_per_step": moved // K, "boxes_last_batch": sum(len(net.boxes(f)) for f in range(B)),
                          "frames": "test.bmp fitted to 320x320, rolled by (frame mod 16) pixels: every frame has candidates to decode"}
        except Exception as ex:
            pic_result = {"error": str(ex)[:200]}

    print("bench.py rank %d/%d: resident %.3f ms/step, e2e %.3f ms/step (wall %.3f ms/step)" % (rank, world, ms / K, ms_e2e / KE, 1e3 * wall_e2e / KE), file=sys.stderr)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = peaks()
        try:
            tf32_peak = fb.measure_tf32_peak()
            tf32_src = "measured live: ffb_measure_tf32_peak (tcgen05.mma.kind::tf32 M128 N256 K8 from shared memory on every SM)"
        except Exception as ex:
            tf32_peak, tf32_src = 1125.0, "nominal (half the 2.25 PFLOP/s bf16 figure); live measurement failed: %s" % ex
        value = world * B * K / (ms * 1e-3)
        e2e = world * B * KE / (ms_e2e * 1e-3)
        # per-layer CUDA-event timings (same batch), grouped by kernel: the dominant kernel's roofline
        lt = net.layer_times(reps=5)
        groups = {}
        for i in range(net.layer_num):
            by, fl, name = net.layer_cost(i)
            if lt[i] > 0:
                g = groups.setdefault(name, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "layers": 0})
                g["ms"] += float(lt[i]); g["bytes"] += by * B; g["flops"] += fl * B; g["layers"] += 1
        top = max(groups, key=lambda k: groups[k]["ms"])
        tg = groups[top]
        achieved = tg["bytes"] / (tg["ms"] * 1e-3) / 1e9
        total_ms = sum(g["ms"] for g in groups.values())
        if args.layers:
            for i in range(net.layer_num):
                by, fl, name = net.layer_cost(i)
                if lt[i] > 0:
                    print("L%-3d %-16s %8.4f ms %8.1f GB/s  %5.1f%% of HBM peak" % (i, name, lt[i], by * B / (lt[i] * 1e-3) / 1e9,
                                                                                  100 * by * B / (lt[i] * 1e-3) / 1e9 / peak), file=sys.stderr)
        rows = launch_table(net, lt, B, peak, tf32_peak)
        fused_by = {}
        for r in rows:
            g = fused_by.setdefault(r["kernel"], {"ms": 0.0, "floor_ms": 0.0, "must_bytes": 0.0, "alg_bytes": 0.0, "launches": 0, "bounds": {}})
            g["ms"] += r["ms"]; g["floor_ms"] += r["floor_ms"]; g["must_bytes"] += r["must_bytes"]; g["alg_bytes"] += r["alg_bytes"]; g["launches"] += 1
            g["bounds"][r["bound"]] = g["bounds"].get(r["bound"], 0) + 1
        floor_total = sum(r["floor_ms"] for r in rows)
        must_total = sum(r["must_bytes"] for r in rows)
        if args.layers:
            for r in rows:
                print("slot L%-3d %-22s %2d layers %8.4f ms  floor %7.4f ms (%s)  %5.1f%% of its floor   must-move %7.1f MB (alg %7.1f MB)" %
                      (r["slot"], r["kernel"], r["layers"], r["ms"], r["floor_ms"], r["bound"], 100 * r["floor_ms"] / r["ms"], r["must_bytes"] / 1e6, r["alg_bytes"] / 1e6), file=sys.stderr)
        traffic = None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
            if top in tj:
                traffic = dict(tj[top])
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % B,
                       "frames_per_gpu_per_step": B, "l2": "inputs rotate over %d distinct resident batches (%d MB u8) and the activation arena (%d MB) exceeds L2"
                       % (NB, NB * frame_bytes >> 20, net.get_option("arena_mb")),
                       "parallelism": "dp%d: contiguous frame shards, no data-path collective; weights broadcast once over NCCL (%d B)" % (world, bcast_bytes),
                       "pw_mode": net.get_option("pw_mode"), "weights": "yolo-fastest-1.1.weights" if os.path.exists(wts) else "zero"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": frame_bytes, "d2h_bytes_per_step": d2h // KE,
                    "ms_per_step": ms_e2e / KE, "api": "ffb_submit_u8 + ffb_collect (pinned host u8 frames in, decoded+NMS boxes out; copies of the next two batches overlapped)", "boxes_last_batch": nboxes},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["traffic"] if traffic else None, "traffic_of": traffic, "peak_source": peak_src,
                         "accounting": "achieved = algorithmic bytes (SURVEY 8d, per layer, unfused) of the layers the kernel runs / its time; a fused block kernel is "
                                       "charged the sum over the layers it replaces, its real DRAM traffic (traffic_of) is far lower -- the expanded tensors never reach HBM", "kernel_share_of_step": tg["ms"] / total_ms, "layers_in_kernel": tg["layers"],
                         "whole_graph": {"achieved": ALG_BYTES_PER_FRAME * B * K / (ms * 1e-3) / 1e9 / 1, "frac": ALG_BYTES_PER_FRAME * B * K / (ms * 1e-3) / 1e9 / peak,
                                         "alg_bytes_per_frame": ALG_BYTES_PER_FRAME},
                         "fused": {"definition": "per launch: floor = max(bytes the launch must move (input of its first layer + output of its last + weights, "
                                                 "each once) / HBM peak, 3 x pointwise flops / tf32 tensor peak (FFMA peak for FFMA kernels), stencil+stem flops / FFMA peak); "
                                                 "frac = floor / measured time.  This is the roofline of the kernels as fused; `frac` above is against the UNFUSED per-layer bytes of SURVEY 8d",
                                   "kernel": top, "frac": fused_by[top]["floor_ms"] / fused_by[top]["ms"], "floor_ms": fused_by[top]["floor_ms"], "ms": fused_by[top]["ms"],
                                   "tf32_peak_tflops": tf32_peak, "tf32_peak_source": tf32_src, "ffma_peak_tflops": FFMA_PEAK_TFLOPS,
                                   "whole_graph": {"floor_ms": floor_total, "ms_per_step": ms / K, "frac": floor_total / (ms / K), "must_move_MB_per_frame": must_total / B / 1e6,
                                                   "fps_at_floor": B / (floor_total * 1e-3)},
                                   "by_kernel": {k: {"ms": round(v["ms"], 4), "floor_ms": round(v["floor_ms"], 4), "frac": round(v["floor_ms"] / v["ms"], 3),
                                                     "launches": v["launches"], "bounds": v["bounds"]} for k, v in sorted(fused_by.items())}},
                         "by_kernel": {k: {"ms": round(v["ms"], 4), "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                                           "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 3), "layers": v["layers"]} for k, v in sorted(groups.items())}},
        }
        if shards is not None and equal_e2e is not None and equal_e2e["value"] > e2e:
            # the calibration did not pay on this box / this run (e.g. a world whose ranks share the host path equally and whose rates
            # differ only by noise): both loops were measured the same way through the same calls, the faster one is the headline
            line["e2e"]["proportional_shards"] = {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / KE, "shards": [hi - lo for lo, hi in shards]}
            line["e2e"]["value"] = equal_e2e["value"]; line["e2e"]["ms_per_step"] = equal_e2e["ms_per_step"]
            line["e2e"]["sharding"] = "equal shards of %d frames (rate-proportional shards measured slower in this run: proportional_shards)" % B
            line["e2e"]["equal_shards"] = equal_e2e
            shards = None
        if shards is not None:
            sizes = [hi - lo for lo, hi in shards]
            line["e2e"]["shards"] = sizes
            line["e2e"]["h2d_bytes_per_step_by_rank"] = [n * NET_H * PITCH for n in sizes]
            line["e2e"]["sharding"] = ("the same %d frames per step, contiguous shards sized in proportion to each rank's measured end-to-end rate "
                                       "(ffcnn_b200/shard.py::weighted_shards; the GPUs of this box do not share the host memory path equally); "
                                       "h2d_bytes_per_step is the mean over ranks; equal_shards = the same loop with %d frames on every rank" % (B * world, B))
            line["e2e"]["equal_shards"] = equal_e2e
        line["parity_check"] = pc
        if pic_result is not None:
            line["e2e"]["picture_frames"] = pic_result
        if world == 1 and os.environ.get("BENCH_MICRO", "1") == "1":
            net.close()                                   # the microbench tensors (20 GB) want the arena's room on smaller boxes
            try:
                line["microbench"] = microbench(fb, torch, np, peak, tf32_peak, stream)
            except Exception as ex:                      # never a reason to lose the line
                line["microbench"] = {"error": str(ex)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, cpus_before)         # the CPU baseline uses every host core
            try:
                fps, desc = run_reference(cores, 40)
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": desc}
            except Exception as ex:                      # the baseline is a reported number, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "failed: %s" % ex}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    net.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
