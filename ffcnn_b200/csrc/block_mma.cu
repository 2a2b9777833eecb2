/*
 * block_mma.cu -- planner, weight preparation and launcher of the fused inverted-residual block kernel
 * (block_mma.cuh).  The kernel is instantiated for the channel/stride combinations of yolo-fastest-1.1
 * (KS1 = ceil(cin/8), NT3 = ceil(cout/8)); other combinations report "unsupported" and the engine keeps
 * the three separate layers.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "block_mma.h"
#include "pw_tc.h"
#include "block_mma.cuh"
#include "block_ws.cuh"

#include "ffb_internal.h"

using namespace ffb;

struct BlkPlan {
    int cin, cexp, cout, S, H, W, OH, OW, res;
    int KS1, NT3, MTW, MINB, G, GC, NC, TH, TW, HH, HW, SEs, xrows, chunk_floats, occ;
    int XH, XW, xo, yo, frame;
    int tc, nmt, tmem_cols;               /* tcgen05 expand: m-tiles of the x tile, TMEM columns */
    int ws, R, XB, xbuf_floats;           /* warp-specialised kernel (block_ws.cuh): weight slots, x buffers, floats per x buffer */
    float slope1, sloped, slope3, slope_res;
    size_t smem;
    float *d_chunks, *d_sb3;
    int row1, rowd, row3;
    int num_sms;
    char desc[128];
};

typedef void (*BlkKernel)(const CUtensorMap, const BlkArgs);
struct BlkInst { int KS1, NT3, S, MTW, GC, MINB; BlkKernel fn; ffb_smem_cfg configured; int tc; };

#define INST(K, N, S, M, G, B) { K, N, S, M, G, B, k_block_mma<K, N, S, M, G, B>, {}, 0 }
#define INST_TC(K, N, S, M, G) { K, N, S, M, G, 2, k_block_mma<K, N, S, M, G, 2, true>, {}, 1 }
#define INST3(K, N, S, G) INST(K, N, S, 1, G, 2), INST(K, N, S, 2, G, 2), INST(K, N, S, 4, G, 2)
#define INST2(K, N, S, G) INST(K, N, S, 1, G, 2), INST(K, N, S, 2, G, 2)
static BlkInst g_inst[] = {
    INST3(1, 1, 1, 1), INST3(1, 1, 1, 2), INST3(1, 1, 1, 3),      /* 8->8->4, 4->8->4, 8->32->8, 8->48->8 */
    INST3(1, 1, 2, 1), INST3(1, 1, 2, 2),                         /* 4->24->8 s2, 8->32->8 s2 */
    INST3(1, 2, 1, 1), INST3(1, 2, 1, 3),                         /* 8->48->16 */
    INST3(2, 2, 1, 1), INST3(2, 2, 1, 2), INST3(2, 2, 1, 3),      /* 16->96->16 */
    INST2(2, 3, 2, 1), INST2(2, 3, 2, 2), INST2(2, 3, 2, 3),      /* 16->96->24 s2 */
    INST2(3, 3, 1, 1), INST2(3, 3, 1, 3),                         /* 24->136->24 */
    INST2(3, 6, 2, 1), INST2(3, 6, 2, 3),                         /* 24->136->48 s2 */
    INST2(6, 6, 1, 1), INST2(6, 6, 1, 2),                         /* 48->224->48 */
    /* expand GEMM on tcgen05 (x split hi/lo into TMEM as the A operand, W1 chunk as SWIZZLE_128B B sub-tiles, accumulators in
       TMEM one chunk ahead of the depthwise stage): the (tile, group) shapes the plan of yolo-fastest-1.1 uses.  Selected per
       block shape where it measured faster (tc_wanted below); FFCNN_BLK_TC=1 forces it wherever an instance exists, -1 disables it. */
    INST_TC(1, 1, 1, 2, 2), INST_TC(1, 1, 1, 2, 1), INST_TC(1, 1, 1, 4, 2), INST_TC(1, 1, 2, 1, 2), INST_TC(1, 1, 1, 2, 3), INST_TC(1, 2, 1, 2, 3),
    INST_TC(2, 2, 1, 2, 1), INST_TC(2, 2, 1, 2, 3), INST_TC(2, 2, 1, 1, 2), INST_TC(3, 3, 1, 2, 1), INST_TC(3, 3, 1, 1, 3),
    INST_TC(2, 2, 1, 2, 2), INST_TC(2, 3, 2, 1, 3), INST_TC(3, 3, 1, 2, 3), INST_TC(3, 6, 2, 1, 1), INST_TC(3, 6, 2, 1, 3), INST_TC(6, 6, 1, 1, 1), INST_TC(6, 6, 1, 1, 2),
};
#undef INST3
#undef INST2
#undef INST
#undef INST_TC

/* warp-specialised kernel (block_ws.cuh): one 16-channel group per chunk, 5 stage-B warps + 3 stage-A warps */
struct WsInst { int KS1, NT3, S, MTW; BlkKernel fn; ffb_smem_cfg configured; };
#define INST_WS(K, N, S, M) { K, N, S, M, k_block_ws<K, N, S, M, 1>, {} }
static WsInst g_ws[] = {
    INST_WS(1, 1, 2, 1),                                          /* 8->32->8 s2 */
    INST_WS(1, 1, 1, 2), INST_WS(1, 2, 1, 2),                     /* 8->48->8, 8->48->16 */
    INST_WS(2, 2, 1, 2),                                          /* 16->96->16 */
};
#undef INST_WS
static WsInst *find_ws(int KS1, int NT3, int S, int MTW)          /* MTW == 0: any */
{
    for (WsInst &i : g_ws) if (i.KS1 == KS1 && i.NT3 == NT3 && i.S == S && (!MTW || i.MTW == MTW)) return &i;
    return nullptr;
}

static BlkInst *find_inst(int KS1, int NT3, int S, int MTW, int GC, int tc = 0)      /* MTW / GC == 0: any */
{
    for (BlkInst &i : g_inst) if (i.tc == tc && i.KS1 == KS1 && i.NT3 == NT3 && i.S == S && (!MTW || i.MTW == MTW) && (!GC || i.GC == GC)) return &i;
    return nullptr;
}

static long long *g_blk_trace = nullptr;
extern "C" void ffb_blk_set_trace(long long *dev_buf) { g_blk_trace = dev_buf; }      /* developer hook, effective only in -DFFB_BLK_TRACE builds */

static float slope_of(int act) { return act == 2 ? 0.1f : act == 1 ? 0.f : 1.f; }

/* Tile search: minimise an estimate of SM cycles per output pixel (tensor pipe: 2.14 clk per m16n8k8 on the SM, measured
 * with tools/micro/mma_sync.cu; FFMA/issue work of the depthwise stage; barrier and tile-start latencies), subject to
 * the shared-memory budget.  FFCNN_BLK_TILE_<OH>_<CEXP>="TH,TW,GC" overrides the choice for blocks with that output height / expanded width. */
static bool plan_tile(BlkPlan *p)
{
    const int S = p->S, KS1 = p->KS1, NT3 = p->NT3, G = p->G, SXs = 8 * KS1 + 4;
    int fTH = 0, fTW = 0, fGC = 0;
    char key[64]; snprintf(key, sizeof key, "FFCNN_BLK_TILE_%d_%d", p->OH, p->cexp);      /* developer override for tile sweeps (tools/blk_sweep.py) */
    if (const char *ov = getenv(key)) sscanf(ov, "%d,%d,%d", &fTH, &fTW, &fGC);
    else {
        /* tiles measured best on a B200 at batch 256 (tools/blk_sweep.py, profiles/r1k_block_tile_sweep.txt); the cost
           model below covers every other shape */
        static const struct { int oh, ow, cexp, s, th, tw, gc; } best_tiles[] = {
            { 80, 80, 32, 1, 16, 16, 2 }, { 40, 40, 32, 2, 4, 20, 2 }, { 40, 40, 48, 1, 8, 20, 3 }, { 40, 40, 96, 1, 8, 20, 2 },
            { 20, 20, 96, 2, 10, 10, 3 }, { 20, 20, 136, 1, 10, 20, 1 },     /* r2v sweep: one group per chunk fits two CTAs per SM (0.0597 vs 0.0638 ms) */
        };
        for (const auto &b : best_tiles)
            if (b.oh == p->OH && b.ow == p->OW && b.cexp == p->cexp && b.s == S) { fTH = b.th; fTW = b.tw; fGC = b.gc; }
    }
    double best = 1e30; bool ok = false;
    for (int GC = 1; GC <= G && GC <= 4; GC++) {
        if (G % GC || (fGC && GC != fGC)) continue;
        const int NC = G / GC, SEs = 16 * GC + (S == 1 ? 8 : 4);
        const BlkChunk off(GC, KS1, NT3, p->tc != 0);
        for (int TW = 2; TW <= p->OW && TW <= 80; TW += 2) {
            if (fTW && TW != fTW) continue;
            for (int TH = 1; TH <= p->OH && TH <= 80; TH++) {
                if (fTH && TH != fTH) continue;
                /* stage-B units: single m-tiles (MTW 1) or 2x2-pixel quads over row pairs (MTW 2, 4) */
                const int M3 = (TH * TW + 15) / 16;
                int MTW = (M3 + BLK_WARPS - 1) / BLK_WARPS;
                if (MTW > 1 && (TH & 1)) continue;                     /* quads need whole row pairs */
                if (MTW > 1) { const int nquads = ((TH + 1) / 2 * TW + 15) / 16; MTW = 2 * ((nquads + BLK_WARPS - 1) / BLK_WARPS); }
                const BlkInst *inst = find_inst(KS1, NT3, S, MTW, GC, p->tc);
                if (!inst) continue;
                const int HH = (TH - 1) * S + 3, HW = (TW - 1) * S + 3;
                const bool frame = TH >= p->OH && TW >= p->OW;        /* the halo ring is all padding: fetch / expand the image only */
                const int XH = frame ? p->H : HH, XW = frame ? p->W : HW;
                if (XH > 256 || XW > 256) continue;
                const int M1 = (XH * XW + 15) / 16, xrows = 32 * ((M1 + 1) / 2);
                const size_t smem = 4 * (size_t)(128 + 2 * xrows + 2 * off.total + 2 * xrows * SXs + HH * HW * SEs) + 128 + (p->tc ? 1024 : 0);
                if (smem > 225 * 1024) continue;
                int occ = (int)((228 * 1024) / (smem + 1024)); if (occ > inst->MINB) occ = inst->MINB; if (occ < 1) occ = 1;
                /* TC: x_hi | x_lo and the double-buffered accumulators of every 128-pixel m-tile live in TMEM (512 columns per SM) */
                const int nmt = (XH * XW + 127) / 128;
                int tmem_cols = 32; while (tmem_cols < nmt * (16 * KS1 + 32 * GC)) tmem_cols *= 2;
                if (p->tc) { if (tmem_cols > 512) continue; if (occ * tmem_cols > 512) occ = 512 / tmem_cols; }
                /* per-tile cost on one SM, three candidate limiters (profiles/r1j: shared-memory wavefronts bind first):
                   wf  = shared-memory wavefronts (1 per clk), mma = m16n8k8 tensor ops (2.14 clk each), ins = issue slots / 4 */
                const int MT = KS1 * GC >= 6 ? 1 : 2, items = (M1 + MT - 1) / MT, rounds = (items + BLK_WARPS - 1) / BLK_WARPS;
                const int units = MTW > 1 ? ((TH + 1) / 2 * TW + 15) / 16 : M3, urounds = (units + BLK_WARPS - 1) / BLK_WARPS;
                const int taps = MTW > 1 ? (S + 3) * (S + 3) : 3 * (S + 3), mper = MTW > 1 ? 2 : 1;
                const double wfA = (double)items * NC * (4 * MT * KS1 + 8 * KS1 * GC + 4 * MT + 8 * MT * GC);
                const double wfB = (double)units * G * 4 * taps + (double)BLK_WARPS * G * (8 * NT3 + 11);
                const double mmaA = (double)rounds * BLK_WARPS * MT * G * KS1 * 6, mmaB = (double)urounds * BLK_WARPS * mper * G * 6 * NT3;
                const double insA = (double)rounds * BLK_WARPS * NC * (KS1 * (14 * MT + GC * (2 + 6 * MT)) + MT * 2 * (8 + 14 * GC));
                const double insB = (double)urounds * BLK_WARPS * G * (taps + mper * (72 + 16 + 24) + 2 * NT3 * (1 + 3 * mper) + 14);
                const double t_wf = (wfA + wfB) / 0.85, t_mma = 2.14 * (mmaA + mmaB), t_ins = (insA + insB) / 4 / 0.8;
                const double t_fix = (NC * 2 * 200.0 + 1500.0) * (occ >= 2 ? 0.5 : 1.0);
                const double t_tile = std::max(t_wf, std::max(t_mma, t_ins)) * (occ >= 2 ? 1.15 : 1.4) + t_fix;
                const double waste = (double)((p->OH + TH - 1) / TH * TH) * ((p->OW + TW - 1) / TW * TW) / ((double)p->OH * p->OW);
                /* tiles per SM: few, big tiles quantise badly when a frame is one or two tiles */
                const double score = t_tile / (TH * TW) * waste;
                if (score < best) {
                    best = score; ok = true;
                    p->GC = GC; p->NC = NC; p->SEs = SEs; p->TH = TH; p->TW = TW; p->HH = HH; p->HW = HW; p->MTW = MTW; p->MINB = inst->MINB;
                    p->xrows = xrows; p->chunk_floats = off.total; p->smem = smem; p->occ = occ;
                    p->XH = XH; p->XW = XW; p->frame = frame; p->xo = p->yo = frame ? 1 : 0; p->nmt = nmt; p->tmem_cols = tmem_cols;
                }
            }
        }
    }
    return ok;
}

/* Tile of the warp-specialised kernel: at most WS_BW stage-B units (one per B warp), two E buffers + the x buffers + the
 * weight slots inside half an SM's shared memory, x split + accumulators of every 96-pixel m-tile inside 256 TMEM columns.
 * Among the feasible tiles: fewest idle B-warp lanes per output pixel, then fewest halo pixels. */
static bool plan_ws(BlkPlan *p)
{
    const int S = p->S, KS1 = p->KS1, NT3 = p->NT3, G = p->G, SXs = 8 * KS1 + 4, GC = 1, NC = G;
    if (NC < 2 || !find_ws(KS1, NT3, S, 0)) return false;
    const int SEs = 16 * GC + (S == 1 ? 8 : 4);
    const BlkChunk off(GC, KS1, NT3, true);
    const int R = NC <= WS_RING ? NC : WS_RING, XB = NC >= 6 ? 2 : 3;   /* a tile's x is split three chunk steps before its first stage B */
    int fTH = 0, fTW = 0, fGC = 0;
    char key[64]; snprintf(key, sizeof key, "FFCNN_BLK_WSTILE_%d_%d", p->OH, p->cexp);
    if (const char *ov = getenv(key)) sscanf(ov, "%d,%d,%d", &fTH, &fTW, &fGC);
    double best = 1e30; bool ok = false;
    for (int TW = 2; TW <= p->OW && TW <= 80; TW += 2) {
        if (fTW && TW != fTW) continue;
        for (int TH = 1; TH <= p->OH && TH <= 80; TH++) {
            if (fTH && TH != fTH) continue;
            for (int MTW = 1; MTW <= 2; MTW++) {
                if (MTW == 2 && (TH & 1)) continue;
                if (!find_ws(KS1, NT3, S, MTW)) continue;
                const int units = MTW == 2 ? (TH / 2 * TW + 15) / 16 : (TH * TW + 15) / 16;
                if (units > WS_BW) continue;
                const int HH = (TH - 1) * S + 3, HW = (TW - 1) * S + 3;
                if (TH >= p->OH && TW >= p->OW) continue;             /* frame mode is k_block_mma's */
                if (HH > 256 || HW > 256) continue;
                const int XP = HH * HW, nmt = (XP + WS_MROWS - 1) / WS_MROWS, xrows = 4 * WS_MROWS;   /* the kernel reads map rows of m-tile pairs */
                const int xbuf = (XP * SXs + 31) / 32 * 32;
                const size_t smem = 4 * (size_t)(128 + 2 * xrows + R * off.total + XB * xbuf + 2 * HH * HW * SEs + SEs) + 1024 + 128;
                if (smem > 113 * 1024) continue;
                int tmem_cols = 32; while (tmem_cols < nmt * (16 * KS1 + 32 * GC)) tmem_cols *= 2;
                if (tmem_cols > 256 || nmt > 4) continue;
                const double waste = (double)((p->OH + TH - 1) / TH * TH) * ((p->OW + TW - 1) / TW * TW) / ((double)p->OH * p->OW);
                /* a chunk step costs the longer of one B unit and the A warps' passes over the halo (about a third of a unit per m-tile) */
                const double step = std::max(1.0, 0.35 * nmt) + 0.08;
                const double score = step * waste / (TH * TW) * (1.0 + 0.02 * XP / (double)(TH * TW));
                if (score < best) {
                    best = score; ok = true;
                    p->GC = GC; p->NC = NC; p->SEs = SEs; p->TH = TH; p->TW = TW; p->HH = HH; p->HW = HW; p->MTW = MTW; p->MINB = 2;
                    p->xrows = xrows; p->chunk_floats = off.total; p->smem = smem; p->occ = 2;
                    p->XH = HH; p->XW = HW; p->frame = 0; p->xo = p->yo = 0; p->nmt = nmt; p->tmem_cols = tmem_cols;
                    p->R = R; p->XB = XB; p->xbuf_floats = xbuf;
                }
            }
        }
    }
    return ok;
}

BlkPlan *blk_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res)
{
    if (cin < 1 || cexp < 1 || cout < 1 || cin % 4 || cexp % 4 || cout % 2 || (stride != 1 && stride != 2)) return nullptr;
    if (res && (stride != 1 || cin != cout)) return nullptr;
    BlkPlan *p = new BlkPlan(); memset(p, 0, sizeof *p);
    p->cin = cin; p->cexp = cexp; p->cout = cout; p->S = stride; p->H = h; p->W = w; p->res = res;
    p->OH = (h - 3 + 2) / stride + 1; p->OW = (w - 3 + 2) / stride + 1;
    if (p->OW % 2 || p->OH < 1) { delete p; return nullptr; }
    p->KS1 = (cin + 7) / 8; p->NT3 = (cout + 7) / 8; p->G = (cexp + 15) / 16;
    /* measured on a B200 at batch 256 (profiles/r2a_block_tc.txt, r2r_block_tc_policy.txt): with the MMAs issued from the warp
       that has no depthwise unit, the tcgen05 expand stage wins on the 96- and 224-channel blocks (L38-L57 -10 %, L58 -17 %,
       L84-L108 -14 % against mma.sync) and on the stride-1 32-channel blocks (L12, L17 -3 %); at 136 channels it tied until the
       stage-A stores lost their branches (r2s: 0.0661 vs 0.0703 ms); it loses on the 48-channel blocks and on L22 */
    static const int env_tc = getenv("FFCNN_BLK_TC") ? atoi(getenv("FFCNN_BLK_TC")) : 0;
    p->tc = env_tc > 0 ? 1 : env_tc < 0 ? 0 : (cexp == 96 || cexp == 224 || cexp == 136 || (cexp == 32 && stride == 1)) ? 1 : 0;
    /* round up to an instantiated (KS1, NT3) pair: zero-padded K / N lanes cost tensor work, not correctness */
    bool found = false;
    for (int k = p->KS1; k <= 6 && !found; k++)
        for (int n = p->NT3; n <= 6 && !found; n++)
            if (find_inst(k, n, stride, 0, 0, p->tc)) { p->KS1 = k; p->NT3 = n; found = true; }
    if (!found && p->tc) {                  /* no tcgen05 instance for this shape: the mma.sync kernel */
        p->tc = 0;
        for (int k = p->KS1; k <= 6 && !found; k++)
            for (int n = p->NT3; n <= 6 && !found; n++)
                if (find_inst(k, n, stride, 0, 0, 0)) { p->KS1 = k; p->NT3 = n; found = true; }
    }
    if (!found) { delete p; return nullptr; }
    p->slope1 = slope_of(act1); p->sloped = slope_of(actd); p->slope3 = slope_of(act3); p->slope_res = slope_of(act_res);
    p->row1 = ((cin + 3) & ~3) + 4; p->rowd = 16; p->row3 = ((cexp + 3) & ~3) + 4;
    if (!plan_tile(p)) {
        if (!p->tc) { delete p; return nullptr; }
        p->tc = 0;                              /* the tcgen05 instances cover fewer tiles */
        if (!plan_tile(p)) { delete p; return nullptr; }
    }
    /* warp-specialised kernel (block_ws.cuh), opt-in with FFCNN_BLK_WS=1 wherever a tile exists.  Built for the blocks whose
       regular tile leaves three of the eight warps without a stage-B unit (the 40x40 maps); parity-green, but measured SLOWER
       than k_block_mma on every one of them (L38: 0.124 vs 0.105 ms, profiles/r2s_block_ws.txt): the fixed roles remove the
       barrier stalls and add instructions (one 16-channel group per chunk, 96-pixel m-tiles, per-chunk GEMM issue), and what
       binds these kernels is instructions issued per cycle at four warps per scheduler, not the idle warps. */
    {
        static const int env_ws = getenv("FFCNN_BLK_WS") ? atoi(getenv("FFCNN_BLK_WS")) : 0;
        const int units = p->MTW > 1 ? ((p->TH + 1) / 2 * p->TW + 15) / 16 : (p->TH * p->TW + 15) / 16;
        if (env_ws > 0 && units <= 2 * WS_BW) {
            BlkPlan w = *p;
            w.KS1 = (cin + 7) / 8; w.NT3 = (cout + 7) / 8;
            bool found_ws = false;
            for (int k = w.KS1; k <= 6 && !found_ws; k++)
                for (int n = w.NT3; n <= 6 && !found_ws; n++)
                    if (find_ws(k, n, stride, 0)) { w.KS1 = k; w.NT3 = n; found_ws = true; }
            if (found_ws && plan_ws(&w)) { *p = w; p->ws = 1; p->tc = 1; }
        }
    }
    p->num_sms = ffb_num_sms();
    snprintf(p->desc, sizeof p->desc, "%d->%d->%d s%d%s tile %dx%d%s gc%d mtw%d smem %zuKB occ%d%s%s", cin, cexp, cout, stride, res ? "+res" : "",
             p->TH, p->TW, p->frame ? " (frame)" : "", p->GC, p->MTW, p->smem >> 10, p->occ, p->tc ? " tcgen05-expand" : "", p->ws ? " warp-specialised" : "");
    return p;
}

void blk_plan_destroy(BlkPlan *p)
{
    if (!p) return;
    cudaFree(p->d_chunks); cudaFree(p->d_sb3);
    delete p;
}

const char *blk_describe(const BlkPlan *p) { return p ? p->desc : ""; }
int blk_uses_tcgen05(const BlkPlan *p) { return p && p->tc; }

int blk_prepare(BlkPlan *p, const float *p1, const float *pd, const float *p3, cudaStream_t st)
{
    const size_t nfl = (size_t)p->NC * p->chunk_floats;
    if (!p->d_chunks && (cudaMalloc(&p->d_chunks, nfl * sizeof(float)) != cudaSuccess || cudaMalloc(&p->d_sb3, 16 * p->NT3 * sizeof(float)) != cudaSuccess)) {
        ffb_set_error("block_mma: cudaMalloc failed"); return -1;
    }
    k_prep_block<<<(int)((nfl + 16 * p->NT3 + 255) / 256), 256, 0, st>>>(p1, p->row1, p->cin, pd, p->rowd, p3, p->row3, p->cexp, p->cout,
                                                                         p->KS1, p->NT3, p->GC, p->NC, p->d_chunks, p->d_sb3, p->tc);
    if (cudaGetLastError() != cudaSuccess) { ffb_set_error("block_mma: weight preparation launch failed"); return -1; }
    return 0;
}

int blk_run(BlkPlan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st)
{
    BlkKernel fn = nullptr;
    if (p->ws) {
        WsInst *wi = find_ws(p->KS1, p->NT3, p->S, p->MTW);
        if (!wi) { ffb_set_error("block_ws: no kernel instance"); return -1; }
        if (ffb_ensure_smem((const void *)wi->fn, p->smem, &wi->configured) != 0) return -1;
        fn = wi->fn;
    } else {
        BlkInst *inst = find_inst(p->KS1, p->NT3, p->S, p->MTW, p->GC, p->tc);
        if (!inst) { ffb_set_error("block_mma: no kernel instance"); return -1; }
        if (ffb_ensure_smem((const void *)inst->fn, p->smem, &inst->configured) != 0) return -1;
        fn = inst->fn;
    }
    BlkArgs a;
    a.x = x; a.y = y; a.wchunks = p->d_chunks; a.sb3 = p->d_sb3;
    a.N = n; a.H = p->H; a.W = p->W; a.OH = p->OH; a.OW = p->OW; a.ldx = ldx; a.ldy = ldy; a.cout = p->cout;
    a.TH = p->TH; a.TW = p->TW; a.HH = p->HH; a.HW = p->HW; a.ntx = (p->OW + p->TW - 1) / p->TW; a.nty = (p->OH + p->TH - 1) / p->TH;
    a.ntiles = (long)n * a.ntx * a.nty;
    a.NC = p->NC; a.xrows = p->xrows;
    a.nmt = p->nmt; a.tmem_cols = (uint32_t)p->tmem_cols;
    a.XH = p->XH; a.XW = p->XW; a.xo = p->xo; a.yo = p->yo; a.frame = p->frame;
    a.R = p->R; a.XB = p->XB; a.xbuf_floats = p->xbuf_floats;
    a.inv_tpf = 1.0f / (float)(a.ntx * a.nty); a.inv_ntx = 1.0f / (float)a.ntx;
    CUtensorMap tm;
    const int SXs = 8 * p->KS1 + 4;
    const unsigned long long dims[4] = { (unsigned long long)ldx, (unsigned long long)p->W, (unsigned long long)p->H, (unsigned long long)n };
    const unsigned long long strides[3] = { (unsigned long long)ldx * 4, (unsigned long long)p->W * ldx * 4, (unsigned long long)p->H * p->W * ldx * 4 };
    const unsigned box[4] = { (unsigned)SXs, (unsigned)p->XW, (unsigned)p->XH, 1u };
    if (ffb_make_tensor_map(&tm, x, 4, dims, strides, box, 0) != 0) return -1;
    a.slope1 = p->slope1; a.sloped = p->sloped; a.slope3 = p->slope3; a.slope_res = p->slope_res; a.res = p->res;
    {   /* FFCNN_BLK_TRACE_SHAPE="cexp,stride" picks the block shape whose launches stamp the developer timeline */
        static int tc = -1, ts = 0;
        if (tc < 0) { tc = 0; if (const char *e = getenv("FFCNN_BLK_TRACE_SHAPE")) sscanf(e, "%d,%d", &tc, &ts); }
        a.trace = (g_blk_trace && p->cexp == tc && p->S == ts) ? g_blk_trace : nullptr;
    }
    const int grid = (int)std::min<long>(a.ntiles, (long)p->num_sms * p->occ);
    cudaError_t e = sm100::launch_pdl(fn, dim3(grid), dim3(BLK_THREADS), p->smem, st, tm, a);
    if (e != cudaSuccess) { ffb_set_error("block_mma launch failed: %s (grid %d smem %zu)", cudaGetErrorString(e), grid, p->smem); return -1; }
    return 0;
}
