/*
 * block_tc.cu -- planner, weight preparation and launcher of the all-tcgen05 fused block kernel (block_tc.cuh).
 * Instantiated for the (cin, cout, stride) classes of yolo-fastest-1.1; other shapes report "unsupported" and the engine
 * falls back to block_mma.cu / the separate layers.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "block_tc.h"
#include "pw_tc.h"
#include "ffb_internal.h"
#include "block_tc.cuh"

using namespace ffb;

struct Blk2Plan {
    int cin, cexp, cout, S, H, W, OH, OW, res;
    int KS1, N3, NC, TH, TW, HH, HW, XH, XW, xo, yo, frame, nmt, xrows, tmem_cols, occ, chunk_floats;
    float slope1, sloped, slope3, slope_res;
    size_t smem;
    float *d_chunks, *d_sb3;
    int row1, rowd, row3, num_sms;
    char desc[128];
};

typedef void (*Blk2Kernel)(const CUtensorMap, const Blk2Args);
struct Blk2Inst { int KS1, N3, S; Blk2Kernel fn; ffb_smem_cfg configured; };
#define INST(K, N, S) { K, N, S, k_block_tc<K, N, S>, {} }
static Blk2Inst g_inst2[] = {
    INST(1, 16, 1), INST(1, 16, 2),          /* 8->32->8, 8->48->8, 8->48->16, 4->24->8 s2, 8->32->8 s2 */
    INST(2, 16, 1), INST(2, 32, 2),          /* 16->96->16, 16->96->24 s2 */
    INST(3, 32, 1), INST(3, 48, 2),          /* 24->136->24, 24->136->48 s2 */
    INST(6, 48, 1),                          /* 48->224->48 */
};
#undef INST

static Blk2Inst *find_inst2(int KS1, int N3, int S)
{
    for (Blk2Inst &i : g_inst2) if (i.KS1 == KS1 && i.N3 == N3 && i.S == S) return &i;
    return nullptr;
}

static float slope_of2(int act) { return act == 2 ? 0.1f : act == 1 ? 0.f : 1.f; }

/* Tile search.  Cost of a tile in "thread-pixel passes": stage A touches nmt*128 halo rows per chunk (~0.45 of a stage-B pass
 * each), stage B 128 rows; fixed per-tile cost for the x split, epilogue and barriers.  One CTA per SM (TMEM > 256 columns or
 * shared memory > 113 KB) hides latency worse: x1.5.  FFCNN_BLK2_TILE_<OH>_<CEXP>="TH,TW" overrides. */
static bool plan_tile2(Blk2Plan *p)
{
    const int S = p->S, KP = 8 * p->KS1, SXs = KP + 4;
    const Blk2Chunk off(p->KS1, p->N3);
    int fTH = 0, fTW = 0;
    char key[64]; snprintf(key, sizeof key, "FFCNN_BLK2_TILE_%d_%d", p->OH, p->cexp);
    if (const char *ov = getenv(key)) sscanf(ov, "%d,%d", &fTH, &fTW);
    double best = 1e30; bool ok = false;
    for (int TH = 1; TH <= p->OH && TH <= 128; TH++) {
        if (fTH && TH != fTH) continue;
        for (int TW = 1; TW <= p->OW && TH * TW <= 128; TW++) {
            if (fTW && TW != fTW) continue;
            const int HH = (TH - 1) * S + 3, HW = (TW - 1) * S + 3;
            const bool frame = TH >= p->OH && TW >= p->OW;
            const int XH = frame ? p->H : HH, XW = frame ? p->W : HW;
            if (XH > 256 || XW > 256) continue;
            const int XP = XH * XW, nmt = (XP + 127) / 128;
            if (nmt > 3) continue;
            const int xrows = nmt * 128;
            int tmem = nmt * (2 * KP + B2_CH) + B2_CH + p->N3, cols = 32; while (cols < tmem) cols *= 2;
            if (cols > 512) continue;
            const size_t smem = 4 * (size_t)(128 + 2 * xrows) + 1024 + 4 * (size_t)(2 * off.total + 128 * 32 + 2 * xrows * SXs + HH * HW * B2_SE) + 64;
            if (smem > 225 * 1024) continue;
            int occ = (int)((228 * 1024) / (smem + 1024)); if (occ > 2) occ = 2; if (occ < 1) occ = 1;
            if (occ * cols > 512) occ = 512 / cols;
            const int tiles = ((p->OH + TH - 1) / TH) * ((p->OW + TW - 1) / TW);
            const double per_tile = p->NC * (nmt * 0.45 + 1.0) + 1.2;
            const double score = tiles * per_tile * (occ >= 2 ? 1.0 : 1.5) + 1e-3 * XP;
            if (score < best) {
                best = score; ok = true;
                p->TH = TH; p->TW = TW; p->HH = HH; p->HW = HW; p->XH = XH; p->XW = XW; p->frame = frame; p->xo = p->yo = frame ? 1 : 0;
                p->nmt = nmt; p->xrows = xrows; p->tmem_cols = cols; p->smem = smem; p->occ = occ; p->chunk_floats = off.total;
            }
        }
    }
    return ok;
}

Blk2Plan *blk2_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res)
{
    if (cin < 1 || cexp < 1 || cout < 1 || cin % 4 || cexp % 4 || cout % 4 || (stride != 1 && stride != 2)) return nullptr;
    if (res && (stride != 1 || cin != cout)) return nullptr;
    Blk2Plan *p = new Blk2Plan(); memset(p, 0, sizeof *p);
    p->cin = cin; p->cexp = cexp; p->cout = cout; p->S = stride; p->H = h; p->W = w; p->res = res;
    p->OH = (h - 3 + 2) / stride + 1; p->OW = (w - 3 + 2) / stride + 1;
    if (p->OH < 1 || p->OW < 1) { delete p; return nullptr; }
    p->KS1 = (cin + 7) / 8; p->N3 = (cout + 15) / 16 * 16; p->NC = (cexp + B2_CH - 1) / B2_CH;
    /* round up to an instantiated (KS1, N3): zero-padded K / N lanes cost tensor work only */
    bool found = false;
    for (int k = p->KS1; k <= 6 && !found; k++)
        for (int n = p->N3; n <= 48 && !found; n += 16)
            if (find_inst2(k, n, stride)) { p->KS1 = k; p->N3 = n; found = true; }
    if (!found || !plan_tile2(p)) { delete p; return nullptr; }
    p->slope1 = slope_of2(act1); p->sloped = slope_of2(actd); p->slope3 = slope_of2(act3); p->slope_res = slope_of2(act_res);
    p->row1 = ((cin + 3) & ~3) + 4; p->rowd = 16; p->row3 = ((cexp + 3) & ~3) + 4;
    p->num_sms = ffb_num_sms();
    snprintf(p->desc, sizeof p->desc, "%d->%d->%d s%d%s tcgen05 tile %dx%d%s halo m-tiles %d chunks %d tmem %d smem %zuKB occ%d", cin, cexp, cout, stride,
             res ? "+res" : "", p->TH, p->TW, p->frame ? " (frame)" : "", p->nmt, p->NC, p->tmem_cols, p->smem >> 10, p->occ);
    return p;
}

void blk2_plan_destroy(Blk2Plan *p)
{
    if (!p) return;
    cudaFree(p->d_chunks); cudaFree(p->d_sb3);
    delete p;
}

const char *blk2_describe(const Blk2Plan *p) { return p ? p->desc : ""; }

int blk2_prepare(Blk2Plan *p, const float *p1, const float *pd, const float *p3, cudaStream_t st)
{
    const size_t nfl = (size_t)p->NC * p->chunk_floats;
    if (!p->d_chunks && (cudaMalloc(&p->d_chunks, nfl * sizeof(float)) != cudaSuccess || cudaMalloc(&p->d_sb3, 2 * p->N3 * sizeof(float)) != cudaSuccess)) {
        ffb_set_error("block_tc: cudaMalloc failed"); return -1;
    }
    k_prep_block2<<<(int)((nfl + 2 * p->N3 + 255) / 256), 256, 0, st>>>(p1, p->row1, p->cin, pd, p->rowd, p3, p->row3, p->cexp, p->cout,
                                                                        p->KS1, p->N3, p->NC, p->d_chunks, p->d_sb3);
    if (cudaGetLastError() != cudaSuccess) { ffb_set_error("block_tc: weight preparation launch failed"); return -1; }
    return 0;
}

int blk2_run(Blk2Plan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st)
{
    Blk2Inst *inst = find_inst2(p->KS1, p->N3, p->S);
    if (!inst) { ffb_set_error("block_tc: no kernel instance"); return -1; }
    if (ffb_ensure_smem((const void *)inst->fn, p->smem, &inst->configured) != 0) return -1;
    Blk2Args a;
    a.x = x; a.y = y; a.wchunks = p->d_chunks; a.sb3 = p->d_sb3;
    a.N = n; a.H = p->H; a.W = p->W; a.OH = p->OH; a.OW = p->OW; a.ldx = ldx; a.ldy = ldy; a.cout = p->cout;
    a.TH = p->TH; a.TW = p->TW; a.HH = p->HH; a.HW = p->HW; a.ntx = (p->OW + p->TW - 1) / p->TW; a.nty = (p->OH + p->TH - 1) / p->TH;
    a.ntiles = (long)n * a.ntx * a.nty;
    a.NC = p->NC; a.xrows = p->xrows; a.XH = p->XH; a.XW = p->XW; a.xo = p->xo; a.yo = p->yo; a.frame = p->frame; a.nmt = p->nmt;
    a.tmem_cols = (uint32_t)p->tmem_cols;
    a.inv_tpf = 1.0f / (float)(a.ntx * a.nty); a.inv_ntx = 1.0f / (float)a.ntx;
    a.slope1 = p->slope1; a.sloped = p->sloped; a.slope3 = p->slope3; a.slope_res = p->slope_res; a.res = p->res;
    CUtensorMap tm;
    const int SXs = 8 * p->KS1 + 4;
    const unsigned long long dims[4] = { (unsigned long long)ldx, (unsigned long long)p->W, (unsigned long long)p->H, (unsigned long long)n };
    const unsigned long long strides[3] = { (unsigned long long)ldx * 4, (unsigned long long)p->W * ldx * 4, (unsigned long long)p->H * p->W * ldx * 4 };
    const unsigned box[4] = { (unsigned)SXs, (unsigned)p->XW, (unsigned)p->XH, 1u };
    if (ffb_make_tensor_map(&tm, x, 4, dims, strides, box, 0) != 0) return -1;
    const int grid = (int)std::min<long>(a.ntiles, (long)p->num_sms * p->occ);
    cudaError_t e = sm100::launch_pdl(inst->fn, dim3(grid), dim3(B2_THREADS), p->smem, st, tm, a);
    if (e != cudaSuccess) { ffb_set_error("block_tc launch failed: %s (grid %d smem %zu)", cudaGetErrorString(e), grid, p->smem); return -1; }
    return 0;
}
