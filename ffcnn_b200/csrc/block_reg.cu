/*
 * block_reg.cu -- planner / launcher of the register-resident fused block kernel (block_reg.cuh), instantiated for the
 * three 160x160 blocks of yolo-fastest-1.1: 8->8->4 (L1-L3), 4->8->4 + shortcut (L4-L8), 4->24->8 stride 2 (L9-L11).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "block_reg.h"
#include "block_reg.cuh"
#include "stem_block.cuh"
#include "block_s2.cuh"
#include "ffb_internal.h"


using namespace ffb;

struct RegPlan {
    int cin, cexp, cout, S, H, W, OH, OW, res, R, kind;
    float slope1, sloped, slope3, slope_res;
    union { RegBlockW<8, 8, 4> w884; RegBlockW<4, 8, 4> w484; RegBlockW<4, 8, 8> w488[3]; } u;     /* 4->24->8: three 8-channel slices */
    RegBlockW<4, 24, 8> w4248;              /* 4->24->8 as one shared-memory-tiled kernel (block_s2.cuh) */
    int s2_tile, ring;
    char desc[96];
};

static float slope_of(int act) { return act == 2 ? 0.1f : act == 1 ? 0.f : 1.f; }

/* weights of the expanded channels [c0, c0 + CEXP) of a block whose expanded width is cexp_total */
template <int CIN, int CEXP, int COUT>
static void fill(RegBlockW<CIN, CEXP, COUT> &w, const float *h1, const float *hd, const float *h3, int c0 = 0, int cexp_total = CEXP)
{
    const int row1 = ((CIN + 3) & ~3) + 4, rowd = 16, row3 = ((cexp_total + 3) & ~3) + 4;
    for (int c = 0; c < CEXP; c++) {
        const float *r1 = h1 + (size_t)(c0 + c) * row1, *rd = hd + (size_t)(c0 + c) * rowd;
        for (int k = 0; k < CIN; k++) w.w1[k][c] = r1[k];
        w.s1[c] = r1[row1 - 4]; w.b1[c] = r1[row1 - 3];
        for (int t = 0; t < 9; t++) w.wd[t][c] = rd[t];
        w.sd[c] = rd[rowd - 4]; w.bd[c] = rd[rowd - 3];
    }
    for (int co = 0; co < COUT; co++) {
        for (int c = 0; c < CEXP; c++) w.w2[c][co] = h3[(size_t)co * row3 + c0 + c];
        w.s3[co] = h3[(size_t)co * row3 + row3 - 4]; w.b3[co] = h3[(size_t)co * row3 + row3 - 3];
    }
}

RegPlan *reg_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res,
                         const float *h1, const float *hd, const float *h3)
{
    if (!h1 || !hd || !h3 || (w & 1)) return nullptr;
    int kind = -1;
    if (cin == 8 && cexp == 8 && cout == 4 && stride == 1 && !res) kind = 0;
    else if (cin == 4 && cexp == 8 && cout == 4 && stride == 1 && res) kind = 1;
    else if (cin == 4 && cexp == 24 && cout == 8 && stride == 2 && !res) kind = 2;
    if (kind < 0) return nullptr;
    RegPlan *p = new RegPlan(); memset(p, 0, sizeof *p);
    p->cin = cin; p->cexp = cexp; p->cout = cout; p->S = stride; p->H = h; p->W = w; p->res = res; p->kind = kind;
    p->OH = (h - 3 + 2) / stride + 1; p->OW = (w - 3 + 2) / stride + 1;
    p->slope1 = slope_of(act1); p->sloped = slope_of(actd); p->slope3 = slope_of(act3); p->slope_res = slope_of(act_res);
    const char *env = getenv("FFCNN_REG_ROWS");
    p->R = env ? atoi(env) : (kind == 2 ? 9 : 16);      /* measured: 16 rows per strip is best for the stride-1 kernels, 9 for the stride-2 slices (3.9 waves of CTAs instead of 2.2) */
    if (p->R < 1) p->R = 16;
    /* x rows through a per-lane cp.async ring in shared memory (block_reg.cuh "row ring") instead of register loads two rows ahead; FFCNN_REG_RING=0 restores the latter */
    p->ring = getenv("FFCNN_REG_RING") ? atoi(getenv("FFCNN_REG_RING")) : 1;
    if (kind == 0) fill(p->u.w884, h1, hd, h3); else if (kind == 1) fill(p->u.w484, h1, hd, h3); else for (int i = 0; i < 3; i++) fill(p->u.w488[i], h1, hd, h3, 8 * i, 24);
    if (kind == 2) {
        fill(p->w4248, h1, hd, h3);
        /* FFCNN_S2_TILE=1: the shared-memory-tiled kernel (block_s2.cuh) instead of the three register-resident slices.  Bit-identical,
           a third fewer instructions, and measured SLOWER (0.228 vs 0.214 ms; 16x8 / 16x4 tiles at 3-6 CTAs per SM all within 0.222-0.228):
           kept as an option and as the record of the experiment */
        p->s2_tile = getenv("FFCNN_S2_TILE") ? atoi(getenv("FFCNN_S2_TILE")) : 0;
    }
    if (kind == 2 && p->s2_tile) snprintf(p->desc, sizeof p->desc, "%d->%d->%d s%d shared-memory tiles %dx%d", cin, cexp, cout, stride, S2_TXO, S2_TYO);
    else snprintf(p->desc, sizeof p->desc, "%d->%d->%d s%d%s register-resident, %d rows per warp strip%s", cin, cexp, cout, stride, res ? "+res" : "", p->R, p->ring ? ", cp.async row ring" : "");
    return p;
}

void reg_plan_destroy(RegPlan *p) { delete p; }
const char *reg_describe(const RegPlan *p) { return p ? p->desc : ""; }
int reg_launches(const RegPlan *p) { return p && p->kind == 2 && !p->s2_tile ? 3 : 1; }

int reg_run(RegPlan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st)
{
    if (ldx != p->cin || ldy != p->cout) { ffb_set_error("block_reg: pixel pitch %d/%d does not match the channel counts %d/%d", ldx, ldy, p->cin, p->cout); return -1; }
    RegBlockArgs a;
    a.x = x; a.y = y; a.N = n; a.H = p->H; a.W = p->W; a.OH = p->OH; a.OW = p->OW; a.R = p->R;
    a.nsx = p->S == 1 ? (p->OW + 59) / 60 : (p->OW + 30) / 31; a.nsy = (p->OH + p->R - 1) / p->R;
    a.slope1 = p->slope1; a.sloped = p->sloped; a.slope3 = p->slope3; a.slope_res = p->slope_res;
    const long strips = (long)n * a.nsx * a.nsy;
    const dim3 grid((unsigned)((strips + REG_WARPS - 1) / REG_WARPS)), block(REG_WARPS * 32);
    cudaError_t e;
    if (p->kind == 0)      e = p->ring ? sm100::launch_pdl(k_block_reg_s1<8, 8, 4, false, true>, grid, block, 0, st, p->u.w884, a)
                                       : sm100::launch_pdl(k_block_reg_s1<8, 8, 4, false>, grid, block, 0, st, p->u.w884, a);
    else if (p->kind == 1) e = p->ring ? sm100::launch_pdl(k_block_reg_s1<4, 8, 4, true, true>, grid, block, 0, st, p->u.w484, a)
                                       : sm100::launch_pdl(k_block_reg_s1<4, 8, 4, true>, grid, block, 0, st, p->u.w484, a);
    else if (p->s2_tile) {
        static ffb_smem_cfg configured;
        if (ffb_ensure_smem((const void *)k_block_s2_tile, S2_SMEM, &configured) != 0) return -1;
        S2Args s; s.x = x; s.y = y; s.N = n; s.H = p->H; s.W = p->W; s.OH = p->OH; s.OW = p->OW; s.slope1 = p->slope1; s.sloped = p->sloped; s.slope3 = p->slope3;
        const dim3 g2((unsigned)((p->OW + S2_TXO - 1) / S2_TXO), (unsigned)((p->OH + S2_TYO - 1) / S2_TYO), (unsigned)n);
        e = sm100::launch_pdl(k_block_s2_tile, g2, dim3(S2_THREADS), S2_SMEM, st, p->w4248, s);
    }
    else if (p->ring) {    /* three 8-channel slices of the expanded tensor, accumulated through y (which stays in L2) */
        e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 1, true>, grid, block, 0, st, p->u.w488[0], a);
        if (e == cudaSuccess) e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 2, true>, grid, block, 0, st, p->u.w488[1], a);
        if (e == cudaSuccess) e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 3, true>, grid, block, 0, st, p->u.w488[2], a);
    }
    else {
        e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 1>, grid, block, 0, st, p->u.w488[0], a);
        if (e == cudaSuccess) e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 2>, grid, block, 0, st, p->u.w488[1], a);
        if (e == cudaSuccess) e = sm100::launch_pdl(k_block_reg_s2<4, 8, 8, 3>, grid, block, 0, st, p->u.w488[2], a);
    }
    if (e != cudaSuccess) { ffb_set_error("block_reg launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

/* ---- stem + first block as one kernel (stem_block.cuh) ---- */
int reg_stem_ok(const RegPlan *p, int ih, int iw, int pitch, const void *frames)
{
    return p && p->kind == 0 && p->H == (ih - 3 + 2) / 2 + 1 && p->W == (iw - 3 + 2) / 2 + 1 && iw % 4 == 0 && pitch % 4 == 0 && ((uintptr_t)frames & 3) == 0;
}

int reg_run_stem(RegPlan *p, const void *stemw, int act0, const unsigned char *frames, int pitch, float *y, int ldy, int n, int ih, int iw,
                 const float *mean, const float *norm, cudaStream_t st)
{
    if (!reg_stem_ok(p, ih, iw, pitch, frames) || ldy != 4) { ffb_set_error("stem_block: shape not supported"); return -1; }
    static ffb_smem_cfg configured;
    if (ffb_ensure_smem((const void *)k_stem_block, SB_SMEM, &configured) != 0) return -1;
    StemBlockArgs a;
    a.frames = frames; a.pitch = pitch; a.y = y; a.H = ih; a.W = iw; a.OH = p->H; a.OW = p->W;
    a.act0 = act0; a.m0 = mean[0]; a.m1 = mean[1]; a.m2 = mean[2]; a.n0 = norm[0]; a.n1 = norm[1]; a.n2 = norm[2];
    a.slope1 = p->slope1; a.sloped = p->sloped; a.slope3 = p->slope3;
    const dim3 grid((unsigned)((p->W + SB_TXO - 1) / SB_TXO), (unsigned)((p->H + SB_TYO - 1) / SB_TYO), (unsigned)n);
    cudaError_t e = sm100::launch_pdl(k_stem_block, grid, dim3(SB_THREADS), SB_SMEM, st, *reinterpret_cast<const StemW *>(stemw), p->u.w884, a);
    if (e != cudaSuccess) { ffb_set_error("stem_block launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}
