/*
 * engine.cu -- device half of the ffcnn.h boundary: the layer loop of net_forward (ffcnn.c:476-520)
 * re-designed for a B200.
 *
 *   reference                                    here
 *   -------------------------------------------  ------------------------------------------------------
 *   one image, planar CHW, malloc/free per       batch of frames, NHWC fp32 in one HBM arena planned once
 *   layer with refcounts (ffcnn.c:481-517)        from the same dependency information (liveness reuse)
 *   switch(type) -> static C loops                one hand-written sm_100a kernel (or alias) per layer,
 *                                                 the whole sequence captured in a CUDA graph per batch size
 *   groupconv() picks a CPU kernel by shape       conv_pick() picks a GPU kernel by the same predicates
 *   (conv-v6.c:481-502)
 *   dropout moves a pointer (ffcnn.c:412-416)     dropout / single-input route are aliases: no launch
 *   yolo decode + nms on the CPU                  GPU candidate filter, exact decode + NMS on the host
 *
 * There is deliberately no CPU execution path in this file: without a CUDA device attach fails.
 */
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "ffb_internal.h"
#include "../../include/conv.h"
#include "kernels.cuh"
#include "pw_ffma.cuh"
#include "dw_tma.cuh"
#include "pw_tc.h"
#include "block_mma.h"
#include "block_reg.h"
#include "block_tc.h"
#include "conv_tc.h"

using namespace ffb;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ffb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

/* Per-device state.  cudaFuncSetAttribute and the SM count are properties of a (function, device) pair, and several nets
 * may live on different devices of one process (ffb_multi_*), so nothing here is a process-wide scalar. */
int ffb_num_sms(void)
{
    static int sms[FFB_MAX_DEVICES];                    /* 0 = not queried yet; racing writers store the same value */
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FFB_MAX_DEVICES) return 148;
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}
int ffb_ensure_smem(const void *func, size_t smem, ffb_smem_cfg *cfg)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FFB_MAX_DEVICES) { ffb_set_error("cudaGetDevice failed"); return -1; }
    if (smem <= cfg->bytes[dev]) return 0;
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        ffb_set_error("cannot raise dynamic shared memory to %zu bytes: %s", smem, cudaGetErrorString(cudaGetLastError())); return -1;
    }
    cfg->bytes[dev] = smem;
    return 0;
}
#define g_num_sms (ffb_num_sms())
namespace sm100 { int g_ffb_pdl = 1; }     /* programmatic dependent launch between the kernels of a forward pass (FFCNN_PDL=0 disables): every kernel triggers at its top and waits after its prologue (barrier init, TMEM allocation, clearing E, weight prefetch), so the prologue of kernel n+1 runs under the tail of kernel n wherever an SM has room.  Round 1 measured nothing in the captured graph (3.44 vs 3.40 ms) and left it off; with 41 shorter launches it is worth 1.6 % (r2v: 2.213 -> 2.177 ms).  ffb_layer_times switches it off: per-kernel times are those of isolated kernels. */
using sm100::launch_pdl;

static inline int grid_for(long total, int block, int waves = 8)
{
    long g = (total + block - 1) / block;
    long cap = (long)g_num_sms * waves;
    return (int)std::max<long>(1, std::min(g, cap));
}

/* =================================================================================== conv operator */

enum ConvKind { CK_GENERIC = 0, CK_PW_FFMA, CK_PW_TC, CK_DW_S1_3, CK_DW_S1_5, CK_DW3_S2, CK_STEM, CK_IGEMM_TC };

struct ffb_conv {
    int ic, groups, pad, stride, fs, fn, act, row, taps;
    int dw5_exact, pw_mode, dw_mode;   /* dw_mode 0 = TMA-fed stencil, 1 = register-window LDG kernel */
    ConvKind kind;
    const float *d_packed;      /* packed reference rows on device */
    float *d_owned_packed;      /* set when this op owns d_packed (standalone ops) */
    float *d_prep;              /* [taps][fn_pad] + scale[fn_pad] + bias[fn_pad] */
    int fn_pad;
    /* pointwise FFMA tiling */
    int TM, TN, NT, TY; size_t smem;
    PwTcPlan *tc;               /* tcgen05 plan (pw_tc.cu), NULL if not used */
    IgPlan *ig;                 /* implicit-GEMM tcgen05 plan (conv_tc.cu): dense k x k convs and pointwise layers too large for pw_tc */
    const float *h_packed;      /* host copy of the packed rows, only dereferenced during conv_prepare */
    StemW stemw;                /* stem weights as a kernel-parameter block (constant-bank FFMA operands) */
    char name[48];
};

static int conv_out_dim(int in, int fs, int pad, int stride) { return (in - fs + 2 * pad) / stride + 1; }

static void conv_pick(ffb_conv *op)
{
    const int cpg = op->ic / op->groups;
    op->kind = CK_GENERIC;
    if (op->pad == 0 && op->fs == 1 && op->stride == 1 && op->groups == 1 && op->ic % 4 == 0) op->kind = CK_PW_FFMA;
    else if (cpg == 1 && op->fn == op->ic && op->ic % 4 == 0 && op->stride == 1 && op->fs == 3 && op->pad == 1) op->kind = CK_DW_S1_3;
    else if (cpg == 1 && op->fn == op->ic && op->ic % 4 == 0 && op->stride == 1 && op->fs == 5 && op->pad == 2) op->kind = CK_DW_S1_5;
    else if (cpg == 1 && op->fn == op->ic && op->ic % 4 == 0 && op->stride == 2 && op->fs == 3 && op->pad == 1) op->kind = CK_DW3_S2;
    else if (op->groups == 1 && op->ic == 3 && op->fn == 8 && op->fs == 3 && op->stride == 2 && op->pad == 1) op->kind = CK_STEM;
}

/* shared-memory footprint of the FFMA pointwise kernel for a given tiling */
static size_t pw_smem(int K, int BN, int TM, int TY) { return ((size_t)K * BN + 2 * BN + 2 * (size_t)TM * TY * (K + 4)) * sizeof(float); }

static void pw_plan(ffb_conv *op)
{
    const int K = op->ic, N = op->fn;
    op->TN = N > 64 ? 8 : 4;
    op->NT = (N + op->TN - 1) / op->TN;
    op->TY = 256 / op->NT;
    op->fn_pad = op->NT * op->TN;
    const int tms[4] = { 8, 4, 2, 1 };
    op->TM = 1;
    for (int pass = 0; pass < 2; pass++) {
        const size_t limit = pass == 0 ? 100 * 1024 : 220 * 1024;
        bool found = false;
        for (int t = 0; t < 4; t++) {
            if (tms[t] * op->TN > 64) continue;
            if (pw_smem(K, op->fn_pad, tms[t], op->TY) <= limit) { op->TM = tms[t]; found = true; break; }
        }
        if (found) break;
    }
    op->smem = pw_smem(K, op->fn_pad, op->TM, op->TY);
}

/* ---- TMA-fed depthwise kernels (3x3 s1, 3x3 s2, 5x5 s1): tile planner + launcher ---- */
struct DwPlan { int CB, TW, TH, RC, nch, IWb, IHb, ntx, nty, ntc, stages; size_t smem; };

/* Pick channel block / output tile / row chunking: one work item (2 px x 4 ch x RC rows) per thread, stage <= 108 KB (2 stages),
 * minimising  halo re-read factor x shared-memory reads per output, with penalties for idle threads and for thin TMA
 * rows when the channel block is not the whole (contiguous) pixel.  FS/stride select the box geometry; max_rc caps RC. */
static bool dw_plan(int C, int OH, int OW, int FS, int stride, int max_rc, DwPlan *p)
{
    /* developer overrides for tile-shape sweeps (tools/op_bench.py): FFCNN_DW_STAGE_KB, FFCNN_DW_STAGES */
    static const int env_kb = getenv("FFCNN_DW_STAGE_KB") ? atoi(getenv("FFCNN_DW_STAGE_KB")) : 0;
    static const int env_st = getenv("FFCNN_DW_STAGES") ? atoi(getenv("FFCNN_DW_STAGES")) : 0;
    const int stages = env_st ? env_st : 2;
    const size_t stage_limit = (size_t)(env_kb ? env_kb : 108) * 1024;   /* measured (profiles/r1f_dw_tile_sweep.txt): few big tiles beat many small ones */
    double best = 1e30; bool ok = false;
    for (int CB = 4; CB <= C && CB <= 256; CB += 4) {
        if (C % CB) continue;
        for (int ntx = 1; ntx <= 40 && ntx <= OW; ntx++) {
            const int TW = (OW + ntx - 1) / ntx, pairs = (TW + 1) / 2, TWp = 2 * pairs;
            const int IWb = (TWp - 1) * stride + FS;
            if (IWb > 256) continue;
            const int per_chunk = pairs * (CB / 4);
            if (per_chunk > DW_THREADS) continue;
            const size_t rowb = (size_t)IWb * CB * 4;
            int ihmax = (int)(stage_limit / rowb); if (ihmax > 256) ihmax = 256;
            int thmax = (ihmax - FS) / stride + 1; if (thmax > OH) thmax = OH;
            if (ihmax < FS || thmax < 1) continue;
            for (int nty = (OH + thmax - 1) / thmax; nty <= OH; nty++) {
                const int TH = (OH + nty - 1) / nty, IHb = (TH - 1) * stride + FS;
                int nch = DW_THREADS / per_chunk; if (nch > TH) nch = TH; if (nch < 1) nch = 1;
                int RC = (TH + nch - 1) / nch;
                if (RC > max_rc) { RC = max_rc; if ((TH + RC - 1) / RC * per_chunk > DW_THREADS) continue; }
                nch = (TH + RC - 1) / RC;
                const double halo = (double)IWb * IHb / ((double)TW * stride * TH * stride);
                const double lds = FS == 3 && stride == 1 ? 2.0 * (RC + 2) / RC : FS == 3 ? 5.0 * (2 * RC + 1) / (2.0 * RC) : 15.0;
                const double idle = (double)DW_THREADS / (per_chunk * nch);
                const double thin = (CB < C && CB * 4 < 256) ? 1.0 + 0.25 * (256.0 / (CB * 4) - 1.0) : 1.0;
                const double score = (halo + 0.35 * lds / stride) * (0.75 + 0.25 * idle) * thin;
                if (score < best) {
                    best = score; ok = true;
                    p->CB = CB; p->TW = TW; p->TH = TH; p->RC = RC; p->nch = nch; p->IWb = IWb; p->IHb = IHb;
                    p->ntx = (OW + TW - 1) / TW; p->nty = nty; p->ntc = C / CB;
                }
                if (RC >= 12 || TH <= 4) break;                               /* shorter tiles only get worse from here */
            }
        }
    }
    if (!ok) return false;
    p->stages = stages;
    const size_t stage = (((size_t)p->IHb * p->IWb * p->CB * 4) + 127) & ~(size_t)127;
    p->smem = p->stages * stage + 64 + 256;
    return true;
}

/* 5x5 kernel (k_dw5s1_tma): the CTA's threads loop over items of 2 px x 2 ch x DW5_RC rows, so any tile whose stage fits is
 * legal; minimise  halo re-read  x  idle thread slots in the last pass  x  the thin-row penalty of a partial channel block. */
static bool dw5_plan(int C, int OH, int OW, DwPlan *p)
{
    static const int env_kb = getenv("FFCNN_DW_STAGE_KB") ? atoi(getenv("FFCNN_DW_STAGE_KB")) : 0;
    const size_t stage_limit = (size_t)(env_kb ? env_kb : 108) * 1024;
    double best = 1e30; bool ok = false;
    int fCB = 0, fTW = 0, fTH = 0;                                   /* developer override for tile sweeps: FFCNN_DW5_TILE_<C>="CB,TW,TH" */
    { char key[48]; snprintf(key, sizeof key, "FFCNN_DW5_TILE_%d", C); if (const char *ov = getenv(key)) sscanf(ov, "%d,%d,%d", &fCB, &fTW, &fTH); }
    for (int CB = 4; CB <= C && CB <= 256; CB += 4) {
        if (C % CB || (fCB && CB != fCB)) continue;
        for (int ntx = 1; ntx <= OW; ntx++) {
            const int TW = (OW + ntx - 1) / ntx, pairs = (TW + 1) / 2, IWb = 2 * pairs + 4;
            if (IWb > 256) continue;
            for (int nty = 1; nty <= OH; nty++) {
                const int TH = (OH + nty - 1) / nty, IHb = TH + 4;
                if ((fTW && TW != fTW) || (fTH && TH != fTH)) continue;
                if (IHb > 256 || (size_t)IWb * IHb * CB * 4 > stage_limit) continue;
                const int nch = (TH + DW5_RC - 1) / DW5_RC, items = nch * pairs * (CB / 2);
                const double halo = (double)IWb * IHb / ((double)TW * TH);
                const double passes = (double)((items + DW_THREADS - 1) / DW_THREADS) * DW_THREADS / items;
                const double rows = (double)nch * DW5_RC / TH;                   /* ragged last row chunk */
                const double thin = (CB < C && CB * 4 < 256) ? 1.0 + 0.25 * (256.0 / (CB * 4) - 1.0) : 1.0;
                const double score = (halo + 0.5) * (0.5 + 0.5 * passes * rows) * thin;
                if (score < best) {
                    best = score; ok = true;
                    p->CB = CB; p->TW = TW; p->TH = TH; p->RC = DW5_RC; p->nch = nch; p->IWb = IWb; p->IHb = IHb;
                    p->ntx = (OW + TW - 1) / TW; p->nty = (OH + TH - 1) / TH; p->ntc = C / CB;
                }
            }
        }
    }
    if (!ok) return false;
    p->stages = 2;
    const size_t stage = (((size_t)p->IHb * p->IWb * p->CB * 4) + 127) & ~(size_t)127;
    p->smem = p->stages * stage + 64 + 256;
    return true;
}

typedef void (*DwKernel)(const CUtensorMap, const DwArgs);

static int dw_launch(DwKernel kernel, ffb_smem_cfg *configured, const float *in, DwArgs &a, const DwPlan &pl, cudaStream_t st)
{
    if (ffb_ensure_smem((const void *)kernel, pl.smem, configured) != 0) return -1;
    CUtensorMap tm;
    const unsigned long long dims[4] = { (unsigned long long)a.C, (unsigned long long)a.W, (unsigned long long)a.H, (unsigned long long)a.N };
    const unsigned long long strides[3] = { (unsigned long long)a.C * 4, (unsigned long long)a.W * a.C * 4, (unsigned long long)a.H * a.W * a.C * 4 };
    const unsigned box[4] = { (unsigned)pl.CB, (unsigned)pl.IWb, (unsigned)pl.IHb, 1u };
    if (ffb_make_tensor_map(&tm, in, 4, dims, strides, box, 0) != 0) return -1;
    a.CB = pl.CB; a.TW = pl.TW; a.TH = pl.TH; a.RC = pl.RC; a.nch = pl.nch; a.ntx = pl.ntx; a.nty = pl.nty; a.ntc = pl.ntc;
    a.IWb = pl.IWb; a.IHb = pl.IHb;
    a.ntiles = (long)a.N * pl.nty * pl.ntx * pl.ntc; a.stages = pl.stages;
    const int grid = (int)std::min<long>(a.ntiles, g_num_sms);
    CK(launch_pdl(kernel, dim3(grid), dim3(DW_THREADS), pl.smem, st, tm, a));
    return 0;
}

template <int TM, int TN, bool RES>
static cudaError_t pw_launch2(const PwArgs &a, int grid, size_t smem, cudaStream_t st)
{
    static ffb_smem_cfg configured;
    if (ffb_ensure_smem((const void *)k_pw_ffma<TM, TN, RES>, smem, &configured) != 0) return cudaErrorInvalidValue;
    return launch_pdl(k_pw_ffma<TM, TN, RES>, dim3(grid), dim3(256), smem, st, a);
}

template <int TM, int TN>
static cudaError_t pw_launch(const PwArgs &a, int grid, size_t smem, cudaStream_t st)
{
    return a.res ? pw_launch2<TM, TN, true>(a, grid, smem, st) : pw_launch2<TM, TN, false>(a, grid, smem, st);
}

static int conv_prepare(ffb_conv *op, cudaStream_t st)
{
    conv_pick(op);
    op->taps = op->fs * op->fs * (op->ic / op->groups);
    op->row  = FFB_ALIGN(op->taps, 4) + 4;
    op->fn_pad = FFB_ALIGN(op->fn, 4);
    op->tc = NULL;
    if (op->ig) { ig_plan_destroy(op->ig); op->ig = NULL; }
    if (op->kind == CK_PW_FFMA) {
        pw_plan(op);
        if (op->pw_mode != 1) {
            op->tc = pw_tc_plan_create(op->ic, op->fn, op->act, op->pw_mode);
            if (op->tc) op->kind = CK_PW_TC;
        }
        /* the FFMA kernel keeps the whole [K][N] weight matrix in shared memory and maps N over <= 256 threads: layers too
           large for that (and for a resident-weight tcgen05 plan) take the implicit-GEMM tcgen05 kernel or the generic one */
        if (op->kind == CK_PW_FFMA && (op->smem > 227 * 1024 || op->NT > 256 || op->TY < 1)) op->kind = CK_GENERIC;
    }
    if (op->kind == CK_GENERIC && op->pw_mode != 1) {
        op->ig = ig_plan_create(op->ic, op->fn, op->fs, op->stride, op->pad, op->groups, op->act);
        if (op->ig) op->kind = CK_IGEMM_TC;
    }
    const char *names[] = { "conv_generic", "pw_ffma", "pw_tcgen05", "dw3x3_s1", "dw5x5_s1", "dw3x3_s2", "stem3x3_s2", "igemm_tcgen05_3xtf32" };
    snprintf(op->name, sizeof op->name, "%s", names[op->kind]);
    if (op->kind == CK_PW_TC) snprintf(op->name, sizeof op->name, "pw_tcgen05_%s", pw_tc_mode_name(op->tc));
    if (op->kind == CK_STEM) {
        if (!op->h_packed) { ffb_set_error("stem conv needs the host copy of its weights"); return -1; }
        for (int o = 0; o < 8; o++) {
            const float *r = op->h_packed + (size_t)o * op->row;
            for (int t = 0; t < 27; t++) op->stemw.w[t * 8 + o] = r[t];
            op->stemw.s[o] = r[op->row - 4]; op->stemw.b[o] = r[op->row - 3];
        }
    }
    if (op->kind == CK_GENERIC) return 0;
    if (op->kind == CK_IGEMM_TC) return ig_prepare(op->ig, op->d_packed, op->row, st);
    const size_t nfl = (size_t)op->taps * op->fn_pad + 2 * op->fn_pad;
    if (!op->d_prep) CK(cudaMalloc(&op->d_prep, nfl * sizeof(float)));
    float *wt = op->d_prep, *sc = wt + (size_t)op->taps * op->fn_pad, *bi = sc + op->fn_pad;
    k_prep_weights<<<grid_for((long)nfl, 256), 256, 0, st>>>(op->d_packed, op->row, op->fn, op->taps, op->fn_pad, wt, sc, bi);
    CK(cudaGetLastError());
    if (op->tc && pw_tc_prepare(op->tc, op->d_packed, op->row, st) != 0) return -1;
    return 0;
}

static void conv_release(ffb_conv *op)
{
    if (!op) return;
    if (op->tc) pw_tc_plan_destroy(op->tc);
    if (op->ig) ig_plan_destroy(op->ig);
    cudaFree(op->d_prep);
    cudaFree(op->d_owned_packed);
    delete op;
}

/* in: [n][ih][iw][ldi], out: [n][oh][ow][ldo] written at channel offset coff */
static int conv_run(ffb_conv *op, const float *in, int ldi, float *out, int ldo, int coff,
                    int n, int ih, int iw, cudaStream_t st, const float *res = nullptr, int ldr = 0, int act2 = 0)
{
    const int oh = conv_out_dim(ih, op->fs, op->pad, op->stride), ow = conv_out_dim(iw, op->fs, op->pad, op->stride);
    const float *wt = op->d_prep, *sc = wt ? wt + (size_t)op->taps * op->fn_pad : NULL, *bi = sc ? sc + op->fn_pad : NULL;
    /* conv-v6.c:499 geometry -> its 5x5 path drops kernel row 0 on output row oh-2 (422-441) */
    const bool v6_dw5 = op->pad == 2 && op->fs == 5 && op->stride == 1 && op->ic / op->groups == 1 && oh >= 4 && ow >= 4;
    const int skip = (v6_dw5 && !op->dw5_exact) ? oh - 2 : -1;
    ConvKind kind = op->kind;
    if ((kind == CK_DW_S1_3 || kind == CK_DW_S1_5 || kind == CK_DW3_S2) && (ldi != op->ic || ldo != op->fn || coff != 0)) kind = CK_GENERIC;
    if (kind == CK_STEM && (ldi != 4 || ldo != 8 || coff != 0)) kind = CK_GENERIC;
    if (kind == CK_IGEMM_TC && !ig_supports(op->ig, ldi, ldo, coff, ih, iw)) kind = CK_GENERIC;
    if (kind == CK_PW_TC && !pw_tc_supports(op->tc, ldo, coff)) kind = op->smem <= 227 * 1024 && op->NT <= 256 && op->TY >= 1 ? CK_PW_FFMA : CK_GENERIC;
    if (res && kind != CK_PW_TC && kind != CK_PW_FFMA) { ffb_set_error("fused shortcut needs a pointwise conv"); return -1; }
    switch (kind) {
    case CK_IGEMM_TC:
        return ig_run(op->ig, in, ldi, out, ldo, coff, n, ih, iw, st);
    case CK_PW_TC:
        return pw_tc_run(op->tc, in, ldi, out, ldo, coff, (long)n * ih * iw, st, res, ldr, act2);
    case CK_PW_FFMA: {
        PwArgs a; a.in = in; a.out = out; a.wt = wt; a.scale = sc; a.bias = bi; a.M = (long)n * ih * iw; a.K = op->ic; a.N = op->fn;
        a.ldi = ldi; a.ldo = ldo; a.coff = coff; a.BN = op->fn_pad; a.NT = op->NT; a.TY = op->TY; a.act = op->act;
        a.res = res; a.ldr = ldr; a.act2 = act2;
        /* the fused-shortcut variant keeps TM*TN/4 skip vectors live: cap TM at 4 so two CTAs still fit the register file */
        const int TM = (res && op->TM > 4) ? 4 : op->TM;
        const size_t smem = pw_smem(op->ic, op->fn_pad, TM, op->TY);
        const long tiles = (a.M + (long)TM * op->TY - 1) / ((long)TM * op->TY);
        const int per_sm = smem <= 100 * 1024 ? 2 : 1;
        const int grid = (int)std::max<long>(1, std::min<long>(tiles, (long)g_num_sms * per_sm));
        cudaError_t e = cudaSuccess;
        if      (TM == 8 && op->TN == 8) e = pw_launch<8, 8>(a, grid, smem, st);
        else if (TM == 4 && op->TN == 8) e = pw_launch<4, 8>(a, grid, smem, st);
        else if (TM == 2 && op->TN == 8) e = pw_launch<2, 8>(a, grid, smem, st);
        else if (TM == 1 && op->TN == 8) e = pw_launch<1, 8>(a, grid, smem, st);
        else if (TM == 8 && op->TN == 4) e = pw_launch<8, 4>(a, grid, smem, st);
        else if (TM == 4 && op->TN == 4) e = pw_launch<4, 4>(a, grid, smem, st);
        else if (TM == 2 && op->TN == 4) e = pw_launch<2, 4>(a, grid, smem, st);
        else                             e = pw_launch<1, 4>(a, grid, smem, st);
        CK(e);
        return 0; }
    case CK_DW_S1_3: case CK_DW_S1_5: case CK_DW3_S2:
        if (op->dw_mode == 0) {
            DwPlan pl; static ffb_smem_cfg cfg_s1, cfg_s2, cfg_5;
            DwArgs a; a.out = out; a.wt = wt; a.scale = sc; a.bias = bi; a.N = n; a.H = ih; a.W = iw; a.C = op->ic; a.OH = oh; a.OW = ow;
            a.act = op->act; a.skip_row0_at = skip;
            if (kind == CK_DW_S1_3 && dw_plan(op->ic, oh, ow, 3, 1, 1 << 20, &pl)) return dw_launch(k_dw3s1_tma, &cfg_s1, in, a, pl, st);
            if (kind == CK_DW3_S2 && dw_plan(op->ic, oh, ow, 3, 2, 1 << 20, &pl)) return dw_launch(k_dw3s2_tma, &cfg_s2, in, a, pl, st);
            if (kind == CK_DW_S1_5 && dw5_plan(op->ic, oh, ow, &pl)) return dw_launch(k_dw5s1_tma, &cfg_5, in, a, pl, st);
        }
        if (kind == CK_DW3_S2) {
            const int R = oh >= 40 ? 10 : oh;
            dim3 grid((ow * op->ic / 4 + 127) / 128, (oh + R - 1) / R, n);
            CK(launch_pdl(k_dw3_s2, grid, dim3(128), 0, st, in, out, wt, sc, bi, ih, iw, op->ic, oh, ow, R, op->act));
            return 0;
        }
        {
        const int R = ih >= 64 ? 16 : ih >= 32 ? 10 : ih;
        dim3 grid((iw * op->ic / 4 + 127) / 128, (ih + R - 1) / R, n);
        if (kind == CK_DW_S1_3) CK(launch_pdl(k_dw_s1<3>, grid, dim3(128), 0, st, in, out, wt, sc, bi, ih, iw, op->ic, R, op->act, -1));
        else                    CK(launch_pdl(k_dw_s1<5>, grid, dim3(128), 0, st, in, out, wt, sc, bi, ih, iw, op->ic, R, op->act, skip));
        return 0; }
    case CK_STEM: {
        dim3 grid((ow + 31) / 32, (oh + 7) / 8, n), block(32, 8);
        CK(launch_pdl(k_stem_f32<32, 8>, grid, block, 0, st, in, out, op->stemw, ih, iw, oh, ow, op->act));
        return 0; }
    default: {
        const long total = (long)n * oh * ow * op->fn;
        CK(launch_pdl(k_conv_generic, dim3(grid_for(total, 256, 16)), dim3(256), 0, st, in, out + coff, op->d_packed, n, ih, iw, op->ic, ldi, oh, ow, op->fn, ldo,
                      op->groups, op->pad, op->stride, op->fs, op->row, op->act, skip));
        return 0; }
    }
}

/* net_input fused into the stem: BGR u8 frames (frame size == net size) -> layer-0 output */
static int stem_run_u8(ffb_conv *op, const unsigned char *frames, int pitch, float *out, int n, int ih, int iw,
                       const float *mean, const float *norm, cudaStream_t st)
{
    const int oh = conv_out_dim(ih, 3, 1, 2), ow = conv_out_dim(iw, 3, 1, 2);
    /* two output pixels per thread with word-wise staging when rows can be read as aligned 32-bit words (FFCNN_STEM_X2=0: the
       one-pixel kernel) */
    static const int x2 = getenv("FFCNN_STEM_X2") ? atoi(getenv("FFCNN_STEM_X2")) : 1;
    if (x2 && iw % 4 == 0 && pitch % 4 == 0 && ((uintptr_t)frames & 3) == 0) {
        /* tile sweep (r2v, batch 256, ms): 16x8 threads 0.111 | 40x4 0.113 | 20x8 0.118 | 40x2 0.124 | 20x4 0.126 | 40x8 0.135; one-pixel kernel 0.140 */
        dim3 grid((ow + 31) / 32, (oh + 7) / 8, n), block(16, 8);
        CK(launch_pdl(k_stem_u8x2<16, 8>, grid, block, 0, st, (const uint8_t *)frames, pitch, out, op->stemw, ih, iw, oh, ow, op->act, mean[0], mean[1], mean[2], norm[0], norm[1], norm[2]));
        return 0;
    }
    dim3 grid((ow + 31) / 32, (oh + 7) / 8, n), block(32, 8);
    CK(launch_pdl(k_stem_u8<32, 8>, grid, block, 0, st, (const uint8_t *)frames, pitch, out, op->stemw, ih, iw, oh, ow, op->act, mean[0], mean[1], mean[2], norm[0], norm[1], norm[2]));
    return 0;
}

/* =================================================================================== engine */

struct Tens { int buf = -1; float *p = nullptr; int h = 0, w = 0, c = 0, ld = 0, coff = 0;   /* coff: channel offset inside a wider (concat) buffer */
              size_t frame_floats() const { return (size_t)h * w * ld; } };

struct Buf { size_t floats = 0, offset = 0; int first = 0, last = 0; };

struct ffb_engine {
    ffb_net *net = nullptr;
    int device = 0, max_batch = 0, batch = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    float *d_packed = nullptr;
    std::vector<ffb_conv *> convs;          /* per layer, NULL for non-conv */
    std::vector<Tens> outs;                 /* per layer output (aliases share buf) */
    std::vector<int> fuse_sc;               /* conv layer i also performs shortcut layer fuse_sc[i] (-1: none) */
    std::vector<char> fused_away;           /* shortcut layer j is computed inside its producer conv */
    int fuse_shortcut = 1;
    /* fused inverted-residual blocks (block_mma.cu): expand conv `first`, depthwise first+1, projection first+2 and, when
       sc >= 0, the shortcut layer sc run as ONE kernel launched in the slot of layer `first` */
    struct Block { int first, sc; BlkPlan *plan; RegPlan *reg; Blk2Plan *tc2; };      /* exactly one of plan (block_mma.cu) / reg (block_reg.cu) / tc2 (block_tc.cu) is set */
    std::vector<Block> blocks;
    std::vector<int> blk_at;                /* per layer: index into blocks if the layer is a block's first conv, else -1 */
    std::vector<char> in_block;             /* per layer: computed inside a block (its own launch slot is empty) */
    int fuse_block = 1, blk2 = 0;           /* blk2: use the all-tcgen05 block kernel (block_tc.cu) where it has a plan */
    /* tail fusions: the SPP block (three max pools of one tensor + their route) as one kernel launched in the route's
       slot, and upsample layers that write straight into the concat tensor of the route that reads them */
    struct Spp { int route, x, r[3], off[3], offx; };
    std::vector<Spp> spps;
    std::vector<int> spp_at, up_into;       /* per layer: index into spps (route layers) / the route an upsample writes into, else -1 */
    std::vector<char> in_spp;               /* pool layers computed by the SPP kernel */
    int fuse_tail = 1, cand_cap = 0;
    /* side branch (option fork_tail): the chain of convs that ends in a yolo head which is not the graph's last layer (L116-L121 of
       yolo-fastest-1.1: five small kernels on 10x10 maps, each a single under-filled wave) runs on a second stream next to the layers
       that follow it in program order (the upsample -> ... -> second head chain), forked after the layer both read and joined at the
       end of the forward pass; inside the captured graph that is a fork / join of two kernel chains */
    int fork_tail = 1, side_a = -1, side_b = -1;
    /* stem + first block as one kernel (stem_block.cuh) when net_input is fused into the stem and L1-L3 is the 8->8->4 register block */
    int fuse_stem = 1; bool stem_block_ok = false;
    cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_side = nullptr;
    Tens input;
    std::vector<Buf> bufs;
    float *d_arena = nullptr; size_t arena_floats = 0;
    int dw5_exact = 0, pw_mode = 0, dw_mode = 0, use_graph = 1, keep_all = 0, fuse_input = 1;
    /* u8 frames of the current batch when net_input is fused into the stem (frame size == net size) */
    const unsigned char *u8_src = nullptr; int u8_pitch = 0; bool input_fused = false; float in_mean[3] = {0, 0, 0}, in_norm[3] = {0, 0, 0};
    bool plan_dirty = true;
    cudaGraphExec_t gexec = nullptr; int graph_batch = -1;
    int launches = 0;
    /* input staging */
    unsigned char *d_frames = nullptr; size_t d_frames_cap = 0;
    float *h_stage = nullptr; size_t h_stage_cap = 0;
    /* detection */
    /* one candidate set per pipeline slot: while the host decodes batch i from one, the GPU may already be filtering the batches queued behind it into the others */
    struct DetSet { Candidate *d_cand = nullptr, *h_cand = nullptr; int *d_count = nullptr, *h_count = nullptr, *h_count_dev = nullptr; int cap = 0, n = 0, s1 = 1, s2 = 1; long want = 0;
                    cudaEvent_t done = nullptr; } det[FFB_SLOTS];
    int det_cur = 0;
    cudaStream_t d2h_stream = nullptr;      /* candidate read-back must not queue behind the next batch's forward pass */
    bool slot_enqueued[FFB_SLOTS] = {};
    std::vector<std::vector<BBOX>> boxes, raw;
    int s1 = 1, s2 = 1; size_t d2h_bytes = 0;
    /* submit/collect pipeline: FFB_SLOTS device frame slots filled on a copy stream while the earlier batches compute */
    cudaStream_t copy_stream2 = nullptr; cudaEvent_t ev_join = nullptr; int h2d_chunks = 1;    /* FFCNN_H2D_CHUNKS: split each batch copy over two copy streams */
    cudaStream_t copy_stream = nullptr; unsigned char *d_slot[FFB_SLOTS] = {}; size_t slot_cap[FFB_SLOTS] = {};
    cudaEvent_t ev_copied[FFB_SLOTS] = {}, ev_free[FFB_SLOTS] = {};
    long submitted = 0, collected = 0;
    struct SlotMeta { int n, w, h, pitch; float mean[3], norm[3]; bool has_mean, has_norm; } slot_meta[FFB_SLOTS];
    /* L2 flush scratch for ffb_layer_times */
    float *d_flush = nullptr; size_t flush_floats = 0;
};

static ffb_engine *engine_of(NET *net)
{
    if (!net) { ffb_set_error("NULL net"); return nullptr; }
    ffb_net *fn = ffb_from_pub(net);
    if (fn->magic != FFB_MAGIC) { ffb_set_error("NET was not created by this library"); return nullptr; }
    if (!fn->engine) { ffb_set_error("net has no GPU engine (ffb_net_attach not called or failed)"); return nullptr; }
    return fn->engine;
}

int ffb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static void engine_free_plan(ffb_engine *e)
{
    if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
    e->graph_batch = -1;
    cudaFree(e->d_arena); e->d_arena = nullptr; e->arena_floats = 0;
    e->bufs.clear();
}

void ffb_engine_destroy(ffb_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    engine_free_plan(e);
    for (ffb_conv *c : e->convs) conv_release(c);
    for (ffb_engine::Block &b : e->blocks) { blk_plan_destroy(b.plan); reg_plan_destroy(b.reg); blk2_plan_destroy(b.tc2); }
    cudaFree(e->d_packed); cudaFree(e->d_frames); cudaFree(e->d_flush);
    cudaFreeHost(e->h_stage);
    for (ffb_engine::DetSet &d : e->det) { cudaFree(d.d_cand); cudaFree(d.d_count); cudaFreeHost(d.h_cand); cudaFreeHost(d.h_count); if (d.done) cudaEventDestroy(d.done); }
    if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
    for (int i = 0; i < FFB_SLOTS; i++) { cudaFree(e->d_slot[i]); if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]); if (e->ev_free[i]) cudaEventDestroy(e->ev_free[i]); }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->copy_stream2) cudaStreamDestroy(e->copy_stream2);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side_stream) cudaStreamDestroy(e->side_stream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_side) cudaEventDestroy(e->ev_side);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

/* through dropout / single-input route aliases: the layer that really produced what layer k outputs */
static int producer_of(const NET *net, int k)
{
    while (k >= 0 && (net->layer_list[k].type == LAYER_TYPE_DROPOUT || (net->layer_list[k].type == LAYER_TYPE_ROUTE && net->layer_list[k].depend_num == 1)))
        k = net->layer_list[k].type == LAYER_TYPE_DROPOUT ? k - 1 : net->layer_list[k].depend_list[0];
    return k;
}

/* how many non-alias layers read each layer's output (the static form of the reference's refcounts, ffcnn.c:481-487) */
static std::vector<int> count_readers(const NET *net)
{
    const int L = net->layer_num;
    std::vector<int> readers(L, 0);
    for (int k = 0; k < L; k++) {
        const LAYER *l = net->layer_list + k;
        if (l->type == LAYER_TYPE_DROPOUT || (l->type == LAYER_TYPE_ROUTE && l->depend_num == 1)) continue;
        if (l->type != LAYER_TYPE_ROUTE && k > 0) { const int p = producer_of(net, k - 1); if (p >= 0) readers[p]++; }
        for (int d = 0; d < l->depend_num; d++) { const int p = producer_of(net, l->depend_list[d]); if (p >= 0) readers[p]++; }
    }
    return readers;
}

/* Block fusion: find every  1x1 conv -> 3x3 depthwise conv -> 1x1 conv [-> dropout -> shortcut from the block
 * input]  chain whose intermediate tensors have no other reader, and plan one fused kernel for it (block_mma.cu). */
static int engine_find_blocks(ffb_engine *e)
{
    NET *net = &e->net->pub; const int L = net->layer_num;
    for (ffb_engine::Block &b : e->blocks) { blk_plan_destroy(b.plan); reg_plan_destroy(b.reg); blk2_plan_destroy(b.tc2); }
    e->blocks.clear(); e->blk_at.assign(L, -1); e->in_block.assign(L, 0);
    if (!e->fuse_block) return 0;
    const std::vector<int> readers = count_readers(net);
    auto is_pw = [](const LAYER *l) { return l->type == LAYER_TYPE_CONV && l->fs == 1 && l->stride == 1 && l->groups == 1 && l->pad == 0; };
    for (int i = 0; i + 2 < L; i++) {
        const LAYER *a = net->layer_list + i, *d = a + 1, *p = a + 2;
        if (!is_pw(a) || !is_pw(p) || d->type != LAYER_TYPE_CONV) continue;
        if (!(d->fs == 3 && d->pad == 1 && d->groups == d->c && d->fn == d->c && (d->stride == 1 || d->stride == 2))) continue;
        if (readers[i] != 1 || readers[i + 1] != 1) continue;
        int sc = -1, act_res = 0;
        for (int j = i + 3; j < L && j <= i + 4; j++) {
            const LAYER *sl = net->layer_list + j;
            if (sl->type == LAYER_TYPE_DROPOUT) continue;
            if (sl->type == LAYER_TYPE_SHORTCUT && i > 0 && readers[i + 2] == 1 && producer_of(net, j - 1) == i + 2 &&
                producer_of(net, sl->depend_list[0]) == producer_of(net, i - 1) && d->stride == 1 && a->c == p->fn) { sc = j; act_res = sl->activation; }
            break;
        }
        /* two fused kernels: the register-resident one (block_reg.cu) for the 160x160 blocks with <= 24 expanded channels,
           the shared-memory / tensor-core one (block_mma.cu) for the rest.  fuse_block 1 (default) uses the latter only for
           the block shapes where it beats the three separate layers on a B200 (measured: every inverted-residual chain of
           yolo-fastest-1.1 since round 2s -- the stride-2 136-channel block L81-L83 lost by 10 % in round 1 and wins by 3 % now,
           0.0518 vs 0.0533 ms); 2: every supported block; 3: block_mma only */
        BlkPlan *plan = nullptr; RegPlan *reg = nullptr; Blk2Plan *tc2 = nullptr;
        if (e->fuse_block != 3)
            reg = reg_plan_create(a->c, a->fn, p->fn, d->stride, a->h, a->w, a->activation, d->activation, p->activation, sc >= 0, act_res,
                                  a->filter, d->filter, p->filter);
        if (!reg && e->blk2) {
            const bool wanted2 = e->blk2 >= 2 || e->fuse_block >= 2 || a->fn >= 32;
            if (wanted2) tc2 = blk2_plan_create(a->c, a->fn, p->fn, d->stride, a->h, a->w, a->activation, d->activation, p->activation, sc >= 0, act_res);
            if (tc2 && blk2_prepare(tc2, e->convs[i]->d_packed, e->convs[i + 1]->d_packed, e->convs[i + 2]->d_packed, e->stream) != 0) { blk2_plan_destroy(tc2); return -1; }
        }
        if (!reg && !tc2) {
            const bool wanted = e->fuse_block >= 2 || a->fn == 32 || a->fn == 48 || a->fn == 96 || a->fn == 136 || a->fn == 224;
            if (!wanted) continue;
            plan = blk_plan_create(a->c, a->fn, p->fn, d->stride, a->h, a->w, a->activation, d->activation, p->activation, sc >= 0, act_res);
            if (!plan) continue;
            if (blk_prepare(plan, e->convs[i]->d_packed, e->convs[i + 1]->d_packed, e->convs[i + 2]->d_packed, e->stream) != 0) { blk_plan_destroy(plan); return -1; }
        }
        e->blk_at[i] = (int)e->blocks.size(); e->in_block[i + 1] = e->in_block[i + 2] = 1;
        if (sc >= 0) e->in_block[sc] = 1;
        e->blocks.push_back({ i, sc, plan, reg, tc2 });
        if (getenv("FFCNN_BLK_VERBOSE")) fprintf(stderr, "ffcnn_b200: block L%d-L%d: %s\n", i, sc >= 0 ? sc : i + 2, tc2 ? blk2_describe(tc2) : plan ? blk_describe(plan) : reg_describe(reg));
        i += 2;
    }
    return 0;
}

/* ---- activation arena: one buffer per produced tensor, reused once its last reader has run ---- */
static int engine_plan(ffb_engine *e)
{
    NET *net = &e->net->pub; const int L = net->layer_num;
    engine_free_plan(e);
    e->outs.assign(L, Tens());
    std::vector<Buf> &bufs = e->bufs;
    auto new_buf = [&](size_t frame_floats, int first) {
        Buf b; b.floats = FFB_ALIGN(frame_floats * (size_t)e->max_batch, 64); b.first = first; b.last = first;
        bufs.push_back(b); return (int)bufs.size() - 1;
    };
    const LAYER *l0 = net->layer_list;
    e->input.h = l0->h; e->input.w = l0->w; e->input.c = l0->c; e->input.ld = FFB_ALIGN(l0->c, 4);
    e->input.buf = new_buf(e->input.frame_floats(), -1);
    auto in_of = [&](int i) -> Tens & { return i == 0 ? e->input : e->outs[i - 1]; };
    auto touch = [&](const Tens &t, int at) { if (t.buf >= 0) bufs[t.buf].last = std::max(bufs[t.buf].last, at); };
    /* shortcut fusion (the cheap half): a pointwise conv whose only reader -- through dropout aliases -- is a
       shortcut layer adds the skip tensor in its own epilogue and writes the shortcut's output; the conv's own output and
       the separate add kernel disappear (2 of the 4 tensor passes).  Off under keep_all (every layer output must exist). */
    e->fuse_sc.assign(L, -1); e->fused_away.assign(L, 0);
    const bool fuse = e->keep_all != 1;                         /* keep_all 1: every layer materialised by its own kernel; 2: fused plan, no buffer reuse */
    const bool use_blocks = fuse && e->fuse_block && (int)e->blk_at.size() == L;
    if (e->fuse_shortcut && fuse) {
        const std::vector<int> readers = count_readers(net);
        for (int j = 1; j < L; j++) {
            const LAYER *sl = net->layer_list + j;
            if (sl->type != LAYER_TYPE_SHORTCUT || (use_blocks && e->in_block[j])) continue;
            const int p = producer_of(net, j - 1), d = sl->depend_list[0];
            if (p < 0 || producer_of(net, d) == p || (use_blocks && (e->in_block[p] || e->blk_at[p] >= 0))) continue;
            const ffb_conv *op = net->layer_list[p].type == LAYER_TYPE_CONV ? e->convs[p] : nullptr;
            if (!op || (op->kind != CK_PW_FFMA && op->kind != CK_PW_TC) || readers[p] != 1 || net->layer_list[p + 1].c % 4) continue;
            /* the skip tensor must have the conv output's shape (the unfused shortcut kernel checks the same and fails) */
            const LAYER *po = net->layer_list + p + 1, *so = net->layer_list + d + 1;
            if (d < 0 || d >= L || so->w != po->w || so->h != po->h || so->c != po->c) continue;
            e->fuse_sc[p] = j; e->fused_away[j] = 1;
        }
    }
    e->spps.clear(); e->spp_at.assign(L, -1); e->up_into.assign(L, -1); e->in_spp.assign(L, 0);
    std::vector<int> up_coff(L, 0), route_buf(L, -1);
    if (fuse && e->fuse_tail) {
        const std::vector<int> readers = count_readers(net);
        for (int j = 1; j < L; j++) {
            const LAYER *rl = net->layer_list + j;
            if (rl->type != LAYER_TYPE_ROUTE || rl->depend_num < 2) continue;
            /* SPP: exactly three stride-1 odd max pools of X plus X itself, each pool read by nothing else */
            int pools[4], npool = 0, xdep = -1, coff[4], acc = 0; bool ok = rl->depend_num == 4;
            for (int d = 0; d < rl->depend_num && ok; d++) {
                const int q = producer_of(net, rl->depend_list[d]); const LAYER *ql = net->layer_list + q;
                coff[d] = acc; acc += net->layer_list[q + 1].c;
                if (ql->type == LAYER_TYPE_MAXPOOL && ql->stride == 1 && (ql->fs & 1) && readers[q] == 1 && ql->c % 4 == 0) pools[npool++] = d;
                else if (xdep < 0) xdep = d; else ok = false;
            }
            if (ok && npool == 3 && xdep >= 0) {
                const int X = producer_of(net, rl->depend_list[xdep]);
                ffb_engine::Spp sp; sp.route = j; sp.x = X; sp.offx = coff[xdep];
                for (int a = 0; a < 3 && ok; a++) {
                    const int q = producer_of(net, rl->depend_list[pools[a]]);
                    if (q < 1 || producer_of(net, q - 1) != X) ok = false;
                    sp.r[a] = (net->layer_list[q].fs - 1) / 2; sp.off[a] = coff[pools[a]];
                }
                for (int a = 0; a < 3; a++) for (int b = a + 1; b < 3; b++) if (sp.r[b] < sp.r[a]) { std::swap(sp.r[a], sp.r[b]); std::swap(sp.off[a], sp.off[b]); }
                if (ok && sp.r[0] < sp.r[1] && sp.r[1] < sp.r[2]) {
                    e->spp_at[j] = (int)e->spps.size(); e->spps.push_back(sp);
                    for (int a = 0; a < 3; a++) e->in_spp[producer_of(net, rl->depend_list[pools[a]])] = 1;
                    continue;
                }
            }
            /* an upsample read only by this route writes its pixels straight into the concat tensor */
            acc = 0;
            for (int d = 0; d < rl->depend_num; d++) {
                const int q = producer_of(net, rl->depend_list[d]); const LAYER *ql = net->layer_list + q;
                if (ql->type == LAYER_TYPE_UPSAMPLE && readers[q] == 1 && acc % 4 == 0 && ql->c % 4 == 0 && e->up_into[q] < 0) { e->up_into[q] = j; up_coff[q] = acc; }
                acc += net->layer_list[q + 1].c;
            }
        }
    }
    int blk_buf = -1;                                           /* output buffer of the block being walked through */
    for (int i = 0; i < L; i++) {
        const LAYER *il = net->layer_list + i, *ol = il + 1;
        Tens &o = e->outs[i];
        o.h = ol->h; o.w = ol->w; o.c = ol->c; o.ld = FFB_ALIGN(ol->c, 4);
        if (use_blocks && e->blk_at[i] >= 0) {
            /* the fused kernel runs in this layer's slot: it reads the block input and writes the projection's
               (or the shortcut's) tensor; the two expanded tensors are never materialised */
            const LAYER *pl = net->layer_list + i + 3;           /* geometry of the projection's output */
            blk_buf = new_buf((size_t)pl->h * pl->w * FFB_ALIGN(pl->c, 4), i);
            touch(in_of(i), i);
            continue;                                            /* o stays without a buffer */
        }
        if (use_blocks && e->in_block[i]) {
            if (il->type == LAYER_TYPE_CONV && net->layer_list[i].fs == 1) { o.buf = blk_buf; touch(o, i); }   /* projection: the block's output */
            else if (il->type == LAYER_TYPE_SHORTCUT) { o = in_of(i); touch(o, i); }                           /* fused shortcut: same tensor */
            continue;                                            /* depthwise: no buffer */
        }
        if (e->in_spp[i]) { touch(in_of(i), i); continue; }      /* pool inside the SPP kernel: never materialised */
        if (e->up_into[i] >= 0) {
            /* the route's concat buffer is created here, at its first writer; this layer's tensor is a channel window of it */
            const int j = e->up_into[i]; const LAYER *jo = net->layer_list + j + 1;
            if (route_buf[j] < 0) route_buf[j] = new_buf((size_t)jo->h * jo->w * FFB_ALIGN(jo->c, 4), i);
            o.buf = route_buf[j]; o.ld = FFB_ALIGN(jo->c, 4); o.coff = up_coff[i];
            touch(in_of(i), i); touch(o, i);
            continue;
        }
        if (il->type == LAYER_TYPE_CONV && e->fuse_sc[i] >= 0) {
            /* the conv writes the shortcut's tensor: one buffer, created now, named by both layers */
            o.buf = new_buf(o.frame_floats(), i); touch(in_of(i), i); touch(e->outs[net->layer_list[e->fuse_sc[i]].depend_list[0]], i); touch(o, i);
            continue;
        }
        if (il->type == LAYER_TYPE_SHORTCUT && e->fused_away[i]) { o = in_of(i); touch(o, i); continue; }
        switch (il->type) {
        case LAYER_TYPE_DROPOUT: o = in_of(i); break;                               /* alias */
        case LAYER_TYPE_ROUTE:
            if (il->depend_num == 1) { o = e->outs[il->depend_list[0]]; break; }   /* alias */
            o.buf = route_buf[i] >= 0 ? route_buf[i] : new_buf(o.frame_floats(), i);
            for (int d = 0; d < il->depend_num; d++) touch(e->outs[il->depend_list[d]], i);
            if (e->spp_at[i] >= 0) touch(e->outs[e->spps[e->spp_at[i]].x], i);
            break;
        case LAYER_TYPE_YOLO: o = Tens(); touch(in_of(i), 1 << 30); break;          /* heads stay alive for ffb_detect */
        case LAYER_TYPE_SHORTCUT:
            o.buf = new_buf(o.frame_floats(), i); touch(in_of(i), i); touch(e->outs[il->depend_list[0]], i); break;
        default:
            o.buf = new_buf(o.frame_floats(), i); touch(in_of(i), i); break;
        }
        if (o.buf >= 0) touch(o, i);
    }
    /* side branch: walk back from the first yolo head that has layers after it, through plain convs whose output nobody else reads;
       the layer where the walk stops because its output has a second reader is the fork point.  Its layers run concurrently with
       everything after the head in program order, so none of the buffers they touch may be recycled before the pass ends. */
    e->side_a = e->side_b = -1;
    if (fuse && e->fork_tail) {
        const std::vector<int> readers = count_readers(net);
        for (int y = 2; y + 1 < L && e->side_a < 0; y++) {
            if (net->layer_list[y].type != LAYER_TYPE_YOLO || net->layer_list[y + 1].type != LAYER_TYPE_ROUTE) continue;
            auto plain = [&](int k) {
                return k >= 1 && net->layer_list[k].type == LAYER_TYPE_CONV && e->blk_at[k] < 0 && !e->in_block[k] && e->fuse_sc[k] < 0 &&
                       net->layer_list[k - 1].type != LAYER_TYPE_ROUTE && net->layer_list[k - 1].type != LAYER_TYPE_DROPOUT;
            };
            int a = y;
            while (plain(a - 1) && (a == y || readers[a - 1] == 1)) a--;       /* layer a - 1 joins while its output feeds only layer a (the head feeds the yolo layer) */
            if (a < y && plain(a) && readers[a - 1] > 1 && y - a >= 2) {
                bool closed = true;                                              /* nothing after the head reads a tensor of the branch */
                for (int j = y + 1; j < L && closed; j++) {
                    const LAYER *jl = net->layer_list + j;
                    if (jl->type != LAYER_TYPE_ROUTE && j > 0) { const int q = producer_of(net, j - 1); if (q >= a && q < y) closed = false; }
                    for (int d = 0; d < jl->depend_num; d++) { const int q = producer_of(net, jl->depend_list[d]); if (q >= a && q < y) closed = false; }
                }
                if (closed) { e->side_a = a; e->side_b = y; }
            }
        }
        if (e->side_a >= 0)
            for (int i = e->side_a; i <= e->side_b; i++) {
                if (e->outs[i].buf >= 0) bufs[e->outs[i].buf].last = 1 << 30;
                if (in_of(i).buf >= 0) bufs[in_of(i).buf].last = 1 << 30;
            }
    }
    e->stem_block_ok = false;
    if (use_blocks && e->fuse_stem && L > 4 && e->blk_at[1] >= 0 && e->blocks[e->blk_at[1]].reg && e->blocks[e->blk_at[1]].sc < 0 &&
        net->layer_list[0].type == LAYER_TYPE_CONV && e->convs[0] && e->convs[0]->kind == CK_STEM && e->outs[0].ld == 8 && e->outs[3].ld == 4) {
        const std::vector<int> readers = count_readers(net);
        e->stem_block_ok = readers[0] == 1;                     /* nobody but the block reads the stem's output */
    }
    if (e->keep_all) for (Buf &b : bufs) b.last = 1 << 30;
    /* greedy best-fit over a free list, in creation order */
    struct Free { size_t off, len; };
    std::vector<Free> fl; size_t top = 0;
    std::vector<int> order(bufs.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::vector<int> live;
    for (int bi : order) {
        Buf &b = bufs[bi];
        /* release buffers whose last reader ran strictly before this buffer is first written */
        for (size_t k = 0; k < live.size();) {
            Buf &lb = bufs[live[k]];
            if (lb.last < b.first) {
                fl.push_back({ lb.offset, lb.floats });
                live.erase(live.begin() + k);
            } else k++;
        }
        std::sort(fl.begin(), fl.end(), [](const Free &x, const Free &y) { return x.off < y.off; });
        for (size_t k = 0; k + 1 < fl.size();) {
            if (fl[k].off + fl[k].len == fl[k + 1].off) { fl[k].len += fl[k + 1].len; fl.erase(fl.begin() + k + 1); } else k++;
        }
        int best = -1;
        for (size_t k = 0; k < fl.size(); k++) if (fl[k].len >= b.floats && (best < 0 || fl[k].len < fl[best].len)) best = (int)k;
        if (best >= 0) { b.offset = fl[best].off; fl[best].off += b.floats; fl[best].len -= b.floats; if (!fl[best].len) fl.erase(fl.begin() + best); }
        else { b.offset = top; top += b.floats; }
        live.push_back(bi);
    }
    e->arena_floats = top;
    CK(cudaMalloc(&e->d_arena, top * sizeof(float)));
    CK(cudaMemsetAsync(e->d_arena, 0, top * sizeof(float), e->stream));
    e->input.p = e->d_arena + bufs[e->input.buf].offset;
    for (Tens &t : e->outs) if (t.buf >= 0) t.p = e->d_arena + bufs[t.buf].offset + t.coff;
    e->plan_dirty = false;
    return 0;
}

static int engine_prepare_weights(ffb_engine *e)
{
    NET *net = &e->net->pub;
    for (int i = 0; i < net->layer_num; i++) {
        const LAYER *l = net->layer_list + i;
        if (l->type != LAYER_TYPE_CONV) continue;
        ffb_conv *op = e->convs[i];
        if (!op) {
            op = new ffb_conv(); memset(op, 0, sizeof *op);
            e->convs[i] = op;
        }
        op->ic = l->c; op->groups = l->groups; op->pad = l->pad; op->stride = l->stride; op->fs = l->fs; op->fn = l->fn;
        op->act = l->activation; op->dw5_exact = e->dw5_exact; op->pw_mode = e->pw_mode; op->dw_mode = e->dw_mode;
        op->d_packed = e->d_packed + (l->filter - net->weight_buf);
        op->h_packed = l->filter;
        if (op->tc) { pw_tc_plan_destroy(op->tc); op->tc = NULL; }
        if (conv_prepare(op, e->stream) != 0) return -1;
    }
    if (engine_find_blocks(e) != 0) return -1;
    e->plan_dirty = true;
    return 0;
}

static int engine_attach_body(ffb_engine *e, NET *net);

int ffb_net_attach(NET *net, int device, int max_batch)
{
    if (!net) { ffb_set_error("NULL net"); return -1; }
    ffb_net *fn = ffb_from_pub(net);
    if (fn->magic != FFB_MAGIC) { ffb_set_error("NET was not created by this library"); return -1; }
    if (max_batch < 1) max_batch = 1;
    int ndev = ffb_device_count();
    if (ndev <= 0) { ffb_set_error("no CUDA device available: libffcnn_b200 has no CPU fallback"); return -1; }
    if (device < 0 || device >= ndev) { ffb_set_error("device %d out of range (%d devices)", device, ndev); return -1; }
    ffb_engine *e = fn->engine;
    if (e && e->device == device) {
        if (max_batch > e->max_batch) { e->max_batch = max_batch; e->plan_dirty = true; }
        return 0;
    }
    if (e) { ffb_engine_destroy(e); fn->engine = nullptr; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { ffb_set_error("device %d is sm_%d%d; this build only carries sm_100a code", device, prop.major, prop.minor); return -1; }
    e = new ffb_engine(); e->net = fn; e->device = device; e->max_batch = max_batch;
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; ffb_set_error("cudaStreamCreate failed"); return -1; }
    e->stream = e->own_stream;
    if (cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_side, cudaEventDisableTiming) != cudaSuccess) { ffb_engine_destroy(e); ffb_set_error("side stream / events: creation failed"); return -1; }
    fn->engine = e;
    if (engine_attach_body(e, net) != 0) {                   /* never leave a half-initialised engine behind */
        ffb_engine_destroy(e); fn->engine = nullptr;
        return -1;
    }
    return 0;
}

static int engine_attach_body(ffb_engine *e, NET *net)
{
    e->convs.assign(net->layer_num, nullptr);
    const char *env;
    if ((env = getenv("FFCNN_DW5_EXACT"))) e->dw5_exact = atoi(env);
    if ((env = getenv("FFCNN_PW_MODE")))   e->pw_mode = atoi(env);
    if ((env = getenv("FFCNN_GRAPH")))     e->use_graph = atoi(env);
    if ((env = getenv("FFCNN_DW_MODE")))   e->dw_mode = atoi(env);
    if ((env = getenv("FFCNN_PDL")))       sm100::g_ffb_pdl = atoi(env);
    if ((env = getenv("FFCNN_FUSE_BLOCK"))) e->fuse_block = atoi(env);
    if ((env = getenv("FFCNN_FORK_TAIL"))) e->fork_tail = atoi(env);
    if ((env = getenv("FFCNN_FUSE_STEM"))) e->fuse_stem = atoi(env);
    if ((env = getenv("FFCNN_BLK2")))      e->blk2 = atoi(env);
    CK(cudaMalloc(&e->d_packed, std::max(1, net->weight_size) * sizeof(float)));
    CK(cudaMemcpyAsync(e->d_packed, net->weight_buf, (size_t)net->weight_size * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    if (engine_prepare_weights(e) != 0) return -1;
    for (ffb_engine::DetSet &d : e->det) {
        CK(cudaMalloc(&d.d_count, 2 * sizeof(int)));              /* [candidates, finished blocks of the last filter launch]: zero between uses (the last block resets them) */
        CK(cudaMemset(d.d_count, 0, 2 * sizeof(int)));
        CK(cudaMallocHost(&d.h_count, sizeof(int)));
        CK(cudaHostGetDevicePointer((void **)&d.h_count_dev, d.h_count, 0));      /* the last filter launch writes the count here */
        *d.h_count = 0;
        CK(cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming));
    }
    CK(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

NET *net_load(char *cfgfile, char *weightsfile, int inputw, int inputh)
{
    NET *net = ffb_net_parse(cfgfile, weightsfile, inputw, inputh);
    const char *dev = getenv("FFCNN_DEVICE");
    if (!net) { fprintf(stderr, "ffcnn_b200: net_load: %s\n", ffb_last_error()); return NULL; }
    if (ffb_net_attach(net, dev ? atoi(dev) : 0, 1) != 0) {
        fprintf(stderr, "ffcnn_b200: net_load: %s\n", ffb_last_error());
        net_free(net);
        return NULL;
    }
    return net;
}

void *ffb_packed_weights_device(NET *net, size_t *nfloats)
{
    ffb_engine *e = engine_of(net);
    if (!e) return NULL;
    if (nfloats) *nfloats = (size_t)net->weight_size;
    return e->d_packed;
}

int ffb_commit_weights(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    CK(cudaSetDevice(e->device));
    /* keep the host copy coherent with what was broadcast into the device buffer */
    CK(cudaMemcpyAsync(net->weight_buf, e->d_packed, (size_t)net->weight_size * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (engine_prepare_weights(e) != 0) return -1;
    CK(cudaStreamSynchronize(e->stream));
    if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; e->graph_batch = -1; }
    return 0;
}

static bool stem_block_active(const ffb_engine *e);

int ffb_set_option(NET *net, const char *name, int value)
{
    ffb_engine *e = engine_of(net);
    if (!e || !name) return -1;
    bool reweight = false;
    if      (!strcmp(name, "dw5_exact")) { reweight = e->dw5_exact != value; e->dw5_exact = value; }
    else if (!strcmp(name, "pw_mode"))   { reweight = e->pw_mode != value; e->pw_mode = value; }
    else if (!strcmp(name, "dw_mode"))   { reweight = e->dw_mode != value; e->dw_mode = value; }
    else if (!strcmp(name, "graph"))     { e->use_graph = value; }
    else if (!strcmp(name, "fuse_input")) { e->fuse_input = value; }
    else if (!strcmp(name, "fuse_shortcut")) { if (e->fuse_shortcut != value) e->plan_dirty = true; e->fuse_shortcut = value; }
    else if (!strcmp(name, "keep_all"))  { if (e->keep_all != value) e->plan_dirty = true; e->keep_all = value; }
    else if (!strcmp(name, "fuse_block")) { reweight = e->fuse_block != value; e->fuse_block = value; }
    else if (!strcmp(name, "blk2"))      { reweight = e->blk2 != value; e->blk2 = value; }
    else if (!strcmp(name, "fuse_tail")) { if (e->fuse_tail != value) e->plan_dirty = true; e->fuse_tail = value; }
    else if (!strcmp(name, "fork_tail")) { if (e->fork_tail != value) e->plan_dirty = true; e->fork_tail = value; }
    else if (!strcmp(name, "fuse_stem")) { if (e->fuse_stem != value) e->plan_dirty = true; e->fuse_stem = value; }
    else if (!strcmp(name, "cand_cap")) { e->cand_cap = value; for (ffb_engine::DetSet &d : e->det) d.want = 0; }   /* test hook: initial candidate capacity */
    else { ffb_set_error("unknown option '%s'", name); return -1; }
    if (reweight) {
        CK(cudaSetDevice(e->device));
        if (engine_prepare_weights(e) != 0) return -1;
        CK(cudaStreamSynchronize(e->stream));
    }
    if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; e->graph_batch = -1; }
    return 0;
}

int ffb_get_option(NET *net, const char *name)
{
    ffb_engine *e = engine_of(net);
    if (!e || !name) return -1;
    if (!strcmp(name, "dw5_exact")) return e->dw5_exact;
    if (!strcmp(name, "pw_mode")) return e->pw_mode;
    if (!strcmp(name, "dw_mode")) return e->dw_mode;
    if (!strcmp(name, "graph")) return e->use_graph;
    if (!strcmp(name, "fuse_input")) return e->fuse_input;
    if (!strcmp(name, "input_fused")) return e->input_fused ? 1 : 0;
    if (!strcmp(name, "fuse_shortcut")) return e->fuse_shortcut;
    if (!strcmp(name, "keep_all")) return e->keep_all;
    if (!strcmp(name, "fuse_block")) return e->fuse_block;
    if (!strcmp(name, "blk2")) return e->blk2;
    if (!strcmp(name, "fuse_tail")) return e->fuse_tail;
    if (!strcmp(name, "fork_tail")) return e->fork_tail;
    if (!strcmp(name, "fuse_stem")) return e->fuse_stem;
    if (!strcmp(name, "stem_block")) return stem_block_active(e) ? 1 : 0;
    if (!strcmp(name, "side_branch")) return e->side_a >= 0 ? e->side_a * 1000 + e->side_b : -1;
    if (!strcmp(name, "blocks")) return (int)e->blocks.size();
    if (!strcmp(name, "max_batch")) return e->max_batch;
    if (!strcmp(name, "batch")) return e->batch;
    if (!strcmp(name, "device")) return e->device;
    if (!strcmp(name, "arena_mb")) return (int)(e->arena_floats * sizeof(float) >> 20);
    return -1;
}

int ffb_set_stream(NET *net, void *s)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    cudaStreamSynchronize(e->stream);
    e->stream = s ? (cudaStream_t)s : e->own_stream;
    if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; e->graph_batch = -1; }
    return 0;
}
void *ffb_get_stream(NET *net) { ffb_engine *e = engine_of(net); return e ? (void *)e->stream : NULL; }

int ffb_sync(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

static int engine_ready(ffb_engine *e, int n)
{
    CK(cudaSetDevice(e->device));
    if (n > e->max_batch) { e->max_batch = n; e->plan_dirty = true; }
    if (e->plan_dirty || !e->d_arena) { CK(cudaStreamSynchronize(e->stream)); if (engine_plan(e) != 0) return -1; }
    return 0;
}

/* ---- input ---- */
int ffb_input_u8(NET *net, const unsigned char *frames, int n, int w, int h, int pitch,
                 const float *mean, const float *norm, int on_device)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (!frames || n < 1 || w < 1 || h < 1 || pitch < 3 * w) { ffb_set_error("ffb_input_u8: bad arguments"); return -1; }
    if (net->layer_list[0].c != 3) { ffb_set_error("ffb_input_u8 needs a 3-channel network input"); return -1; }
    if (engine_ready(e, n) != 0) return -1;
    static const float zero3[3] = { 0, 0, 0 }, unit3[3] = { 1 / 255.f, 1 / 255.f, 1 / 255.f };
    if (!mean) mean = zero3;
    if (!norm) norm = unit3;
    const unsigned char *src = frames;
    const size_t bytes = (size_t)n * h * pitch;
    if (!on_device) {
        if (bytes > e->d_frames_cap) {
            CK(cudaStreamSynchronize(e->stream));
            cudaFree(e->d_frames); e->d_frames = nullptr; e->d_frames_cap = 0;
            CK(cudaMalloc(&e->d_frames, bytes)); e->d_frames_cap = bytes;
        }
        CK(cudaMemcpyAsync(e->d_frames, frames, bytes, cudaMemcpyHostToDevice, e->stream));
        src = e->d_frames;
    }
    int sw, sh, s1, s2;
    ffb_fit_geometry(w, h, e->input.w, e->input.h, &sw, &sh, &s1, &s2);
    e->s1 = s1; e->s2 = s2; e->batch = n;
    net->s1 = s1; net->s2 = s2;
    for (int i = 0; i < 3; i++) { e->in_mean[i] = mean[i]; e->in_norm[i] = norm[i]; }
    /* no resize (frame == net size): the stem reads the u8 frames itself and the fp32 input tensor is never materialised */
    e->input_fused = e->fuse_input && e->keep_all != 1 && w == e->input.w && h == e->input.h && net->layer_num > 0 &&
                     net->layer_list[0].type == LAYER_TYPE_CONV && e->convs[0] && e->convs[0]->kind == CK_STEM && e->outs[0].ld == 8;
    e->u8_src = src; e->u8_pitch = pitch;
    if (e->input_fused) return 0;
    const long total = (long)n * e->input.h * e->input.w;
    CK(launch_pdl(k_input_u8, dim3(grid_for(total, 256, 16)), dim3(256), 0, e->stream, (const uint8_t *)src, e->input.p, n, w, h, pitch, e->input.w, e->input.h, sw, sh, s1, s2,
                  mean[0], mean[1], mean[2], norm[0], norm[1], norm[2]));
    return 0;
}

int ffb_input_chw(NET *net, const float *chw, int n, int s1, int s2)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (!chw || n < 1) { ffb_set_error("ffb_input_chw: bad arguments"); return -1; }
    if (engine_ready(e, n) != 0) return -1;
    const Tens &t = e->input;
    const size_t fl = (size_t)n * t.frame_floats();
    CK(cudaStreamSynchronize(e->stream));                 /* staging buffer may still be in flight */
    if (fl > e->h_stage_cap) { cudaFreeHost(e->h_stage); e->h_stage = nullptr; CK(cudaMallocHost(&e->h_stage, fl * sizeof(float))); e->h_stage_cap = fl; }
    const size_t plane = (size_t)t.h * t.w;
    for (int f = 0; f < n; f++)
        for (size_t p = 0; p < plane; p++)
            for (int c = 0; c < t.ld; c++)
                e->h_stage[((size_t)f * plane + p) * t.ld + c] = c < t.c ? chw[((size_t)f * t.c + c) * plane + p] : 0.f;
    CK(cudaMemcpyAsync(t.p, e->h_stage, fl * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    e->batch = n; e->s1 = s1; e->s2 = s2; net->s1 = s1; net->s2 = s2;
    e->input_fused = false;
    return 0;
}

/* stem + first block in one kernel: the plan allows it and this batch's frames are read by the stem directly */
static bool stem_block_active(const ffb_engine *e)
{
    return e->stem_block_ok && e->input_fused && e->keep_all != 1 && e->fuse_block && e->blk_at.size() > 1 && e->blk_at[1] >= 0 &&
           reg_stem_ok(e->blocks[e->blk_at[1]].reg, e->input.h, e->input.w, e->u8_pitch, e->u8_src);
}

/* ---- the layer loop ---- */
static int run_layer(ffb_engine *e, int i, cudaStream_t st, int *launches)
{
    NET *net = &e->net->pub; const LAYER *il = net->layer_list + i;
    const Tens &in = i == 0 ? e->input : e->outs[i - 1]; const Tens &o = e->outs[i];
    const int n = e->batch;
    if (e->keep_all != 1 && e->fuse_block && (int)e->blk_at.size() > i) {
        if (e->in_block[i]) return 0;                          /* computed by the block kernel launched in an earlier slot */
        if (e->blk_at[i] >= 0) {
            const ffb_engine::Block &b = e->blocks[e->blk_at[i]]; const Tens &y = e->outs[i + 2];
            if (i == 1 && stem_block_active(e)) return 0;      /* computed together with the stem in layer 0's slot */
            if ((b.tc2 ? blk2_run(b.tc2, in.p, in.ld, y.p, y.ld, n, st) : b.plan ? blk_run(b.plan, in.p, in.ld, y.p, y.ld, n, st) : reg_run(b.reg, in.p, in.ld, y.p, y.ld, n, st)) != 0) return -1;
            *launches += b.reg ? reg_launches(b.reg) : 1;
            return 0;
        }
    }
    switch (il->type) {
    case LAYER_TYPE_CONV:
        if (i == 0 && stem_block_active(e)) {
            const ffb_engine::Block &b = e->blocks[e->blk_at[1]]; const Tens &y = e->outs[3];
            if (reg_run_stem(b.reg, &e->convs[0]->stemw, e->convs[0]->act, e->u8_src, e->u8_pitch, y.p, y.ld, n, in.h, in.w, e->in_mean, e->in_norm, st) != 0) return -1;
        } else if (i == 0 && e->input_fused) {
            if (stem_run_u8(e->convs[0], e->u8_src, e->u8_pitch, o.p, n, in.h, in.w, e->in_mean, e->in_norm, st) != 0) return -1;
        } else if (e->fuse_sc[i] >= 0) {
            const LAYER *sl = net->layer_list + e->fuse_sc[i]; const Tens &r = e->outs[sl->depend_list[0]];
            if (conv_run(e->convs[i], in.p, in.ld, o.p, o.ld, 0, n, in.h, in.w, st, r.p, r.ld, sl->activation) != 0) return -1;
        } else if (conv_run(e->convs[i], in.p, in.ld, o.p, o.ld, 0, n, in.h, in.w, st) != 0) return -1;
        (*launches)++;
        break;
    case LAYER_TYPE_MAXPOOL: case LAYER_TYPE_AVGPOOL:
        if (e->keep_all != 1 && (int)e->in_spp.size() > i && e->in_spp[i]) break;      /* computed by the SPP kernel in the route's slot */
        if (in.c % 4)
            CK(launch_pdl(k_pool_scalar, dim3(grid_for((long)n * o.h * o.w * o.c, 256)), dim3(256), 0, st, (const float *)in.p, o.p, n, in.h, in.w, in.c, in.ld, o.h, o.w, o.ld, 0,
                          il->fs, il->stride, (int)(il->type == LAYER_TYPE_MAXPOOL)));
        else
        CK(launch_pdl(k_pool, dim3(grid_for((long)n * o.h * o.w * (o.c / 4), 128)), dim3(128), 0, st, (const float *)in.p, o.p, n, in.h, in.w, in.c, in.ld, o.h, o.w, o.ld, 0,
                      il->fs, il->stride, (int)(il->type == LAYER_TYPE_MAXPOOL)));
        (*launches)++;
        break;
    case LAYER_TYPE_UPSAMPLE:
        if (in.c % 4)
            CK(launch_pdl(k_upsample_scalar, dim3(grid_for((long)n * o.h * o.w * o.c, 256, 64)), dim3(256), 0, st, (const float *)in.p, o.p, n, in.h, in.w, in.c, in.ld, o.ld, 0, il->stride));
        else
        CK(launch_pdl(k_upsample, dim3(std::min(n * in.h, g_num_sms * 16)), dim3(std::min(256, (in.w * (in.c / 4) + 31) / 32 * 32)), 0, st, (const float *)in.p, o.p, n, in.h, in.w, in.c, in.ld, o.ld, 0, il->stride));
        (*launches)++;
        break;
    case LAYER_TYPE_SHORTCUT: {
        if (e->fused_away[i]) break;                           /* done in the producer conv's epilogue */
        const Tens &s = e->outs[il->depend_list[0]];
        if (s.h != in.h || s.w != in.w || s.c != in.c) { ffb_set_error("shortcut layer %d: shape mismatch", i); return -1; }
        const long n4 = (long)n * o.frame_floats() / 4;       /* ld is a multiple of 4 and identical for all three */
        CK(launch_pdl(k_shortcut, dim3(grid_for(n4, 256)), dim3(256), 0, st, (const float *)in.p, (const float *)s.p, o.p, n4, il->activation));
        (*launches)++;
        break; }
    case LAYER_TYPE_ROUTE:
        if (il->depend_num > 1 && e->keep_all != 1 && (int)e->spp_at.size() > i && e->spp_at[i] >= 0) {
            const ffb_engine::Spp &sp = e->spps[e->spp_at[i]]; const Tens &x = e->outs[sp.x];
            const size_t spp_bytes = (size_t)4 * x.h * x.w * x.c * sizeof(float);
            static ffb_smem_cfg spp_configured;
            if (spp_bytes <= 200 * 1024) {                     /* frame fits in shared memory: separable version, one CTA per frame */
                if (ffb_ensure_smem((const void *)k_spp_smem, spp_bytes, &spp_configured) != 0) return -1;
                CK(launch_pdl(k_spp_smem, dim3(n), dim3(256), spp_bytes, st, (const float *)x.p, o.p, x.h, x.w, x.c, x.ld, o.ld,
                              sp.r[0], sp.r[1], sp.r[2], sp.off[0], sp.off[1], sp.off[2], sp.offx));
            } else
                CK(launch_pdl(k_spp, dim3(grid_for((long)n * x.h * x.w * (x.c / 4), 128, 16)), dim3(128), 0, st, (const float *)x.p, o.p, n, x.h, x.w, x.c, x.ld, o.ld,
                              sp.r[0], sp.r[1], sp.r[2], sp.off[0], sp.off[1], sp.off[2], sp.offx));
            (*launches)++;
        } else if (il->depend_num > 1) {
            int coff = 0;
            for (int d = 0; d < il->depend_num; d++) {
                const Tens &s = e->outs[il->depend_list[d]];
                if (s.buf == o.buf && s.buf >= 0) { coff += s.c; continue; }          /* written in place by its producer (upsample into concat) */
                const long px = (long)n * s.h * s.w;
                if (s.c % 4 == 0 && coff % 4 == 0) CK(launch_pdl(k_concat, dim3(grid_for(px * (s.c / 4), 256)), dim3(256), 0, st, (const float *)s.p, o.p, px, s.c, s.ld, o.ld, coff));
                else CK(launch_pdl(k_copy_strided, dim3(grid_for(px * s.c, 256)), dim3(256), 0, st, (const float *)s.p, o.p, px, s.c, s.ld, o.ld, coff));
                coff += s.c; (*launches)++;
            }
        }
        break;
    default: break;                                            /* dropout, yolo: nothing to launch */
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { ffb_set_error("layer %d launch failed: %s", i, cudaGetErrorString(err)); return -1; }
    return 0;
}

static int run_all(ffb_engine *e, cudaStream_t st, int first = 0)
{
    int launches = first;
    const bool side = e->side_a > first && e->side_stream;
    for (int i = first; i < e->net->pub.layer_num; i++) {
        cudaStream_t s = st;
        if (side && i >= e->side_a && i <= e->side_b) {
            if (i == e->side_a) { CK(cudaEventRecord(e->ev_fork, st)); CK(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0)); }
            s = e->side_stream;
        }
        if (run_layer(e, i, s, &launches) != 0) return -1;
    }
    if (side) { CK(cudaEventRecord(e->ev_side, e->side_stream)); CK(cudaStreamWaitEvent(st, e->ev_side, 0)); }
    e->launches = launches;
    return 0;
}

int ffb_forward(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (e->batch < 1 || !e->d_arena || e->plan_dirty) { ffb_set_error("ffb_forward: no input set for the current plan"); return -1; }
    CK(cudaSetDevice(e->device));
    if (!e->use_graph) return run_all(e, e->stream);
    /* with net_input fused into the stem, layer 0 reads the caller's frame pointer, which changes from batch to batch:
       it is launched eagerly and the graph covers layers 1.. */
    const int first = e->input_fused ? 1 : 0;
    const int gkey = e->batch * 2 + first;
    if (!e->gexec || e->graph_batch != gkey) {
        if (e->gexec) { cudaGraphExecDestroy(e->gexec); e->gexec = nullptr; }
        /* one eager pass first: lazily-set function attributes must not happen inside capture */
        if (run_all(e, e->stream) != 0) return -1;
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int rc = run_all(e, e->stream, first);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc != 0 || ce != cudaSuccess) { if (g) cudaGraphDestroy(g); if (ce != cudaSuccess) ffb_set_error("graph capture failed: %s", cudaGetErrorString(ce)); return -1; }
        ce = cudaGraphInstantiate(&e->gexec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { ffb_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); e->gexec = nullptr; return -1; }
        e->graph_batch = gkey;
        return 0;                                               /* the eager pass above already produced this batch's outputs */
    }
    if (first) { int l = 0; if (run_layer(e, 0, e->stream, &l) != 0) return -1; }
    CK(cudaGraphLaunch(e->gexec, e->stream));
    return 0;
}

int ffb_launches_per_forward(NET *net) { ffb_engine *e = engine_of(net); return e ? e->launches : -1; }

/* ---- detection ---- */
struct Head { int layer, cells, gw, gh, key_base; };

static std::vector<Head> yolo_heads(ffb_engine *e, int *total_keys)
{
    NET *net = &e->net->pub; std::vector<Head> heads; int key = 0;
    for (int i = 1; i < net->layer_num; i++) if (net->layer_list[i].type == LAYER_TYPE_YOLO) {
        const Tens &t = e->outs[i - 1];
        heads.push_back({ i, t.h * t.w, t.w, t.h, key }); key += t.h * t.w * 3;
    }
    if (total_keys) *total_keys = key;
    return heads;
}

/* GPU half: candidate filter kernels + async copy of the candidate count. No synchronisation. */
int ffb_detect_enqueue(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    CK(cudaSetDevice(e->device));
    const int n = e->batch;
    if (n < 1) { ffb_set_error("ffb_detect: no batch"); return -1; }
    int key = 0; std::vector<Head> heads = yolo_heads(e, &key);
    /* Candidate capacity: one slot per (frame, cell, anchor) -- the filter can then never overflow -- unless that exceeds
       64 MB (huge inputs x big batches): then 2048 per frame, and an overflow is reported by ffb_detect_finish, which also
       records the size that would have been needed so the next enqueue allocates it (ffb_detect retries by itself).
       The reference keeps at most bbox_max boxes per frame in scan order (ffcnn.c:453); the host decode applies that cap. */
    ffb_engine::DetSet &d = e->det[e->det_cur];
    const long full = (long)key * n;
    long cap_l = full * (long)sizeof(Candidate) <= (64L << 20) ? full : std::max<long>(4096, 2048L * n);
    cap_l = std::min<long>(std::max<long>(std::max<long>(cap_l, d.want), 4096), full > 4096 ? full : 4096);
    if (e->cand_cap > 0 && !d.want) cap_l = e->cand_cap;
    const int cap = (int)cap_l;
    if (cap != d.cap && (cap > d.cap || e->cand_cap > 0)) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(d.d_cand); cudaFreeHost(d.h_cand); d.d_cand = nullptr; d.h_cand = nullptr;
        CK(cudaMalloc(&d.d_cand, (size_t)cap * sizeof(Candidate)));
        CK(cudaMallocHost(&d.h_cand, (size_t)cap * sizeof(Candidate)));
        d.cap = cap;
    }
    d.n = n; d.s1 = e->s1; d.s2 = e->s2;
    if (heads.empty()) *d.h_count = 0;                         /* no yolo layer: nothing will write the count */
    size_t hk = 0;
    for (const Head &h : heads) {
        const LAYER *yl = net->layer_list + h.layer; const Tens &t = e->outs[h.layer - 1];
        if (t.c != 3 * (5 + yl->class_num)) {
            cudaMemsetAsync(d.d_count, 0, 2 * sizeof(int), e->stream);       /* an earlier head may have counted: leave the set clean */
            ffb_set_error("yolo layer %d: %d channels, expected %d", h.layer, t.c, 3 * (5 + yl->class_num)); return -1;
        }
        const long threads = (long)n * h.cells;                 /* one thread per grid cell */
        CK(launch_pdl(k_yolo_filter, dim3((int)((threads + 255) / 256)), dim3(256), 0, e->stream, (const float *)t.p, n, h.cells, t.ld, yl->class_num, 0, h.key_base,
                      yl->ignore_thres, d.d_cand, d.d_count, d.cap, ++hk == heads.size() ? d.h_count_dev : (int *)nullptr));
    }
    CK(cudaEventRecord(d.done, e->stream));
    return (int)heads.size();
}

/* Host half: wait, fetch the candidates, exact decode (reference scan order) and per-frame NMS. */
int ffb_detect_finish(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    CK(cudaSetDevice(e->device));
    ffb_engine::DetSet &d = e->det[e->det_cur];
    const int n = d.n;
    std::vector<Head> heads = yolo_heads(e, nullptr);
    CK(cudaEventSynchronize(d.done));                      /* this batch only: a look-ahead batch may already be queued behind it */
    if (*d.h_count > d.cap) {
        /* the filter found more candidates than the buffer holds: which ones were dropped depends on atomic order, so the
           result would not be the reference's -- fail loudly instead of returning a truncated box set */
        d.want = (long)*d.h_count + (*d.h_count >> 2);
        e->boxes.assign(n, std::vector<BBOX>()); e->raw.assign(n, std::vector<BBOX>());
        ffb_set_error("yolo candidate overflow: %d candidates, capacity %d (the next ffb_detect_enqueue allocates %ld)", *d.h_count, d.cap, d.want);
        return FFB_E_OVERFLOW;
    }
    int cnt = *d.h_count;
    e->d2h_bytes = sizeof(int) + (size_t)cnt * sizeof(Candidate);
    if (cnt > 0) {
        CK(cudaMemcpyAsync(d.h_cand, d.d_cand, (size_t)cnt * sizeof(Candidate), cudaMemcpyDeviceToHost, e->d2h_stream));
        CK(cudaStreamSynchronize(e->d2h_stream));
    }
    Candidate *h_cand = d.h_cand;
    std::sort(h_cand, h_cand + cnt, [](const Candidate &a, const Candidate &b) { return a.frame != b.frame ? a.frame < b.frame : a.key < b.key; });
    e->boxes.assign(n, std::vector<BBOX>()); e->raw.assign(n, std::vector<BBOX>());
    const int netw = net->layer_list[0].w, neth = net->layer_list[0].h;
    /* exact decode (the reference's double-precision exp arithmetic, host_decode.c) + NMS of the frames [f0, f1), whose candidates are
       h_cand[k0, k1): frames are independent, every frame's vectors are touched by one caller only */
    auto decode_range = [&](int k0, int k1, int f0, int f1) {
        for (int k = k0; k < k1; k++) {
            const Candidate &c = h_cand[k];
            if (c.frame < 0 || c.frame >= n) continue;
            const Head *hd = nullptr;
            for (const Head &h : heads) if (c.key >= h.key_base && c.key < h.key_base + h.cells * 3) hd = &h;
            if (!hd) continue;
            ffb_candidate hc; hc.frame = c.frame; hc.key = c.key; hc.cls = c.cls; hc.bs = c.bs; hc.cs = c.cs; hc.tx = c.tx; hc.ty = c.ty; hc.tw = c.tw; hc.th = c.th;
            BBOX b; const int rel = c.key - hd->key_base;
            if ((int)e->raw[c.frame].size() < net->bbox_max &&
                ffb_decode_candidate(net->layer_list + hd->layer, netw, neth, hd->gw, hd->gh, rel / 3, rel % 3, &hc, &b)) e->raw[c.frame].push_back(b);
        }
        for (int f = f0; f < f1; f++) {
            e->boxes[f] = e->raw[f];
            const int m = ffb_nms(e->boxes[f].data(), (int)e->boxes[f].size(), 0.5f, 1, d.s1, d.s2);
            e->boxes[f].resize(m);
        }
    };
    /* A batch of picture frames brings thousands of candidates, six libm exp() each: 0.76 ms per 256 frames on one core
       (tools/e2e_probe.py).  Large batches are split at frame boundaries over a few short-lived threads; the result is the same
       vectors in the same order.  Small candidate sets (and FFCNN_DECODE_THREADS=1) stay on the calling thread. */
    static const int max_threads = [] { const char *v = getenv("FFCNN_DECODE_THREADS"); const int t = v ? atoi(v) : 4; return t < 1 ? 1 : t > 16 ? 16 : t; }();
    const int T = (cnt >= 2048 && n >= 8) ? std::min(max_threads, n / 4) : 1;
    if (T <= 1) { decode_range(0, cnt, 0, n); return 0; }
    std::vector<int> ks(T + 1), fs(T + 1);
    ks[0] = 0; fs[0] = 0; ks[T] = cnt; fs[T] = n;
    for (int t = 1; t < T; t++) {
        int k = std::max(ks[t - 1], (int)((long)cnt * t / T));
        while (k > 0 && k < cnt && h_cand[k].frame == h_cand[k - 1].frame) k++;          /* to the next frame boundary */
        ks[t] = k;
        fs[t] = k < cnt ? std::min(std::max(h_cand[k].frame, fs[t - 1]), n) : n;       /* candidates are sorted by frame */
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) {
        try { pool.emplace_back(decode_range, ks[t], ks[t + 1], fs[t], fs[t + 1]); }
        catch (...) { decode_range(ks[t], ks[t + 1], fs[t], fs[t + 1]); }          /* no thread to be had: this range on the calling thread (nothing may unwind through the C ABI) */
    }
    decode_range(ks[0], ks[1], fs[0], fs[1]);
    for (std::thread &th : pool) th.join();
    return 0;
}

int ffb_detect(NET *net)
{
    if (ffb_detect_enqueue(net) < 0) return -1;
    int rc = ffb_detect_finish(net);
    if (rc == FFB_E_OVERFLOW) {                                /* heads are still resident: filter again into the grown buffer */
        if (ffb_detect_enqueue(net) < 0) return -1;
        rc = ffb_detect_finish(net);
    }
    return rc;
}

long ffb_last_d2h_bytes(NET *net) { ffb_engine *e = engine_of(net); return e ? (long)e->d2h_bytes : -1; }

int ffb_boxes(NET *net, int frame, BBOX **boxes)
{
    ffb_engine *e = engine_of(net);
    if (!e || frame < 0 || frame >= (int)e->boxes.size()) { if (e) ffb_set_error("ffb_boxes: frame %d out of range", frame); return -1; }
    if (boxes) *boxes = e->boxes[frame].data();
    return (int)e->boxes[frame].size();
}

int ffb_raw_boxes(NET *net, int frame, BBOX **boxes)
{
    ffb_engine *e = engine_of(net);
    if (!e || frame < 0 || frame >= (int)e->raw.size()) { if (e) ffb_set_error("ffb_raw_boxes: frame %d out of range", frame); return -1; }
    if (boxes) *boxes = e->raw[frame].data();
    return (int)e->raw[frame].size();
}

int ffb_detect_batch_u8(NET *net, const unsigned char *frames_host, int n, int w, int h, int pitch,
                        const float *mean, const float *norm)
{
    if (ffb_input_u8(net, frames_host, n, w, h, pitch, mean, norm, 0) != 0) return -1;
    if (ffb_forward(net) != 0) return -1;
    return ffb_detect(net);
}

/* ---- pipelined end-to-end path: H2D of batch i+1 overlaps the forward pass and the host decode of batch i ---- */
int ffb_submit_u8(NET *net, const unsigned char *frames_host, int n, int w, int h, int pitch, const float *mean, const float *norm)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (!frames_host || n < 1 || w < 1 || h < 1 || pitch < 3 * w) { ffb_set_error("ffb_submit_u8: bad arguments"); return -1; }
    if (e->submitted - e->collected >= FFB_SLOTS) { ffb_set_error("ffb_submit_u8: %d batches already in flight, call ffb_collect first", FFB_SLOTS); return -1; }
    CK(cudaSetDevice(e->device));
    if (!e->copy_stream) {
        CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < FFB_SLOTS; i++) { CK(cudaEventCreateWithFlags(&e->ev_copied[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&e->ev_free[i], cudaEventDisableTiming)); }
        const char *ch = getenv("FFCNN_H2D_CHUNKS");
        e->h2d_chunks = ch ? std::max(1, std::min(64, atoi(ch))) : 1;
        if (e->h2d_chunks > 1) { CK(cudaStreamCreateWithFlags(&e->copy_stream2, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming)); }
    }
    const int slot = (int)(e->submitted % FFB_SLOTS);
    const size_t bytes = (size_t)n * h * pitch;
    if (bytes > e->slot_cap[slot]) {
        /* sized for the largest batch the engine is planned for, so a caller that changes its batch size from one submit to the
           next (rate-proportional shards) never comes back here: the reallocation synchronises both streams */
        const size_t cap = std::max(bytes, (size_t)std::max(n, e->max_batch) * h * pitch);
        CK(cudaStreamSynchronize(e->stream)); CK(cudaStreamSynchronize(e->copy_stream));
        cudaFree(e->d_slot[slot]); e->d_slot[slot] = nullptr; e->slot_cap[slot] = 0;
        CK(cudaMalloc(&e->d_slot[slot], cap)); e->slot_cap[slot] = cap;
    }
    if (e->submitted >= FFB_SLOTS) CK(cudaStreamWaitEvent(e->copy_stream, e->ev_free[slot], 0));     /* the batch that used this slot has consumed it */
    if (e->h2d_chunks > 1) {
        /* several smaller copies alternating over two streams keep two DMA transfers in flight (developer knob; measured in
           profiles/r2h_e2e_8gpu.txt) */
        if (e->submitted >= FFB_SLOTS) CK(cudaStreamWaitEvent(e->copy_stream2, e->ev_free[slot], 0));
        const size_t chunk = ((bytes + e->h2d_chunks - 1) / e->h2d_chunks + 4095) & ~(size_t)4095;
        int k = 0;
        for (size_t off = 0; off < bytes; off += chunk, k++)
            CK(cudaMemcpyAsync(e->d_slot[slot] + off, frames_host + off, std::min(chunk, bytes - off), cudaMemcpyHostToDevice, (k & 1) ? e->copy_stream2 : e->copy_stream));
        CK(cudaEventRecord(e->ev_join, e->copy_stream2));
        CK(cudaStreamWaitEvent(e->copy_stream, e->ev_join, 0));
    } else
        CK(cudaMemcpyAsync(e->d_slot[slot], frames_host, bytes, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(e->ev_copied[slot], e->copy_stream));
    ffb_engine::SlotMeta &m = e->slot_meta[slot];
    m.n = n; m.w = w; m.h = h; m.pitch = pitch; m.has_mean = mean != nullptr; m.has_norm = norm != nullptr;
    for (int i = 0; i < 3; i++) { m.mean[i] = mean ? mean[i] : 0.f; m.norm[i] = norm ? norm[i] : 0.f; }
    e->submitted++;
    return 0;
}

/* enqueue everything batch `slot` needs on the GPU: wait for its copy, net_input + layer loop, candidate filter */
static int collect_enqueue(NET *net, ffb_engine *e, int slot)
{
    const ffb_engine::SlotMeta &m = e->slot_meta[slot];
    CK(cudaStreamWaitEvent(e->stream, e->ev_copied[slot], 0));
    if (ffb_input_u8(net, e->d_slot[slot], m.n, m.w, m.h, m.pitch, m.has_mean ? m.mean : nullptr, m.has_norm ? m.norm : nullptr, 1) != 0) return -1;
    if (ffb_forward(net) != 0) return -1;
    CK(cudaEventRecord(e->ev_free[slot], e->stream));
    e->det_cur = slot;
    if (ffb_detect_enqueue(net) < 0) return -1;
    e->slot_enqueued[slot] = true;
    return 0;
}

int ffb_collect(NET *net)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (e->collected >= e->submitted) { ffb_set_error("ffb_collect: nothing submitted"); return -1; }
    CK(cudaSetDevice(e->device));
    const int slot = (int)(e->collected % FFB_SLOTS);
    /* look-ahead: every submitted batch is queued (in order) before the host blocks on the oldest one, so the GPU goes straight
       from batch i to batch i+1 while the host decodes batch i (each batch's candidates sit in its own detection set).  With
       FFB_SLOTS = 3 the caller can keep two batches queued behind the one it collects: a host hiccup of up to one whole step
       (a scheduler tick, an NVML query taking the driver lock) no longer drains the GPU */
    for (long k = e->collected; k < e->submitted; k++) {
        const int s = (int)(k % FFB_SLOTS);
        if (!e->slot_enqueued[s] && collect_enqueue(net, e, s) != 0) return -1;
    }
    e->det_cur = slot;
    e->slot_enqueued[slot] = false;
    e->collected++;
    return ffb_detect_finish(net);
}

/* net_forward(): host CHW input tensor -> boxes, one frame (the reference's own entry point) */
int ffb_engine_forward_single(ffb_net *fn)
{
    NET *net = &fn->pub;
    const int s1 = net->s1 ? net->s1 : 1, s2 = net->s2 ? net->s2 : 1;
    if (ffb_input_chw(net, net->layer_list[0].data, 1, s1, s2) != 0) return -1;
    if (ffb_forward(net) != 0) return -1;
    if (ffb_detect(net) != 0) return -1;
    BBOX *b = nullptr; int m = ffb_boxes(net, 0, &b);
    if (m < 0) return -1;
    if (m > net->bbox_max) m = net->bbox_max;
    memset(net->bbox_list, 0, sizeof(BBOX) * (size_t)net->bbox_max);
    if (m) memcpy(net->bbox_list, b, sizeof(BBOX) * (size_t)m);
    net->bbox_num = m;
    return 0;
}

/* ---- inspection ---- */
long ffb_layer_output(NET *net, int layer, int frame, float *chw, long capacity)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (layer < -1 || layer >= net->layer_num || frame < 0 || frame >= e->batch) { ffb_set_error("ffb_layer_output: bad layer/frame"); return -1; }
    const Tens &t = layer < 0 ? e->input : e->outs[layer];
    /* under the default plan buffers are recycled: only tensors that stay alive to the end (the yolo heads) are readable */
    if (!e->keep_all && !(t.buf >= 0 && e->bufs[t.buf].last >= (1 << 30))) { ffb_set_error("ffb_layer_output: layer %d is not kept by the default plan (set option keep_all=1 or 2 before ffb_forward)", layer); return -1; }
    if (!t.p) return 0;
    const long count = (long)t.c * t.h * t.w;
    if (!chw) return count;
    if (capacity < count) { ffb_set_error("ffb_layer_output: buffer too small"); return -1; }
    CK(cudaSetDevice(e->device));
    std::vector<float> tmp(t.frame_floats());
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(tmp.data(), t.p + (size_t)frame * t.frame_floats(), (tmp.size() - t.coff) * sizeof(float), cudaMemcpyDeviceToHost));
    const size_t plane = (size_t)t.h * t.w;
    for (int c = 0; c < t.c; c++) for (size_t p = 0; p < plane; p++) chw[c * plane + p] = tmp[p * t.ld + c];
    return count;
}

static thread_local bool g_cost_recursing = false;
int ffb_layer_cost(NET *net, int i, double *bytes, double *flops, char *kname, int cap)
{
    if (!net || i < 0 || i >= net->layer_num) { ffb_set_error("ffb_layer_cost: bad layer"); return -1; }
    const LAYER *a = net->layer_list + i, *b = a + 1;
    const double in = (double)a->w * a->h * a->c, out = (double)b->w * b->h * b->c;
    double by = 0, fl = 0; const char *nm = "alias";
    switch (a->type) {
    case LAYER_TYPE_CONV: {
        const double taps = (double)a->fs * a->fs * (a->c / a->groups);
        by = 4 * (in + out + a->fn * (taps + 2)); fl = 2 * taps * out; nm = "conv"; break; }
    case LAYER_TYPE_MAXPOOL: case LAYER_TYPE_AVGPOOL: by = 4 * (in + out); fl = out * a->fs * a->fs; nm = "pool"; break;
    case LAYER_TYPE_UPSAMPLE: by = 4 * (in + out); nm = "upsample"; break;
    case LAYER_TYPE_SHORTCUT: by = 4 * 3 * out; fl = out; nm = "shortcut"; break;
    case LAYER_TYPE_ROUTE: by = 4 * 2 * out; nm = a->depend_num > 1 ? "concat" : "alias"; break;
    case LAYER_TYPE_YOLO: by = 4 * in; nm = "yolo_filter"; break;
    default: break;
    }
    ffb_net *fn = ffb_from_pub(net);
    if (a->type == LAYER_TYPE_CONV && fn->engine && fn->engine->convs[i]) nm = fn->engine->convs[i]->name;
    if (fn->engine && fn->engine->keep_all != 1 && (int)fn->engine->spp_at.size() > i && !g_cost_recursing) {
        ffb_engine *e = fn->engine;
        if (e->in_spp[i]) { by = 0; fl = 0; nm = "in_spp"; }
        else if (e->spp_at[i] >= 0) {                          /* the SPP kernel: its three pools + the route, reported in the route's slot */
            g_cost_recursing = true;
            for (int k = 0; k < net->layer_num; k++) if (e->in_spp[k]) { double b2 = 0, f2 = 0; ffb_layer_cost(net, k, &b2, &f2, NULL, 0); by += b2; fl += f2; }
            g_cost_recursing = false;
            nm = "spp_fused";
        }
    }
    if (fn->engine && fn->engine->keep_all != 1 && fn->engine->fuse_block && (int)fn->engine->blk_at.size() > i && !g_cost_recursing) {
        /* a fused block is reported in its first layer's slot with the summed (unfused) algorithmic cost of its layers */
        ffb_engine *e = fn->engine;
        if (e->in_block[i] || (i == 1 && stem_block_active(e))) { by = 0; fl = 0; nm = "in_block"; }
        else if (i == 0 && stem_block_active(e)) {             /* stem + block L1-L3: reported in the stem's slot */
            g_cost_recursing = true;
            for (int k = 1; k <= 3; k++) { double b2 = 0, f2 = 0; ffb_layer_cost(net, k, &b2, &f2, NULL, 0); by += b2; fl += f2; }
            g_cost_recursing = false;
            nm = "block_stem_reg_fp32";
        }
        else if (e->blk_at[i] >= 0) {
            const ffb_engine::Block &b = e->blocks[e->blk_at[i]];
            g_cost_recursing = true;
            for (int k = i + 1; k <= (b.sc >= 0 ? b.sc : i + 2); k++) { double b2 = 0, f2 = 0; ffb_layer_cost(net, k, &b2, &f2, NULL, 0); by += b2; fl += f2; }
            g_cost_recursing = false;
            nm = b.tc2 ? "block_tcgen05_3xtf32" : b.plan ? (blk_uses_tcgen05(b.plan) ? "block_mma_tcgen05_3xtf32" : "block_mma_3xtf32") : "block_reg_fp32";
        }
    }
    if (bytes) *bytes = by;
    if (flops) *flops = fl;
    if (kname && cap > 0) snprintf(kname, (size_t)cap, "%s", nm);
    return 0;
}

int ffb_layer_times(NET *net, float *ms, int nlayers, int reps, int flush_l2)
{
    ffb_engine *e = engine_of(net);
    if (!e) return -1;
    if (e->batch < 1 || !e->d_arena || e->plan_dirty) { ffb_set_error("ffb_layer_times: run ffb_forward first"); return -1; }
    CK(cudaSetDevice(e->device));
    if (reps < 1) reps = 1;
    if (flush_l2 && !e->d_flush) { e->flush_floats = (size_t)64 << 20; CK(cudaMalloc(&e->d_flush, e->flush_floats * sizeof(float))); }
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int L = std::min(nlayers, net->layer_num);
    struct PdlOff { int saved; PdlOff() : saved(sm100::g_ffb_pdl) { sm100::g_ffb_pdl = 0; } ~PdlOff() { sm100::g_ffb_pdl = saved; } } pdl_off;   /* isolated-kernel times */
    for (int i = 0; i < L; i++) {
        int launches = 0;
        if (run_layer(e, i, e->stream, &launches) != 0) return -1;          /* untimed warm-up */
        if (!launches) { ms[i] = 0.f; continue; }
        float total = 0;
        if (flush_l2) {                                                       /* cold: one launch per event pair, L2 evicted before each */
            for (int r = 0; r < reps; r++) {
                k_fill<<<g_num_sms * 8, 256, 0, e->stream>>>(e->d_flush, (long)e->flush_floats, 0.f);
                CK(cudaEventRecord(a, e->stream));
                if (run_layer(e, i, e->stream, &launches) != 0) return -1;
                CK(cudaEventRecord(b, e->stream));
                CK(cudaEventSynchronize(b));
                float t = 0; CK(cudaEventElapsedTime(&t, a, b)); total += t;
            }
        } else {                                                              /* back to back: launch latency hidden, as inside the graph */
            CK(cudaEventRecord(a, e->stream));
            for (int r = 0; r < reps; r++) if (run_layer(e, i, e->stream, &launches) != 0) return -1;
            CK(cudaEventRecord(b, e->stream));
            CK(cudaEventSynchronize(b));
            CK(cudaEventElapsedTime(&total, a, b));
        }
        ms[i] = total / reps;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

/* =================================================================================== standalone op API */

static int ensure_device()
{
    if (ffb_device_count() <= 0) { ffb_set_error("no CUDA device available: libffcnn_b200 has no CPU fallback"); return -1; }
    return 0;
}

ffb_conv *ffb_conv_create(const float *packed, int ic, int groups, int pad, int stride, int fs, int fn, int act, int flags)
{
    if (!packed || ic < 1 || groups < 1 || fs < 1 || fn < 1 || stride < 1 || ic % groups || fn % groups) { ffb_set_error("ffb_conv_create: bad arguments"); return NULL; }
    if (ensure_device() != 0) return NULL;
    ffb_conv *op = new ffb_conv(); memset(op, 0, sizeof *op);
    op->ic = ic; op->groups = groups; op->pad = pad; op->stride = stride; op->fs = fs; op->fn = fn; op->act = act;
    op->dw5_exact = flags & 1; op->pw_mode = (flags >> 8) & 0xff; op->dw_mode = (flags >> 16) & 0xff;
    const int row = FFB_ALIGN(fs * fs * (ic / groups), 4) + 4;
    const size_t bytes = (size_t)fn * row * sizeof(float);
    if (cudaMalloc(&op->d_owned_packed, bytes) != cudaSuccess || cudaMemcpy(op->d_owned_packed, packed, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        ffb_set_error("ffb_conv_create: device allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError())); conv_release(op); return NULL;
    }
    op->d_packed = op->d_owned_packed;
    op->h_packed = packed;
    if (conv_prepare(op, 0) != 0 || cudaDeviceSynchronize() != cudaSuccess) { conv_release(op); return NULL; }
    return op;
}

void ffb_conv_destroy(ffb_conv *op) { conv_release(op); }
const char *ffb_conv_kernel_name(ffb_conv *op) { return op ? op->name : ""; }

int ffb_conv_run(ffb_conv *op, const float *in, float *out, int n, int ih, int iw, void *stream)
{
    if (!op || !in || !out) { ffb_set_error("ffb_conv_run: bad arguments"); return -1; }
    return conv_run(op, in, FFB_ALIGN(op->ic, 4), out, FFB_ALIGN(op->fn, 4), 0, n, ih, iw, (cudaStream_t)stream);
}

void *ffb_dev_alloc(size_t bytes) { void *p = NULL; if (cudaMalloc(&p, bytes) != cudaSuccess) { ffb_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError())); return NULL; } return p; }
void  ffb_dev_free(void *p) { cudaFree(p); }
int   ffb_copy_h2d(void *d, const void *s, size_t b) { CK(cudaMemcpy(d, s, b, cudaMemcpyHostToDevice)); return 0; }
int   ffb_copy_d2h(void *d, const void *s, size_t b) { CK(cudaMemcpy(d, s, b, cudaMemcpyDeviceToHost)); return 0; }
void *ffb_host_alloc_pinned_wc(size_t bytes) { void *p = NULL; if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) { ffb_set_error("cudaHostAlloc(write-combined) failed: %s", cudaGetErrorString(cudaGetLastError())); return NULL; } return p; }
void *ffb_host_alloc_pinned(size_t bytes) { void *p = NULL; if (cudaMallocHost(&p, bytes) != cudaSuccess) { ffb_set_error("cudaMallocHost failed: %s", cudaGetErrorString(cudaGetLastError())); return NULL; } return p; }
void  ffb_host_free_pinned(void *p) { cudaFreeHost(p); }

int ffb_chw_to_nhwc(const float *src, float *dst, int n, int c, int h, int w, void *stream)
{
    const int ld = FFB_ALIGN(c, 4);
    k_chw_to_nhwc<<<grid_for((long)n * h * w * ld, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, dst, n, c, h, w, ld);
    CK(cudaGetLastError());
    return 0;
}

int ffb_nhwc_to_chw(const float *src, float *dst, int n, int c, int h, int w, void *stream)
{
    const int ld = FFB_ALIGN(c, 4);
    k_nhwc_to_chw<<<grid_for((long)n * c * h * w, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, dst, n, c, h, w, ld);
    CK(cudaGetLastError());
    return 0;
}

/* =================================================================================== operator seam */

/* conv.h:4-7 backed by the GPU: host CHW in -> device NHWC -> kernel -> host CHW out. */
extern "C" void groupconv(float *datai, float *dataf, float *datao, int iw, int ih, int ic, int ig, int ipad, int istride,
                          int fs, int fn, int ow, int oh, int oc, int activation, float **gc_buffer, int *gc_bufsize)
{
    (void)gc_buffer; (void)gc_bufsize; (void)oc;
    const char *ex = getenv("FFCNN_DW5_EXACT"), *pm = getenv("FFCNN_PW_MODE");
    const char *dm = getenv("FFCNN_DW_MODE");
    const int flags = (ex && atoi(ex) ? 1 : 0) | ((pm ? atoi(pm) : 0) << 8) | ((dm ? atoi(dm) : 0) << 16);
    ffb_conv *op = ffb_conv_create(dataf, ic, ig, ipad, istride, fs, fn, activation, flags);
    if (!op) { fprintf(stderr, "ffcnn_b200: groupconv: %s\n", ffb_last_error()); return; }
    const int ldi = FFB_ALIGN(ic, 4), ldo = FFB_ALIGN(fn, 4);
    /* every sub-buffer starts 256-byte aligned (the kernels use 128-bit accesses) */
    const size_t in_chw = FFB_ALIGN((size_t)ic * ih * iw, 64), out_chw = FFB_ALIGN((size_t)fn * oh * ow, 64);
    const size_t in_n = FFB_ALIGN((size_t)ih * iw * ldi, 64), out_n = FFB_ALIGN((size_t)oh * ow * ldo, 64);
    float *d = NULL;
    if (cudaMalloc(&d, (in_chw + out_chw + in_n + out_n) * sizeof(float)) != cudaSuccess) {
        fprintf(stderr, "ffcnn_b200: groupconv: device allocation failed\n"); cudaGetLastError(); conv_release(op); return;
    }
    float *d_in_chw = d, *d_out_chw = d + in_chw, *d_in = d_out_chw + out_chw, *d_out = d_in + in_n;
    bool ok = cudaMemcpy(d_in_chw, datai, (size_t)ic * ih * iw * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess
           && cudaMemset(d_out, 0, out_n * sizeof(float)) == cudaSuccess
           && ffb_chw_to_nhwc(d_in_chw, d_in, 1, ic, ih, iw, 0) == 0
           && conv_run(op, d_in, ldi, d_out, ldo, 0, 1, ih, iw, 0) == 0
           && ffb_nhwc_to_chw(d_out, d_out_chw, 1, fn, oh, ow, 0) == 0
           && cudaMemcpy(datao, d_out_chw, (size_t)fn * oh * ow * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) fprintf(stderr, "ffcnn_b200: groupconv failed: %s / %s\n", ffb_last_error(), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    conv_release(op);
}
