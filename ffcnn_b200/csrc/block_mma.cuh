/*
 * block_mma.cuh -- one kernel per inverted-residual block of the graph (block fusion):
 *
 *     x --1x1 expand, BN, act--> e --3x3 depthwise (stride 1|2), BN, act--> d --1x1 project, BN, act--> [+ x] --> y
 *
 * i.e. three groupconv calls (conv-v6.c:46-91, 96-287) plus the optional dropout/shortcut pair (ffcnn.c:412-423) that
 * net_forward (ffcnn.c:476-520) runs as separate layers.  Unfused, the expanded tensors e and d (6x the channels of x
 * and y) make four trips through HBM; here they never leave the SM: a CTA owns a TH x TW tile of output pixels of one
 * frame, keeps x's halo tile in shared memory, and walks the expanded channels in chunks of 16*GC:
 *
 *   stage A (expand)   E[halo px][chunk] = act(s1 * (X . W1^T) + b1)  on the tensor cores (mma.sync m16n8k8 tf32,
 *                      3xTF32 split, fp32 accumulate), written to shared memory; pixels outside the image are never
 *                      computed -- their rows stay zero, which is exactly the depthwise conv's zero padding
 *   stage B (dw+proj)  every lane computes the 3x3 depthwise outputs of 2 pixels x 4 channels from shared memory
 *                      (FFMA, tap order ky,kx as conv-v0.c:16-25), applies BN+act, and the results ARE the A fragments
 *                      of the projection GEMM (rows = pixels, K = expanded channels): they are split hi/lo in
 *                      registers and multiplied into per-warp accumulators that persist across chunks
 *   epilogue           act(s3 * acc + b3) [+ x from the shared-memory tile, act] -> float2 stores
 *
 * Weights travel pre-arranged in fragment order ("chunks", built once by k_prep_block) and stream through a
 * double-buffered cp.async ring, so shared memory holds one chunk of E, not the whole expanded tile: 60-110 KB per CTA,
 * two CTAs per SM, one in its tensor-heavy stage while the other runs FFMAs.
 *
 * The fragment <-> tensor index maps are free permutations of M (pixels), N and K (channels); they are chosen so that
 * every shared-memory access is a conflict-free 64/128-bit access:
 *   expand   A row g   <-> tile pixel 2g,  row g+8 <-> pixel 2g+1;  k col t <-> cin 8ks+2t, col t+4 <-> cin 8ks+2t+1
 *            n-tile pair (2 x 8 cols) <-> 16 channels: col 2t'+j of tile ntl <-> channel 4t' + 2ntl + j
 *   project  A row g/g+8 <-> output pixels (ty, tx), (ty, tx+1);  k-step kk col t <-> channel 4t+2kk, col t+4 <-> 4t+2kk+1
 *
 * Numerics: 3xTF32 with a rounded split (hi = rna_tf32(x), lo = x - hi exact; D = lo*hi' + hi*lo' + hi*hi'), the same
 * scheme as pw_tc.cu -- fp32-equivalent (see DESIGN.md "Numerics of the tensor-core path").
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda.h>
#include "sm100.cuh"

namespace ffb {

using sm100::pdl_trigger;
using sm100::pdl_wait;

__device__ __forceinline__ void blk_fma4(float4 &acc, const float4 v, const float4 w)
{
    acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
    acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
}
__device__ __forceinline__ float4 blk_zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

/* Four channels as two packed fp32 pairs (sm100.cuh: FFMA2 / FMUL2 / FADD2 do two IEEE fp32 operations per issue slot, results
 * bit-identical to the scalar instructions).  The depthwise stage is bound by instruction issue on the warps that own a unit,
 * so its 9 taps, BN and the lo half of the tf32 split run on pairs. */
using sm100::f32x2; using sm100::f4p; using sm100::lds128p;
__device__ __forceinline__ f4p blk_zero4p() { return sm100::zero4p(); }
__device__ __forceinline__ f4p ld4p(const float *p) { const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p); f4p r; r.a = t.x; r.b = t.y; return r; }
__device__ __forceinline__ void blk_fma4p(f4p &acc, const f4p v, const f4p w) { sm100::fma4p(acc, v, w); }
__device__ __forceinline__ f4p blk_mul4p(const f4p v, const f4p w) { f4p r; r.a = sm100::f2_mul(v.a, w.a); r.b = sm100::f2_mul(v.b, w.b); return r; }

constexpr int BLK_THREADS = 256;
constexpr int BLK_WARPS = BLK_THREADS / 32;

struct BlkArgs {
    const float *x; float *y;
    const float *wchunks;                 /* [NC][chunk_floats], fragment order (k_prep_block) */
    const float *sb3;                     /* [2][8*NT3] projection scale, bias (zero padded) */
    int N, H, W, OH, OW, ldx, ldy, cout;
    int TH, TW, HH, HW, ntx, nty; long ntiles;
    int NC, xrows;
    int nmt; uint32_t tmem_cols;          /* TC: 128-pixel m-tiles of the x tile, TMEM columns to allocate (power of 2) */
    int XH, XW, xo, yo, frame;            /* x-tile box and its offset inside the halo; frame: one tile covers the whole image */
    int R, XB, xbuf_floats;               /* k_block_ws: weight slots (all chunks resident when NC <= R), x buffers, floats per x buffer */
    float inv_tpf, inv_ntx;
    float slope1, sloped, slope3, slope_res; int res;
    long long *trace;                     /* developer timeline (-DFFB_BLK_TRACE, tools/blk_trace.py): CTA 0, warps 0 and 7 stamp clock64 per stage */
};

#ifdef FFB_BLK_TRACE
#define BTRACE(ev) do { if (a.trace && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 7) && tr_n < 250) a.trace[(warp ? 1 : 0) * 512 + 2 * tr_n] = (ev), a.trace[(warp ? 1 : 0) * 512 + 2 * tr_n + 1] = clock64(), tr_n++; } while (0)
#else
#define BTRACE(ev) do { } while (0)
#endif

/* per-chunk section offsets (floats) */
struct BlkChunk {
    int w1, s1, b1, wd, sd, bd, w2, total;
    /* TC (tcgen05 expand, experimental): the w1 section is the UMMA B operand instead of mma.sync fragments -- hi and lo
       parts, each (KS1+3)/4 K-major SWIZZLE_128B sub-tiles of [16*GC channels x 32 cin] floats (128-byte rows, 8-row atoms
       of 1024 B), so the section starts the chunk and the chunk is padded to a multiple of 1024 B */
    __host__ __device__ constexpr BlkChunk(int GC, int KS1, int NT3, bool TC = false)
        : w1(0), s1(TC ? 2 * ((KS1 + 3) / 4) * 16 * GC * 32 : GC * 2 * KS1 * 128), b1(s1 + GC * 16), wd(b1 + GC * 16), sd(wd + GC * 144),
          bd(sd + GC * 16), w2(bd + GC * 16), total(TC ? (w2 + GC * 2 * NT3 * 128 + 255) / 256 * 256 : w2 + GC * 2 * NT3 * 128) {}
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

/* x = hi + lo, hi = x rounded to nearest tf32 (two integer ops), lo exact */
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}


/* activation with negative-side slope in [0, 1] (leaky 0.1, relu 0, linear 1; utils.h:15-23): max(v, slope*v) */
__device__ __forceinline__ float slope_act(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float4 bn_act4(float4 a, float4 s, float4 b, float slope)
{
    float4 r;
    r.x = slope_act(fmaf(a.x, s.x, b.x), slope); r.y = slope_act(fmaf(a.y, s.y, b.y), slope);
    r.z = slope_act(fmaf(a.z, s.z, b.z), slope); r.w = slope_act(fmaf(a.w, s.w, b.w), slope);
    return r;
}

/* BN + activation of four channels held as pairs; the result comes back as scalars (the max is a scalar instruction) */
__device__ __forceinline__ float4 bn_act4p(f4p a, f4p s, f4p b, f32x2 slope2)
{
    const f32x2 va = sm100::f2_fma(a.a, s.a, b.a), vb = sm100::f2_fma(a.b, s.b, b.b);
    const f32x2 ma = sm100::f2_mul(va, slope2), mb = sm100::f2_mul(vb, slope2);
    return make_float4(fmaxf(sm100::f2_lo(va), sm100::f2_lo(ma)), fmaxf(sm100::f2_hi(va), sm100::f2_hi(ma)),
                       fmaxf(sm100::f2_lo(vb), sm100::f2_lo(mb)), fmaxf(sm100::f2_hi(vb), sm100::f2_hi(mb)));
}
/* max(v, slope * v) on a pair */
__device__ __forceinline__ float2 act2(f32x2 v, f32x2 slope2)
{
    const f32x2 m = sm100::f2_mul(v, slope2);
    return make_float2(fmaxf(sm100::f2_lo(v), sm100::f2_lo(m)), fmaxf(sm100::f2_hi(v), sm100::f2_hi(m)));
}
/* tf32 split of two values that are adjacent in an mma fragment: hi by integer rounding, lo = x - hi as one packed subtract */
__device__ __forceinline__ void split_tf32x2(float x0, float x1, uint32_t &h0, uint32_t &h1, uint32_t &l0, uint32_t &l1)
{
    h0 = (__float_as_uint(x0) + 0x1000u) & 0xffffe000u; h1 = (__float_as_uint(x1) + 0x1000u) & 0xffffe000u;
    const f32x2 l = sm100::f2_sub(sm100::f2_pack(x0, x1), (f32x2)h0 | ((f32x2)h1 << 32));
    l0 = (uint32_t)l; l1 = (uint32_t)(l >> 32);
}

/* 1-D bulk copy global -> shared, completion (bytes) on an mbarrier; size multiple of 16, both addresses 16-byte aligned */
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar)) : "memory");
}

struct BlkTile { int n, oy0, ox0, th, tw, iy0, ix0; };

template <int S>
__device__ __forceinline__ BlkTile blk_tile(const BlkArgs &a, long tile)
{
    BlkTile g;
    const int tiles_per_frame = a.ntx * a.nty;
    g.n = (int)(((float)tile + 0.5f) * a.inv_tpf);                        /* exact: tile < 2^22, |error| << 0.5 / tiles_per_frame */
    const int tr = (int)(tile - (long)g.n * tiles_per_frame), tyi = (int)(((float)tr + 0.5f) * a.inv_ntx), txi = tr - tyi * a.ntx;
    g.oy0 = tyi * a.TH; g.ox0 = txi * a.TW;
    g.th = min(a.TH, a.OH - g.oy0); g.tw = min(a.TW, a.OW - g.ox0);
    g.iy0 = g.oy0 * S - 1; g.ix0 = g.ox0 * S - 1;                        /* image coordinates of halo pixel (0,0) */
    return g;
}

/*
 * Shared-memory x tile = one TMA box [XH][XW][SXs] of the NHWC tensor at (ix0 + xo, iy0 + yo): out-of-image pixels and the
 * channel lanes beyond the tensor's (box wider than the tensor) arrive as zeros, which gives the padded, bank-conflict-free
 * pixel stride SXs = 8*KS1 + 4 for free.  Normally the box is the whole halo (xo = yo = 0); when one tile covers the whole
 * image ("frame mode") the halo ring lies entirely outside the image, the box is just the image (xo = yo = 1) and the ring
 * rows of E are cleared once per kernel.
 *
 * Output-pixel <-> fragment-row map of stage B.  MTW == 1: m-tile mt = 16 consecutive pixels of the row-major tile, lane
 * (g) owns pixels 2g, 2g+1.  MTW >= 2: the tile is walked in ROW PAIRS -- m-tiles 2k and 2k+1 cover the same 16 positions
 * of the (row pair, x) sequence in the upper and the lower row -- so a lane owns a 2x2 pixel quad and its 3x3 stencils
 * share loads: (S+3)^2 shared-memory reads per quad instead of 2 * 3 * (S+3).
 */
template <int KS1, int NT3, int S, int MTW, int GC, int MINB, bool TC = false>
__global__ void __launch_bounds__(BLK_THREADS, MINB) k_block_mma(const __grid_constant__ CUtensorMap tmX, const BlkArgs a)
{
    extern __shared__ __align__(128) float4 blk_smem4[];
    float *smem = reinterpret_cast<float *>(blk_smem4);
    constexpr int CIN_P = 8 * KS1, SXs = CIN_P + 4, COUT_P = 8 * NT3;
    constexpr int SEs = 16 * GC + (S == 1 ? 8 : 4);       /* E pixel stride (floats): conflict-free 128-bit stencil loads */
    constexpr int MT = (KS1 * GC >= 6) ? 1 : 2;           /* m-tiles per stage-A work item (bounds the accumulator registers; 1 also spreads the 7 m-tiles of a 10x10 frame over 7 warps) */
    constexpr bool QUAD = MTW >= 2;
    constexpr int NQ = QUAD ? MTW / 2 : 1;                /* quads (or single m-tiles) per warp */
    constexpr BlkChunk off(GC, KS1, NT3, TC);
    /* the thread that issues the tcgen05 expand GEMMs: lane 0 of the LAST warp -- on the 40x40 and 10x10 blocks that warp has no
       stage-B unit (5 / 7 units for 8 warps), so the ~0.5-1.5 k cycles of MMA issue per chunk leave the critical path
       (profiles/r2q_blockmma_timeline.txt: with thread 0 issuing, warp 0 -- a stage-B worker -- reached every stage that much later) */
    constexpr int ISSUER = (BLK_WARPS - 1) * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int HW = a.HW;
    float    *sSB3 = smem;                                          /* 96 floats */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 96);       /* full_x[2], full_w[2] */
    int2     *sMap = reinterpret_cast<int2 *>(smem + 128);          /* [xrows]: x-tile pixel -> { byte offset of its E row or -1, hy | hx << 16 } */
    float    *sW = smem + 128 + 2 * a.xrows;
    if constexpr (TC) sW += ((1024u - (sm100::smem_u32(sW) & 1023u)) & 1023u) >> 2;   /* UMMA SWIZZLE_128B atoms are 1024-byte aligned */
    float    *sXB = sW + 2 * off.total;                             /* [2][xrows * SXs] */
    float    *sE = sXB + 2 * a.xrows * SXs;
    uint64_t *full_x = bars, *full_w = bars + 2;
    uint64_t *dfull = bars + 4;                                     /* TC: expand accumulator buffer [2] written (tcgen05.commit) */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    const uint32_t sE_addr = sm100::smem_u32(sE), sW_addr = sm100::smem_u32(sW);
    const uint32_t x_bytes = (uint32_t)a.XH * a.XW * SXs * 4;
    constexpr uint32_t w_bytes = (uint32_t)off.total * 4;
    const int XP = a.XH * a.XW, nitems = (((XP + 15) >> 4) + MT - 1) / MT;

    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmX);
        for (int i = 0; i < (TC ? 6 : 4); i++) sm100::mbar_init(bars + i, 1);
        sm100::fence_barrier_init();
    }
    if (TC && warp == 0) sm100::tmem_alloc(tmem_slot, a.tmem_cols);
    if (tid < 2 * COUT_P) sSB3[tid] = a.sb3[tid];
    for (int xp = tid; xp < a.xrows; xp += BLK_THREADS) {
        const int ry = xp / a.XW, rx = xp - ry * a.XW, hy = ry + a.yo, hx = rx + a.xo;
        sMap[xp] = make_int2(xp < XP ? (hy * HW + hx) * SEs * 4 : -1, hy | (hx << 16));
    }
    for (int i = tid; i < a.HH * HW * SEs / 4; i += BLK_THREADS) reinterpret_cast<float4 *>(sE)[i] = blk_zero4();
    /* tile-independent lane geometry: top-left output pixel (ty, tx) of this lane's quad (or pixel pair) per owned unit */
    uint32_t dwbase[NQ]; int tyx[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) {
        const int unit = warp + BLK_WARPS * qi;                                  /* quad index (QUAD) or m-tile index */
        const int p = unit * 16 + 2 * g;
        int ty, tx; bool valid;
        if (QUAD) { const int np = (a.TH + 1) / 2 * a.TW, pc = min(p, np - 2); const int rp = pc / a.TW; tx = pc - rp * a.TW; ty = 2 * rp; valid = p < np; }
        else      { const int np = a.TH * a.TW, pc = min(p, np - 2); ty = pc / a.TW; tx = pc - ty * a.TW; valid = p < np; }
        dwbase[qi] = sE_addr + (uint32_t)(((ty * S) * HW + tx * S) * SEs + 4 * t) * 4;
        tyx[qi] = valid ? (ty << 16 | tx) : (0x7fff << 16);                      /* invalid units fail the ty < th test */
    }
    const int nunits = QUAD ? ((a.TH + 1) / 2 * a.TW + 15) >> 4 : (a.TH * a.TW + 15) >> 4;
    const int nq = warp < nunits ? (nunits - warp + BLK_WARPS - 1) / BLK_WARPS : 0;     /* units this warp owns (warp-uniform) */
    const uint32_t rowpitch = (uint32_t)HW * SEs * 4;
    if (TC) sm100::tc_fence_before_sync();
    __syncthreads();
    if (TC) sm100::tc_fence_after_sync();
    pdl_trigger();                                    /* the wait comes after the first weight chunk is requested (weights do not depend on the previous kernel) */
    const f32x2 slope1_2 = sm100::f2_pack(a.slope1, a.slope1), sloped2 = sm100::f2_pack(a.sloped, a.sloped);
    /* ---- TC: expand GEMM on tcgen05.  TMEM columns: per 128-pixel m-tile mt the A operand [x_hi (KP cols) | x_lo (KP cols)]
       at mt * 2KP, then the accumulators D[mt][buf] (16*GC cols each, double buffered over chunks) ---- */
    constexpr int KP = 8 * KS1, KC = (KS1 + 3) / 4;
    const uint32_t tmem_base = TC ? *tmem_slot : 0u;
    const uint32_t tq_addr = (uint32_t)((warp & 3) * 32) << 16;                 /* this warp's TMEM lane quarter */
    const uint32_t dcol0 = tmem_base + (uint32_t)a.nmt * 2 * KP;
    auto issue_expand = [&](uint32_t wslot, uint32_t buf) {                     /* one thread: chunk in weight slot wslot -> D[.][buf] */
        constexpr uint32_t sub = 16 * GC * 128;                                 /* bytes of one [16*GC x 32] B sub-tile */
        constexpr uint32_t idesc = sm100::umma_idesc_tf32(128, 16 * GC);
        const uint32_t bh = sW_addr + wslot * (uint32_t)off.total * 4, bl = bh + KC * sub;
        /* descriptors once per chunk: a k-step / sub-tile only moves the start-address field (bytes >> 4); building every
           descriptor from its address cost the issuing thread ~100 cycles per MMA (profiles/r2k_pwtc_trace.txt) */
        const uint64_t dbh = sm100::umma_desc_sw128(bh), dbl = sm100::umma_desc_sw128(bl);
        for (int mt = 0; mt < a.nmt; mt++) {
            const uint32_t d = dcol0 + (uint32_t)(mt * 2 + buf) * 16 * GC;
            const uint32_t ahi = tmem_base + (uint32_t)mt * 2 * KP, alo = ahi + KP;
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                    /* x_lo . W_hi */
                sm100::mma_tf32_ts(d, alo + 8 * ks, dbh + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, ks > 0);
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                    /* x_hi . W_lo */
                sm100::mma_tf32_ts(d, ahi + 8 * ks, dbl + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, 1);
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                    /* x_hi . W_hi */
                sm100::mma_tf32_ts(d, ahi + 8 * ks, dbh + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, 1);
        }
        sm100::tc_commit(dfull + buf);
    };

    auto load_x = [&](long tile, int b) {                                       /* one thread */
        const BlkTile q = blk_tile<S>(a, tile);
        sm100::mbar_arrive_expect_tx(full_x + b, x_bytes);
        sm100::tma_load_4d(sXB + b * a.xrows * SXs, &tmX, 0, q.ix0 + a.xo, q.iy0 + a.yo, q.n, full_x + b);
    };
    auto load_chunk = [&](int c, int wb) {                                      /* one thread */
        sm100::mbar_arrive_expect_tx(full_w + wb, w_bytes);
        bulk_load(sW_addr + (uint32_t)wb * w_bytes, a.wchunks + (long)c * off.total, w_bytes, full_w + wb);
    };
    if (tid == 0 && (long)blockIdx.x < a.ntiles) load_chunk(0, 0);
    pdl_wait();                                       /* from here on the kernel touches the previous layer's output */
    if (tid == 0 && (long)blockIdx.x < a.ntiles) load_x(blockIdx.x, 0);
    /* weights: with a single chunk they stay resident for the whole kernel; otherwise the two-slot ring runs continuously
       across tiles (chunk (c+1) % NC is requested while chunk c is being used), so no tile ever waits for L2 */
    const bool w_resident = a.NC == 1;
    [[maybe_unused]] int tr_n = 0;
    uint32_t it = 0, cs = 0;                                                    /* tiles / weight chunks consumed so far by this CTA */
    for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
        const BlkTile q = blk_tile<S>(a, tile);
        const int xb = it & 1;
        const float *sX = sXB + xb * a.xrows * SXs;
        const bool border = !a.frame && (q.iy0 < 0 || q.ix0 < 0 || q.iy0 + a.HH > a.H || q.ix0 + HW > a.W);

        BTRACE(1);
        __syncthreads();                                  /* the previous tile no longer reads sW / sE / the other x buffer */
        BTRACE(2);
        const bool last_tile = tile + gridDim.x >= a.ntiles;
        if (tid == 0 && !last_tile) load_x(tile + gridDim.x, xb ^ 1);
        float pacc[MTW][NT3][4];
#pragma unroll
        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
            for (int nt = 0; nt < NT3; nt++)
#pragma unroll
                for (int j = 0; j < 4; j++) pacc[mi][nt][j] = 0.f;
        sm100::mbar_wait(full_x + xb, (it >> 1) & 1);
        BTRACE(3);

        if constexpr (TC) {
            /* x tile -> TMEM as the A operand, split hi/lo (thread = pixel = TMEM lane); x itself stays untouched in shared
               memory for the shortcut.  All expand MMAs of the previous tile have completed (their dfull waits). */
            for (int mt = warp >> 2; mt < a.nmt; mt += BLK_WARPS / 4) {
                const int p = mt * 128 + (warp & 3) * 32 + lane;
                const float *xr = sX + p * SXs;
                const uint32_t acol = tmem_base + tq_addr + (uint32_t)mt * 2 * KP;
#pragma unroll
                for (int ks = 0; ks < KS1; ks++) {
                    float4 x0 = blk_zero4(), x1 = blk_zero4();
                    if (p < XP) { x0 = *reinterpret_cast<const float4 *>(xr + 8 * ks); x1 = *reinterpret_cast<const float4 *>(xr + 8 * ks + 4); }
                    uint32_t hi[8], lo[8];
                    split_tf32x2(x0.x, x0.y, hi[0], hi[1], lo[0], lo[1]); split_tf32x2(x0.z, x0.w, hi[2], hi[3], lo[2], lo[3]);
                    split_tf32x2(x1.x, x1.y, hi[4], hi[5], lo[4], lo[5]); split_tf32x2(x1.z, x1.w, hi[6], hi[7], lo[6], lo[7]);
                    sm100::tmem_st8(acol + 8 * ks, hi);
                    sm100::tmem_st8(acol + KP + 8 * ks, lo);
                }
            }
            sm100::tmem_st_wait();
            sm100::tc_fence_before_sync();
            __syncthreads();
            if (tid == ISSUER) {                          /* chunk 0 of this tile: its weights were requested during the previous tile */
                const uint32_t ws0 = (a.NC == 1) ? 0u : (cs & 1u);
                sm100::mbar_wait(full_w + ws0, (a.NC == 1) ? 0u : ((cs >> 1) & 1u));
                sm100::tc_fence_after_sync();
                issue_expand(ws0, cs & 1u);
            }
        }

        for (int c = 0; c < a.NC; c++, cs++) {
            const int wb = w_resident ? 0 : cs & 1;
            sm100::mbar_wait(full_w + wb, w_resident ? 0 : (cs >> 1) & 1);
            BTRACE(10);
            if (c > 0) __syncthreads();                   /* every warp is done with chunk c-1: E and the other weight buffer are free */
            BTRACE(11);
            if (tid == 0 && !w_resident && !(last_tile && c + 1 == a.NC)) load_chunk(c + 1 < a.NC ? c + 1 : 0, wb ^ 1);
            const float *wc = sW + wb * off.total;
            const float *wl4 = wc + lane * 4, *wt4 = wc + 4 * t;

            [[maybe_unused]] bool expand_pending = false;
            if constexpr (TC) {
                /* ---------------- stage A (TC): the chunk's expand accumulators TMEM -> BN + act -> E rows ---------------- */
                const uint32_t buf = cs & 1u;
                sm100::mbar_wait(dfull + buf, (cs >> 1) & 1u);
                BTRACE(12);
                sm100::tc_fence_after_sync();
                for (int mt = warp >> 2; mt < a.nmt; mt += BLK_WARPS / 4) {
                    const int p = mt * 128 + (warp & 3) * 32 + lane;
                    const int2 mp = p < a.xrows ? sMap[p] : make_int2(-1, 0);
                    bool inside = true;
                    if (border && mp.x >= 0) {            /* halo pixels outside the image are the depthwise conv's zero padding */
                        const int iy = q.iy0 + (mp.y & 0xffff), ix = q.ix0 + (mp.y >> 16);
                        inside = (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
                    }
                    const uint32_t dcol = dcol0 + tq_addr + (uint32_t)(mt * 2 + buf) * 16 * GC;
#pragma unroll
                    for (int gr = 0; gr < GC; gr++) {
                        uint32_t r[16];
                        sm100::tmem_ld16(dcol + gr * 16, r);
                        /* the group's scale / bias first, as eight independent loads under the TMEM load's latency: the stores below are
                           ordered (volatile asm), and one load pair per store made this stage a chain of exposed shared-memory latencies */
                        f4p s1v[4], b1v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            s1v[j] = ld4p(wc + off.s1 + gr * 16 + 4 * j);
                            b1v[j] = ld4p(wc + off.b1 + gr * 16 + 4 * j);
                        }
                        sm100::tmem_ld_wait();
                        if (mp.x >= 0) {
                            /* pixels outside the image are zeroed with a mask, not a branch: the four float4 stay independent streams */
                            const uint32_t msk = inside ? 0xffffffffu : 0u;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                f4p acc; acc.a = (f32x2)r[4 * j] | ((f32x2)r[4 * j + 1] << 32); acc.b = (f32x2)r[4 * j + 2] | ((f32x2)r[4 * j + 3] << 32);
                                float4 v = bn_act4p(acc, s1v[j], b1v[j], slope1_2);
                                v.x = __uint_as_float(__float_as_uint(v.x) & msk); v.y = __uint_as_float(__float_as_uint(v.y) & msk);
                                v.z = __uint_as_float(__float_as_uint(v.z) & msk); v.w = __uint_as_float(__float_as_uint(v.w) & msk);
                                sm100::sts128(sE_addr + mp.x + (gr * 16 + 4 * j) * 4, v);
                            }
                        }
                    }
                }
                sm100::tc_fence_before_sync();
                expand_pending = c + 1 < a.NC;
            } else
            /* ---------------- stage A: expand GEMM; work item = MT m-tiles x all GC groups of the chunk ----------------
               the A fragments (x, split hi/lo on the fly) are loaded once per k-step and reused by every group */
            for (int item = warp; item < nitems; item += BLK_WARPS) {
                const int mt0 = item * MT;
                float acc[GC][MT][2][4];
#pragma unroll
                for (int gr = 0; gr < GC; gr++)
#pragma unroll
                    for (int m = 0; m < MT; m++)
#pragma unroll
                        for (int ntl = 0; ntl < 2; ntl++)
#pragma unroll
                            for (int j = 0; j < 4; j++) acc[gr][m][ntl][j] = 0.f;
                const float *xr = sX + (mt0 * 16 + 2 * g) * SXs + 2 * t;
#pragma unroll
                for (int ks = 0; ks < KS1; ks++) {
                    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
                    for (int m = 0; m < MT; m++) {
                        const float2 x0 = *reinterpret_cast<const float2 *>(xr + m * 16 * SXs + 8 * ks), x1 = *reinterpret_cast<const float2 *>(xr + m * 16 * SXs + SXs + 8 * ks);
                        split_tf32(x0.x, ah[m][0], al[m][0]); split_tf32(x1.x, ah[m][1], al[m][1]);
                        split_tf32(x0.y, ah[m][2], al[m][2]); split_tf32(x1.y, ah[m][3], al[m][3]);
                    }
#pragma unroll
                    for (int gr = 0; gr < GC; gr++) {
                        uint32_t bh[2][2], bl[2][2];
#pragma unroll
                        for (int ntl = 0; ntl < 2; ntl++) {
                            const float4 b = *reinterpret_cast<const float4 *>(wl4 + off.w1 + ((gr * 2 + ntl) * KS1 + ks) * 128);   /* hi0 hi1 lo0 lo1 */
                            bh[ntl][0] = __float_as_uint(b.x); bh[ntl][1] = __float_as_uint(b.y); bl[ntl][0] = __float_as_uint(b.z); bl[ntl][1] = __float_as_uint(b.w);
                        }
                        /* the three 3xTF32 terms of one accumulator are dependent: each term is issued across the
                           independent accumulators so the tensor pipe always has independent work in flight */
#pragma unroll
                        for (int m = 0; m < MT; m++)
#pragma unroll
                            for (int ntl = 0; ntl < 2; ntl++) mma_tf32(acc[gr][m][ntl], al[m], bh[ntl][0], bh[ntl][1]);
#pragma unroll
                        for (int m = 0; m < MT; m++)
#pragma unroll
                            for (int ntl = 0; ntl < 2; ntl++) mma_tf32(acc[gr][m][ntl], ah[m], bl[ntl][0], bl[ntl][1]);
#pragma unroll
                        for (int m = 0; m < MT; m++)
#pragma unroll
                            for (int ntl = 0; ntl < 2; ntl++) mma_tf32(acc[gr][m][ntl], ah[m], bh[ntl][0], bh[ntl][1]);
                    }
                }
#pragma unroll
                for (int m = 0; m < MT; m++)
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const int2 mp = sMap[(mt0 + m) * 16 + 2 * g + r];
                        if (mp.x >= 0) {
                            bool inside = true;
                            if (border) {                 /* halo pixels outside the image are the depthwise conv's zero padding */
                                const int iy = q.iy0 + (mp.y & 0xffff), ix = q.ix0 + (mp.y >> 16);
                                inside = (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
                            }
                            const uint32_t msk = inside ? 0xffffffffu : 0u;        /* a mask, not a branch: the groups stay independent streams */
#pragma unroll
                            for (int gr = 0; gr < GC; gr++) {
                                f4p av; av.a = sm100::f2_pack(acc[gr][m][0][2 * r], acc[gr][m][0][2 * r + 1]); av.b = sm100::f2_pack(acc[gr][m][1][2 * r], acc[gr][m][1][2 * r + 1]);
                                float4 v = bn_act4p(av, ld4p(wt4 + off.s1 + gr * 16), ld4p(wt4 + off.b1 + gr * 16), slope1_2);
                                v.x = __uint_as_float(__float_as_uint(v.x) & msk); v.y = __uint_as_float(__float_as_uint(v.y) & msk);
                                v.z = __uint_as_float(__float_as_uint(v.z) & msk); v.w = __uint_as_float(__float_as_uint(v.w) & msk);
                                sm100::sts128(sE_addr + mp.x + (gr * 16 + 4 * t) * 4, v);
                            }
                        }
                    }
            }
            BTRACE(13);
            __syncthreads();
            BTRACE(14);
            if constexpr (TC) {
                /* the next chunk's expand GEMM runs on the tensor core while the warps do stage B of this one: D[.][buf ^ 1] was
                   drained before the previous barrier pair, its weights were requested at the top of this chunk */
                if (tid == ISSUER && expand_pending && sm100::mbar_try_wait(full_w + (wb ^ 1), ((cs + 1) >> 1) & 1u)) {
                    sm100::tc_fence_after_sync();
                    issue_expand((uint32_t)(wb ^ 1), (cs + 1) & 1u);
                    expand_pending = false;
                }
            }

            /* ---------------- stage B: depthwise 3x3 in registers -> projection GEMM ---------------- */
#pragma unroll
            for (int grp = 0; grp < GC; grp++) {
                f4p wd[9];
#pragma unroll
                for (int k = 0; k < 9; k++) wd[k] = ld4p(wt4 + off.wd + (grp * 9 + k) * 16);
                const f4p sd = ld4p(wt4 + off.sd + grp * 16), bd = ld4p(wt4 + off.bd + grp * 16);
                uint32_t ah[MTW][2][4], al[MTW][2][4];
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    constexpr uint32_t px = SEs * 4;
                    constexpr int NR = QUAD ? S + 3 : 3, NCOL = S + 3;             /* input rows / columns the unit touches */
                    f4p d[2][2];                                                   /* [row of the quad][pixel] */
#pragma unroll
                    for (int h = 0; h < 2; h++) { d[h][0] = blk_zero4p(); d[h][1] = blk_zero4p(); }
                    if (qi < nq) {
#pragma unroll
                        for (int r = 0; r < NR; r++) {
                            const uint32_t row = dwbase[qi] + grp * 64 + r * rowpitch;
                            f4p e[NCOL];
#pragma unroll
                            for (int k = 0; k < NCOL; k++) e[k] = lds128p(row + k * px);
#pragma unroll
                            for (int h = 0; h < (QUAD ? 2 : 1); h++) {
                                const int dy = r - h * S;                              /* tap row of this input row for quad row h */
                                if (dy == 0) {            /* first tap: a product, not 0 + product -- no accumulator clearing */
                                    d[h][0] = blk_mul4p(e[0], wd[0]); d[h][1] = blk_mul4p(e[S], wd[0]);
                                    blk_fma4p(d[h][0], e[1], wd[1]); blk_fma4p(d[h][0], e[2], wd[2]);
                                    blk_fma4p(d[h][1], e[S + 1], wd[1]); blk_fma4p(d[h][1], e[S + 2], wd[2]);
                                } else if (dy > 0 && dy < 3) {
                                    blk_fma4p(d[h][0], e[0], wd[dy * 3]); blk_fma4p(d[h][0], e[1], wd[dy * 3 + 1]); blk_fma4p(d[h][0], e[2], wd[dy * 3 + 2]);
                                    blk_fma4p(d[h][1], e[S], wd[dy * 3]); blk_fma4p(d[h][1], e[S + 1], wd[dy * 3 + 1]); blk_fma4p(d[h][1], e[S + 2], wd[dy * 3 + 2]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < (QUAD ? 2 : 1); h++) {
                        const int mi = QUAD ? 2 * qi + h : qi;
                        const float4 d0 = bn_act4p(d[h][0], sd, bd, sloped2), d1 = bn_act4p(d[h][1], sd, bd, sloped2);
                        /* fragment registers 0/1 and 2/3 hold the same channel of the lane's two pixels: split them as pairs */
                        split_tf32x2(d0.x, d1.x, ah[mi][0][0], ah[mi][0][1], al[mi][0][0], al[mi][0][1]);
                        split_tf32x2(d0.y, d1.y, ah[mi][0][2], ah[mi][0][3], al[mi][0][2], al[mi][0][3]);
                        split_tf32x2(d0.z, d1.z, ah[mi][1][0], ah[mi][1][1], al[mi][1][0], al[mi][1][1]);
                        split_tf32x2(d0.w, d1.w, ah[mi][1][2], ah[mi][1][3], al[mi][1][2], al[mi][1][3]);
                    }
                }
                if (nq > 0) {
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        uint32_t bh[NT3][2], bl[NT3][2];
#pragma unroll
                        for (int nt = 0; nt < NT3; nt++) {
                            const float4 b = *reinterpret_cast<const float4 *>(wl4 + off.w2 + ((grp * 2 + kk) * NT3 + nt) * 128);
                            bh[nt][0] = __float_as_uint(b.x); bh[nt][1] = __float_as_uint(b.y); bl[nt][0] = __float_as_uint(b.z); bl[nt][1] = __float_as_uint(b.w);
                        }
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                            for (int nt = 0; nt < NT3; nt++) mma_tf32(pacc[mi][nt], al[mi][kk], bh[nt][0], bh[nt][1]);
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                            for (int nt = 0; nt < NT3; nt++) mma_tf32(pacc[mi][nt], ah[mi][kk], bl[nt][0], bl[nt][1]);
#pragma unroll
                        for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                            for (int nt = 0; nt < NT3; nt++) mma_tf32(pacc[mi][nt], ah[mi][kk], bh[nt][0], bh[nt][1]);
                    }
                }
            }
            BTRACE(15);
            if constexpr (TC) {
                if (tid == ISSUER && expand_pending) {    /* the next chunk's weights had not landed when stage B started */
                    sm100::mbar_wait(full_w + (wb ^ 1), ((cs + 1) >> 1) & 1u);
                    sm100::tc_fence_after_sync();
                    issue_expand((uint32_t)(wb ^ 1), (cs + 1) & 1u);
                }
            }
        }

        /* ---------------- block epilogue: BN + act [+ shortcut from the resident x tile] -> y ---------------- */
        BTRACE(20);
        const f32x2 slope3_2 = sm100::f2_pack(a.slope3, a.slope3), sloper_2 = sm100::f2_pack(a.slope_res, a.slope_res);
#pragma unroll
        for (int mi = 0; mi < MTW; mi++) {
            const int qi = QUAD ? mi >> 1 : mi;
            const int ty = (tyx[qi] >> 16) + (QUAD ? (mi & 1) : 0), tx = tyx[qi] & 0xffff;
            if (ty < q.th && tx < q.tw) {
                float *yp = a.y + (((long)q.n * a.OH + q.oy0 + ty) * a.OW + q.ox0 + tx) * a.ldy;
                const float *xc = sX + ((ty + 1 - a.yo) * a.XW + tx + 1 - a.xo) * SXs;   /* centre pixel; S == 1 whenever res is set */
#pragma unroll
                for (int nt = 0; nt < NT3; nt++) {
                    const int co = 8 * nt + 2 * t;
                    if (co < a.cout) {
                        const float2 s3 = *reinterpret_cast<const float2 *>(sSB3 + co), b3 = *reinterpret_cast<const float2 *>(sSB3 + COUT_P + co);
                        const f32x2 s3p = sm100::f2_pack(s3.x, s3.y), b3p = sm100::f2_pack(b3.x, b3.y);
                        float2 v0 = act2(sm100::f2_fma(sm100::f2_pack(pacc[mi][nt][0], pacc[mi][nt][1]), s3p, b3p), slope3_2);
                        float2 v1 = act2(sm100::f2_fma(sm100::f2_pack(pacc[mi][nt][2], pacc[mi][nt][3]), s3p, b3p), slope3_2);
                        if (a.res) {
                            const float2 r0 = *reinterpret_cast<const float2 *>(xc + co), r1 = *reinterpret_cast<const float2 *>(xc + SXs + co);
                            v0 = act2(sm100::f2_add(sm100::f2_pack(v0.x, v0.y), sm100::f2_pack(r0.x, r0.y)), sloper_2);
                            v1 = act2(sm100::f2_add(sm100::f2_pack(v1.x, v1.y), sm100::f2_pack(r1.x, r1.y)), sloper_2);
                        }
                        *reinterpret_cast<float2 *>(yp + co) = v0;
                        *reinterpret_cast<float2 *>(yp + a.ldy + co) = v1;               /* tw is even: pixel tx+1 is inside the tile */
                    }
                }
            }
        }
    }
    if constexpr (TC) {
        sm100::tc_fence_before_sync();
        __syncthreads();
        if (warp == 0) { sm100::tc_fence_after_sync(); sm100::tmem_dealloc(tmem_base, a.tmem_cols); }
    }
}

/* Build the fragment-ordered weight chunks of one block from the three convs' packed reference rows
 * (ffcnn.c:218-234: [weights..., pad, scale', bias', mean, var] per filter). */
__global__ void k_prep_block(const float *__restrict__ p1, int row1, int cin, const float *__restrict__ pd, int rowd,
                             const float *__restrict__ p3, int row3, int cexp, int cout, int KS1, int NT3, int GC, int NC,
                             float *__restrict__ chunks, float *__restrict__ sb3, int tc)
{
    const BlkChunk off(GC, KS1, NT3, tc != 0);
    const long total = (long)NC * off.total;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total + 16 * NT3; i += (long)gridDim.x * blockDim.x) {
        if (i >= total) {                                  /* projection scale / bias */
            const int k = (int)(i - total), co = k % (8 * NT3), which = k / (8 * NT3);
            sb3[k] = co < cout ? p3[(long)co * row3 + row3 - 4 + which] : 0.f;
            continue;
        }
        const int c = (int)(i / off.total), r = (int)(i - (long)c * off.total);
        float v = 0.f; bool is_w = false; int lohalf = 0;
        if (r < off.s1 && tc) {
            /* UMMA B operand: [hi | lo] x KC sub-tiles of [16*GC rows (channels) x 32 floats (cin)], SWIZZLE_128B:
               element (n, k) of a sub-tile sits at float n*32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3) */
            const int KC = (KS1 + 3) / 4, subf = 16 * GC * 32;
            const int sub = r / subf, rem = r - sub * subf, n = rem >> 5, pos = rem & 31;
            const int k = ((((pos >> 2) ^ (n & 7)) << 2) | (pos & 3)), part = sub / KC, kc = sub - part * KC;
            const int ch = c * GC * 16 + n, ci = kc * 32 + k;
            if (ch < cexp && ci < cin) v = p1[(long)ch * row1 + ci];
            is_w = true; lohalf = part;
        } else if (r < off.s1) {
            const int lane4 = r % 128; int rest = r / 128;
            const int ks = rest % KS1; rest /= KS1;
            const int ntl = rest % 2, grp = rest / 2, lane = lane4 >> 2, j = lane4 & 1, g = lane >> 2, t = lane & 3;
            const int ch = (c * GC + grp) * 16 + 4 * (g >> 1) + 2 * ntl + (g & 1), ci = 8 * ks + 2 * t + j;
            if (ch < cexp && ci < cin) v = p1[(long)ch * row1 + ci];
            is_w = true; lohalf = (lane4 >> 1) & 1;
        } else if (r < off.wd) {
            const int rr = r - off.s1, which = rr / (GC * 16), ch = c * GC * 16 + rr % (GC * 16);
            if (ch < cexp) v = p1[(long)ch * row1 + row1 - 4 + which];
        } else if (r < off.sd) {
            const int rr = r - off.wd, grp = rr / 144, tap = (rr % 144) / 16, ch = (c * GC + grp) * 16 + rr % 16;
            if (ch < cexp) v = pd[(long)ch * rowd + tap];
        } else if (r < off.w2) {
            const int rr = r - off.sd, which = rr / (GC * 16), ch = c * GC * 16 + rr % (GC * 16);
            if (ch < cexp) v = pd[(long)ch * rowd + rowd - 4 + which];
        } else {
            const int rr = r - off.w2, lane4 = rr % 128; int rest = rr / 128;
            const int nt = rest % NT3; rest /= NT3;
            const int kk = rest % 2, grp = rest / 2, lane = lane4 >> 2, j = lane4 & 1, g = lane >> 2, t = lane & 3;
            const int co = 8 * nt + g, ch = (c * GC + grp) * 16 + 4 * t + 2 * kk + j;
            if (co < cout && ch < cexp) v = p3[(long)co * row3 + ch];
            is_w = true; lohalf = (lane4 >> 1) & 1;
        }
        if (is_w) {                                        /* fragment floats are stored pre-split: hi0 hi1 lo0 lo1 per lane */
            uint32_t hi, lo; split_tf32(v, hi, lo);
            v = __uint_as_float(lohalf ? lo : hi);
        }
        chunks[i] = v;
    }
}

} // namespace ffb
