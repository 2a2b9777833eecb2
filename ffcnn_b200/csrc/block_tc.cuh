/*
 * block_tc.cuh -- the inverted-residual block of the graph as ONE kernel with BOTH pointwise convs on the 5th-generation
 * tensor cores (tcgen05, accumulators in tensor memory):
 *
 *     x --1x1 expand, BN, act--> e --3x3 depthwise (stride 1|2), BN, act--> d --1x1 project, BN, act--> [+ x] --> y
 *
 * = three groupconv calls (conv-v6.c:46-91, 96-287) + the dropout/shortcut pair (ffcnn.c:412-423) of net_forward's layer
 * loop (ffcnn.c:476-520).  Successor of block_mma.cuh (mma.sync fragments, 5 work units for 8 warps in its depthwise stage,
 * 29 % of the samples stalled on barriers): here no warp ever touches a GEMM fragment, and every CUDA-core stage is mapped
 * thread = pixel, so all 256 threads have the same amount of work.
 *
 * A CTA owns a TH x TW tile (<= 128 pixels = one UMMA M tile) of one frame and walks the expanded channels in chunks of 32:
 *
 *   x tile      TMA box (halo included; out-of-image pixels arrive as zeros) -> shared memory, double buffered over tiles
 *   split       thread = halo pixel: x -> hi = rna_tf32(x), lo = x - hi, both written to TENSOR MEMORY (tcgen05.st): the A
 *               operand of the expand GEMM, M = 128 halo pixels per m-tile (x itself stays exact in smem for the shortcut)
 *   expand      one thread issues  D1[m-tile] = x_lo.W1_hi + x_hi.W1_lo + x_hi.W1_hi  (tcgen05.mma kind::tf32, A from TMEM,
 *               B = the chunk's W1 as K-major SWIZZLE_128B tiles that arrive pre-split by cp.async.bulk), N = 32
 *   stage A     thread = halo pixel: tcgen05.ld D1 -> act(s1*v + b1) -> E[pixel][32 ch] in shared memory (zeros for pixels
 *               outside the image = the depthwise conv's padding).  D1 is free again -> the NEXT chunk's expand GEMM is
 *               issued now and runs under stage B
 *   stage B     thread = output pixel x 16 channels: 3x3 depthwise from E (FFMA, tap order ky,kx as conv-v0.c:16-25),
 *               BN + act, split hi/lo: hi -> shared memory as the K-major SWIZZLE_128B A tile of the projection GEMM,
 *               lo -> tensor memory (tcgen05.st; row = TMEM lane = this thread)
 *   project     one thread issues  D2 += d_lo.W2_hi + d_hi.W2_lo + d_hi.W2_hi  (M = 128 output pixels, N = cout, K = 32);
 *               D2 accumulates over the chunks in tensor memory
 *   epilogue    thread = output pixel: tcgen05.ld D2 -> act(s3*v + b3) [+ x, act] -> 64-byte contiguous stores
 *
 * Tensor memory per CTA (<= 256 columns, two CTAs per SM): [x_hi|x_lo per m-tile][D1 per m-tile: 32][d_lo: 32][D2: cout].
 * Numerics: the 3xTF32 scheme of pw_tc.cu / block_mma.cuh (rounded split, fp32 accumulation) -- fp32-equivalent.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda.h>
#include "sm100.cuh"

namespace ffb {

constexpr int B2_THREADS = 256;
constexpr int B2_CH = 32;                   /* expanded channels per chunk = K of one projection step = one 128-byte row */
constexpr int B2_SE = B2_CH + 4;            /* E pixel stride (floats): 144 B -> conflict-free 128-bit access with thread = pixel */

struct Blk2Args {
    const float *x; float *y;
    const float *wchunks;                   /* [NC][chunk floats] (k_prep_block2) */
    const float *sb3;                       /* [2][N3] projection scale, bias (zero padded) */
    int N, H, W, OH, OW, ldx, ldy, cout;
    int TH, TW, HH, HW, ntx, nty; long ntiles;
    int NC, xrows, XH, XW, xo, yo, frame, nmt;
    uint32_t tmem_cols;
    float inv_tpf, inv_ntx;
    float slope1, sloped, slope3, slope_res; int res;
};

/* chunk sections (floats).  w1: [hi|lo] x KC sub-tiles [32 ch x 32 cin] SW128;  w2: [hi|lo] x [N3 cout x 32 ch] SW128 */
struct Blk2Chunk {
    int w1, w2, s1, b1, wd, sd, bd, total;
    __host__ __device__ constexpr Blk2Chunk(int KS1, int N3)
        : w1(0), w2(2 * ((KS1 + 3) / 4) * 1024), s1(w2 + 2 * N3 * 32), b1(s1 + 32), wd(b1 + 32), sd(wd + 9 * 32), bd(sd + 32),
          total((bd + 32 + 255) / 256 * 256) {}
};

__device__ __forceinline__ void b2_split(float x, uint32_t &hi, uint32_t &lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ float b2_act(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float4 b2_bn_act(float4 a, float4 s, float4 b, float slope)
{
    float4 r;
    r.x = b2_act(fmaf(a.x, s.x, b.x), slope); r.y = b2_act(fmaf(a.y, s.y, b.y), slope);
    r.z = b2_act(fmaf(a.z, s.z, b.z), slope); r.w = b2_act(fmaf(a.w, s.w, b.w), slope);
    return r;
}
__device__ __forceinline__ void b2_fma4(float4 &acc, const float4 v, const float4 w)
{
    acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
}
__device__ __forceinline__ void b2_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar)) : "memory");
}

struct B2Tile { int n, oy0, ox0, th, tw, iy0, ix0; };

template <int S>
__device__ __forceinline__ B2Tile b2_tile(const Blk2Args &a, long tile)
{
    B2Tile g;
    const int tpf = a.ntx * a.nty;
    g.n = (int)(((float)tile + 0.5f) * a.inv_tpf);                       /* exact: tile < 2^22 */
    const int tr = (int)(tile - (long)g.n * tpf), tyi = (int)(((float)tr + 0.5f) * a.inv_ntx), txi = tr - tyi * a.ntx;
    g.oy0 = tyi * a.TH; g.ox0 = txi * a.TW;
    g.th = min(a.TH, a.OH - g.oy0); g.tw = min(a.TW, a.OW - g.ox0);
    g.iy0 = g.oy0 * S - 1; g.ix0 = g.ox0 * S - 1;
    return g;
}

/* KS1 = ceil(cin / 8) k-steps of the expand GEMM, N3 = cout rounded up to 16 (N of the projection GEMM), S = stride */
template <int KS1, int N3, int S>
__global__ void __launch_bounds__(B2_THREADS, 2) k_block_tc(const __grid_constant__ CUtensorMap tmX, const Blk2Args a)
{
    extern __shared__ __align__(128) float4 b2_smem4[];
    float *smem = reinterpret_cast<float *>(b2_smem4);
    constexpr int KP = 8 * KS1, SXs = KP + 4, KC = (KS1 + 3) / 4;
    constexpr Blk2Chunk off(KS1, N3);
    constexpr uint32_t w_bytes = (uint32_t)off.total * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = a.HW, XP = a.XH * a.XW;

    /* shared memory: [sb3 2*N3][barriers][sMap xrows int2] | 1024-aligned: [W ring 2 chunks][A2 hi tile 16 KB] | [x 2 tiles][E] */
    float    *sSB3 = smem;                                                  /* <= 96 floats */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 96);               /* full_x[2] full_w[2] dfull pbar */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    int2     *sMap = reinterpret_cast<int2 *>(smem + 128);                  /* x-tile pixel -> { byte offset of its E row or -1, hy | hx << 16 } */
    float    *sW = smem + 128 + 2 * a.xrows;
    sW += ((1024u - (sm100::smem_u32(sW) & 1023u)) & 1023u) >> 2;           /* UMMA SWIZZLE_128B atoms are 1024-byte aligned */
    float    *sA2 = sW + 2 * off.total;                                     /* [128 rows][32 floats] SW128 */
    float    *sXB = sA2 + 128 * 32;                                         /* [2][xrows * SXs] */
    float    *sE = sXB + 2 * a.xrows * SXs;                                 /* [HH*HW][B2_SE] */
    uint64_t *full_x = bars, *full_w = bars + 2, *dfull = bars + 4, *pbar = bars + 5;
    const uint32_t sE_addr = sm100::smem_u32(sE), sW_addr = sm100::smem_u32(sW), sA2_addr = sm100::smem_u32(sA2);
    const uint32_t x_bytes = (uint32_t)a.XH * a.XW * SXs * 4;

    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmX);
        for (int i = 0; i < 6; i++) sm100::mbar_init(bars + i, 1);
        sm100::fence_barrier_init();
    }
    if (warp == 0) sm100::tmem_alloc(tmem_slot, a.tmem_cols);
    if (tid < 2 * N3) sSB3[tid] = a.sb3[tid];
    for (int xp = tid; xp < a.xrows; xp += B2_THREADS) {
        const int ry = xp / a.XW, rx = xp - ry * a.XW, hy = ry + a.yo, hx = rx + a.xo;
        sMap[xp] = make_int2(xp < XP ? (hy * HW + hx) * B2_SE * 4 : -1, hy | (hx << 16));
    }
    for (int i = tid; i < a.HH * HW * B2_SE / 4; i += B2_THREADS) reinterpret_cast<float4 *>(sE)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 128 * 32 / 4; i += B2_THREADS) reinterpret_cast<float4 *>(sA2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    sm100::fence_proxy_async_smem();
    sm100::tc_fence_before_sync();
    __syncthreads();
    sm100::tc_fence_after_sync();
    sm100::pdl_trigger(); sm100::pdl_wait();

    /* tensor memory columns */
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tq = (uint32_t)((warp & 3) * 32) << 16;                  /* this warp's TMEM lane quarter */
    const uint32_t colX = tmem_base, colD1 = colX + (uint32_t)a.nmt * 2 * KP, colA2 = colD1 + (uint32_t)a.nmt * B2_CH, colD2 = colA2 + B2_CH;

    /* stage-B / epilogue geometry: this thread's output pixel (row of the projection m-tile = its TMEM lane) and channel half */
    const int row = (warp & 3) * 32 + lane, hb = warp >> 2;
    const int ty = row / a.TW, tx = row - ty * a.TW;
    const bool row_in_tile = row < a.TH * a.TW;
    const uint32_t e_base = sE_addr + (uint32_t)(((ty * S) * HW + tx * S) * B2_SE + 16 * hb) * 4;
    const uint32_t e_row = (uint32_t)HW * B2_SE * 4;

    auto issue_expand = [&](uint32_t wslot) {                               /* one thread: chunk in weight slot wslot -> D1 */
        constexpr uint32_t idesc = sm100::umma_idesc_tf32(128, B2_CH);
        const uint32_t bh = sW_addr + wslot * w_bytes, bl = bh + KC * 4096;
        const uint64_t dbh = sm100::umma_desc_sw128(bh), dbl = sm100::umma_desc_sw128(bl);      /* k-steps only move the start-address field */
        for (int mt = 0; mt < a.nmt; mt++) {
            const uint32_t d = colD1 + (uint32_t)mt * B2_CH, ahi = colX + (uint32_t)mt * 2 * KP, alo = ahi + KP;
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                /* x_lo . W1_hi (small terms first) */
                sm100::mma_tf32_ts(d, alo + 8 * ks, dbh + (((ks >> 2) * 4096 + (ks & 3) * 32) >> 4), idesc, ks > 0);
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                /* x_hi . W1_lo */
                sm100::mma_tf32_ts(d, ahi + 8 * ks, dbl + (((ks >> 2) * 4096 + (ks & 3) * 32) >> 4), idesc, 1);
#pragma unroll
            for (int ks = 0; ks < KS1; ks++)                                /* x_hi . W1_hi */
                sm100::mma_tf32_ts(d, ahi + 8 * ks, dbh + (((ks >> 2) * 4096 + (ks & 3) * 32) >> 4), idesc, 1);
        }
        sm100::tc_commit(dfull);
    };
    auto issue_project = [&](uint32_t wslot, bool first) {                  /* one thread: D2 (+)= d . W2 of the chunk in slot wslot */
        constexpr uint32_t idesc = sm100::umma_idesc_tf32(128, N3);
        const uint32_t bh = sW_addr + wslot * w_bytes + (uint32_t)off.w2 * 4, bl = bh + N3 * 128;
        const uint64_t dbh = sm100::umma_desc_sw128(bh), dbl = sm100::umma_desc_sw128(bl), da = sm100::umma_desc_sw128(sA2_addr);
#pragma unroll
        for (int kk = 0; kk < 4; kk++)                                      /* d_lo (TMEM) . W2_hi */
            sm100::mma_tf32_ts(colD2, colA2 + 8 * kk, dbh + 2 * kk, idesc, !(first && kk == 0));
#pragma unroll
        for (int kk = 0; kk < 4; kk++)                                      /* d_hi . W2_lo */
            sm100::mma_tf32_ss(colD2, da + 2 * kk, dbl + 2 * kk, idesc, 1);
#pragma unroll
        for (int kk = 0; kk < 4; kk++)                                      /* d_hi . W2_hi */
            sm100::mma_tf32_ss(colD2, da + 2 * kk, dbh + 2 * kk, idesc, 1);
        sm100::tc_commit(pbar);
    };
    auto load_x = [&](long tile, int b) {                                   /* one thread */
        const B2Tile q = b2_tile<S>(a, tile);
        sm100::mbar_arrive_expect_tx(full_x + b, x_bytes);
        sm100::tma_load_4d(sXB + b * a.xrows * SXs, &tmX, 0, q.ix0 + a.xo, q.iy0 + a.yo, q.n, full_x + b);
    };
    auto load_chunk = [&](int c, int wb) {                                  /* one thread */
        sm100::mbar_arrive_expect_tx(full_w + wb, w_bytes);
        b2_bulk_load(sW_addr + (uint32_t)wb * w_bytes, a.wchunks + (long)c * off.total, w_bytes, full_w + wb);
    };

    if (tid == 0 && (long)blockIdx.x < a.ntiles) { load_x(blockIdx.x, 0); load_chunk(0, 0); }
    const bool w_resident = a.NC == 1;
    uint32_t it = 0, cs = 0;                                                /* tiles / chunks consumed so far by this CTA */
    for (long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, it++) {
        const B2Tile q = b2_tile<S>(a, tile);
        const int xb = it & 1;
        const float *sX = sXB + xb * a.xrows * SXs;
        const bool border = !a.frame && (q.iy0 < 0 || q.ix0 < 0 || q.iy0 + a.HH > a.H || q.ix0 + HW > a.W);
        const bool last_tile = tile + gridDim.x >= a.ntiles;

        __syncthreads();                                  /* the previous tile's epilogue no longer reads the other x buffer */
        if (tid == 0 && !last_tile) load_x(tile + gridDim.x, xb ^ 1);
        sm100::mbar_wait(full_x + xb, (it >> 1) & 1);

        /* ---- x tile -> tensor memory as the A operand of the expand GEMM, split hi / lo (thread = halo pixel = TMEM lane).
           Every MMA that read the previous tile's operand has completed (its dfull / pbar waits). ---- */
        for (int mt = warp >> 2; mt < a.nmt; mt += B2_THREADS / 128) {
            const int p = mt * 128 + (warp & 3) * 32 + lane;
            const float *xr = sX + p * SXs;
            const uint32_t acol = colX + tq + (uint32_t)mt * 2 * KP;
#pragma unroll
            for (int ks = 0; ks < KS1; ks++) {
                float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
                if (p < XP) { x0 = *reinterpret_cast<const float4 *>(xr + 8 * ks); x1 = *reinterpret_cast<const float4 *>(xr + 8 * ks + 4); }
                uint32_t hi[8], lo[8];
                b2_split(x0.x, hi[0], lo[0]); b2_split(x0.y, hi[1], lo[1]); b2_split(x0.z, hi[2], lo[2]); b2_split(x0.w, hi[3], lo[3]);
                b2_split(x1.x, hi[4], lo[4]); b2_split(x1.y, hi[5], lo[5]); b2_split(x1.z, hi[6], lo[6]); b2_split(x1.w, hi[7], lo[7]);
                sm100::tmem_st8(acol + 8 * ks, hi);
                sm100::tmem_st8(acol + KP + 8 * ks, lo);
            }
        }
        sm100::tmem_st_wait();
        sm100::tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {                                   /* chunk 0 of this tile: its weights were requested during the previous tile */
            const uint32_t ws0 = w_resident ? 0u : (cs & 1u);
            sm100::mbar_wait(full_w + ws0, w_resident ? 0u : ((cs >> 1) & 1u));
            sm100::tc_fence_after_sync();
            issue_expand(ws0);
        }

        for (int c = 0; c < a.NC; c++, cs++) {
            const uint32_t wb = w_resident ? 0u : (cs & 1u);
            const bool more = c + 1 < a.NC;
            /* the other weight slot held chunk cs-1, whose projection GEMM may still be reading W2: wait for it, then request
               chunk cs+1 (of this tile, or chunk 0 of the next tile) into that slot */
            if (tid == 0 && !w_resident && !(last_tile && !more)) {
                if (cs > 0) sm100::mbar_wait(pbar, (cs - 1) & 1u);
                load_chunk(more ? c + 1 : 0, wb ^ 1u);
            }
            sm100::mbar_wait(full_w + wb, w_resident ? 0u : ((cs >> 1) & 1u));
            const float *wc = sW + wb * off.total;

            /* ---------------- stage A: D1 (tensor memory) -> BN + act -> E rows ---------------- */
            sm100::mbar_wait(dfull, cs & 1u);
            sm100::tc_fence_after_sync();
            for (int mt = warp >> 2; mt < a.nmt; mt += B2_THREADS / 128) {
                const int p = mt * 128 + (warp & 3) * 32 + lane;
                const int2 mp = p < a.xrows ? sMap[p] : make_int2(-1, 0);
                bool inside = true;
                if (border && mp.x >= 0) {                /* halo pixels outside the image are the depthwise conv's zero padding */
                    const int iy = q.iy0 + (mp.y & 0xffff), ix = q.ix0 + (mp.y >> 16);
                    inside = (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
                }
                const uint32_t dcol = colD1 + tq + (uint32_t)mt * B2_CH;
                uint32_t r0[16], r1[16];
                sm100::tmem_ld16(dcol, r0);
                sm100::tmem_ld16(dcol + 16, r1);
                sm100::tmem_ld_wait();
                if (mp.x >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 s1 = *reinterpret_cast<const float4 *>(wc + off.s1 + 4 * j);
                        const float4 b1 = *reinterpret_cast<const float4 *>(wc + off.b1 + 4 * j);
                        const uint32_t *r = j < 4 ? r0 + 4 * j : r1 + 4 * (j - 4);
                        float4 v = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                        v = inside ? b2_bn_act(v, s1, b1, a.slope1) : make_float4(0.f, 0.f, 0.f, 0.f);
                        sm100::sts128(sE_addr + mp.x + 16 * j, v);
                    }
                }
            }
            sm100::tc_fence_before_sync();
            __syncthreads();                              /* E complete, D1 drained */
            bool expand_pending = more;
            if (tid == 0 && expand_pending && sm100::mbar_try_wait(full_w + (wb ^ 1u), ((cs + 1) >> 1) & 1u)) {
                sm100::tc_fence_after_sync();
                issue_expand(wb ^ 1u);                    /* the next chunk's expand GEMM runs under stage B */
                expand_pending = false;
            }

            /* ---------------- stage B: 3x3 depthwise (thread = output pixel x 16 channels) -> A operand of the projection ---------------- */
            if (cs > 0) sm100::mbar_wait(pbar, (cs - 1) & 1u);          /* the previous projection GEMM has finished reading A2 (hi tile and TMEM) */
            sm100::tc_fence_after_sync();
            {
                uint32_t lo[16];
#pragma unroll
                for (int jc = 0; jc < 4; jc++) {
                    const int c4 = 4 * hb + jc;                            /* 16-byte chunk (4 channels) of the 128-byte row */
                    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row_in_tile) {
#pragma unroll
                        for (int ky = 0; ky < 3; ky++)
#pragma unroll
                            for (int kx = 0; kx < 3; kx++) {
                                const float4 e = sm100::lds128(e_base + ky * e_row + (uint32_t)(kx * B2_SE * 4) + jc * 16);
                                const float4 w = *reinterpret_cast<const float4 *>(wc + off.wd + (ky * 3 + kx) * 32 + 4 * c4);
                                b2_fma4(d, e, w);
                            }
                        const float4 sd = *reinterpret_cast<const float4 *>(wc + off.sd + 4 * c4);
                        const float4 bd = *reinterpret_cast<const float4 *>(wc + off.bd + 4 * c4);
                        d = b2_bn_act(d, sd, bd, a.sloped);
                    }
                    uint32_t hi[4];
                    b2_split(d.x, hi[0], lo[4 * jc]); b2_split(d.y, hi[1], lo[4 * jc + 1]); b2_split(d.z, hi[2], lo[4 * jc + 2]); b2_split(d.w, hi[3], lo[4 * jc + 3]);
                    sm100::sts128(sA2_addr + row * 128 + ((c4 ^ (row & 7)) << 4),
                                  make_float4(__uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3])));
                }
                uint32_t l0[8], l1[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { l0[i] = lo[i]; l1[i] = lo[8 + i]; }
                sm100::tmem_st8(colA2 + tq + 16 * hb, l0);
                sm100::tmem_st8(colA2 + tq + 16 * hb + 8, l1);
            }
            sm100::fence_proxy_async_smem();              /* hi tile: generic-proxy writes -> visible to the tensor core */
            sm100::tmem_st_wait();
            sm100::tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                sm100::tc_fence_after_sync();
                if (expand_pending) {                     /* the next chunk's weights had not landed when stage B started */
                    sm100::mbar_wait(full_w + (wb ^ 1u), ((cs + 1) >> 1) & 1u);
                    issue_expand(wb ^ 1u);
                }
                issue_project(wb, c == 0);
            }
        }

        /* ---------------- block epilogue: D2 -> BN + act [+ shortcut from the resident x tile] -> y ---------------- */
        sm100::mbar_wait(pbar, (cs - 1) & 1u);
        sm100::tc_fence_after_sync();
        if (hb * 16 < N3) {
            const bool valid = row_in_tile && ty < q.th && tx < q.tw;
            float *yp = a.y + (((long)q.n * a.OH + q.oy0 + ty) * a.OW + q.ox0 + tx) * a.ldy;
            const float *xc = sX + ((ty + 1 - a.yo) * a.XW + tx + 1 - a.xo) * SXs;       /* centre pixel; S == 1 whenever res is set */
#pragma unroll
            for (int j = 0; j < N3 / 16; j++) {
                if ((j & 1) != hb && N3 > 16) continue;                    /* 16-column blocks alternate between the two warp halves */
                uint32_t r[16];
                sm100::tmem_ld16(colD2 + tq + 16 * j, r);
                sm100::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int co = 16 * j + 4 * k;
                        if (co < a.cout) {                                  /* cout is a multiple of 4 */
                            const float4 s3 = *reinterpret_cast<const float4 *>(sSB3 + co), b3 = *reinterpret_cast<const float4 *>(sSB3 + N3 + co);
                            float4 v = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
                            v = b2_bn_act(v, s3, b3, a.slope3);
                            if (a.res) {
                                const float4 xr = *reinterpret_cast<const float4 *>(xc + co);
                                v.x = b2_act(v.x + xr.x, a.slope_res); v.y = b2_act(v.y + xr.y, a.slope_res);
                                v.z = b2_act(v.z + xr.z, a.slope_res); v.w = b2_act(v.w + xr.w, a.slope_res);
                            }
                            *reinterpret_cast<float4 *>(yp + co) = v;
                        }
                    }
                }
            }
        }
        sm100::tc_fence_before_sync();                    /* D2 drained before the next tile's first projection overwrites it */
    }
    sm100::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) { sm100::tc_fence_after_sync(); sm100::tmem_dealloc(tmem_base, a.tmem_cols); }
}

/* Build the chunk images of one block from the three convs' packed reference rows (ffcnn.c:218-234:
 * [weights..., pad, scale', bias', mean, var] per filter).  Element (n, k) of a SWIZZLE_128B tile [rows][32 floats] sits at
 * float  n*32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3). */
__global__ void k_prep_block2(const float *__restrict__ p1, int row1, int cin, const float *__restrict__ pd, int rowd,
                              const float *__restrict__ p3, int row3, int cexp, int cout, int KS1, int N3, int NC,
                              float *__restrict__ chunks, float *__restrict__ sb3)
{
    const Blk2Chunk off(KS1, N3);
    const long total = (long)NC * off.total;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total + 2 * N3; i += (long)gridDim.x * blockDim.x) {
        if (i >= total) {                                  /* projection scale / bias */
            const int k = (int)(i - total), co = k % N3, which = k / N3;
            sb3[k] = co < cout ? p3[(long)co * row3 + row3 - 4 + which] : 0.f;
            continue;
        }
        const int c = (int)(i / off.total), r = (int)(i - (long)c * off.total);
        float v = 0.f; int split = 0;                      /* 0: plain value, 1: hi part, 2: lo part */
        if (r < off.w2) {                                  /* W1: [hi | lo] x KC sub-tiles of [32 channels x 32 cin] */
            const int KC = (KS1 + 3) / 4, sub = r >> 10, rem = r & 1023, n = rem >> 5, pos = rem & 31;
            const int k = (((pos >> 2) ^ (n & 7)) << 2) | (pos & 3), part = sub / KC, kc = sub - part * KC;
            const int ch = c * B2_CH + n, ci = kc * 32 + k;
            if (ch < cexp && ci < cin) v = p1[(long)ch * row1 + ci];
            split = 1 + part;
        } else if (r < off.s1) {                           /* W2: [hi | lo] x [N3 cout x 32 channels] */
            const int rr = r - off.w2, part = rr / (N3 * 32), rem = rr - part * N3 * 32, n = rem >> 5, pos = rem & 31;
            const int k = (((pos >> 2) ^ (n & 7)) << 2) | (pos & 3), ch = c * B2_CH + k;
            if (n < cout && ch < cexp) v = p3[(long)n * row3 + ch];
            split = 1 + part;
        } else if (r < off.wd) {                           /* expand scale, bias */
            const int rr = r - off.s1, which = rr >> 5, ch = c * B2_CH + (rr & 31);
            if (ch < cexp) v = p1[(long)ch * row1 + row1 - 4 + which];
        } else if (r < off.sd) {                           /* depthwise taps [9][32] */
            const int rr = r - off.wd, tap = rr >> 5, ch = c * B2_CH + (rr & 31);
            if (ch < cexp) v = pd[(long)ch * rowd + tap];
        } else if (r < off.bd + 32) {                      /* depthwise scale, bias */
            const int rr = r - off.sd, which = rr >> 5, ch = c * B2_CH + (rr & 31);
            if (ch < cexp) v = pd[(long)ch * rowd + rowd - 4 + which];
        }
        if (split) { uint32_t hi, lo; b2_split(v, hi, lo); v = __uint_as_float(split == 1 ? hi : lo); }
        chunks[i] = v;
    }
}

} // namespace ffb
