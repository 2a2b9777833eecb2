/*
 * dw_tma.cuh -- depthwise 3x3, stride 1, pad 1 (conv-v6.c:96-229) as a TMA-fed shared-memory stencil.
 *
 * 20 of the 84 conv layers and the largest single share of the forward pass.  2.2 FLOP per byte: pure HBM streaming.
 * The first version (k_dw_s1 in kernels.cuh: a register window fed by LDG) measured 30 % of DRAM peak with exactly
 * algorithmic DRAM traffic (profiles/r1b): latency bound -- 16 resident warps x 3 loads, two of them L1 re-reads, leave
 * ~8 KB of unique bytes in flight per SM where Little's law wants ~32 KB.  Here the bytes in flight are decoupled from
 * warps and registers:
 *
 *   - the NHWC activation tensor is described to the TMA unit as a 4-D tensor (C, W, H, N); one thread pulls
 *     [(TH+2) x (TW+2) x CB] input boxes (output tile + halo) into a 3-stage shared-memory ring (mbarrier completion).
 *     Boxes start at (x0 - 1, y0 - 1): the hardware's out-of-bounds ZERO FILL is the convolution's zero padding on all
 *     four image borders, so the compute loop has no edge cases;
 *   - every thread owns one fixed work item of the tile geometry -- 2 adjacent output pixels x 4 channels x RC rows --
 *     for the whole persistent loop, so its 9 weight vectors / scale / bias and all index arithmetic are hoisted out;
 *     walking down the rows it keeps a 3-row x 4-column register window (slot = row mod 3 resolved at compile time),
 *     i.e. 4 LDS.128 per row for 2 outputs; consecutive lanes read consecutive 16 B -> conflict free;
 *   - outputs leave as float4 stores with the fused epilogue act(fma(sum, scale, bias)).
 *
 * Accumulation order kernel-row -> kernel-column as conv-v0.c:17-24.
 */
#pragma once
#include "kernels.cuh"
#include "sm100.cuh"

namespace ffb {

constexpr int DW_THREADS = 384;

/* The stencil loops keep four channels as two packed fp32 pairs (sm100.cuh: FFMA2 does two IEEE fp32 FMAs per issue slot, results
 * bit-identical to fmaf), loaded from the staged tile with explicit 128-bit shared loads. */
using sm100::f4p; using sm100::lds128p; using sm100::ldg128p; using sm100::fma4p; using sm100::zero4p;
__device__ __forceinline__ float4 epilogue4p(const f4p a, const f4p s, const f4p b, int act)
{
    const sm100::f32x2 lo = sm100::f2_fma(a.a, s.a, b.a), hi = sm100::f2_fma(a.b, s.b, b.b);
    return make_float4(act_apply(sm100::f2_lo(lo), act), act_apply(sm100::f2_hi(lo), act), act_apply(sm100::f2_lo(hi), act), act_apply(sm100::f2_hi(hi), act));
}

struct DwArgs {
    float *out; const float *wt, *scale, *bias;     /* wt: [FS*FS][C] tap-major */
    int N, H, W, C, OH, OW;                          /* input and output geometry (OH = H, OW = W for stride 1) */
    int IWb, IHb, skip_row0_at;                      /* TMA box extent in pixels; conv-v6 5x5 quirk row (-1: none) */
    int CB, TW, TH, RC, nch;                         /* channel block, output tile, rows per work item, row chunks */
    int ntx, nty, ntc; long ntiles;
    int stages, act;
};

__global__ void __launch_bounds__(DW_THREADS)
k_dw3s1_tma(const __grid_constant__ CUtensorMap tmIn, const DwArgs a)
{
    extern __shared__ uint8_t dw_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
    const int S = a.stages, IWb = a.IWb, IHb = a.IHb;            /* box = output tile (whole pixel pairs) + 1-pixel halo */
    const uint32_t stage_bytes = (uint32_t)IHb * IWb * a.CB * 4;
    const uint32_t stage_stride = (stage_bytes + 127) & ~127u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_stride);
    const int tid = threadIdx.x;

    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmIn);
        for (int s = 0; s < S; s++) sm100::mbar_init(full + s, 1);
        sm100::fence_barrier_init();
    }
    __syncthreads();
    pdl_trigger(); pdl_wait();

    auto issue = [&](long tile, int s) {
        long r = tile;
        const int tc = (int)(r % a.ntc); r /= a.ntc;
        const int tx = (int)(r % a.ntx); r /= a.ntx;
        const int ty = (int)(r % a.nty); const int n = (int)(r / a.nty);
        sm100::mbar_arrive_expect_tx(full + s, stage_bytes);
        sm100::tma_load_4d(smem + (size_t)s * stage_stride, &tmIn, tc * a.CB, tx * a.TW - 1, ty * a.TH - 1, n, full + s);
    };
    const long first = blockIdx.x, step = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < S - 1; s++) if (first + s * step < a.ntiles) issue(first + s * step, s);

    /* this thread's fixed work item: row chunk `ch`, pixel pair `xp`, channel quad `c` */
    const int cb4 = a.CB / 4, pairs = (a.TW + 1) / 2;
    const int per_chunk = pairs * cb4;
    const bool worker = tid < per_chunk * a.nch;
    const int ch = tid / per_chunk, jj0 = tid - ch * per_chunk;
    const int xp = jj0 / cb4, c = (jj0 - xp * cb4) * 4;
    const int xl = 2 * xp;                                        /* local x of the first of the two outputs */
    const int yl0 = ch * a.RC, yl1_tile = min(a.TH, yl0 + a.RC);
    const int srow = IWb * a.CB;                                  /* floats per staged input row */
    const int col_off = xl * a.CB + c;                            /* top-left tap of output (yl = 0, xl) inside a stage */

    f4p wv[9], sc, bi;
    int wc0 = -1;

    long it = 0;
    for (long tile = first; tile < a.ntiles; tile += step, it++) {
        const int s = (int)(it % S); const uint32_t ph = (uint32_t)((it / S) & 1);
        if (tid == 0) { const long nx = tile + (long)(S - 1) * step; if (nx < a.ntiles) issue(nx, (int)((it + S - 1) % S)); }
        long r = tile;
        const int tc = (int)(r % a.ntc); r /= a.ntc;
        const int tx = (int)(r % a.ntx); r /= a.ntx;
        const int ty = (int)(r % a.nty); const int n = (int)(r / a.nty);
        const int ox = tx * a.TW + xl, oy0 = ty * a.TH, c0 = tc * a.CB + c;
        if (worker && c0 != wc0) {                                /* weights change only when the channel block does */
#pragma unroll
            for (int t = 0; t < 9; t++) wv[t] = ldg128p(a.wt + t * a.C + c0);
            sc = ldg128p(a.scale + c0); bi = ldg128p(a.bias + c0);
            wc0 = c0;
        }
        sm100::mbar_wait(full + s, ph);
        if (worker && ox < a.W) {
            const uint32_t col = sm100::smem_u32(smem + (size_t)s * stage_stride) + (uint32_t)col_off * 4;
            const int yl1 = min(yl1_tile, a.H - oy0);
            const bool two = (xl + 1 < a.TW) && (ox + 1 < a.W);
            float *dst = a.out + (((long)n * a.H + oy0) * a.W + ox) * a.C + c0;
            const long orow = (long)a.W * a.C;
            f4p win[3][4];
            auto load_row = [&](int slot, int yrow) {
                const uint32_t rp = col + (uint32_t)(yrow * srow) * 4;
#pragma unroll
                for (int k = 0; k < 4; k++) win[slot][k] = lds128p(rp + (uint32_t)(k * a.CB) * 4);
            };
            load_row(0, yl0); load_row(1, yl0 + 1);
            for (int yb = yl0; yb < yl1; yb += 3) {
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int yl = yb + u;
                    if (yl < yl1) {
                        load_row((u + 2) % 3, yl + 2);
                        f4p acc0 = zero4p(), acc1 = zero4p();
#pragma unroll
                        for (int j = 0; j < 3; j++)
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                fma4p(acc0, win[(u + j) % 3][k], wv[j * 3 + k]);
                                fma4p(acc1, win[(u + j) % 3][k + 1], wv[j * 3 + k]);
                            }
                        float *o = dst + yl * orow;
                        *reinterpret_cast<float4 *>(o) = epilogue4p(acc0, sc, bi, a.act);
                        if (two) *reinterpret_cast<float4 *>(o + a.C) = epilogue4p(acc1, sc, bi, a.act);
                    }
                }
            }
        }
        __syncthreads();                                          /* everyone is done with stage s before it is refilled */
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Common prologue of the TMA-fed stencils: barrier ring, tile decode, TMA issue.  ORGX/ORGY map an output tile origin to
 * the input box origin (stride * origin - pad).
 * ---------------------------------------------------------------------------------------------------------------- */
struct DwTile { int tc, tx, ty, n; };
__device__ __forceinline__ DwTile dw_decode(long tile, const DwArgs &a)
{
    DwTile t; long r = tile;
    t.tc = (int)(r % a.ntc); r /= a.ntc;
    t.tx = (int)(r % a.ntx); r /= a.ntx;
    t.ty = (int)(r % a.nty); t.n = (int)(r / a.nty);
    return t;
}

/* Depthwise 3x3, stride 2, pad 1 (conv-v6.c:233-287).  Same structure as k_dw3s1_tma; the input box of an output tile
 * TW x TH is (2*TW+1) x (2*TH+1) pixels starting at (2*x0-1, 2*y0-1).  A thread owns 2 adjacent output pixels x 4 channels
 * x RC rows: per output row it loads two new 5-column input rows (the third is carried over from the row above). */
__global__ void __launch_bounds__(DW_THREADS)
k_dw3s2_tma(const __grid_constant__ CUtensorMap tmIn, const DwArgs a)
{
    extern __shared__ uint8_t dw_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
    const int S = a.stages;
    const uint32_t stage_bytes = (uint32_t)a.IHb * a.IWb * a.CB * 4;
    const uint32_t stage_stride = (stage_bytes + 127) & ~127u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_stride);
    const int tid = threadIdx.x;
    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmIn);
        for (int s = 0; s < S; s++) sm100::mbar_init(full + s, 1);
        sm100::fence_barrier_init();
    }
    __syncthreads();
    pdl_trigger(); pdl_wait();
    auto issue = [&](long tile, int s) {
        const DwTile t = dw_decode(tile, a);
        sm100::mbar_arrive_expect_tx(full + s, stage_bytes);
        sm100::tma_load_4d(smem + (size_t)s * stage_stride, &tmIn, t.tc * a.CB, 2 * t.tx * a.TW - 1, 2 * t.ty * a.TH - 1, t.n, full + s);
    };
    const long first = blockIdx.x, step = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < S - 1; s++) if (first + s * step < a.ntiles) issue(first + s * step, s);

    const int cb4 = a.CB / 4, pairs = (a.TW + 1) / 2, per_chunk = pairs * cb4;
    const bool worker = tid < per_chunk * a.nch;
    const int ch = tid / per_chunk, jj0 = tid - ch * per_chunk;
    const int xp = jj0 / cb4, c = (jj0 - xp * cb4) * 4;
    const int xl = 2 * xp;
    const int yl0 = ch * a.RC, yl1_tile = min(a.TH, yl0 + a.RC);
    const int srow = a.IWb * a.CB;
    const int col_off = (2 * xl) * a.CB + c;                      /* input column 2*xl of box row 0 */
    f4p wv[9], sc, bi; int wc0 = -1;

    long it = 0;
    for (long tile = first; tile < a.ntiles; tile += step, it++) {
        const int s = (int)(it % S); const uint32_t ph = (uint32_t)((it / S) & 1);
        if (tid == 0) { const long nx = tile + (long)(S - 1) * step; if (nx < a.ntiles) issue(nx, (int)((it + S - 1) % S)); }
        const DwTile t = dw_decode(tile, a);
        const int ox = t.tx * a.TW + xl, oy0 = t.ty * a.TH, c0 = t.tc * a.CB + c;
        if (worker && c0 != wc0) {
#pragma unroll
            for (int k = 0; k < 9; k++) wv[k] = ldg128p(a.wt + k * a.C + c0);
            sc = ldg128p(a.scale + c0); bi = ldg128p(a.bias + c0); wc0 = c0;
        }
        sm100::mbar_wait(full + s, ph);
        if (worker && ox < a.OW) {
            const uint32_t col = sm100::smem_u32(smem + (size_t)s * stage_stride) + (uint32_t)col_off * 4;
            const int yl1 = min(yl1_tile, a.OH - oy0);
            const bool two = (xl + 1 < a.TW) && (ox + 1 < a.OW);
            float *dst = a.out + (((long)t.n * a.OH + oy0) * a.OW + ox) * a.C + c0;
            const long orow = (long)a.OW * a.C;
            f4p top[5], mid[5], bot[5];
            auto load_row = [&](f4p (&r)[5], int yrow) {
                const uint32_t rp = col + (uint32_t)(yrow * srow) * 4;
#pragma unroll
                for (int k = 0; k < 5; k++) r[k] = lds128p(rp + (uint32_t)(k * a.CB) * 4);
            };
            load_row(top, 2 * yl0);
            for (int yl = yl0; yl < yl1; yl++) {
                load_row(mid, 2 * yl + 1);
                load_row(bot, 2 * yl + 2);
                f4p acc0 = zero4p(), acc1 = zero4p();
#pragma unroll
                for (int k = 0; k < 3; k++) { fma4p(acc0, top[k], wv[k]);     fma4p(acc1, top[k + 2], wv[k]); }
#pragma unroll
                for (int k = 0; k < 3; k++) { fma4p(acc0, mid[k], wv[3 + k]); fma4p(acc1, mid[k + 2], wv[3 + k]); }
#pragma unroll
                for (int k = 0; k < 3; k++) { fma4p(acc0, bot[k], wv[6 + k]); fma4p(acc1, bot[k + 2], wv[6 + k]); }
                float *o = dst + yl * orow;
                *reinterpret_cast<float4 *>(o) = epilogue4p(acc0, sc, bi, a.act);
                if (two) *reinterpret_cast<float4 *>(o + a.C) = epilogue4p(acc1, sc, bi, a.act);
#pragma unroll
                for (int k = 0; k < 5; k++) top[k] = bot[k];
            }
        }
        __syncthreads();
    }
}

/* Depthwise 5x5, stride 1, pad 2 (conv-v6.c:291-465).  A work item is 2 adjacent output pixels x 2 channels (one packed fp32
 * pair) x DW5_RC output rows; the threads of the CTA loop over the items of a tile.  Two channels instead of four is what lets
 * the item keep ALL 25 weight pairs in registers (50) and walk the INPUT rows once: each of the DW5_RC + 4 rows is read with
 * six 64-bit shared loads and feeds every output row it belongs to, 12 loads per output pixel pair instead of the 30 of the
 * first version (kernel-row outer loop over a 4-channel item, which re-read every row five times and was bound by
 * shared-memory wavefronts: 29 % of the HBM roofline).  Per output the taps are still added in the reference's order, kernel row
 * -> kernel column (conv-v0.c:17-24).  skip_row0_at: conv-v6 drops kernel row 0 on output row oh-2 (422-441). */
constexpr int DW5_RC = 4;
__device__ __forceinline__ sm100::f32x2 dw5_lds64(uint32_t addr)
{
    sm100::f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ sm100::f32x2 dw5_ldg64(const float *p)
{
    sm100::f32x2 v;
    asm("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(DW_THREADS)
k_dw5s1_tma(const __grid_constant__ CUtensorMap tmIn, const DwArgs a)
{
    using sm100::f32x2; using sm100::f2_fma; using sm100::f2_lo; using sm100::f2_hi;
    extern __shared__ uint8_t dw_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
    const int S = a.stages;
    const uint32_t stage_bytes = (uint32_t)a.IHb * a.IWb * a.CB * 4;
    const uint32_t stage_stride = (stage_bytes + 127) & ~127u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * stage_stride);
    const int tid = threadIdx.x;
    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmIn);
        for (int s = 0; s < S; s++) sm100::mbar_init(full + s, 1);
        sm100::fence_barrier_init();
    }
    __syncthreads();
    pdl_trigger(); pdl_wait();
    auto issue = [&](long tile, int s) {
        const DwTile t = dw_decode(tile, a);
        sm100::mbar_arrive_expect_tx(full + s, stage_bytes);
        sm100::tma_load_4d(smem + (size_t)s * stage_stride, &tmIn, t.tc * a.CB, t.tx * a.TW - 2, t.ty * a.TH - 2, t.n, full + s);
    };
    const long first = blockIdx.x, step = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < S - 1; s++) if (first + s * step < a.ntiles) issue(first + s * step, s);

    const int cb2 = a.CB / 2, pairs = (a.TW + 1) / 2;
    const int nitems = a.nch * pairs * cb2;                       /* a.RC == DW5_RC rows per item */
    const uint32_t srow = (uint32_t)a.IWb * a.CB * 4, spx = (uint32_t)a.CB * 4;   /* bytes per staged row / pixel */

    long it = 0;
    for (long tile = first; tile < a.ntiles; tile += step, it++) {
        const int s = (int)(it % S); const uint32_t ph = (uint32_t)((it / S) & 1);
        if (tid == 0) { const long nx = tile + (long)(S - 1) * step; if (nx < a.ntiles) issue(nx, (int)((it + S - 1) % S)); }
        const DwTile t = dw_decode(tile, a);
        const int oy0 = t.ty * a.TH;
        const uint32_t stage = sm100::smem_u32(smem + (size_t)s * stage_stride);
        sm100::mbar_wait(full + s, ph);
        for (int item = tid; item < nitems; item += DW_THREADS) {
            const int c2 = item % cb2, rest = item / cb2, xp = rest % pairs, ch = rest / pairs;
            const int xl = 2 * xp, ox = t.tx * a.TW + xl, c0 = t.tc * a.CB + 2 * c2;
            const int yl0 = ch * DW5_RC, yl1 = min(min(a.TH, yl0 + DW5_RC), a.H - oy0);
            if (ox >= a.W || yl0 >= yl1) continue;
            const bool two = (xl + 1 < a.TW) && (ox + 1 < a.W);
            f32x2 w[25];
#pragma unroll
            for (int k = 0; k < 25; k++) w[k] = dw5_ldg64(a.wt + k * a.C + c0);
            const uint32_t col = stage + (uint32_t)(yl0 * a.IWb + xl) * spx + (uint32_t)c2 * 8;
            f32x2 acc[DW5_RC][2];
#pragma unroll
            for (int r = 0; r < DW5_RC; r++) { acc[r][0] = 0ull; acc[r][1] = 0ull; }
#pragma unroll
            for (int iy = 0; iy < DW5_RC + 4; iy++) {             /* input row yl0 + iy of the staged box */
                if (yl0 + (iy > 4 ? iy - 4 : 0) < yl1) {          /* some valid output row of the item uses it */
                    f32x2 v[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) v[k] = dw5_lds64(col + iy * srow + k * spx);
#pragma unroll
                    for (int r = 0; r < DW5_RC; r++) {
                        const int j = iy - r;                     /* kernel row of this input row for output row r (compile time) */
                        if (j >= 0 && j < 5) {
                            if (yl0 + r < yl1 && !(j == 0 && oy0 + yl0 + r == a.skip_row0_at)) {
#pragma unroll
                                for (int k = 0; k < 5; k++) { acc[r][0] = f2_fma(v[k], w[j * 5 + k], acc[r][0]); acc[r][1] = f2_fma(v[k + 1], w[j * 5 + k], acc[r][1]); }
                            }
                        }
                    }
                }
            }
            const f32x2 sc = dw5_ldg64(a.scale + c0), bi = dw5_ldg64(a.bias + c0);
            float *dst = a.out + (((long)t.n * a.H + oy0 + yl0) * a.W + ox) * a.C + c0;
            const long orow = (long)a.W * a.C;
#pragma unroll
            for (int r = 0; r < DW5_RC; r++) {
                if (yl0 + r < yl1) {
                    const f32x2 o0 = f2_fma(acc[r][0], sc, bi), o1 = f2_fma(acc[r][1], sc, bi);
                    *reinterpret_cast<float2 *>(dst + r * orow) = make_float2(act_apply(f2_lo(o0), a.act), act_apply(f2_hi(o0), a.act));
                    if (two) *reinterpret_cast<float2 *>(dst + r * orow + a.C) = make_float2(act_apply(f2_lo(o1), a.act), act_apply(f2_hi(o1), a.act));
                }
            }
        }
        __syncthreads();
    }
}

} // namespace ffb
