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

struct DwArgs {
    float *out; const float *wt, *scale, *bias;     /* wt: [9][C] tap-major */
    int N, H, W, C;
    int CB, TW, TH, RC, nch;                         /* channel block, output tile, rows per work item, row chunks */
    int ntx, nty, ntc; long ntiles;
    int stages, act;
};

__global__ void __launch_bounds__(DW_THREADS)
k_dw3s1_tma(const __grid_constant__ CUtensorMap tmIn, const DwArgs a)
{
    extern __shared__ uint8_t dw_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
    const int S = a.stages, IWb = 2 * ((a.TW + 1) / 2) + 2, IHb = a.TH + 2;   /* box width covers whole pixel pairs */
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

    float4 wv[9], sc, bi;
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
            for (int t = 0; t < 9; t++) wv[t] = ldg4(a.wt + t * a.C + c0);
            sc = ldg4(a.scale + c0); bi = ldg4(a.bias + c0);
            wc0 = c0;
        }
        sm100::mbar_wait(full + s, ph);
        if (worker && ox < a.W) {
            const float *col = reinterpret_cast<const float *>(smem + (size_t)s * stage_stride) + col_off;
            const int yl1 = min(yl1_tile, a.H - oy0);
            const bool two = (xl + 1 < a.TW) && (ox + 1 < a.W);
            float *dst = a.out + (((long)n * a.H + oy0) * a.W + ox) * a.C + c0;
            const long orow = (long)a.W * a.C;
            float4 win[3][4];
            auto load_row = [&](int slot, int yrow) {
                const float *rp = col + yrow * srow;
#pragma unroll
                for (int k = 0; k < 4; k++) win[slot][k] = *reinterpret_cast<const float4 *>(rp + k * a.CB);
            };
            load_row(0, yl0); load_row(1, yl0 + 1);
            for (int yb = yl0; yb < yl1; yb += 3) {
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int yl = yb + u;
                    if (yl < yl1) {
                        load_row((u + 2) % 3, yl + 2);
                        float4 acc0 = zero4(), acc1 = zero4();
#pragma unroll
                        for (int j = 0; j < 3; j++)
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                fma4(acc0, win[(u + j) % 3][k], wv[j * 3 + k]);
                                fma4(acc1, win[(u + j) % 3][k + 1], wv[j * 3 + k]);
                            }
                        float *o = dst + yl * orow;
                        *reinterpret_cast<float4 *>(o) = epilogue4(acc0, sc, bi, a.act);
                        if (two) *reinterpret_cast<float4 *>(o + a.C) = epilogue4(acc1, sc, bi, a.act);
                    }
                }
            }
        }
        __syncthreads();                                          /* everyone is done with stage s before it is refilled */
    }
}

} // namespace ffb
