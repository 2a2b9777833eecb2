/*
 * block_s2.cuh -- the stride-2 block 4->24->8 of yolo-fastest-1.1 (L9-L11: 1x1 expand 4->24, 3x3 depthwise stride 2, 1x1 project 24->8
 * at 160x160 -> 80x80) as ONE shared-memory-tiled kernel.
 *
 * The register-resident kernel (block_reg.cuh::k_block_reg_s2) keeps a pixel's whole expanded vector in a lane's registers, which
 * for 24 channels takes three launches of 8 channels each, every one re-deriving its addresses, predicates and shuffles and going
 * through y: ~1900 instructions per output pixel (ncu r2v).  Here a CTA owns a 16 x 8 tile of output pixels:
 *
 *   phase B   one thread per input pixel of the 33 x 17 halo tile: x (one float4) straight from global memory, expanded to the 24
 *             channels in registers, BN + act, zero outside the image (the depthwise conv's padding), six float4 into shared memory;
 *   phase C   two threads per output pixel, 12 channels each: 9 taps x 3 float4 from shared memory, BN + act; the pair exchanges its
 *             depthwise results by shuffle so that each thread projects all 24 channels, in order, onto four of the eight outputs.
 *
 * e of input pixel (r, c) and channel quad q lives at E[r][c & 1][q][c >> 1]: the threads of phase C read columns two apart, so
 * consecutive threads touch consecutive float4 of one plane; phase B writes two runs of consecutive float4 per warp.
 *
 * Arithmetic and its order are k_block_reg_s2's (expand: input channel ascending, first term a product; depthwise: ky -> kx, first tap
 * a product; projection: expanded channel 0..23 ascending from zero -- which is what the three slices accumulating through y amount
 * to), as packed fp32 pairs: the output is BIT-IDENTICAL to the three-launch path (tests/test_gpu_parity.py).
 */
#pragma once
#include "block_reg.cuh"

namespace ffb {

constexpr int S2_TXO = 16, S2_TYO = 8;                    /* output tile */
constexpr int S2_IW = 2 * S2_TXO + 1, S2_IH = 2 * S2_TYO + 1;   /* halo tile of input pixels */
constexpr int S2_COLS = S2_TXO + 1;                       /* columns per parity plane */
constexpr int S2_THREADS = 288;                           /* 561 input pixels = 2 passes; 256 threads in phase C */
constexpr int S2_E4 = S2_IH * 2 * 6 * S2_COLS;            /* float4 of the E tile */

struct S2Args {
    const float *x; float *y;
    int N, H, W, OH, OW;
    float slope1, sloped, slope3;
};

/* the parts of the block's weights phase C indexes per thread (channel half, output half): shared memory */
struct S2Shared { float wd[9][24], sd[24], bd[24], w2[24][8], s3[8], b3[8]; };

__global__ void __launch_bounds__(S2_THREADS, 3)
k_block_s2_tile(const __grid_constant__ RegBlockW<4, 24, 8> w, const S2Args a)
{
    extern __shared__ float4 s2_smem[];
    float4 *E = s2_smem;
    S2Shared &sw = *reinterpret_cast<S2Shared *>(s2_smem + S2_E4);
    const int tid = threadIdx.x;
    {   /* weights of phase C -> shared memory (constant data: before the dependency wait) */
        float *dst = reinterpret_cast<float *>(&sw);
        for (int i = tid; i < 9 * 24; i += S2_THREADS) dst[i] = w.wd[i / 24][i % 24];
        for (int i = tid; i < 24; i += S2_THREADS) { sw.sd[i] = w.sd[i]; sw.bd[i] = w.bd[i]; }
        for (int i = tid; i < 24 * 8; i += S2_THREADS) sw.w2[i / 8][i % 8] = w.w2[i / 8][i % 8];
        if (tid < 8) { sw.s3[tid] = w.s3[tid]; sw.b3[tid] = w.b3[tid]; }
    }
    sm100::pdl_trigger(); sm100::pdl_wait();
    const int ox0 = blockIdx.x * S2_TXO, oy0 = blockIdx.y * S2_TYO;
    const long n = blockIdx.z;
    const float *xf = a.x + n * (long)a.H * a.W * 4;
    const f32x2 slope1 = f2_pack(a.slope1, a.slope1);

    /* ---------------- phase B: expand every input pixel of the halo tile ---------------- */
#pragma unroll
    for (int pass = 0; pass < (S2_IW * S2_IH + S2_THREADS - 1) / S2_THREADS; pass++) {
        const int p = tid + pass * S2_THREADS;
        if (p < S2_IW * S2_IH) {
            const int r = p / S2_IW, c = p - r * S2_IW;
            const int iy = 2 * oy0 - 1 + r, ix = 2 * ox0 - 1 + c;
            const bool inside = (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W;
            float4 *dst = E + ((r * 2 + (c & 1)) * 6) * S2_COLS + (c >> 1);
            if (inside) {
                const float4 xv = __ldg(reinterpret_cast<const float4 *>(xf + ((long)iy * a.W + ix) * 4));
                const float x[4] = { xv.x, xv.y, xv.z, xv.w };
                f32x2 e[4];
                rb_expand<0>(w, x, true, slope1, e);
                dst[0 * S2_COLS] = make_float4(f2_lo(e[0]), f2_hi(e[0]), f2_lo(e[1]), f2_hi(e[1]));
                dst[1 * S2_COLS] = make_float4(f2_lo(e[2]), f2_hi(e[2]), f2_lo(e[3]), f2_hi(e[3]));
                rb_expand<8>(w, x, true, slope1, e);
                dst[2 * S2_COLS] = make_float4(f2_lo(e[0]), f2_hi(e[0]), f2_lo(e[1]), f2_hi(e[1]));
                dst[3 * S2_COLS] = make_float4(f2_lo(e[2]), f2_hi(e[2]), f2_lo(e[3]), f2_hi(e[3]));
                rb_expand<16>(w, x, true, slope1, e);
                dst[4 * S2_COLS] = make_float4(f2_lo(e[0]), f2_hi(e[0]), f2_lo(e[1]), f2_hi(e[1]));
                dst[5 * S2_COLS] = make_float4(f2_lo(e[2]), f2_hi(e[2]), f2_lo(e[3]), f2_hi(e[3]));
            } else {
#pragma unroll
                for (int q = 0; q < 6; q++) dst[q * S2_COLS] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncthreads();

    /* ---------------- phase C: depthwise stride 2 (12 channels per thread) -> exchange -> projection (4 outputs per thread) ---------------- */
    if (tid < 2 * S2_TXO * S2_TYO) {
        const int px = tid >> 1, half = tid & 1, ty = px / S2_TXO, tx = px - ty * S2_TXO;
        const int oy = oy0 + ty, ox = ox0 + tx;
        f32x2 d[6];                                       /* channel pairs 6*half .. 6*half+5 */
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float4 *src = E + (((2 * ty + r) * 2 + (k & 1)) * 6 + 3 * half) * S2_COLS + tx + (k >> 1);
                const float *wt = sw.wd[r * 3 + k] + 12 * half;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const float4 v = src[q * S2_COLS];
                    const float4 wv = *reinterpret_cast<const float4 *>(wt + 4 * q);
                    const f32x2 v0 = f2_pack(v.x, v.y), v1 = f2_pack(v.z, v.w), w0 = f2_pack(wv.x, wv.y), w1 = f2_pack(wv.z, wv.w);
                    if (r == 0 && k == 0) { d[2 * q] = f2_mul(v0, w0); d[2 * q + 1] = f2_mul(v1, w1); }       /* first tap: a product (rb_taps SET) */
                    else { d[2 * q] = f2_fma(v0, w0, d[2 * q]); d[2 * q + 1] = f2_fma(v1, w1, d[2 * q + 1]); }
                }
            }
        const f32x2 sloped = f2_pack(a.sloped, a.sloped);
        float dd[12];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const float *sp = sw.sd + 12 * half + 2 * i, *bp = sw.bd + 12 * half + 2 * i;
            const f32x2 t = rb_act2(f2_fma(d[i], f2_pack(sp[0], sp[1]), f2_pack(bp[0], bp[1])), sloped);
            dd[2 * i] = f2_lo(t); dd[2 * i + 1] = f2_hi(t);
        }
        /* both threads of the pixel need all 24 depthwise results, in channel order: channels 0-11 come from the even lane, 12-23 from the odd one */
        const int lane = tid & 31;
        f32x2 o[2] = { 0ull, 0ull };                      /* output channel pairs 2*half, 2*half+1 */
#pragma unroll
        for (int hh = 0; hh < 2; hh++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const float dc = __shfl_sync(0xffffffffu, dd[i], (lane & ~1) | hh);
                const float *wp = sw.w2[12 * hh + i] + 4 * half;
                o[0] = f2_fma(f2_pack(dc, dc), f2_pack(wp[0], wp[1]), o[0]);
                o[1] = f2_fma(f2_pack(dc, dc), f2_pack(wp[2], wp[3]), o[1]);
            }
        const f32x2 slope3 = f2_pack(a.slope3, a.slope3);
        float out[4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float *sp = sw.s3 + 4 * half + 2 * i, *bp = sw.b3 + 4 * half + 2 * i;
            const f32x2 t = rb_act2(f2_fma(o[i], f2_pack(sp[0], sp[1]), f2_pack(bp[0], bp[1])), slope3);
            out[2 * i] = f2_lo(t); out[2 * i + 1] = f2_hi(t);
        }
        if (oy < a.OH && ox < a.OW)
            *reinterpret_cast<float4 *>(a.y + ((n * a.OH + oy) * (long)a.OW + ox) * 8 + 4 * half) = make_float4(out[0], out[1], out[2], out[3]);
    }
}

constexpr size_t S2_SMEM = (size_t)S2_E4 * 16 + sizeof(S2Shared) + 16;

} // namespace ffb
