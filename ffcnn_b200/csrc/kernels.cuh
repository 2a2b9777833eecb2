/*
 * kernels.cuh -- hand-written sm_100a kernels for the HBM-bound half of ffcnn's hot path.
 *
 * Data layout: batched NHWC fp32, `ld` floats between consecutive pixels (ld = ALIGN(c,4); the
 * 255-channel yolo heads therefore sit at ld = 256 so every pixel row stays 16-byte aligned).
 * One image row is a flat run of W*C floats, so the depthwise stencils below treat an image as
 * [H][W*C] and each thread owns one float4 (4 channels of one pixel) of that run: consecutive
 * lanes touch consecutive 16 B -> fully coalesced 128-bit accesses; a tap to the left/right is an
 * offset of -/+C floats, to the row above/below -/+W*C.
 *
 * Every conv kernel ends in the reference's fused epilogue  act(sum * scale + bias)
 * (conv-v0.c:27 / conv-v6.c:38,72-75; activate(): utils.h:15-23); BN is pre-folded into
 * (scale, bias) at load (ffcnn.c:222-233).  sum*scale+bias is one FMA here, as in the
 * reference's shipped -Ofast build.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sm100.cuh"
#include "stem_common.cuh"

namespace ffb {

using sm100::pdl_trigger;
using sm100::pdl_wait;


__device__ __forceinline__ float4 epilogue4(float4 a, float4 s, float4 b, int act)
{
    float4 r;
    r.x = act_apply(fmaf(a.x, s.x, b.x), act); r.y = act_apply(fmaf(a.y, s.y, b.y), act);
    r.z = act_apply(fmaf(a.z, s.z, b.z), act); r.w = act_apply(fmaf(a.w, s.w, b.w), act);
    return r;
}

__device__ __forceinline__ void fma4(float4 &acc, const float4 v, const float4 w)
{
    acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
    acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

/* ------------------------------------------------------------------------------------------------
 * net_input, batched (ffcnn.c:259-289): BGR u8 frames -> NHWC fp32 [n][H][W][ld=4] (channel 3 = 0),
 * nearest-neighbour fit to the top-left sw x sh corner, (px - mean) * norm, zero elsewhere.
 * One thread per output pixel; the float4 store is coalesced, the 3 byte loads hit L1/L2 sectors
 * shared with the neighbouring lanes.
 * ---------------------------------------------------------------------------------------------- */
__global__ void k_input_u8(const uint8_t *__restrict__ frames, float *__restrict__ out,
                           int n, int w, int h, int pitch, int W, int H, int sw, int sh, int s1, int s2,
                           float m0, float m1, float m2, float n0, float n1, float n2)
{
    pdl_trigger(); pdl_wait();
    const long total = (long)n * H * W;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H); const long f = i / ((long)W * H);
        float4 v = zero4();
        if (x < sw && y < sh) {
            const uint8_t *px = frames + f * (long)h * pitch + (long)(y * s1 / s2) * pitch + (x * s1 / s2) * 3;
            v.x = ((float)px[2] - m0) * n0;          /* R */
            v.y = ((float)px[1] - m1) * n1;          /* G */
            v.z = ((float)px[0] - m2) * n2;          /* B */
        }
        reinterpret_cast<float4 *>(out)[i] = v;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Stem: dense 3x3 stride-2 pad-1 conv, 3 -> 8 channels -- the only layer that takes the reference's generic im2row
 * path (conv-v6.c:9-42).  Each CTA stages the (2*TY+1) x (2*TX+1) input halo tile in shared memory as float4
 * (R,G,B,0), then every thread produces one output pixel x 8 channels (two float4 stores).
 * The 216 weights + scale/bias travel as a __grid_constant__ kernel parameter: every lane uses the same weight at the
 * same time, so each FFMA takes it straight from the constant bank -- no shared-memory or register traffic for weights
 * (the first version read them with broadcast LDS and was shared-memory bound at 30 % of the HBM roofline).
 * Accumulation order channel -> ky -> kx as conv-v0.c:16-25.
 *   k_stem_f32: input is the fp32 NHWC tensor (ld 4) written by k_input_u8 / ffb_input_chw.
 *   k_stem_u8 : net_input fused in -- reads the BGR u8 frames directly (no resize: frame size == net size) and applies
 *               (px - mean) * norm while staging, the same float arithmetic as ffcnn.c:281-283, 4x fewer input bytes.
 * ---------------------------------------------------------------------------------------------- */

template <int TX, int TY>
__device__ __forceinline__ void stem_compute(const float4 (*tile)[2 * TX + 1], const StemW &sw, float *__restrict__ out,
                                             long f, int OH, int OW, int ox, int oy, int act)
{
    if (ox >= OW || oy >= OH) return;
    /* output channels as four packed pairs (FFMA2: two fp32 FMAs per issue slot, bit-identical to fmaf; the weight pairs come
       from the constant bank through uniform registers, the pixel value is a broadcast operand) */
    using sm100::f32x2; using sm100::f2_pack; using sm100::f2_fma; using sm100::f2_lo; using sm100::f2_hi;
    f32x2 acc[4];
#pragma unroll
    for (int o = 0; o < 4; o++) acc[o] = 0ull;
    float4 p[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) p[j][k] = tile[2 * threadIdx.y + j][2 * threadIdx.x + k];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float v = c == 0 ? p[j][k].x : c == 1 ? p[j][k].y : p[j][k].z;
                const float *wt = sw.w + ((c * 3 + j) * 3 + k) * 8;
#pragma unroll
                for (int o = 0; o < 4; o++) acc[o] = f2_fma(f2_pack(v, v), f2_pack(wt[2 * o], wt[2 * o + 1]), acc[o]);
            }
    float rr[8];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const f32x2 t = f2_fma(acc[o], f2_pack(sw.s[2 * o], sw.s[2 * o + 1]), f2_pack(sw.b[2 * o], sw.b[2 * o + 1]));
        rr[2 * o] = act_apply(f2_lo(t), act); rr[2 * o + 1] = act_apply(f2_hi(t), act);
    }
    const float4 r0 = make_float4(rr[0], rr[1], rr[2], rr[3]), r1 = make_float4(rr[4], rr[5], rr[6], rr[7]);
    float4 *o = reinterpret_cast<float4 *>(out + (f * (long)OH * OW + (long)oy * OW + ox) * 8);
    o[0] = r0; o[1] = r1;
}

template <int TX, int TY>
__global__ void __launch_bounds__(TX * TY)
k_stem_f32(const float *__restrict__ in, float *__restrict__ out, const __grid_constant__ StemW sw,
           int H, int W, int OH, int OW, int act)
{
    pdl_trigger(); pdl_wait();
    constexpr int IW = 2 * TX + 1, IH = 2 * TY + 1;
    __shared__ float4 tile[IH][IW];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
    const long f = blockIdx.z;
    const float *img = in + f * (long)H * W * 4;
    const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
    for (int i = tid; i < IW * IH; i += TX * TY) {
        const int tx = i % IW, ty = i / IW, ix = ix0 + tx, iy = iy0 + ty;
        tile[ty][tx] = (ix >= 0 && ix < W && iy >= 0 && iy < H) ? ldg4(img + ((long)iy * W + ix) * 4) : zero4();
    }
    __syncthreads();
    stem_compute<TX, TY>(tile, sw, out, f, OH, OW, ox0 + threadIdx.x, oy0 + threadIdx.y, act);
}

template <int TX, int TY>
__global__ void __launch_bounds__(TX * TY)
k_stem_u8(const uint8_t *__restrict__ frames, int pitch, float *__restrict__ out, const __grid_constant__ StemW sw,
          int H, int W, int OH, int OW, int act, float m0, float m1, float m2, float n0, float n1, float n2)
{
    pdl_trigger(); pdl_wait();
    constexpr int IW = 2 * TX + 1, IH = 2 * TY + 1;
    __shared__ float4 tile[IH][IW];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
    const long f = blockIdx.z;
    const uint8_t *img = frames + f * (long)H * pitch;
    const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
    /* all byte loads of this thread are issued before the first is consumed: one DRAM latency per CTA, not one per pixel */
    constexpr int NIT = (IW * IH + TX * TY - 1) / (TX * TY);
    unsigned char bgr[NIT][3]; bool ok[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int i = tid + it * TX * TY, tx = i % IW, ty = i / IW, ix = ix0 + tx, iy = iy0 + ty;
        ok[it] = i < IW * IH && ix >= 0 && ix < W && iy >= 0 && iy < H;
        const uint8_t *px = img + (long)(ok[it] ? iy : 0) * pitch + (ok[it] ? ix : 0) * 3;
        bgr[it][0] = __ldg(px); bgr[it][1] = __ldg(px + 1); bgr[it][2] = __ldg(px + 2);
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int i = tid + it * TX * TY;
        if (i < IW * IH) {
            float4 v = zero4();
            if (ok[it]) {
                v.x = ((float)bgr[it][2] - m0) * n0;         /* R */
                v.y = ((float)bgr[it][1] - m1) * n1;         /* G */
                v.z = ((float)bgr[it][0] - m2) * n2;         /* B */
            }
            tile[i / IW][i % IW] = v;
        }
    }
    __syncthreads();
    stem_compute<TX, TY>(tile, sw, out, f, OH, OW, ox0 + threadIdx.x, oy0 + threadIdx.y, act);
}

/* k_stem_u8x2: the u8 stem with two output pixels per thread and word-wise staging (round 2v).  ncu on k_stem_u8 (r2t): 590
 * instructions per output pixel, issue-bound, 40 % of them the staging -- three byte loads, three I2F and ~10 index / predicate
 * instructions per input pixel, 4.3 input pixels per output.  Here
 *   - a tile row is fetched as aligned 32-bit words (the row of input pixels 2*ox0-1 ... starts one byte after a 4-byte boundary
 *     whenever the tile's first output column is even, which the launcher guarantees together with a 4-byte aligned frame pointer
 *     and pitch); a byte becomes a float by PRMT into the mantissa of 2^23 and one subtraction (exact, = (float)byte, no
 *     conversion pipe), then (f - mean) * norm exactly as ffcnn.c:281-283;
 *   - a thread produces two adjacent output pixels: every weight pair fetched from the constant bank feeds two FFMA2, the 3x5 input
 *     window is 15 shared loads for two pixels instead of 18.
 * Accumulation order per output stays channel -> ky -> kx (conv-v0.c:16-25): results are bit-identical to k_stem_u8. */
template <int TX, int TY>
__global__ void __launch_bounds__(TX * TY, TX * TY <= 160 ? 5 : 2)
k_stem_u8x2(const uint8_t *__restrict__ frames, int pitch, float *__restrict__ out, const __grid_constant__ StemW sw,
            int H, int W, int OH, int OW, int act, float m0, float m1, float m2, float n0, float n1, float n2)
{
    using sm100::f32x2; using sm100::f2_pack; using sm100::f2_fma; using sm100::f2_lo; using sm100::f2_hi;
    pdl_trigger(); pdl_wait();
    constexpr int OWT = 2 * TX, IW = 2 * OWT + 1, IH = 2 * TY + 1, NT = TX * TY;
    constexpr int GPR = (IW - 1) / 4, NG = GPR * IH;                /* 4-pixel groups per tile row (tile pixels 1 .. IW-1), per tile */
    /* tile pixel px of row ty lives at tile[ty][px & 3][px >> 2]: a thread stages / reads pixels 4 apart, so consecutive threads touch
       consecutive float4 of one plane (with a plain [IH][IW] tile every 128-bit access was a 4-way bank conflict and the kernel lost 40 %) */
    __shared__ float4 tile[IH][4][GPR + 1];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int ox0 = blockIdx.x * OWT, oy0 = blockIdx.y * TY;
    const long f = blockIdx.z;
    const uint8_t *img = frames + f * (long)H * pitch;
    const int ix0 = 2 * ox0 - 1, iy0 = 2 * oy0 - 1;
    /* tile pixel 1 = image pixel 2*ox0 starts at byte 6*ox0 of its row: 4-byte aligned (ox0 is a multiple of 2*TX), so tile pixels
       1+4g .. 4+4g are three aligned words [B0 G0 R0 B1][G1 R1 B2 G2][R2 B3 G3 R3]; W % 4 == 0 puts a group wholly inside or outside */
    constexpr int NIT = (NG + NT - 1) / NT;
    uint32_t w0[NIT], w1[NIT], w2[NIT]; bool ok[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {                              /* all loads of this thread are issued before the first is consumed */
        const int g = tid + it * NT, ty = g / GPR, gx = g - ty * GPR, iy = iy0 + ty, ix = ix0 + 1 + 4 * gx;
        ok[it] = g < NG && (unsigned)iy < (unsigned)H && ix + 3 < W;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(img + (long)(ok[it] ? iy : 0) * pitch + 3 * (ok[it] ? ix : 0));
        w0[it] = ok[it] ? __ldg(src) : 0u; w1[it] = ok[it] ? __ldg(src + 1) : 0u; w2[it] = ok[it] ? __ldg(src + 2) : 0u;
    }
    /* a byte becomes a float by PRMT into the mantissa of 2^23 and one subtraction: exact, = (float)byte, no conversion pipe */
    auto cvt = [&](uint32_t word, int k, float mean, float norm) {
        return (__uint_as_float(__byte_perm(word, 0x4b000000u, 0x7540 + k)) - 8388608.0f - mean) * norm;
    };
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int g = tid + it * NT, ty = g / GPR, gx = g - ty * GPR;
        if (g < NG) {
            if (ok[it]) {
                tile[ty][1][gx] = make_float4(cvt(w0[it], 2, m0, n0), cvt(w0[it], 1, m1, n1), cvt(w0[it], 0, m2, n2), 0.f);
                tile[ty][2][gx] = make_float4(cvt(w1[it], 1, m0, n0), cvt(w1[it], 0, m1, n1), cvt(w0[it], 3, m2, n2), 0.f);
                tile[ty][3][gx] = make_float4(cvt(w2[it], 0, m0, n0), cvt(w1[it], 3, m1, n1), cvt(w1[it], 2, m2, n2), 0.f);
                tile[ty][0][gx + 1] = make_float4(cvt(w2[it], 3, m0, n0), cvt(w2[it], 2, m1, n1), cvt(w2[it], 1, m2, n2), 0.f);
            } else {
                tile[ty][1][gx] = tile[ty][2][gx] = tile[ty][3][gx] = tile[ty][0][gx + 1] = zero4();
            }
        }
    }
    if (tid < IH) {                                                 /* tile pixel 0 = image pixel 2*ox0 - 1, the left halo column */
        const int iy = iy0 + tid;
        float4 v = zero4();
        if ((unsigned)iy < (unsigned)H && ix0 >= 0) {
            const uint8_t *px = img + (long)iy * pitch + 3 * ix0;
            v.x = ((float)__ldg(px + 2) - m0) * n0; v.y = ((float)__ldg(px + 1) - m1) * n1; v.z = ((float)__ldg(px) - m2) * n2;
        }
        tile[tid][0][0] = v;
    }
    __syncthreads();
    const int ox = ox0 + 2 * threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox >= OW || oy >= OH) return;
    f32x2 acc[2][4];
#pragma unroll
    for (int o = 0; o < 4; o++) { acc[0][o] = 0ull; acc[1][o] = 0ull; }
    float4 p[3][5];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 5; k++) p[j][k] = tile[2 * threadIdx.y + j][k & 3][threadIdx.x + (k >> 2)];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float v0 = c == 0 ? p[j][k].x : c == 1 ? p[j][k].y : p[j][k].z;
                const float v1 = c == 0 ? p[j][k + 2].x : c == 1 ? p[j][k + 2].y : p[j][k + 2].z;
                const float *wt = sw.w + ((c * 3 + j) * 3 + k) * 8;
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    const f32x2 wp = f2_pack(wt[2 * o], wt[2 * o + 1]);
                    acc[0][o] = f2_fma(f2_pack(v0, v0), wp, acc[0][o]);
                    acc[1][o] = f2_fma(f2_pack(v1, v1), wp, acc[1][o]);
                }
            }
    float4 *dst = reinterpret_cast<float4 *>(out + (f * (long)OH * OW + (long)oy * OW + ox) * 8);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (ox + q < OW) {
            float rr[8];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const f32x2 t = f2_fma(acc[q][o], f2_pack(sw.s[2 * o], sw.s[2 * o + 1]), f2_pack(sw.b[2 * o], sw.b[2 * o + 1]));
                rr[2 * o] = act_apply(f2_lo(t), act); rr[2 * o + 1] = act_apply(f2_hi(t), act);
            }
            dst[2 * q] = make_float4(rr[0], rr[1], rr[2], rr[3]); dst[2 * q + 1] = make_float4(rr[4], rr[5], rr[6], rr[7]);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Depthwise FSxFS, stride 1, pad FS/2 (conv-v6.c:96-229 for 3x3, 291-465 for 5x5).
 * Thread = one float4 of the flattened row (pixel x, channels c..c+3); it walks R output rows downwards
 * keeping the last FS input rows in registers (slot = row mod FS, resolved at compile time by unrolling
 * the row loop FS-fold), so each input row is loaded once per thread: FS float4 loads per output float4.
 * wt is [FS*FS][C] (tap-major) so the 4 channel weights of a tap are one float4.
 * skip_row0_at: output row index whose kernel row 0 is ignored (-1 = never): conv-v6.c:422-441 forgets
 * kernel row 0 on output row oh-2 of its 5x5 path; default builds reproduce that (the named oracle).
 * ---------------------------------------------------------------------------------------------- */
template <int FS>
__global__ void __launch_bounds__(128)
k_dw_s1(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ wt,
        const float *__restrict__ scale, const float *__restrict__ bias,
        int H, int W, int C, int R, int act, int skip_row0_at)
{
    pdl_trigger(); pdl_wait();
    constexpr int P = FS / 2;
    const int rowlen = W * C;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) * 4;           /* float offset inside a row */
    if (q >= rowlen) return;
    const int c = q % C, x = q / C;
    const long f = blockIdx.z;
    const float *img = in + f * (long)H * rowlen;
    float *dst = out + f * (long)H * rowlen;
    const int y0 = blockIdx.y * R, y1 = min(H, y0 + R);

    float4 wv[FS * FS];
#pragma unroll
    for (int t = 0; t < FS * FS; t++) wv[t] = ldg4(wt + t * C + c);
    const float4 sc = ldg4(scale + c), bi = ldg4(bias + c);
    bool okx[FS];
#pragma unroll
    for (int k = 0; k < FS; k++) okx[k] = (unsigned)(x + k - P) < (unsigned)W;

    float4 win[FS][FS];                                                   /* [slot][kx] */
    auto load_row = [&](int slot, int iy) {
        const bool oky = (unsigned)iy < (unsigned)H;
        const float *rp = img + (long)iy * rowlen + q;
#pragma unroll
        for (int k = 0; k < FS; k++) win[slot][k] = (oky && okx[k]) ? ldg4(rp + (k - P) * C) : zero4();
    };
    /* slot(row) = (row - (y0 - P)) mod FS */
#pragma unroll
    for (int j = 0; j < FS - 1; j++) load_row(j, y0 - P + j);

    for (int yb = y0; yb < y1; yb += FS) {
#pragma unroll
        for (int u = 0; u < FS; u++) {
            const int y = yb + u;
            if (y < y1) {
                load_row((u + FS - 1) % FS, y + P);
                float4 acc = zero4();
                const bool skip0 = (y == skip_row0_at);
#pragma unroll
                for (int j = 0; j < FS; j++) {
                    if (j == 0 && skip0) continue;
#pragma unroll
                    for (int k = 0; k < FS; k++) fma4(acc, win[(u + j) % FS][k], wv[j * FS + k]);
                }
                *reinterpret_cast<float4 *>(dst + (long)y * rowlen + q) = epilogue4(acc, sc, bi, act);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Depthwise 3x3, stride 2, pad 1 (conv-v6.c:233-287).  Thread = one output float4 (ox, c..c+3), walking
 * R output rows; input row 2*oy+1 of one step is row 2*(oy+1)-1 of the next, so it is carried in
 * registers: 6 float4 loads per output float4 for 9 taps.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128)
k_dw3_s2(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ wt,
         const float *__restrict__ scale, const float *__restrict__ bias,
         int H, int W, int C, int OH, int OW, int R, int act)
{
    pdl_trigger(); pdl_wait();
    const int orow = OW * C, irow = W * C;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q >= orow) return;
    const int c = q % C, ox = q / C;
    const long f = blockIdx.z;
    const float *img = in + f * (long)H * irow;
    float *dst = out + f * (long)OH * orow;
    const int y0 = blockIdx.y * R, y1 = min(OH, y0 + R);
    float4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; t++) wv[t] = ldg4(wt + t * C + c);
    const float4 sc = ldg4(scale + c), bi = ldg4(bias + c);
    const int ix = 2 * ox - 1;
    const bool okl = ix >= 0, okr = ix + 2 < W;
    auto load_row = [&](float4 (&r)[3], int iy) {
        const bool oky = (unsigned)iy < (unsigned)H;
        const float *rp = img + (long)iy * irow + (long)ix * C + c;
        r[0] = (oky && okl) ? ldg4(rp) : zero4();
        r[1] = oky ? ldg4(rp + C) : zero4();
        r[2] = (oky && okr) ? ldg4(rp + 2 * C) : zero4();
    };
    float4 top[3], mid[3], bot[3];
    load_row(top, 2 * y0 - 1);
    for (int oy = y0; oy < y1; oy++) {
        load_row(mid, 2 * oy);
        load_row(bot, 2 * oy + 1);
        float4 acc = zero4();
#pragma unroll
        for (int k = 0; k < 3; k++) fma4(acc, top[k], wv[k]);
#pragma unroll
        for (int k = 0; k < 3; k++) fma4(acc, mid[k], wv[3 + k]);
#pragma unroll
        for (int k = 0; k < 3; k++) fma4(acc, bot[k], wv[6 + k]);
        *reinterpret_cast<float4 *>(dst + (long)oy * orow + q) = epilogue4(acc, sc, bi, act);
#pragma unroll
        for (int k = 0; k < 3; k++) top[k] = bot[k];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Generic grouped convolution, any geometry (conv-v0.c:7-31 semantics): one thread per output element.
 * The slow-but-always-right path behind the groupconv seam for shapes the specialised kernels do not
 * cover.  flt = packed reference rows (row floats each, scale/bias at row-4/row-3).
 * ---------------------------------------------------------------------------------------------- */
__global__ void k_conv_generic(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ flt,
                               int n, int H, int W, int C, int ldi, int OH, int OW, int OC, int ldo,
                               int groups, int pad, int stride, int fs, int row, int act, int skip_row0_at)
{
    pdl_trigger(); pdl_wait();
    const int cpg = C / groups, opg = OC / groups;
    const long total = (long)n * OH * OW * OC;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int o = (int)(i % OC); long p = i / OC;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH); const long f = p / OH;
        const int g = o / opg;
        const float *w = flt + (long)o * row;
        const float *img = in + f * (long)H * W * ldi + g * cpg;
        float sum = 0.f;
        for (int c = 0; c < cpg; c++)
            for (int j = (oy == skip_row0_at ? 1 : 0); j < fs; j++) {
                const int iy = oy * stride - pad + j;
                if ((unsigned)iy >= (unsigned)H) continue;
                for (int k = 0; k < fs; k++) {
                    const int ix = ox * stride - pad + k;
                    if ((unsigned)ix >= (unsigned)W) continue;
                    sum = fmaf(__ldg(img + ((long)iy * W + ix) * ldi + c), __ldg(w + (c * fs + j) * fs + k), sum);
                }
            }
        out[(f * (long)OH * OW + (long)oy * OW + ox) * ldo + o] = act_apply(fmaf(sum, w[row - 4], w[row - 3]), act);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Pool with the reference's clamped window (ffcnn.c:337-372,381-394): window [x-(fs-1)/2, +fs) cut to
 * the image; max, or sum / fs^2.  One thread per output float4.
 * ---------------------------------------------------------------------------------------------- */
__global__ void k_pool(const float *__restrict__ in, float *__restrict__ out, int n, int H, int W, int C, int ldi,
                       int OH, int OW, int ldo, int coff, int fs, int stride, int is_max)
{
    pdl_trigger(); pdl_wait();
    const int c4n = C / 4;
    const long total = (long)n * OH * OW * c4n;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4; long p = i / c4n;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH); const long f = p / OH;
        const int xa = max(ox * stride - (fs - 1) / 2, 0), xb = min(ox * stride - (fs - 1) / 2 + fs, W);
        const int ya = max(oy * stride - (fs - 1) / 2, 0), yb = min(oy * stride - (fs - 1) / 2 + fs, H);
        const float *img = in + f * (long)H * W * ldi + c;
        float4 m = is_max ? ldg4(img + ((long)ya * W + xa) * ldi) : zero4();
        for (int y = ya; y < yb; y++)
            for (int x = xa; x < xb; x++) {
                const float4 v = ldg4(img + ((long)y * W + x) * ldi);
                if (is_max) { m.x = m.x < v.x ? v.x : m.x; m.y = m.y < v.y ? v.y : m.y; m.z = m.z < v.z ? v.z : m.z; m.w = m.w < v.w ? v.w : m.w; }
                else        { m.x += v.x; m.y += v.y; m.z += v.z; m.w += v.w; }
            }
        if (!is_max) { const float d = (float)(fs * fs); m.x /= d; m.y /= d; m.z /= d; m.w /= d; }
        *reinterpret_cast<float4 *>(out + (f * (long)OH * OW + (long)oy * OW + ox) * ldo + coff + c) = m;
    }
}

/* ------------------------------------------------------------------------------------------------
 * SPP block in one pass: three stride-1 max pools of the SAME tensor (radii r1 < r2 < r3, reference windows
 * [x-(fs-1)/2, +fs) cut to the image, ffcnn.c:354-394) plus the route that concatenates them with the tensor itself
 * (ffcnn.c:425-434) -- layers L109-L114 of yolo-fastest-1.1.  One thread per (pixel, 4 channels): the (2*r3+1)^2 window
 * is read once (L1-resident: a 10x10x48 frame is 19 KB) and feeds the three nested maxima; the four results go straight
 * to their channel offsets of the concat tensor.  max is exact, so nesting the windows changes nothing.
 * ---------------------------------------------------------------------------------------------- */
__global__ void k_spp(const float *__restrict__ in, float *__restrict__ out, int n, int H, int W, int C, int ldi, int ldo,
                      int r1, int r2, int r3, int off1, int off2, int off3, int offx)
{
    pdl_trigger(); pdl_wait();
    const int c4n = C / 4;
    const long total = (long)n * H * W * c4n;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4; long p = i / c4n;
        const int x = (int)(p % W); p /= W;
        const int y = (int)(p % H); const long f = p / H;
        const float *img = in + f * (long)H * W * ldi + c;
        const float4 ctr = ldg4(img + ((long)y * W + x) * ldi);
        float4 m1 = ctr, m2 = ctr, m3 = ctr;
        const int ya = max(y - r3, 0), yb = min(y + r3, H - 1), xa = max(x - r3, 0), xb = min(x + r3, W - 1);
        for (int yy = ya; yy <= yb; yy++) {
            const int ady = abs(yy - y);
            for (int xx = xa; xx <= xb; xx++) {
                const float4 v = ldg4(img + ((long)yy * W + xx) * ldi);
                const int d = max(ady, abs(xx - x));
                m3.x = fmaxf(m3.x, v.x); m3.y = fmaxf(m3.y, v.y); m3.z = fmaxf(m3.z, v.z); m3.w = fmaxf(m3.w, v.w);
                if (d <= r2) { m2.x = fmaxf(m2.x, v.x); m2.y = fmaxf(m2.y, v.y); m2.z = fmaxf(m2.z, v.z); m2.w = fmaxf(m2.w, v.w); }
                if (d <= r1) { m1.x = fmaxf(m1.x, v.x); m1.y = fmaxf(m1.y, v.y); m1.z = fmaxf(m1.z, v.z); m1.w = fmaxf(m1.w, v.w); }
            }
        }
        float *o = out + (f * (long)H * W + (long)y * W + x) * ldo + c;
        *reinterpret_cast<float4 *>(o + off1) = m1; *reinterpret_cast<float4 *>(o + off2) = m2;
        *reinterpret_cast<float4 *>(o + off3) = m3; *reinterpret_cast<float4 *>(o + offx) = ctr;
    }
}

/* Same operation, one CTA per frame, separable: the frame is staged in shared memory, a horizontal pass leaves the three
 * row maxima (radii r1 < r2 < r3), a vertical pass finishes them -- 9 + 17 shared-memory reads per (pixel, 4 channels)
 * instead of 81 global ones.  Needs 4 * H*W*C floats of shared memory (77 KB for the 10x10x48 SPP input). */
__global__ void __launch_bounds__(256) k_spp_smem(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, int ldi, int ldo,
                                                  int r1, int r2, int r3, int off1, int off2, int off3, int offx)
{
    extern __shared__ float4 spp_smem[];
    pdl_trigger(); pdl_wait();
    const int c4n = C / 4, items = H * W * c4n;
    float4 *sIn = spp_smem, *sH1 = sIn + items, *sH2 = sH1 + items, *sH3 = sH2 + items;
    const float *img = in + (long)blockIdx.x * H * W * ldi;
    float *o = out + (long)blockIdx.x * H * W * ldo;
    auto mx = [](float4 a, const float4 b) { a.x = fmaxf(a.x, b.x); a.y = fmaxf(a.y, b.y); a.z = fmaxf(a.z, b.z); a.w = fmaxf(a.w, b.w); return a; };
    for (int i = threadIdx.x; i < items; i += blockDim.x) { const int c = i % c4n, p = i / c4n; sIn[i] = ldg4(img + (long)p * ldi + 4 * c); }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = i % c4n, p = i / c4n, x = p % W, rowbase = (p - x) * c4n + c;
        float4 m = sIn[i];
        for (int d = 1; d <= r3; d++) {
            if (x - d >= 0) m = mx(m, sIn[rowbase + (x - d) * c4n]);
            if (x + d < W)  m = mx(m, sIn[rowbase + (x + d) * c4n]);
            if (d == r1) sH1[i] = m;
            if (d == r2) sH2[i] = m;
        }
        sH3[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int c = i % c4n, p = i / c4n, x = p % W, y = p / W, colbase = x * c4n + c, rs = W * c4n;
        float4 m1 = sH1[i], m2 = sH2[i], m3 = sH3[i];
        for (int d = 1; d <= r3; d++) {
            if (y - d >= 0) { const int j = colbase + (y - d) * rs; m3 = mx(m3, sH3[j]); if (d <= r2) m2 = mx(m2, sH2[j]); if (d <= r1) m1 = mx(m1, sH1[j]); }
            if (y + d < H)  { const int j = colbase + (y + d) * rs; m3 = mx(m3, sH3[j]); if (d <= r2) m2 = mx(m2, sH2[j]); if (d <= r1) m1 = mx(m1, sH1[j]); }
        }
        float *op = o + (long)p * ldo + 4 * c;
        *reinterpret_cast<float4 *>(op + off1) = m1; *reinterpret_cast<float4 *>(op + off2) = m2;
        *reinterpret_cast<float4 *>(op + off3) = m3; *reinterpret_cast<float4 *>(op + offx) = sIn[i];
    }
}

/* nearest upsample (ffcnn.c:396-410): out[y][x] = in[y/s][x/s].  A CTA takes input rows: every input float4 is read once
 * (coalesced) and written to its s x s output pixels; 32-bit index arithmetic only (the first version decoded a flat 64-bit
 * index per output float4 -- three 64-bit divisions each -- and ran at 36 % of the HBM roofline, bound by the divisions). */
__global__ void k_upsample(const float *__restrict__ in, float *__restrict__ out, int n, int H, int W, int C, int ldi,
                           int ldo, int coff, int s)
{
    pdl_trigger(); pdl_wait();
    const int c4n = C / 4, OH = H * s, OW = W * s, row_items = W * c4n;
    for (int r = blockIdx.x; r < n * H; r += gridDim.x) {
        const int f = r / H, y = r - f * H;
        for (int i = threadIdx.x; i < row_items; i += blockDim.x) {
            const int x = i / c4n, c = (i - x * c4n) * 4;
            const float4 v = ldg4(in + ((long)r * W + x) * ldi + c);
            float *o = out + (((long)f * OH + (long)y * s) * OW + (long)x * s) * ldo + coff + c;
            for (int dy = 0; dy < s; dy++)
                for (int dx = 0; dx < s; dx++) *reinterpret_cast<float4 *>(o + ((long)dy * OW + dx) * ldo) = v;
        }
    }
}

/* shortcut (ffcnn.c:418-423): out = act(a + b), both operands dense (ld == c), flat float4 stream */
__global__ void k_shortcut(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long n4, int act)
{
    pdl_trigger(); pdl_wait();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 x = ldg4(a + 4 * i), y = ldg4(b + 4 * i);
        float4 r;
        r.x = act_apply(x.x + y.x, act); r.y = act_apply(x.y + y.y, act);
        r.z = act_apply(x.z + y.z, act); r.w = act_apply(x.w + y.w, act);
        reinterpret_cast<float4 *>(out)[i] = r;
    }
}

/* route (ffcnn.c:425-434): copy one source into channel range [coff, coff+C) of the concat tensor */
__global__ void k_concat(const float *__restrict__ in, float *__restrict__ out, long pixels, int C, int ldi, int ldo, int coff)
{
    pdl_trigger(); pdl_wait();
    const int c4n = C / 4;
    const long total = pixels * c4n;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4; const long p = i / c4n;
        *reinterpret_cast<float4 *>(out + p * ldo + coff + c) = ldg4(in + p * ldi + c);
    }
}

/* scalar variants for channel counts that are not a multiple of 4 (only reachable through odd cfgs) */
__global__ void k_copy_strided(const float *__restrict__ in, float *__restrict__ out, long pixels, int C, int ldi, int ldo, int coff)
{
    pdl_trigger(); pdl_wait();
    const long total = pixels * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const long p = i / C;
        out[p * ldo + coff + c] = in[p * ldi + c];
    }
}

/* layout converters for the CHW boundary (groupconv seam, ffb_input_chw): [n][c][h][w] <-> [n][h][w][ld] */
/* pool / upsample for channel counts that are not a multiple of 4: one thread per output float */
__global__ void k_pool_scalar(const float *__restrict__ in, float *__restrict__ out, int n, int H, int W, int C, int ldi,
                              int OH, int OW, int ldo, int coff, int fs, int stride, int is_max)
{
    pdl_trigger(); pdl_wait();
    const long total = (long)n * OH * OW * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long p = i / C;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH); const long f = p / OH;
        const int xa = max(ox * stride - (fs - 1) / 2, 0), xb = min(ox * stride - (fs - 1) / 2 + fs, W);
        const int ya = max(oy * stride - (fs - 1) / 2, 0), yb = min(oy * stride - (fs - 1) / 2 + fs, H);
        const float *img = in + f * (long)H * W * ldi + c;
        float m = is_max ? __ldg(img + ((long)ya * W + xa) * ldi) : 0.f;
        for (int y = ya; y < yb; y++)
            for (int x = xa; x < xb; x++) {
                const float v = __ldg(img + ((long)y * W + x) * ldi);
                if (is_max) m = m < v ? v : m; else m += v;
            }
        if (!is_max) m /= (float)(fs * fs);
        out[(f * (long)OH * OW + (long)oy * OW + ox) * ldo + coff + c] = m;
    }
}

__global__ void k_upsample_scalar(const float *__restrict__ in, float *__restrict__ out, int n, int H, int W, int C, int ldi,
                                  int ldo, int coff, int s)
{
    pdl_trigger(); pdl_wait();
    const int OH = H * s, OW = W * s;
    const long total = (long)n * OH * OW * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long p = i / C;
        const int ox = (int)(p % OW); p /= OW;
        const int oy = (int)(p % OH); const long f = p / OH;
        out[(f * (long)OH * OW + (long)oy * OW + ox) * ldo + coff + c] = __ldg(in + (f * (long)H * W + (long)(oy / s) * W + ox / s) * ldi + c);
    }
}

__global__ void k_chw_to_nhwc(const float *__restrict__ src, float *__restrict__ dst, int n, int C, int H, int W, int ld)
{
    const long total = (long)n * H * W * ld;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ld); long p = i / ld;
        const int x = (int)(p % W); p /= W;
        const int y = (int)(p % H); const long f = p / H;
        dst[i] = c < C ? src[((f * C + c) * H + y) * (long)W + x] : 0.f;
    }
}

__global__ void k_nhwc_to_chw(const float *__restrict__ src, float *__restrict__ dst, int n, int C, int H, int W, int ld)
{
    const long total = (long)n * C * H * W;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W); long p = i / W;
        const int y = (int)(p % H); p /= H;
        const int c = (int)(p % C); const long f = p / C;
        dst[i] = src[((f * H + y) * (long)W + x) * ld + c];
    }
}

/* weights: packed reference rows -> tap-major [taps][fn_pad] (+ scale[fn_pad], bias[fn_pad], zero padded) */
__global__ void k_prep_weights(const float *__restrict__ flt, int row, int fn, int taps, int fn_pad,
                               float *__restrict__ wt, float *__restrict__ scale, float *__restrict__ bias)
{
    const int total = taps * fn_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + fn_pad; i += gridDim.x * blockDim.x) {
        if (i < total) {
            const int t = i / fn_pad, o = i % fn_pad;
            wt[i] = o < fn ? flt[(long)o * row + t] : 0.f;
        } else {
            const int o = i - total;
            scale[o] = o < fn ? flt[(long)o * row + row - 4] : 0.f;
            bias[o]  = o < fn ? flt[(long)o * row + row - 3] : 0.f;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * yolo candidate filter (first half of ffcnn.c:438-452 on the GPU).  Each anchor's first arg-max over the
 * class logits is found with a shuffle reduction (ties -> lowest class index, like the sequential
 * `cs < val` scan), and lane 0 tests a float estimate of the reference confidence against thresh - margin.
 * Survivors are appended (atomic counter) as raw logits; the host re-does the exact double-precision decode.
 * ---------------------------------------------------------------------------------------------- */
struct Candidate { int frame, key, cls; float bs, cs, tx, ty, tw, th; };

/* Two levels.  (1) One THREAD per grid cell reads only the three objectness logits: the reference confidence
 * 1 / (1 + e^-bs (1 + e^-cs)) can never exceed sigmoid(bs), so an anchor whose sigmoid(bs) is below thresh - margin is out
 * without its 80 class logits ever being read (almost every anchor of almost every cell: the heads are then read at ~1/10
 * of their size).  (2) The survivors of the 32 cells of a warp are taken one by one by the WHOLE warp: coalesced read of
 * the class logits, shuffle arg-max, exact-form float confidence test by lane 0, append. */
__global__ void k_yolo_filter(const float *__restrict__ head, int n, int cells, int ld, int classes, int head_index,
                              int key_base, float thresh, Candidate *__restrict__ list, int *__restrict__ counter, int cap,
                              int *host_count = nullptr)
{
    pdl_trigger(); pdl_wait();
    const int lane = threadIdx.x & 31;
    const long cell_id = blockIdx.x * (long)blockDim.x + threadIdx.x;         /* frame * cells + cell */
    const long total = (long)n * cells;
    const int per = 5 + classes;
    unsigned pass = 0;
    if (cell_id < total) {
        const float *v = head + cell_id * ld;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float bs = __ldg(v + a * per + 4);
            const float ub = 1.0f / (1.0f + expf(-bs));                        /* upper bound of the confidence */
            if (ub >= thresh - 1e-3f || !(ub == ub)) pass |= 1u << a;
        }
    }
    for (int a = 0; a < 3; a++) {
        unsigned todo = __ballot_sync(0xffffffffu, (pass >> a) & 1u);
        while (todo) {
            const int src = __ffs(todo) - 1; todo &= todo - 1;
            const long cid = __shfl_sync(0xffffffffu, cell_id, src);
            const float *pv = head + cid * ld + a * per;
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int l = lane; l < classes; l += 32) {
                const float s = __ldg(pv + 5 + l);
                if (s > best) { best = s; bi = l; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, d);
                const int   oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                const float bs = __ldg(pv + 4);
                const float conf = 1.0f / (1.0f + expf(-bs) * (1.0f + expf(-best)));
                if (conf >= thresh - 1e-3f || !(conf == conf)) {
                    const int slot = atomicAdd(counter, 1);
                    if (slot < cap) {
                        Candidate c;
                        c.frame = (int)(cid / cells); c.key = key_base + (int)(cid % cells) * 3 + a; c.cls = bi == 0x7fffffff ? 0 : bi;
                        c.bs = bs; c.cs = best; c.tx = __ldg(pv); c.ty = __ldg(pv + 1); c.tw = __ldg(pv + 2); c.th = __ldg(pv + 3);
                        list[slot] = c;
                    }
                }
            }
        }
    }
    (void)head_index;
    /* The launch for the LAST head also delivers the candidate count to the host: its last block to finish (ticket in counter[1])
       writes it into pinned host memory.  A separate 4-byte device-to-host copy in the stream put a copy-engine hand-over (and its
       semaphore round trip) between every two batches. */
    if (host_count) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(counter + 1, 1) == (int)gridDim.x - 1) {
                __threadfence();
                *reinterpret_cast<volatile int *>(host_count) = *reinterpret_cast<volatile int *>(counter);
                __threadfence_system();
                counter[0] = 0; counter[1] = 0;            /* ready for the next batch that uses this detection set: no memset in the stream either */
            }
        }
    }
}

__global__ void k_fill(float *p, long n, float v)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

} // namespace ffb
