/*
 * stem_block.cuh -- the stem and the first inverted-residual block of yolo-fastest-1.1 as ONE kernel:
 *
 *     u8 frame --net_input (ffcnn.c:259-289)--> 3x3 s2 conv 3->8, BN, leaky (L0, conv-v6.c:9-42) --> x [160x160x8]
 *       --1x1 8->8, BN, leaky (L1)--> e --3x3 depthwise, BN, leaky (L2)--> d --1x1 8->4, BN, linear (L3)--> y [160x160x4]
 *
 * Unfused, x (210 MB per batch of 256) is written by the stem kernel and read back by the block kernel; here it never
 * exists: a CTA owns a TXO x TYO tile of y, stages the u8 frame tile it needs (k_stem_u8x2's word-wise staging into four
 * pixel planes), computes the stem output at the (TXO+2) x (TYO+2) positions of the tile + the depthwise halo, expands it
 * on the spot and keeps only e in shared memory (zero at positions outside the image: the depthwise conv's padding); after
 * one barrier every thread finishes two adjacent output pixels (9 taps, BN, act, projection, BN).
 *
 * Every arithmetic step is the one k_stem_u8x2 and k_block_reg_s1<8, 8, 4> perform, in the same order (stem: channel -> ky
 * -> kx; expand: input channel ascending, first term a product; depthwise: ky -> kx, first tap a product; projection:
 * expanded channel ascending from zero), as packed fp32 pairs over adjacent channels -- the result is BIT-IDENTICAL to the
 * two-kernel path (tests/test_gpu_parity.py::test_stem_block_fusion_is_bit_identical).
 */
#pragma once
#include "stem_common.cuh"
#include "block_reg.cuh"

namespace ffb {

constexpr int SB_TXO = 32, SB_TYO = 8;                  /* output tile.  Sweep (B200, batch 256, ms for L0-L3): 32x8 at five CTAs per SM 0.199 | 32x16 (3) 0.205 |
                                                           32x12 (4) 0.208 | 32x20 (2) 0.229; the two-kernel path 0.224 */
constexpr int SB_THREADS = 192;                         /* 170 position pairs in phase B, 128 output-pixel pairs in phase C */
/* Shared-memory geometry (float4 units).  Bank conflicts decide this kernel (l1tex is its busiest unit), so both arrays are skewed:
 *  - staged row r starts at r * 4 * IWG + SB_TILE_SKEW(r): phase B's threads walk 17 position pairs per row of positions, a quarter-warp
 *    that straddles two position rows (= two staged rows apart) needs those rows 1 float4 apart mod 8 to stay conflict-free, phase A's
 *    stores (18 groups per staged row) want consecutive rows 2 apart: rows alternate +2 / +7 (sum 9 = 1 mod 8);
 *  - a row of e is padded from 68 to 73 float4 (1 mod 8) for phase B's stores; phase C reads whole quarter-warps inside one row. */
constexpr int SB_IWG = (2 * SB_TXO + 6 + 3) / 4, SB_IH = 2 * SB_TYO + 5, SB_HXN = SB_TXO + 2, SB_HYN = SB_TYO + 2;
__host__ __device__ constexpr int SB_TILE_SKEW(int r) { return (r >> 1) * 9 + (r & 1) * 2; }
constexpr int SB_TILE_F4 = SB_IH * 4 * SB_IWG + SB_TILE_SKEW(SB_IH);
constexpr int SB_E_ROW = ((4 * (SB_HXN / 2) + 7) / 8) * 8 + 1;
constexpr size_t SB_SMEM = ((size_t)SB_TILE_F4 + (size_t)SB_HYN * SB_E_ROW) * 16;

struct StemBlockArgs {
    const uint8_t *frames; int pitch;
    float *y;
    int H, W, OH, OW;                                   /* frame size, size of the stem output / block tensors */
    int act0; float m0, m1, m2, n0, n1, n2;             /* stem activation, net_input mean / norm (R, G, B) */
    float slope1, sloped, slope3;
};

#ifndef SB_MINB
#define SB_MINB 4      /* 80 registers, no spills: 0.1763 ms against 0.1808 at five CTAs per SM of 64 registers (28 B of spills, a tenth of the stream MOVs) */
#endif
__global__ void __launch_bounds__(SB_THREADS, SB_MINB)
k_stem_block(const __grid_constant__ StemW sw, const __grid_constant__ RegBlockW<8, 8, 4> w, const StemBlockArgs a)
{
    sm100::pdl_trigger(); sm100::pdl_wait();
    constexpr int HXN = SB_TXO + 2, HYN = SB_TYO + 2;   /* positions of x / e a tile needs */
    constexpr int IWG = (2 * SB_TXO + 6 + 3) / 4;       /* 4-pixel groups per staged row: image pixels 2*ox0-4 ... (aligned start) */
    constexpr int IH = 2 * SB_TYO + 5;                  /* staged rows: image rows 2*oy0-3 ... 2*oy0+2*TYO+1 */
    constexpr int NG = IWG * IH;
    /* staged pixel p of row r (image pixel 2*ox0-4+p) lives at tile[r][p & 3][p >> 2]; e of position (hy, hx), channel half h,
       lives at E[hy][h][hx & 1][hx >> 1]: the threads of both phases walk pixels 4 (positions 2) apart, so consecutive threads
       touch consecutive float4 of one plane */
    extern __shared__ float4 sb_smem[];
    static_assert(IWG == SB_IWG && IH == SB_IH && HXN == SB_HXN && HYN == SB_HYN, "geometry");
    auto tile = [&](int r, int plane, int gx) -> float4 & { return sb_smem[r * (4 * IWG) + SB_TILE_SKEW(r) + plane * IWG + gx]; };
    float4 *Eb = sb_smem + SB_TILE_F4;
    auto E = [&](int hy, int h, int q, int i) -> float4 & { return Eb[hy * SB_E_ROW + (h * 2 + q) * (HXN / 2) + i]; };
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * SB_TXO, oy0 = blockIdx.y * SB_TYO;
    const long f = blockIdx.z;
    const uint8_t *img = a.frames + f * (long)a.H * a.pitch;
    const int ixs = 2 * ox0 - 4, iys = 2 * oy0 - 3;     /* image coordinates of staged pixel (0, 0) */

    /* ---------------- phase A: u8 frame tile -> fp32 planes (k_stem_u8x2's staging) ---------------- */
    constexpr int NIT = (NG + SB_THREADS - 1) / SB_THREADS;
    uint32_t w0[NIT], w1[NIT], w2[NIT]; bool ok[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int g = tid + it * SB_THREADS, ty = g / IWG, gx = g - ty * IWG, iy = iys + ty, ix = ixs + 4 * gx;
        ok[it] = g < NG && (unsigned)iy < (unsigned)a.H && ix >= 0 && ix + 3 < a.W;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(img + (long)(ok[it] ? iy : 0) * a.pitch + 3 * (ok[it] ? ix : 0));
        w0[it] = ok[it] ? __ldg(src) : 0u; w1[it] = ok[it] ? __ldg(src + 1) : 0u; w2[it] = ok[it] ? __ldg(src + 2) : 0u;
    }
    auto cvt = [&](uint32_t word, int k, float mean, float norm) {
        return (__uint_as_float(__byte_perm(word, 0x4b000000u, 0x7540 + k)) - 8388608.0f - mean) * norm;
    };
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int g = tid + it * SB_THREADS, ty = g / IWG, gx = g - ty * IWG;
        if (g < NG) {
            if (ok[it]) {
                tile(ty, 0, gx) = make_float4(cvt(w0[it], 2, a.m0, a.n0), cvt(w0[it], 1, a.m1, a.n1), cvt(w0[it], 0, a.m2, a.n2), 0.f);
                tile(ty, 1, gx) = make_float4(cvt(w1[it], 1, a.m0, a.n0), cvt(w1[it], 0, a.m1, a.n1), cvt(w0[it], 3, a.m2, a.n2), 0.f);
                tile(ty, 2, gx) = make_float4(cvt(w2[it], 0, a.m0, a.n0), cvt(w1[it], 3, a.m1, a.n1), cvt(w1[it], 2, a.m2, a.n2), 0.f);
                tile(ty, 3, gx) = make_float4(cvt(w2[it], 3, a.m0, a.n0), cvt(w2[it], 2, a.m1, a.n1), cvt(w2[it], 1, a.m2, a.n2), 0.f);
            } else {
                tile(ty, 0, gx) = tile(ty, 1, gx) = tile(ty, 2, gx) = tile(ty, 3, gx) = zero4();
            }
        }
    }
    __syncthreads();

    /* ---------------- phase B: stem conv + expand at two adjacent positions (hy, 2i), (hy, 2i+1) -> e ---------------- */
    if (tid < HYN * (HXN / 2)) {
        const int hy = tid / (HXN / 2), i = tid - hy * (HXN / 2);
        /* position (hy, hx) = stem output (oy0-1+hy, ox0-1+hx); its taps are staged rows 2*hy + j, staged pixels 1 + 2*hx + k */
        f32x2 acc[2][4];
#pragma unroll
        for (int o = 0; o < 4; o++) { acc[0][o] = 0ull; acc[1][o] = 0ull; }
        float4 p[3][5];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 5; k++) p[j][k] = tile(2 * hy + j, (1 + k) & 3, i + ((1 + k) >> 2));
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
        const f32x2 slope1 = f2_pack(a.slope1, a.slope1);
        const int Y = oy0 - 1 + hy;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int X = ox0 - 1 + 2 * i + q;
            const bool inside = (unsigned)Y < (unsigned)a.OH && (unsigned)X < (unsigned)a.OW;
            float x[8];                                  /* the stem's output at this position (k_stem_u8x2's epilogue) */
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const f32x2 t = f2_fma(acc[q][o], f2_pack(sw.s[2 * o], sw.s[2 * o + 1]), f2_pack(sw.b[2 * o], sw.b[2 * o + 1]));
                x[2 * o] = act_apply(f2_lo(t), a.act0); x[2 * o + 1] = act_apply(f2_hi(t), a.act0);
            }
            f32x2 e[4];                                  /* rb_expand<0> of block_reg.cuh */
            rb_expand<0>(w, x, inside, slope1, e);
            E(hy, 0, q, i) = make_float4(f2_lo(e[0]), f2_hi(e[0]), f2_lo(e[1]), f2_hi(e[1]));
            E(hy, 1, q, i) = make_float4(f2_lo(e[2]), f2_hi(e[2]), f2_lo(e[3]), f2_hi(e[3]));
        }
    }
    __syncthreads();

    /* ---------------- phase C: depthwise 3x3 + BN + act + projection + BN at two adjacent output pixels ---------------- */
    if (tid < SB_TYO * (SB_TXO / 2)) {
        const int ty = tid / (SB_TXO / 2), j = tid - ty * (SB_TXO / 2);
        const int oy = oy0 + ty, ox = ox0 + 2 * j;
        if (oy < a.OH && ox < a.OW) {
            f32x2 d0[4], d1[4];                          /* the two pixels' depthwise sums, channel pairs */
#pragma unroll
            for (int r = 0; r < 3; r++) {
                /* positions (ty + r, 2j + c), c = 0..3: e of the three taps of pixel 0 (c = 0..2) and pixel 1 (c = 1..3) */
                f32x2 v[4][4];
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        /* one 128-bit load as two channel pairs: left to itself the compiler fetched the pairs with two 64-bit loads, and
                           64-bit loads of lanes 16 bytes apart are 2-way bank conflicts (half of this kernel's shared-memory wavefronts) */
                        const sm100::f4p t = sm100::lds128p(sm100::smem_u32(&E(ty + r, h, c & 1, j + (c >> 1))));
                        v[c][2 * h] = t.a; v[c][2 * h + 1] = t.b;
                    }
                if (r == 0) { rb_taps<0, true, 0>(w, d0, v[0], v[1], v[2]); rb_taps<0, true, 0>(w, d1, v[1], v[2], v[3]); }
                if (r == 1) { rb_taps<1, false, 0>(w, d0, v[0], v[1], v[2]); rb_taps<1, false, 0>(w, d1, v[1], v[2], v[3]); }
                if (r == 2) { rb_taps<2, false, 0>(w, d0, v[0], v[1], v[2]); rb_taps<2, false, 0>(w, d1, v[1], v[2], v[3]); }
            }
            RegBlockArgs ra;
            ra.sloped = a.sloped; ra.slope3 = a.slope3; ra.slope_res = 1.f; ra.slope1 = a.slope1;
            const float dummy[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float *yp = a.y + ((f * a.OH + oy) * (long)a.OW + ox) * 4;
            rb_finish<false, 0>(w, ra, d0, dummy, yp);
            if (ox + 1 < a.OW) rb_finish<false, 0>(w, ra, d1, dummy, yp + 4);
        }
    }
}

} // namespace ffb
