/*
 * pw_tc.cu -- pointwise (1x1) convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
 *
 * Reference path: convolution_pad0_fs1_stride1_all (conv-v6.c:46-91): out[px][oc] = act(s[oc] * sum_ic W[oc][ic] *
 * in[px][ic] + b[oc]).  In batched NHWC this is one GEMM  D[M x N] = A[M x K] * W[N x K]^T  with M = n*h*w pixels,
 * K = ic, N = oc, both operands K-major -- a genuine dense contraction, so it goes to the tensor pipe:
 *
 *   warp 0  TMA producer : A sub-tiles [128 px x 32 ch] (128-byte rows, SWIZZLE_128B) through an S-slot mbarrier ring --
 *                          the ring is K-CHUNK granular (16 KB slots), so even K = 192 next to 150 KB of resident weights
 *                          keeps several loads in flight; the layer's weights once per CTA (resident for the whole loop)
 *   warp 1  MMA issuer   : one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=NS, K=8 per instruction),
 *                          accumulators live in TMEM (double buffered), completion via tcgen05.commit -> mbarrier
 *   warps 2-5            : (a) 3xTF32 split of the landed activation tile, (b) epilogue: tcgen05.ld -> act(fma(acc, s, b))
 *                          [+ fused shortcut add] -> swizzled smem staging -> TMA store (coalesced 128-byte rows, clipped
 *                          at N and M).  (Separate split / epilogue warp groups were tried and measured slower: r1d.)
 *
 * Numerics.  One TF32 pass keeps 10 mantissa bits per operand and misses the box tolerance by ~0.4 px, so
 * the default is the 3xTF32 split: x = hi + lo with hi = rna_tf32(x) (exactly representable in tf32, so it does not
 * matter that the tensor core truncates its inputs -- measured: it does) and lo = x - hi (exact, zero-mean), and
 * D = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi accumulated in fp32.  Round-to-nearest matters: a truncating split leaves an
 * error of constant sign (~2^-22 per product) that compounds over the 55 pointwise layers of the graph; the rounded
 * split is zero-mean (~2^-24).
 * W_hi / W_lo are split once at load; A_hi overwrites the TMA-landed tile in place and A_lo goes straight to TENSOR
 * MEMORY (tcgen05.st) and is consumed by the A-from-TMEM form of tcgen05.mma, so the split costs no extra shared memory.
 * Mode 3 (1xTF32) skips the split (raw fp32 bits are fed to the tensor core).
 *
 * Everything is HBM-bound by design: per 128-pixel tile the MMA work (K*N/16 clk per pass) is far below the time
 * the tile's bytes need at 1/148th of HBM bandwidth, so the pipeline exists to keep ~S*Kc*16 KB of loads in flight per SM.
 */
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include "pw_tc.h"
#include "sm100.cuh"

#include "ffb_internal.h"

using namespace sm100;

namespace {

constexpr int BM = 128;                 /* pixels per tile = UMMA M */
constexpr int A_SUB = BM * 128;         /* bytes of one [128 x 32 fp32] sub-tile */
constexpr int EPI_THREADS = 256;          /* one split+epilogue group = 8 warps, two threads per tile row */
constexpr int MAX_GROUPS = 2;             /* groups take alternate tiles (group g owns accumulator buffer g) */
constexpr int MAX_THREADS = 64 + MAX_GROUPS * EPI_THREADS;
constexpr int SEP_EPI_THREADS = 128;      /* NG == 1: four more warps (10-13) run the epilogue, warps 2-9 only split */

struct TcArgs {
    long M;
    int K, Kc, ksteps_total, NS, nsl, S, G, act, split;   /* G = number of split+epilogue warp groups */
    int tiles;                          /* M tiles */
    uint32_t tmem_cols;
    const float *scale, *bias;          /* [nsl*NS], zero padded */
    const float *res; int ldr, act2, N; /* optional fused shortcut (ffcnn.c:418-423): out = act2(conv + res[m][n]) */
    int allwait;                        /* developer A/B: every split thread waits on every chunk's barrier */
    float *out; int ldo, coff, direct;  /* direct != 0: the epilogue stores straight from registers (no staging tiles: their 32 KB go to the A ring) */
    long long *trace;                   /* developer timeline (tools/tc_trace.py): CTA 0 stamps [role][iteration][event] */
};

#ifdef FFB_TC_TRACE
#define TRACE(role, it, ev) do { if (a.trace && blockIdx.x == 0 && (it) < 64) a.trace[((role) * 64 + (it)) * 4 + (ev)] = clock64(); } while (0)
#else
#define TRACE(role, it, ev) do { } while (0)
#endif

/* round-to-nearest (ties away) to tf32: the result is an fp32 value whose low 13 mantissa bits are zero */
__device__ __forceinline__ float tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

/* Same rounding with two full-rate integer ops (cvt.rna.tf32 issues on the quarter-rate conversion pipe and was the
 * bottleneck of the split warps): adding half a tf32 ulp to the magnitude bits and clearing the low 13 rounds to nearest,
 * ties away from zero. */
__device__ __forceinline__ float tf32_round(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

/* activation as a negative-side slope: linear 1, relu 0, leaky 0.1 (utils.h:15-23) -- branch free in the epilogue */
__device__ __forceinline__ float act_slope(int act) { return act == 2 ? 0.1f : act == 1 ? 0.f : 1.f; }
__device__ __forceinline__ float act_apply(float v, float slope) { return fmaxf(v, v * slope); }     /* slope in [0, 1]: max(v, slope * v) == (v > 0 ? v : slope * v) */

/* NG = split + epilogue warp groups (1 or 2): a template parameter only so that the one-group launch (320 threads) may use
   up to 200 registers -- under the two-group bound (576 threads, 112 registers) the epilogue spilled 212 bytes */
template <int NG>
__device__ __forceinline__ void pw_tc_body(const CUtensorMap &tmA, const CUtensorMap &tmBh, const CUtensorMap &tmBl, const CUtensorMap &tmD, const TcArgs &a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int Kc = a.Kc, NS = a.NS, S = a.S;
    const uint32_t b_sub = (uint32_t)NS * 128;                       /* bytes of one [NS x 32] weight sub-tile */
    uint8_t *sBh = smem;
    uint8_t *sBl = sBh + (size_t)Kc * b_sub;
    uint8_t *sA  = sBl + (a.split ? (size_t)Kc * b_sub : 0);
    uint8_t *sO  = sA + (size_t)S * A_SUB;
    float   *sSc = reinterpret_cast<float *>(sO + (a.direct ? 0 : (size_t)(NG == 2 ? 16 : 4) * 4096));
    float   *sBi = sSc + NS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBi + NS);
    uint64_t *full = bars, *empty = bars + S, *conv = bars + 2 * S, *tfull = bars + 3 * S, *tempty = tfull + 2, *bfull = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TRACE(5, 1, 0);
    const int slice = blockIdx.x % a.nsl, group = blockIdx.x / a.nsl, ngroups = gridDim.x / a.nsl;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmD);
        if (a.split) tma_prefetch_desc(&tmBl);
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(conv + s, NG == 1 ? EPI_THREADS / 2 : EPI_THREADS); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, NG == 1 ? SEP_EPI_THREADS : EPI_THREADS); }
        mbar_init(bfull, 1);
        fence_barrier_init();
        /* the layer's weights do not depend on the previous kernel: fetch them before the PDL wait */
        mbar_arrive_expect_tx(bfull, (uint32_t)((a.split ? 2 : 1) * Kc) * b_sub);
        for (int kc = 0; kc < Kc; kc++) {
            tma_load_2d(sBh + (size_t)kc * b_sub, &tmBh, kc * 32, slice * NS, bfull);
            if (a.split) tma_load_2d(sBl + (size_t)kc * b_sub, &tmBl, kc * 32, slice * NS, bfull);
        }
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    for (int i = threadIdx.x; i < NS; i += blockDim.x) { sSc[i] = a.scale[slice * NS + i]; sBi[i] = a.bias[slice * NS + i]; }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_trigger(); pdl_wait();             /* from here on the kernel touches the previous layer's output */
    if (threadIdx.x == 0) TRACE(5, 1, 1);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_col0 = 0, alo_col0 = 2 * NS;

    if (warp == 0) {
        /* ===================== TMA producer ===================== */
        if (elect_one()) {
            uint32_t cq = 0;                                 /* K chunks issued so far: slot = cq % S */
            int it = 0;
            for (int t = group; t < a.tiles; t += ngroups, it++) {
                for (int kc = 0; kc < Kc; kc++, cq++) {
                    const int s = cq % S; const uint32_t ph = (cq / S) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    if (kc == 0) TRACE(0, it, 0);
                    mbar_arrive_expect_tx(full + s, (uint32_t)A_SUB);
                    tma_load_2d(sA + (size_t)s * A_SUB, &tmA, kc * 32, t * BM, full + s);
                }
            }
        }
    } else if (warp == 1) {
        /* ===================== MMA issuer ===================== */
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_tf32(BM, NS);
            mbar_wait(bfull, 0);
            uint32_t cq = 0;
            int it = 0;
            for (int t = group; t < a.tiles; t += ngroups, it++) {
                const int ab = it & 1; const uint32_t aph = (it >> 1) & 1;
                const uint32_t d = tmem_base + acc_col0 + ab * NS;
                const uint32_t bh_base = smem_u32(sBh), bl_base = smem_u32(sBl);
                uint32_t accum = 0;
                for (int kc = 0; kc < Kc; kc++, cq++) {
                    const int s = cq % S; const uint32_t ph = (cq / S) & 1;
                    mbar_wait(a.split ? conv + s : full + s, ph);
                    if (kc == 0) { TRACE(1, it, 0); mbar_wait(tempty + ab, aph ^ 1); TRACE(1, it, 1); }
                    tc_fence_after_sync();
                    const uint32_t a_base = smem_u32(sA + (size_t)s * A_SUB);
                    const int nk = min(4, a.ksteps_total - kc * 4);            /* k-steps (of 8) in this chunk */
                    /* descriptors once per chunk; a k-step advances the start-address field (bytes >> 4) by 32 B = 2.  Building
                       each descriptor from its address cost ~13 uniform-pipe instructions per MMA, and the single issuing thread
                       -- ~97 cycles per MMA against 48 of tensor time at N = 96 -- paced the whole CTA (profiles/r2k_pwtc_trace.txt) */
                    const uint64_t da = umma_desc_sw128(a_base), dbh = umma_desc_sw128(bh_base + kc * b_sub), dbl = umma_desc_sw128(bl_base + kc * b_sub);
                    if (a.split) {
                        const uint32_t alo = tmem_base + alo_col0 + s * 32;
                        if (nk == 4) {
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) mma_tf32_ts(d, alo + kk * 8, dbh + 2 * kk, idesc, kk ? 1u : accum);      /* A_lo (TMEM) x W_hi */
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) mma_tf32_ss(d, da + 2 * kk, dbl + 2 * kk, idesc, 1);                    /* A_hi x W_lo */
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) mma_tf32_ss(d, da + 2 * kk, dbh + 2 * kk, idesc, 1);                    /* A_hi x W_hi */
                        } else {
                            for (int kk = 0; kk < nk; kk++) mma_tf32_ts(d, alo + kk * 8, dbh + 2 * kk, idesc, kk ? 1u : accum);
                            for (int kk = 0; kk < nk; kk++) mma_tf32_ss(d, da + 2 * kk, dbl + 2 * kk, idesc, 1);
                            for (int kk = 0; kk < nk; kk++) mma_tf32_ss(d, da + 2 * kk, dbh + 2 * kk, idesc, 1);
                        }
                    } else {
                        for (int kk = 0; kk < nk; kk++) mma_tf32_ss(d, da + 2 * kk, dbh + 2 * kk, idesc, kk ? 1u : accum);         /* raw x raw in 1xTF32 mode */
                    }
                    accum = 1;
                    tc_commit(empty + s);        /* ring slot (and its A_lo columns) reusable once these MMAs retire */
                }
                tc_commit(tfull + ab);           /* accumulator ready for the epilogue */
                TRACE(1, it, 2);
            }
        }
    } else {
        /* ===================== split + epilogue warps (two threads per tile row) ===================== */
        /* One group (NG == 1): warps 2-9 only split, warps 10-13 only run the epilogue (thread = tile row, every column chunk) --
           on the same warps the two stages were serial, 3.7 k + 1.9 k cycles per (tile, slice) against 2.3 k of MMA time
           (profiles/r2k_pwtc_trace.txt).  Two groups (NG == 2, K <= 96): each group of 8 warps splits its tile and then runs the
           epilogue of its previous one. */
        const bool epi_warp = NG == 1 && warp >= 10;
        const int q = warp & 3;                              /* TMEM lane quarter this warp may access */
        const int grp = NG == 1 ? 0 : (warp - 2) >> 3;       /* warp group: takes the tiles with it % G == grp */
        const int half = epi_warp ? 0 : ((warp - 2) >> 2) & 1;   /* which alternate chunk / unit this warp handles */
        const int row = q * 32 + lane;
        const int et = NG == 1 ? ((threadIdx.x == 64 || threadIdx.x == 320) ? 0 : 1) : threadIdx.x - 64 - grp * EPI_THREADS + (grp ? 1024 : 0);   /* 0 for the tracing thread(s) */
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int nchunks = (NS + 31) / 32;
        const float slope1 = act_slope(a.act), slope2 = act_slope(a.act2);
        /* warp-private staging: this warp's [32 rows x 128 B] box, 1024-byte aligned, SWIZZLE_128B like the tensor map.
           No CTA-level barrier anywhere in the epilogue: a warp stages its rows, __syncwarp()s and stores its own box. */
        uint8_t *stage = sO + (size_t)(epi_warp ? warp - 10 : warp - 2) * 4096;
        const uint32_t stage_addr = smem_u32(stage), sc_addr = smem_u32(sSc), bi_addr = smem_u32(sBi);

        auto epilogue = [&](int t, int it) {
            const int ab = it & 1; const uint32_t aph = (it >> 1) & 1;
            const long m = (long)t * BM + row;
            const bool has_res = a.res != nullptr && m < a.M;
            if (et == 0) TRACE(3, it, 0);
            mbar_wait(tfull + ab, aph);
            if (et == 0) TRACE(3, it, 1);
            tc_fence_after_sync();
            for (int j = half; j < nchunks; j += (NG == 1 ? 1 : 2)) {        /* this warp's 32-column chunks of its 32 rows */
                const int ncols = min(32, NS - j * 32);
                uint32_t r[2][16];
                tmem_ld16(tmem_base + lane_addr + acc_col0 + ab * NS + j * 32, r[0]);
                if (ncols > 16) tmem_ld16(tmem_base + lane_addr + acc_col0 + ab * NS + j * 32 + 16, r[1]);
                if (!a.direct) { if (lane == 0) tma_store_wait_read<0>(); __syncwarp(); }    /* the previous box of this warp has left shared memory */
                if (et == 0) TRACE(4, it, j == half ? 0 : 2);
                tmem_ld_wait();
                if (et == 0) TRACE(4, it, j == half ? 1 : 3);
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const int cl = j * 32 + hf * 16;
                    if (hf * 16 < ncols) {
                        float4 rv[4];
                        if (has_res) {                       /* skip tensor of the fused shortcut */
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const int n0 = slice * NS + cl + 4 * c;
                                rv[c] = n0 < a.N ? __ldg(reinterpret_cast<const float4 *>(a.res + m * a.ldr + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                        /* scale / bias of the 16 columns first, as eight independent read-only loads: the shared-memory accessors are
                           volatile asm (ordered), and with one load pair per group of 4 columns the whole epilogue ran as a chain of
                           exposed latencies -- 2.5 k cycles per 32 x 32 chunk (tools/tc_trace.py, profiles/r2k_pwtc_trace.txt) */
                        float4 scv[4], biv[4];
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            scv[c] = __ldg(reinterpret_cast<const float4 *>(a.scale + slice * NS + cl + 4 * c));
                            biv[c] = __ldg(reinterpret_cast<const float4 *>(a.bias + slice * NS + cl + 4 * c));
                        }
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const float4 sc = scv[c], bi = biv[c];
                            float4 v;
                            v.x = act_apply(fmaf(__uint_as_float(r[hf][4 * c + 0]), sc.x, bi.x), slope1);
                            v.y = act_apply(fmaf(__uint_as_float(r[hf][4 * c + 1]), sc.y, bi.y), slope1);
                            v.z = act_apply(fmaf(__uint_as_float(r[hf][4 * c + 2]), sc.z, bi.z), slope1);
                            v.w = act_apply(fmaf(__uint_as_float(r[hf][4 * c + 3]), sc.w, bi.w), slope1);
                            if (has_res) {
                                v.x = act_apply(v.x + rv[c].x, slope2); v.y = act_apply(v.y + rv[c].y, slope2);
                                v.z = act_apply(v.z + rv[c].z, slope2); v.w = act_apply(v.w + rv[c].w, slope2);
                            }
                            if (a.direct) {                  /* 64 contiguous bytes per thread and half: whole sectors, merged in L2 */
                                const int n0 = slice * NS + cl + 4 * c;
                                if (m < a.M && n0 < a.N) *reinterpret_cast<float4 *>(a.out + m * a.ldo + a.coff + n0) = v;
                            } else {
                                const int chunk = hf * 4 + c;
                                sts128(stage_addr + lane * 128 + ((chunk ^ (lane & 7)) << 4), v);
                            }
                        }
                    }
                }
                if (!a.direct) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmD, stage, slice * NS + j * 32, t * BM + q * 32); tma_store_commit(); }
                }
            }
            if (et == 0) TRACE(5, it + 2, 3);
            tc_fence_before_sync();
            mbar_arrive(tempty + ab);                        /* accumulator drained (all 256 threads arrive) */
            if (et == 0) TRACE(3, it, 2);
        };

        int it = 0, prev_t = -1, prev_it = -1;
        if (epi_warp) {
            for (int t = group; t < a.tiles; t += ngroups, it++) epilogue(t, it);
        } else
        for (int t = group; t < a.tiles; t += ngroups, it++) {
            if (it % a.G != grp) continue;
            if (a.split) {
                for (int kc = 0; kc < Kc; kc++) {
                    const uint32_t cq = (uint32_t)it * Kc + kc;
                    const int s = cq % S; const uint32_t ph = (cq / S) & 1;
                    if (et == 0 && kc == 0) TRACE(2, it, 0);
                    /* the K chunks of a tile alternate between the two warp quartets of the group (thread = tile row, all four
                       8-channel units of the chunk): a chunk's split is one latency chain -- loads, stores, proxy fence,
                       tcgen05.wait::st, arrive: ~1 k cycles (profiles/r2k_pwtc_trace.txt) -- so two chunks are kept in flight */
                    /* One group (NG == 1, K > 96): the K chunks of a tile alternate between the two warp quartets (thread = tile row,
                       all four 8-channel units): a chunk's split is one latency chain -- loads, stores, proxy fence, tcgen05.wait::st,
                       arrive: ~1 k cycles (profiles/r2k_pwtc_trace.txt) -- so two chunks are kept in flight.
                       Two groups (96 registers): both quartets share every chunk, two units per thread. */
                    constexpr int NU = NG == 1 ? 4 : 2;
                    const bool mine = NG != 1 || (kc & 1) == half;
                    if (mine || a.allwait) mbar_wait(full + s, ph);
                    if (!mine) continue;
                    if (et == 0 && kc == 0) TRACE(2, it, 1);
                    const uint32_t arow = smem_u32(sA + (size_t)s * A_SUB + row * 128);
                    const uint32_t alo = tmem_base + lane_addr + alo_col0 + s * 32;
                    const int nu = min(4, (a.K - kc * 32 + 7) / 8);             /* valid units (8 consecutive k each) in this chunk */
                    float4 x[NU][2]; uint32_t pp[NU][2];
#pragma unroll
                    for (int i = 0; i < NU; i++) {                              /* all loads issued first */
                        const int c2 = NG == 1 ? i : half + 2 * i;
                        pp[i][0] = arow + (((2 * c2) ^ (row & 7)) << 4);
                        pp[i][1] = arow + (((2 * c2 + 1) ^ (row & 7)) << 4);
                        if (c2 < nu) { x[i][0] = lds128(pp[i][0]); x[i][1] = lds128(pp[i][1]); }
                    }
#pragma unroll
                    for (int i = 0; i < NU; i++) {
                        const int c2 = NG == 1 ? i : half + 2 * i;
                        if (c2 < nu) {
                            const float4 x0 = x[i][0], x1 = x[i][1];
                            float4 h0, h1; uint32_t lo[8];
                            h0.x = tf32_round(x0.x); h0.y = tf32_round(x0.y); h0.z = tf32_round(x0.z); h0.w = tf32_round(x0.w);
                            h1.x = tf32_round(x1.x); h1.y = tf32_round(x1.y); h1.z = tf32_round(x1.z); h1.w = tf32_round(x1.w);
                            /* lo = x - hi is exact, symmetric about 0 and has <= 13 significant bits; the tensor core drops the last two */
                            lo[0] = __float_as_uint(x0.x - h0.x); lo[1] = __float_as_uint(x0.y - h0.y);
                            lo[2] = __float_as_uint(x0.z - h0.z); lo[3] = __float_as_uint(x0.w - h0.w);
                            lo[4] = __float_as_uint(x1.x - h1.x); lo[5] = __float_as_uint(x1.y - h1.y);
                            lo[6] = __float_as_uint(x1.z - h1.z); lo[7] = __float_as_uint(x1.w - h1.w);
                            sts128(pp[i][0], h0); sts128(pp[i][1], h1);
                            tmem_st8(alo + c2 * 8, lo);
                        }
                    }
                    fence_proxy_async_smem();                /* in-place A_hi writes -> visible to the tensor core (async proxy) */
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(conv + s);
                }
                if (et == 0) TRACE(2, it, 2);
            }
            if (NG == 1) continue;                           /* the epilogue has its own warps */
            if (prev_t >= 0) epilogue(prev_t, prev_it);
            prev_t = t; prev_it = it;
        }
        if (NG != 1 && prev_t >= 0) epilogue(prev_t, prev_it);
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) TRACE(5, 1, 2);
    if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem_base, a.tmem_cols); }
}

template <int NG> __global__ void k_pw_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                                          const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmD, const TcArgs a);
template <> __global__ void __launch_bounds__(64 + EPI_THREADS + SEP_EPI_THREADS, 1)
k_pw_tc<1>(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
           const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmD, const TcArgs a) { pw_tc_body<1>(tmA, tmBh, tmBl, tmD, a); }
/* 576 threads = 18 warps, allocated as 20 (granularity 4): 96 registers is the most that launches (104 and 112 via __maxnreg__
   compile but the launch is refused), so this variant keeps a small spill (the two-group plan only serves K <= 96 layers) */
template <> __global__ void __launch_bounds__(64 + 2 * EPI_THREADS, 1)
k_pw_tc<2>(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
           const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmD, const TcArgs a) { pw_tc_body<2>(tmA, tmBh, tmBl, tmD, a); }

__global__ void k_split_weights(const float *__restrict__ flt, int row, int N, int K, int Kld, float *__restrict__ hi, float *__restrict__ lo)
{
    const int total = N * Kld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / Kld, k = i - n * Kld;
        const float w = k < K ? flt[(long)n * row + k] : 0.f;
        const float h = tf32_rna(w);
        hi[i] = h; lo[i] = tf32_rna(w - h);
    }
}

__global__ void k_scale_bias(const float *__restrict__ flt, int row, int N, int NP, float *__restrict__ sc, float *__restrict__ bi)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NP; i += gridDim.x * blockDim.x) {
        sc[i] = i < N ? flt[(long)i * row + row - 4] : 0.f;
        bi[i] = i < N ? flt[(long)i * row + row - 3] : 0.f;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}

/* fp32 [rows][cols] with row stride `stride_floats`; box = 32 cols x box_rows, 128-byte swizzle, OOB reads give 0 */
int make_map(CUtensorMap *m, const void *base, uint64_t cols, uint64_t rows, uint64_t stride_floats, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { ffb_set_error("cuTensorMapEncodeTiled unavailable"); return -1; }
    cuuint64_t gdim[2] = { cols, rows }, gstr[1] = { stride_floats * 4 };
    cuuint32_t box[2] = { 32, box_rows }, estr[2] = { 1, 1 };
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ffb_set_error("cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu stride=%llu box_rows=%u base=%p", (int)r,
                                           (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)stride_floats, box_rows, base); return -1; }
    return 0;
}

} // namespace

int ffb_make_tensor_map(CUtensorMap *m, const void *base, int rank, const unsigned long long *dims,
                        const unsigned long long *strides_bytes, const unsigned *box, int swizzle128)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { ffb_set_error("cuTensorMapEncodeTiled unavailable"); return -1; }
    cuuint64_t gdim[5], gstr[4]; cuuint32_t bx[5], estr[5] = { 1, 1, 1, 1, 1 };
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; if (i + 1 < rank) gstr[i] = strides_bytes[i]; }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ffb_set_error("cuTensorMapEncodeTiled failed (%d), rank %d, box0 %u", (int)r, rank, box[0]); return -1; }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Measured tensor-pipe peak for the roofline denominators ("measure TF32 peak on device"): every SM issues
 * back-to-back tcgen05.mma.kind::tf32 M=128 N=256 K=8 instructions on operand tiles resident in shared memory (contents
 * irrelevant: zeros), four independent TMEM accumulators... the tensor pipe is the only thing working.  Not a product
 * path: bench.py calls it once to report `roofline.fused` against a measured number instead of a nominal one. */
__global__ void __launch_bounds__(128, 1) k_tf32_peak(int iters)
{
    extern __shared__ uint8_t peak_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(peak_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar; __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<float4 *>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = umma_idesc_tf32(128, 256);
        const uint32_t a0 = smem_u32(sm), b0 = a0 + 16384;            /* A: [128 x 32] SW128, B: [256 x 32] SW128 */
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
                mma_tf32_ss(tmem + (uint32_t)((it & 1) * 256), umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc, 1);
        tc_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

extern "C" int ffb_measure_tf32_peak(double *tflops)
{
    static ffb_smem_cfg cfg;
    const size_t smem = 16384 + 32768 + 1024;
    if (ffb_ensure_smem((const void *)k_tf32_peak, smem, &cfg) != 0) return -1;
    const int sms = ffb_num_sms(), iters = 20000;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { ffb_set_error("cudaEventCreate failed"); return -1; }
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a, 0);
        k_tf32_peak<<<sms, 128, smem, 0>>>(iters);
        cudaEventRecord(b, 0);
        if (cudaEventSynchronize(b) != cudaSuccess) { ffb_set_error("tf32 peak kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (tflops) *tflops = (double)sms * iters * 4.0 * 2.0 * 128 * 256 * 8 / (best * 1e-3) / 1e12;
    return 0;
}

/* general form: optional element strides (a stride-S conv reads every S-th pixel: box = S x the elements wanted), swizzle */
int ffb_make_tensor_map_ex(CUtensorMap *m, const void *base, int rank, const unsigned long long *dims,
                           const unsigned long long *strides_bytes, const unsigned *box, const unsigned *elem_strides, int swizzle128)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) { ffb_set_error("cuTensorMapEncodeTiled unavailable"); return -1; }
    cuuint64_t gdim[5], gstr[4]; cuuint32_t bx[5], estr[5] = { 1, 1, 1, 1, 1 };
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; if (elem_strides) estr[i] = elem_strides[i]; if (i + 1 < rank) gstr[i] = strides_bytes[i]; }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ffb_set_error("cuTensorMapEncodeTiled failed (%d), rank %d, box %u x %u, stride %u", (int)r, rank, box[0], rank > 1 ? box[1] : 0, estr[1]); return -1; }
    return 0;
}

static long long *g_tc_trace = nullptr;
extern "C" void ffb_tc_set_trace(long long *dev_buf) { g_tc_trace = dev_buf; }   /* developer hook, effective only in -DFFB_TC_TRACE builds */

struct PwTcPlan {
    int K, N, act, mode, split;
    int Kc, ksteps_total, NS, nsl, S, OB, NP, Kld, direct;
    uint32_t tmem_cols; size_t smem;
    float *d_bhi, *d_blo, *d_scb;
    CUtensorMap tmBh, tmBl;
    int num_sms;
};

static bool plan_tiling_with(PwTcPlan *p, int direct);

/* The epilogue's staging tiles (TMA store, 32 KB per warp group) compete with the A ring for shared memory; a direct-store
 * epilogue (registers -> global, no staging) buys 2 more 16 KB ring slots.  FFCNN_PW_DIRECT = 1 forces it. */
static bool plan_tiling(PwTcPlan *p)
{
    static const int env = getenv("FFCNN_PW_DIRECT") ? atoi(getenv("FFCNN_PW_DIRECT")) : -1;
    if (env == 1) return plan_tiling_with(p, 1);
    /* measured (profiles/r2g_pw_direct.txt): the direct-store epilogue is 10-25 % SLOWER on every shape despite the deeper
       ring -- the ring depth was not the limiter (the epilogue's latency chain was, profiles/r2k_pwtc_trace.txt) -- so it is
       only a fallback for shapes whose staged plan does not fit */
    return plan_tiling_with(p, 0) || plan_tiling_with(p, 1);
}

static bool plan_tiling_with(PwTcPlan *p, int direct)
{
    const int N16 = (p->N + 15) & ~15;
    const size_t limit = 227 * 1024 - 2048;          /* dynamic smem ceiling minus alignment slack */
    /* Few N slices (every slice re-reads and re-splits the activation tile) against a deep A ring (the chain
       TMA -> split -> MMA -> commit -> refill has ~3 k cycles of latency per slot): take the first slice count whose resident
       weights leave `want` 16 KB slots -- half a tile's K chunks, at least 2 -- else the deepest ring any slice count allows.
       Measured (profiles/r2l_pw_tc.txt): 192 -> 192 (6 chunks) runs 0.87 ms with 3 slices / 5 slots against 1.00 ms with 2 slices /
       2 slots; 120 -> 255 (4 chunks) 0.057 ms with 2 slices / 3 slots against 0.067 ms with 3 slices / 6 slots. */
    static const int env_nsl = getenv("FFCNN_PW_NSL") ? atoi(getenv("FFCNN_PW_NSL")) : 0;       /* developer override: minimum number of N slices */
    const int maxS = std::min(8, std::max(2, 2 * p->Kc)), want = std::min(maxS, std::max(2, (p->Kc + 1) / 2));
    bool have = false; PwTcPlan best = *p;
    for (int nsl = std::max(1, env_nsl); nsl <= 8; nsl++) {
        int NS = (N16 + nsl - 1) / nsl;
        NS = nsl > 1 ? (NS + 31) & ~31 : (NS + 15) & ~15;
        if (NS > 256) continue;
        const size_t B = (size_t)(p->split ? 2 : 1) * p->Kc * NS * 128;
        bool found = false;
        for (int S = maxS; S >= 1 && !found; S--)
            for (int OB = (p->Kc <= 3 ? MAX_GROUPS : 1); OB >= 1 && !found; OB--) {   /* OB = warp groups (a second one pays off when tiles are small); staging = 8 warps x 4 KB each */
                /* two groups take alternate tiles and each waits only on its own tiles' ring slots: a slot must then
                   always serve the same group (S a multiple of 2 * Kc), or a group would skip mbarrier phases and its
                   parity wait would alias (found as a hang on 48 -> 224 with S = 2, Kc = 2) */
                if (OB == 2 && S % (2 * p->Kc) != 0) continue;
                const size_t smem = B + (size_t)S * A_SUB + (direct ? 0 : (size_t)(OB == 2 ? 16 : 4) * 4096) + 2 * NS * 4 + (3 * S + 5) * 8 + 16;   /* staging: 4 KB per epilogue warp */
                const int tmem = 2 * NS + (p->split ? S * 32 : 0);
                if (smem <= limit && tmem <= 512) {
                    found = true;
                    if (!have || S > best.S) {
                        best = *p;
                        best.NS = NS; best.nsl = nsl; best.S = S; best.OB = OB; best.NP = nsl * NS; best.direct = direct;
                        best.smem = smem + 1024;
                        if (best.smem < 120 * 1024) best.smem = 120 * 1024;       /* one CTA per SM: TMEM is allocated per CTA */
                        uint32_t c = 32; while ((int)c < tmem) c <<= 1;
                        best.tmem_cols = c;
                        have = true;
                    }
                }
            }
        if (have && best.S >= want) break;
    }
    if (have) { *p = best; return true; }
    return false;
}

PwTcPlan *pw_tc_plan_create(int K, int N, int act, int mode)
{
    if (K % 4 || K < 8 || N < 8 || N > 2048) return nullptr;
    /* auto: measured on B200 (profiles/r1e_pw_kernel_choice.txt) -- for K <= 8 or N <= 8 a 128-pixel tile carries so few bytes
       that this kernel's fixed per-tile latencies (mbarrier hand-offs, proxy fence, TMA store issue) exceed what the FFMA
       streaming kernel needs; from K, N >= 16 the tensor pipe wins */
    if (mode == 0 && (K < 16 || N < 16)) return nullptr;
    if (!encode_fn()) return nullptr;
    PwTcPlan *p = new PwTcPlan(); memset(p, 0, sizeof *p);
    p->K = K; p->N = N; p->act = act; p->mode = mode == 3 ? 3 : 2; p->split = p->mode == 2;
    p->Kc = (K + 31) / 32;
    p->ksteps_total = (p->Kc - 1) * 4 + ((K - 32 * (p->Kc - 1)) + 7) / 8;
    p->Kld = (K + 3) & ~3;
    if (!plan_tiling(p)) { delete p; return nullptr; }
    p->num_sms = ffb_num_sms();
    return p;
}

void pw_tc_plan_destroy(PwTcPlan *p)
{
    if (!p) return;
    cudaFree(p->d_bhi); cudaFree(p->d_blo); cudaFree(p->d_scb);
    delete p;
}

/* a direct-store plan writes whole groups of 4 channels: N % 4 != 0 needs the tensor's own zero pad lanes behind it */
bool pw_tc_supports(const PwTcPlan *p, int ldo, int coff)
{
    if (!p || !p->direct) return true;
    if (ldo % 4 || coff % 4) return false;
    return p->N % 4 == 0 || (coff == 0 && ldo == ((p->N + 3) & ~3));
}

const char *pw_tc_mode_name(const PwTcPlan *p) { return p && p->mode == 3 ? "1xtf32" : "3xtf32"; }

int pw_tc_prepare(PwTcPlan *p, const float *d_packed, int row, cudaStream_t st)
{
    if (!p->d_scb && cudaMalloc(&p->d_scb, 2 * (size_t)p->NP * sizeof(float)) != cudaSuccess) { ffb_set_error("pw_tc: cudaMalloc failed"); return -1; }
    k_scale_bias<<<(p->NP + 255) / 256, 256, 0, st>>>(d_packed, row, p->N, p->NP, p->d_scb, p->d_scb + p->NP);
    if (p->split) {
        const size_t n = (size_t)p->N * p->Kld;
        if (!p->d_bhi && (cudaMalloc(&p->d_bhi, n * sizeof(float)) != cudaSuccess || cudaMalloc(&p->d_blo, n * sizeof(float)) != cudaSuccess)) { ffb_set_error("pw_tc: cudaMalloc failed"); return -1; }
        k_split_weights<<<(int)((n + 255) / 256), 256, 0, st>>>(d_packed, row, p->N, p->K, p->Kld, p->d_bhi, p->d_blo);
        if (make_map(&p->tmBh, p->d_bhi, p->K, p->N, p->Kld, p->NS) != 0) return -1;
        if (make_map(&p->tmBl, p->d_blo, p->K, p->N, p->Kld, p->NS) != 0) return -1;
    } else {
        /* 1xTF32: the packed reference rows ARE the K-major weight matrix (row stride K+4 floats) */
        if (make_map(&p->tmBh, d_packed, p->K, p->N, row, p->NS) != 0) return -1;
        p->tmBl = p->tmBh;
    }
    static ffb_smem_cfg attr_set;
    if (ffb_ensure_smem((const void *)k_pw_tc<1>, 227 * 1024, &attr_set) != 0) return -1;
    static ffb_smem_cfg attr_set2;
    if (ffb_ensure_smem((const void *)k_pw_tc<2>, 227 * 1024, &attr_set2) != 0) return -1;
    if (cudaGetLastError() != cudaSuccess) { ffb_set_error("pw_tc: weight preparation launch failed"); return -1; }
    return 0;
}

int pw_tc_run(PwTcPlan *p, const float *in, int ldi, float *out, int ldo, int coff, long M, cudaStream_t st,
              const float *res, int ldr, int act2)
{
    CUtensorMap tmA, tmD;
    if (make_map(&tmA, in, p->K, (uint64_t)M, ldi, BM) != 0) return -1;
    if (make_map(&tmD, out + coff, p->N, (uint64_t)M, ldo, 32) != 0) return -1;      /* one warp's rows per store box */
    TcArgs a;
    a.M = M; a.K = p->K; a.Kc = p->Kc; a.ksteps_total = p->ksteps_total; a.NS = p->NS; a.nsl = p->nsl; a.S = p->S; a.G = p->OB;
    a.act = p->act; a.split = p->split; a.tiles = (int)((M + BM - 1) / BM); a.tmem_cols = p->tmem_cols;
    a.scale = p->d_scb; a.bias = p->d_scb + p->NP;
    a.res = res; a.ldr = ldr; a.act2 = act2; a.N = p->N;
    a.out = out; a.ldo = ldo; a.coff = coff; a.direct = p->direct;
    /* every split thread observes every phase of every ring slot (3 % on the 192 -> 192 microbenchmark): a quartet that only
       waited on its own chunks could test a slot's barrier one phase early, and mbarrier parity waits alias modulo 2 */
    { static const int aw = getenv("FFCNN_PW_ALLWAIT") ? atoi(getenv("FFCNN_PW_ALLWAIT")) : 1; a.allwait = aw; }
    if (p->direct && !pw_tc_supports(p, ldo, coff)) { ffb_set_error("pw_tc: direct-store plan cannot write %d channels at offset %d of a %d-float pixel", p->N, coff, ldo); return -1; }
    a.trace = g_tc_trace;
    long want = (long)a.tiles * p->nsl;
    int grid = (int)(want < p->num_sms ? want : p->num_sms);
    grid -= grid % p->nsl;
    if (grid < p->nsl) grid = p->nsl;
    cudaError_t e = p->OB == 2 ? launch_pdl(k_pw_tc<2>, dim3(grid), dim3(64 + 2 * EPI_THREADS), p->smem, st, tmA, p->tmBh, p->tmBl, tmD, a)
                               : launch_pdl(k_pw_tc<1>, dim3(grid), dim3(64 + EPI_THREADS + SEP_EPI_THREADS), p->smem, st, tmA, p->tmBh, p->tmBl, tmD, a);
    if (e != cudaSuccess) { ffb_set_error("pw_tc launch failed: %s (grid %d smem %zu)", cudaGetErrorString(e), grid, p->smem); return -1; }
    return 0;
}
