/*
 * block_reg.cuh -- fused inverted-residual block for the high-resolution, low-channel layers (expanded width <= 24):
 *
 *     x --1x1 expand, BN, act--> e --3x3 depthwise (stride 1|2), BN, act--> d --1x1 project, BN, act--> [+ x] --> y
 *
 * (conv-v6.c:46-91 pointwise, 96-287 depthwise, ffcnn.c:412-423 dropout/shortcut -- the layers L1-L3, L4-L8, L9-L11 of
 * yolo-fastest-1.1.)  At 160x160 these layers have so few channels that one pixel's whole expanded vector fits in a
 * lane's registers, so nothing is staged anywhere: a warp owns a strip of 64 input columns and sweeps it top to bottom.
 * Per input row each lane loads its two pixels of x, expands them with FFMAs whose weight operands come straight from
 * the constant bank (the block's ~170-620 weights travel as a __grid_constant__ kernel parameter), fetches the
 * expanded vectors of the two neighbouring columns with warp shuffles, and adds the row's taps to three rolling
 * depthwise accumulators (the output rows above, at and below it).  The output row that just became complete is
 * BN+activated, projected, optionally added to x (still in registers), and stored.  HBM sees x once and y once.
 *
 * Arithmetic is plain fp32 FFMA in the reference's tap order (ky, then kx: conv-v0.c:16-25).  Lanes 0 and 31 (lane 0
 * only for stride 2) are halo lanes: they expand pixels for their neighbours but produce no output.
 *
 * The kernels are bound by instruction issue, not by the FMA pipe (ncu: 56 % of the warp instructions were FFMA, the
 * schedulers issued on 60-67 % of their cycles with 2-3 warps each), so every channel-wise operation runs on PAIRS of
 * adjacent channels with Blackwell's packed fp32 instructions (fma/mul/add.rn.f32x2 -> FFMA2/FMUL2/FADD2): the same IEEE
 * operation per component -- results are bit-identical to the scalar form -- in half the issue slots.  ptxas feeds them
 * the weight pairs straight from uniform registers (LDCU.128 = two pairs) and a pixel's scalar as a broadcast operand.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "sm100.cuh"

namespace ffb {

using sm100::f32x2;
using sm100::f2_pack; using sm100::f2_lo; using sm100::f2_hi; using sm100::f2_fma; using sm100::f2_mul; using sm100::f2_add;

template <int CIN, int CEXP, int COUT>
struct alignas(16) RegBlockW {
    /* every matrix is stored with the OUTPUT channel innermost: consecutive FFMAs of the unrolled loops then use
       consecutive constants (one 128-bit uniform load feeds four) and independent accumulators (no dependent chains) */
    float w1[CIN][CEXP], s1[CEXP], b1[CEXP];
    float wd[9][CEXP], sd[CEXP], bd[CEXP];
    float w2[CEXP][COUT], s3[COUT], b3[COUT];
};
/* channel pair i (channels 2i, 2i+1) of a row of weights */
__device__ __forceinline__ f32x2 rb_pr(const float *row, int i) { return reinterpret_cast<const f32x2 *>(row)[i]; }

struct RegBlockArgs {
    const float *x; float *y;
    int N, H, W, OH, OW, R, nsx, nsy;       /* R output rows per warp strip; nsx x nsy strips per frame */
    float slope1, sloped, slope3, slope_res;
};

constexpr int REG_WARPS = 4;

/* utils.h:15-23 as max(v, slope * v), on a channel pair */
__device__ __forceinline__ f32x2 rb_act2(f32x2 v, f32x2 slope2)
{
    const f32x2 m = f2_mul(v, slope2);
    return f2_pack(fmaxf(f2_lo(v), f2_lo(m)), fmaxf(f2_hi(v), f2_hi(m)));
}

constexpr int RB_CH = 8;            /* expanded channels processed at a time: bounds the transient registers of a row step */
constexpr int RB_P = RB_CH / 2;     /* ... as channel pairs */

/* expand channels [C0, C0 + RB_CH) of one pixel: e[c] = act(s1[c] * sum_k x[k] * w1[c][k] + b1[c]), or 0 outside the image
 * (the depthwise conv's zero padding) */
template <int C0, int CIN, int CEXP, int COUT>
__device__ __forceinline__ void rb_expand(const RegBlockW<CIN, CEXP, COUT> &w, const float (&x)[CIN], bool inside, f32x2 slope2, f32x2 (&e)[RB_P])
{
#pragma unroll
    for (int c = 0; c < RB_P; c++) e[c] = f2_mul(f2_pack(x[0], x[0]), rb_pr(w.w1[0], C0 / 2 + c));
#pragma unroll
    for (int k = 1; k < CIN; k++)
#pragma unroll
        for (int c = 0; c < RB_P; c++) e[c] = f2_fma(f2_pack(x[k], x[k]), rb_pr(w.w1[k], C0 / 2 + c), e[c]);
#pragma unroll
    for (int c = 0; c < RB_P; c++) e[c] = inside ? rb_act2(f2_fma(e[c], rb_pr(w.s1, C0 / 2 + c), rb_pr(w.b1, C0 / 2 + c)), slope2) : 0ull;
}

template <int CIN>
__device__ __forceinline__ void rb_load(const float *p, bool ok, float (&x)[CIN])
{
#pragma unroll
    for (int v = 0; v < CIN / 4; v++) {
        const float4 t = ok ? __ldg(reinterpret_cast<const float4 *>(p) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[4 * v] = t.x; x[4 * v + 1] = t.y; x[4 * v + 2] = t.z; x[4 * v + 3] = t.w;
    }
}

/* ---- row ring (RING variants): x rows travel global -> shared memory by cp.async, several rows ahead of their use, into slots
 * that are PRIVATE to a lane (a lane reads back only the 2 pixels it copied itself, so cp.async.wait_group is the only
 * synchronisation).  Why not plain register loads two rows ahead (the non-RING variants): ncu showed a fifth of all warp samples
 * waiting at the first use of a row that had been requested ~400 instructions earlier -- the consumer waits on a scoreboard
 * that the NEWEST row's loads (issued just before it) also count on, so every row paid a full DRAM latency regardless of the
 * lookahead.  cp.async groups complete in order and wait_group<D> names exactly the row that is needed. ---- */
__device__ __forceinline__ void rb_cp16(uint32_t dst, const float *src, bool ok)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void rb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void rb_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float4 rb_lds(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
/* the lane's two adjacent pixels (2 * CIN contiguous floats at p) of one row -> ring slot; !ok writes zeros (rows / columns outside the image) */
template <int CIN>
__device__ __forceinline__ void rb_ring_issue(uint32_t slot_addr, const float *p, bool ok)
{
#pragma unroll
    for (int v = 0; v < CIN / 2; v++) rb_cp16(slot_addr + v * 512, p + 4 * v, ok);
    rb_commit();
}
template <int CIN>
__device__ __forceinline__ void rb_ring_take(uint32_t slot_addr, float (&x0)[CIN], float (&x1)[CIN])
{
#pragma unroll
    for (int v = 0; v < CIN / 4; v++) {
        const float4 t = rb_lds(slot_addr + v * 512), u = rb_lds(slot_addr + (CIN / 4 + v) * 512);
        x0[4 * v] = t.x; x0[4 * v + 1] = t.y; x0[4 * v + 2] = t.z; x0[4 * v + 3] = t.w;
        x1[4 * v] = u.x; x1[4 * v + 1] = u.y; x1[4 * v + 2] = u.z; x1[4 * v + 3] = u.w;
    }
}

/* the channel chunk of a neighbouring lane (delta = -1: the lane to the left, +1: to the right) */
template <int DELTA>
__device__ __forceinline__ void rb_neighbour(const f32x2 (&e)[RB_P], f32x2 (&o)[RB_P])
{
#pragma unroll
    for (int c = 0; c < RB_P; c++) {
        const float lo = DELTA < 0 ? __shfl_up_sync(0xffffffffu, f2_lo(e[c]), 1) : __shfl_down_sync(0xffffffffu, f2_lo(e[c]), 1);
        const float hi = DELTA < 0 ? __shfl_up_sync(0xffffffffu, f2_hi(e[c]), 1) : __shfl_down_sync(0xffffffffu, f2_hi(e[c]), 1);
        o[c] = f2_pack(lo, hi);
    }
}

/* d[C0 ..] (+)= taps of kernel row DY applied to the three horizontally adjacent expanded pixels l, m, r (channel chunk C0) */
template <int DY, bool SET, int C0, int CIN, int CEXP, int COUT>
__device__ __forceinline__ void rb_taps(const RegBlockW<CIN, CEXP, COUT> &w, f32x2 (&d)[CEXP / 2], const f32x2 (&l)[RB_P], const f32x2 (&m)[RB_P], const f32x2 (&r)[RB_P])
{
#pragma unroll
    for (int c = 0; c < RB_P; c++) d[C0 / 2 + c] = SET ? f2_mul(l[c], rb_pr(w.wd[DY * 3], C0 / 2 + c)) : f2_fma(l[c], rb_pr(w.wd[DY * 3], C0 / 2 + c), d[C0 / 2 + c]);
#pragma unroll
    for (int c = 0; c < RB_P; c++) d[C0 / 2 + c] = f2_fma(m[c], rb_pr(w.wd[DY * 3 + 1], C0 / 2 + c), d[C0 / 2 + c]);
#pragma unroll
    for (int c = 0; c < RB_P; c++) d[C0 / 2 + c] = f2_fma(r[c], rb_pr(w.wd[DY * 3 + 2], C0 / 2 + c), d[C0 / 2 + c]);
}

/* finish one output pixel: BN + act of the depthwise sum, projection, BN + act, optional shortcut, store.
 * PART splits a wide block into channel slices run as consecutive launches (the projection is a sum over expanded channels):
 * 0 = whole block; 1 = first slice, store the raw partial sums; 2 = middle slice, add to them; 3 = last slice, add, BN + act. */
template <bool RES, int PART, int CIN, int CEXP, int COUT>
__device__ __forceinline__ void rb_finish(const RegBlockW<CIN, CEXP, COUT> &w, const RegBlockArgs &a, const f32x2 (&d)[CEXP / 2], const float (&xc)[CIN], float *yp,
                                          uint32_t ypre = 0)           /* PART >= 2: shared-memory address of the prefetched partial sums (planes of 32 float4), 0 = read y */
{
    const f32x2 sloped2 = f2_pack(a.sloped, a.sloped);
    f32x2 dd[CEXP / 2];
#pragma unroll
    for (int c = 0; c < CEXP / 2; c++) dd[c] = rb_act2(f2_fma(d[c], rb_pr(w.sd, c), rb_pr(w.bd, c)), sloped2);
    f32x2 o[COUT / 2];
    if (PART >= 2) {
#pragma unroll
        for (int v = 0; v < COUT / 4; v++) {
            const float4 t = ypre ? rb_lds(ypre + v * 512) : reinterpret_cast<const float4 *>(yp)[v];
            o[2 * v] = f2_pack(t.x, t.y); o[2 * v + 1] = f2_pack(t.z, t.w);
        }
    }
    if (PART < 2) {
#pragma unroll
        for (int co = 0; co < COUT / 2; co++) o[co] = 0ull;
    }
#pragma unroll
    for (int c = 0; c < CEXP; c++) {
        const float dc = (c & 1) ? f2_hi(dd[c / 2]) : f2_lo(dd[c / 2]);
#pragma unroll
        for (int co = 0; co < COUT / 2; co++) o[co] = f2_fma(f2_pack(dc, dc), rb_pr(w.w2[c], co), o[co]);
    }
#pragma unroll
    for (int co = 0; co < COUT / 2; co++) {
        f32x2 s = o[co];
        if (PART == 0 || PART == 3) s = rb_act2(f2_fma(s, rb_pr(w.s3, co), rb_pr(w.b3, co)), f2_pack(a.slope3, a.slope3));
        if (RES) s = rb_act2(f2_add(s, f2_pack(xc[2 * co < CIN ? 2 * co : 0], xc[2 * co + 1 < CIN ? 2 * co + 1 : 0])), f2_pack(a.slope_res, a.slope_res));
        o[co] = s;
    }
#pragma unroll
    for (int v = 0; v < COUT / 4; v++) reinterpret_cast<float4 *>(yp)[v] = make_float4(f2_lo(o[2 * v]), f2_hi(o[2 * v]), f2_lo(o[2 * v + 1]), f2_hi(o[2 * v + 1]));
}

/* ------------------------------------------------------------------ stride 1: two output pixels per lane, 60 per warp */
template <int CIN, int CEXP, int COUT, bool RES, bool RING = false>
__global__ void __launch_bounds__(REG_WARPS * 32, 2) k_block_reg_s1(const __grid_constant__ RegBlockW<CIN, CEXP, COUT> w, const RegBlockArgs a)
{
    constexpr int CP = CEXP / 2;
    constexpr int ST = 3, D = ST - 1;                  /* RING: slots per warp, rows of lookahead (the row loop is unrolled by ST: slot numbers are constants) */
    constexpr uint32_t SLOT = (2 * CIN / 4) * 512;      /* bytes per slot: 2 pixels x CIN floats per lane, stored as planes of 32 float4 (conflict-free) */
    __shared__ float4 ring_s[RING ? REG_WARPS * ST * (2 * CIN / 4) * 32 : 1];
    const int lane = threadIdx.x & 31;
    [[maybe_unused]] const uint32_t ring = sm100::smem_u32(ring_s) + (threadIdx.x >> 5) * ST * SLOT + lane * 16;
    const long strip = (long)blockIdx.x * REG_WARPS + (threadIdx.x >> 5);
    const int per_frame = a.nsx * a.nsy;
    if (strip >= (long)a.N * per_frame) return;
    const int n = (int)(strip / per_frame), sr = (int)(strip - (long)n * per_frame), sy = sr / a.nsx, sx = sr - sy * a.nsx;
    const int ox_lo = sx * 60, oy0 = sy * a.R, oy1 = min(oy0 + a.R, a.OH);
    const int ix0 = ox_lo - 2 + 2 * lane, ix1 = ix0 + 1;                       /* this lane's two columns */
    const bool in0 = ix0 >= 0 && ix0 < a.W, in1 = ix1 >= 0 && ix1 < a.W;
    const bool writes = lane >= 1 && lane <= 30 && ix0 < a.W;                   /* W is even: ix1 is inside whenever ix0 is */
    const float *xf = a.x + (long)n * a.H * a.W * CIN;
    float *yf = a.y + (long)n * a.OH * a.OW * COUT;
    const f32x2 slope1 = f2_pack(a.slope1, a.slope1);
    sm100::pdl_trigger(); sm100::pdl_wait();

    f32x2 A0[CP], A1[CP], B0[CP], B1[CP], C0[CP], C1[CP];                       /* three rolling output rows x two pixels, channel pairs */
    float xc0[CIN], xc1[CIN], xn0[CIN], xn1[CIN], xp0[CIN], xp1[CIN];           /* x of the row in flight / the next row / the previous row */
#pragma unroll
    for (int c = 0; c < CP; c++) { A0[c] = A1[c] = B0[c] = B1[c] = C0[c] = C1[c] = 0ull; }
#pragma unroll
    for (int k = 0; k < CIN; k++) xp0[k] = xp1[k] = 0.f;
    float xm0[CIN], xm1[CIN];                                                   /* x of the row after next: the loads run two rows ahead of their use */
    /* RING: row rn -> slot; ix0 is even and W is even, so the lane's two pixels are inside or outside together and contiguous */
    auto ring_issue = [&](int rn, int slot) {
        const bool rok = rn >= 0 && rn < a.H && rn <= oy1;
        rb_ring_issue<CIN>(ring + slot * SLOT, xf + ((long)min(max(rn, 0), a.H - 1) * a.W + max(ix0, 0)) * CIN, rok && in0);
    };
    if constexpr (RING) {
#pragma unroll
        for (int k = 0; k < D; k++) ring_issue(oy0 - 1 + k, k);
    } else {
        const int r = oy0 - 1; const bool rok = r >= 0;
        rb_load<CIN>(xf + ((long)max(r, 0) * a.W + max(ix0, 0)) * CIN, rok && in0, xn0);
        rb_load<CIN>(xf + ((long)max(r, 0) * a.W + max(ix1, 0)) * CIN, rok && in1, xn1);
        rb_load<CIN>(xf + ((long)min(oy0, a.H - 1) * a.W + max(ix0, 0)) * CIN, in0, xm0);
        rb_load<CIN>(xf + ((long)min(oy0, a.H - 1) * a.W + max(ix1, 0)) * CIN, in1, xm1);
    }
    /* one input row: dm = output row r-1 (gets kernel row 2 and is finished), d0 = row r (kernel row 1), dp = row r+1 (kernel row 0, first contribution) */
    auto step = [&](int r, auto slot_c, f32x2 (&dm0)[CEXP / 2], f32x2 (&dm1)[CEXP / 2], f32x2 (&d00)[CEXP / 2], f32x2 (&d01)[CEXP / 2], f32x2 (&dp0)[CEXP / 2], f32x2 (&dp1)[CEXP / 2]) {   /* CEXP / 2 spelled out: cicc crashes on a generic lambda whose parameter type names a local constexpr */
        constexpr int slot = decltype(slot_c)::value;
        /* expand -> neighbour exchange -> taps, RB_CH channels at a time */
        auto chunk = [&](auto c0, bool rin) {
            constexpr int C0 = decltype(c0)::value;
            f32x2 e0[RB_P], e1[RB_P], el[RB_P], er[RB_P];
            rb_expand<C0>(w, xc0, rin && in0, slope1, e0);
            rb_expand<C0>(w, xc1, rin && in1, slope1, e1);
            rb_neighbour<-1>(e1, el); rb_neighbour<1>(e0, er);
            rb_taps<2, false, C0>(w, dm0, el, e0, e1); rb_taps<2, false, C0>(w, dm1, e0, e1, er);
            rb_taps<1, false, C0>(w, d00, el, e0, e1); rb_taps<1, false, C0>(w, d01, e0, e1, er);
            rb_taps<0, true, C0>(w, dp0, el, e0, e1);  rb_taps<0, true, C0>(w, dp1, e0, e1, er);
        };
        if constexpr (RING) {
            /* request row r + D into the slot whose row (r - 1) was consumed by the previous step, then take row r: at most the D newer rows stay in flight */
            ring_issue(r + D, (slot + D) % ST);
            rb_wait<D>();
            rb_ring_take<CIN>(ring + slot * SLOT, xc0, xc1);
        } else {
#pragma unroll
            for (int k = 0; k < CIN; k++) { xc0[k] = xn0[k]; xc1[k] = xn1[k]; xn0[k] = xm0[k]; xn1[k] = xm1[k]; }
            /* prefetch x two rows ahead (ncu: with one row of lookahead a quarter of the warp samples waited for these loads) */
            const int rn = r + 2; const bool rok = rn < a.H && rn <= oy1;
            rb_load<CIN>(xf + ((long)min(rn, a.H - 1) * a.W + max(ix0, 0)) * CIN, rok && in0, xm0);
            rb_load<CIN>(xf + ((long)min(rn, a.H - 1) * a.W + max(ix1, 0)) * CIN, rok && in1, xm1);
        }
        const bool rin = r >= 0 && r < a.H;
        chunk(std::integral_constant<int, 0>(), rin);
        if constexpr (CEXP > 8)  chunk(std::integral_constant<int, 8>(), rin);
        if constexpr (CEXP > 16) chunk(std::integral_constant<int, 16>(), rin);
        const int oy = r - 1;
        if (writes && oy >= oy0 && oy < oy1) {
            float *yp = yf + ((long)oy * a.OW + ix0) * COUT;
            rb_finish<RES, 0>(w, a, dm0, xp0, yp);
            rb_finish<RES, 0>(w, a, dm1, xp1, yp + COUT);
        }
#pragma unroll
        for (int k = 0; k < CIN; k++) { xp0[k] = xc0[k]; xp1[k] = xc1[k]; }
    };
    /* rows are taken three at a time with no branch around the steps (a warp shuffle under a branch the compiler cannot
       prove uniform costs four extra synchronisation instructions each); rows past the strip load zeros and write nothing */
    for (int r = oy0 - 1; r <= oy1; r += 3) {
        step(r, std::integral_constant<int, 0>(), A0, A1, B0, B1, C0, C1);
        step(r + 1, std::integral_constant<int, 1>(), B0, B1, C0, C1, A0, A1);
        step(r + 2, std::integral_constant<int, 2>(), C0, C1, A0, A1, B0, B1);
    }
}

/* ------------------------------------------------------------------ stride 2: one output pixel per lane, 31 per warp */
template <int CIN, int CEXP, int COUT, int PART, bool RING = false>
__global__ void __launch_bounds__(REG_WARPS * 32, 3) k_block_reg_s2(const __grid_constant__ RegBlockW<CIN, CEXP, COUT> w, const RegBlockArgs a)
{
    constexpr int CP = CEXP / 2;
    constexpr int ST = 4, D = ST - 1;                  /* RING: slots per warp, rows of lookahead (four rows per trip of the row loop: slot numbers are constants) */
    constexpr uint32_t SLOT = (2 * CIN / 4) * 512;
    __shared__ float4 ring_s[RING ? REG_WARPS * ST * (2 * CIN / 4) * 32 : 1];
    /* PART >= 2 (a later channel slice adds to the partial sums in y): the y row an odd input row closes travels in that row's
       cp.async group -- the read-modify-write no longer starts with an exposed L2 round trip per output row */
    constexpr bool YPRE = RING && PART >= 2;
    constexpr uint32_t YSLOT = (COUT / 4) * 512;
    __shared__ float4 yring_s[YPRE ? REG_WARPS * 2 * (COUT / 4) * 32 : 1];
    const int lane = threadIdx.x & 31;
    [[maybe_unused]] const uint32_t ring = sm100::smem_u32(ring_s) + (threadIdx.x >> 5) * ST * SLOT + lane * 16;
    [[maybe_unused]] const uint32_t yring = sm100::smem_u32(yring_s) + (threadIdx.x >> 5) * 2 * YSLOT + lane * 16;
    const long strip = (long)blockIdx.x * REG_WARPS + (threadIdx.x >> 5);
    const int per_frame = a.nsx * a.nsy;
    if (strip >= (long)a.N * per_frame) return;
    const int n = (int)(strip / per_frame), sr = (int)(strip - (long)n * per_frame), sy = sr / a.nsx, sx = sr - sy * a.nsx;
    const int ox = sx * 31 + lane - 1, oy0 = sy * a.R, oy1 = min(oy0 + a.R, a.OH);
    const int ix0 = 2 * ox, ix1 = ix0 + 1;                                      /* centre and right tap columns; the left tap comes from lane - 1 */
    const bool in0 = ix0 >= 0 && ix0 < a.W, in1 = ix1 >= 0 && ix1 < a.W;
    const bool writes = lane >= 1 && ox < a.OW;
    const float *xf = a.x + (long)n * a.H * a.W * CIN;
    float *yf = a.y + (long)n * a.OH * a.OW * COUT;
    const f32x2 slope1 = f2_pack(a.slope1, a.slope1);
    sm100::pdl_trigger(); sm100::pdl_wait();

    f32x2 A[CP], B[CP];                                                         /* output rows oy and oy + 1, channel pairs */
    float xc0[CIN], xc1[CIN], xn0[CIN], xn1[CIN];
#pragma unroll
    for (int c = 0; c < CP; c++) A[c] = B[c] = 0ull;
    const int r_first = 2 * oy0 - 1, r_last = 2 * (oy1 - 1) + 1;
    float xm0[CIN], xm1[CIN];                                                   /* the row after next: the loads run two rows ahead of their use */
    auto ring_issue = [&](int rn, int slot) {                                   /* ix0 and W are even: both pixels are inside or outside together, and contiguous */
        const bool rok = rn >= 0 && rn < a.H && rn <= r_last;
        if constexpr (YPRE) {
            if (!(slot & 1)) {                                                  /* even slots hold the odd input rows 2oy - 1: they close output row oy - 1 */
                const int oyp = (rn - 1) >> 1; const bool yok = writes && oyp >= oy0 && oyp < oy1;
                const float *yp = yf + ((long)min(max(oyp, 0), a.OH - 1) * a.OW + min(max(ox, 0), a.OW - 1)) * COUT;
#pragma unroll
                for (int v = 0; v < COUT / 4; v++) rb_cp16(yring + (slot >> 1) * YSLOT + v * 512, yp + 4 * v, yok);
            }
        }
        rb_ring_issue<CIN>(ring + slot * SLOT, xf + ((long)min(max(rn, 0), a.H - 1) * a.W + max(ix0, 0)) * CIN, rok && in0);
    };
    if constexpr (RING) {
#pragma unroll
        for (int k = 0; k < D; k++) ring_issue(r_first + k, k);
    } else {
        const bool rok = r_first >= 0;
        rb_load<CIN>(xf + ((long)max(r_first, 0) * a.W + max(ix0, 0)) * CIN, rok && in0, xn0);
        rb_load<CIN>(xf + ((long)max(r_first, 0) * a.W + max(ix1, 0)) * CIN, rok && in1, xn1);
        const int r2 = r_first + 1; const bool rok2 = r2 < a.H && r2 <= r_last;
        rb_load<CIN>(xf + ((long)min(r2, a.H - 1) * a.W + max(ix0, 0)) * CIN, rok2 && in0, xm0);
        rb_load<CIN>(xf + ((long)min(r2, a.H - 1) * a.W + max(ix1, 0)) * CIN, rok2 && in1, xm1);
    }
    float dummy[CIN];
#pragma unroll
    for (int k = 0; k < CIN; k++) dummy[k] = 0.f;
    auto fetch_row = [&](int r, auto slot_c) {                                  /* xc <- row r, prefetch row r + 2 (RING: request row r + D) */
        constexpr int slot = decltype(slot_c)::value;
        if constexpr (RING) {
            ring_issue(r + D, (slot + D) % ST);
            rb_wait<D>();
            rb_ring_take<CIN>(ring + slot * SLOT, xc0, xc1);
        } else {
#pragma unroll
            for (int k = 0; k < CIN; k++) { xc0[k] = xn0[k]; xc1[k] = xn1[k]; xn0[k] = xm0[k]; xn1[k] = xm1[k]; }
            const int rn = r + 2; const bool rok = rn < a.H && rn <= r_last;
            rb_load<CIN>(xf + ((long)min(rn, a.H - 1) * a.W + max(ix0, 0)) * CIN, rok && in0, xm0);
            rb_load<CIN>(xf + ((long)min(rn, a.H - 1) * a.W + max(ix1, 0)) * CIN, rok && in1, xm1);
        }
    };
    /* odd input row 2oy-1: kernel row 2 of output row oy-1 (cur) and kernel row 0 of output row oy (nxt), RB_CH channels at a time */
    auto odd_chunk = [&](auto c0, bool rin, f32x2 (&cur)[CEXP / 2], f32x2 (&nxt)[CEXP / 2]) {
        constexpr int C0 = decltype(c0)::value;
        f32x2 e0[RB_P], e1[RB_P], el[RB_P];
        rb_expand<C0>(w, xc0, rin && in0, slope1, e0);
        rb_expand<C0>(w, xc1, rin && in1, slope1, e1);
        rb_neighbour<-1>(e1, el);
        rb_taps<2, false, C0>(w, cur, el, e0, e1);
        rb_taps<0, true, C0>(w, nxt, el, e0, e1);
    };
    auto even_chunk = [&](auto c0, bool rin, f32x2 (&nxt)[CEXP / 2]) {                /* even input row 2oy: kernel row 1 of output row oy */
        constexpr int C0 = decltype(c0)::value;
        f32x2 e0[RB_P], e1[RB_P], el[RB_P];
        rb_expand<C0>(w, xc0, rin && in0, slope1, e0);
        rb_expand<C0>(w, xc1, rin && in1, slope1, e1);
        rb_neighbour<-1>(e1, el);
        rb_taps<1, false, C0>(w, nxt, el, e0, e1);
    };
    /* rows come in (odd, even) pairs: odd row 2oy-1 closes output row oy-1 and opens row oy; even row 2oy is row oy's centre */
    auto pair = [&](int oy, auto slot_c, f32x2 (&cur)[CEXP / 2], f32x2 (&nxt)[CEXP / 2]) {
        constexpr int slot = decltype(slot_c)::value;
        {
            const int r = 2 * oy - 1; const bool rin = r >= 0 && r < a.H;
            fetch_row(r, std::integral_constant<int, slot>());
            /* the closing taps of `cur` must be complete before it is finished, the opening taps of `nxt` are independent:
               chunks are run for both, then `cur` is finished */
            odd_chunk(std::integral_constant<int, 0>(), rin, cur, nxt);
            if constexpr (CEXP > 8)  odd_chunk(std::integral_constant<int, 8>(), rin, cur, nxt);
            if constexpr (CEXP > 16) odd_chunk(std::integral_constant<int, 16>(), rin, cur, nxt);
            if (writes && oy - 1 >= oy0 && oy - 1 < oy1) rb_finish<false, PART>(w, a, cur, dummy, yf + ((long)(oy - 1) * a.OW + ox) * COUT, YPRE ? yring + (slot >> 1) * YSLOT : 0u);
        }
        {
            const int r = 2 * oy; const bool rin = r < a.H && oy < oy1;
            fetch_row(r, std::integral_constant<int, slot + 1>());
            even_chunk(std::integral_constant<int, 0>(), rin, nxt);
            if constexpr (CEXP > 8)  even_chunk(std::integral_constant<int, 8>(), rin, nxt);
            if constexpr (CEXP > 16) even_chunk(std::integral_constant<int, 16>(), rin, nxt);
        }
    };
    for (int oy = oy0; oy <= oy1; oy += 2) {                                   /* no branch around the pairs: see k_block_reg_s1 */
        pair(oy, std::integral_constant<int, 0>(), A, B);
        pair(oy + 1, std::integral_constant<int, 2>(), B, A);
    }
}

} // namespace ffb
