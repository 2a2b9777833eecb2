/*
 * block_ws.cuh -- warp-specialised variant of the fused inverted-residual block kernel (block_mma.cuh) for the block
 * shapes whose tiles give the depthwise stage only five work units (the 40x40 maps of yolo-fastest-1.1: L22-L57).
 *
 * In k_block_mma every chunk of expanded channels is two CTA-wide phases -- stage A (expand accumulators -> BN + act -> E in
 * shared memory) and stage B (3x3 depthwise -> projection GEMM) -- separated by __syncthreads.  A 8x20 tile has 5 stage-B
 * units (2 rows x 16 positions each) for 8 warps, so three warps idle through the longer phase (ncu: 26 % of all warp samples
 * were barrier stalls; profiles/r2q_blockmma_timeline.txt: stage B 4.3 k cycles on warp 0, 0.6 k on warp 7).  Here the roles
 * are fixed instead:
 *
 *   warps 0-4  (B warps)   stage B of chunk k from E[k & 1]; block epilogue after a tile's last chunk
 *   warps 5-7  (A warps)   stage A of chunk k+1 into E[(k+1) & 1] while the B warps work on chunk k, then the x split of the
 *                          next tile when it is due and the tcgen05 expand GEMM of chunk k+2 (each A warp issues the MMAs of
 *                          its own m-tiles); lane 0 of warp 5 also issues every TMA / bulk load
 *
 * with ONE __syncthreads per chunk as the hand-over.  The A warps own TMEM lanes 32-127 only (a warp reaches the lane quarter
 * warp % 4), so the x tile is cut into m-tiles of 96 pixels; rows 0-31 of each 128-row MMA are never read.  The expand
 * accumulators are double buffered in TMEM: the GEMM of chunk k+2 is issued right after the barrier that opens step k (its
 * buffer was drained by stage A of chunk k during step k-1), so its latency and the ~100 cycles each tcgen05.mma costs the
 * issuing thread never sit between a drain and a barrier.
 * Weight chunks (one 16-channel group each, so that two E buffers fit next to the x tiles) stay resident when the block has at
 * most four of them and stream through a four-slot ring otherwise: a chunk is read by the GEMM, by stage A one chunk later and
 * by stage B one chunk after that.
 *
 * Arithmetic, fragment maps, weight layout (k_prep_block with tc = 1) and numerics are those of k_block_mma<TC = true>.
 */
#pragma once
#include "block_mma.cuh"

namespace ffb {

constexpr int WS_BW = 5;                                /* stage-B warps */
constexpr int WS_ATHREADS = (BLK_WARPS - WS_BW) * 32;   /* stage-A threads */
constexpr int WS_MROWS = WS_ATHREADS;                   /* pixels per m-tile */
constexpr int WS_RING = 4;

template <int KS1, int NT3, int S, int MTW, int GC>
__global__ void __launch_bounds__(BLK_THREADS, 2) k_block_ws(const __grid_constant__ CUtensorMap tmX, const BlkArgs a)
{
    extern __shared__ __align__(128) float4 blk_smem4[];
    float *smem = reinterpret_cast<float *>(blk_smem4);
    constexpr int CIN_P = 8 * KS1, SXs = CIN_P + 4, COUT_P = 8 * NT3;
    constexpr int SEs = 16 * GC + (S == 1 ? 8 : 4);
    constexpr bool QUAD = MTW >= 2;
    constexpr BlkChunk off(GC, KS1, NT3, true);
    constexpr int KP = 8 * KS1, KC = (KS1 + 3) / 4;
    static_assert(MTW == 1 || MTW == 2, "a B warp owns one unit");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int HW = a.HW, NC = a.NC;
    const bool resident = NC <= a.R;

    float    *sSB3 = smem;                                          /* 96 floats */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 96);       /* full_x[3], full_w[4], dfull[2] */
    uint64_t *full_x = bars, *full_w = bars + 3, *dfull = bars + 7;   /* dfull[2] */
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    int2     *sMap = reinterpret_cast<int2 *>(smem + 128);          /* [xrows]: m-tile row -> { byte offset of its E row or -1, hy | hx << 16 } */
    float    *sW = smem + 128 + 2 * a.xrows;
    sW += ((1024u - (sm100::smem_u32(sW) & 1023u)) & 1023u) >> 2;   /* UMMA SWIZZLE_128B atoms are 1024-byte aligned */
    float    *sXB = sW + a.R * off.total;                           /* [XB][xbuf_floats] */
    float    *sE = sXB + a.XB * a.xbuf_floats;                      /* [2][HH * HW * SEs] */
    const uint32_t sE_addr = sm100::smem_u32(sE), sW_addr = sm100::smem_u32(sW);
    const uint32_t e_bytes = (uint32_t)a.HH * HW * SEs * 4;
    const uint32_t trash = sE_addr + 2 * e_bytes;                    /* scratch row (SEs floats) behind the two E buffers */
    const uint32_t x_bytes = (uint32_t)a.XH * a.XW * SXs * 4;
    constexpr uint32_t w_bytes = (uint32_t)off.total * 4;
    const int XP = a.XH * a.XW;

    if (tid == 0) {
        sm100::tma_prefetch_desc(&tmX);
        for (int i = 0; i < 7; i++) sm100::mbar_init(bars + i, 1);
        sm100::mbar_init(dfull, BLK_WARPS - WS_BW); sm100::mbar_init(dfull + 1, BLK_WARPS - WS_BW);   /* one commit per A warp and chunk */
        sm100::fence_barrier_init();
    }
    if (warp == 0) sm100::tmem_alloc(tmem_slot, a.tmem_cols);
    if (tid < 2 * COUT_P) sSB3[tid] = a.sb3[tid];
    for (int xp = tid; xp < a.xrows; xp += BLK_THREADS) {
        const int ry = xp / a.XW, rx = xp - ry * a.XW, hy = ry + a.yo, hx = rx + a.xo;
        sMap[xp] = make_int2(xp < XP ? (hy * HW + hx) * SEs * 4 : -1, hy | (hx << 16));
    }
    for (int i = tid; i < 2 * a.HH * HW * SEs / 4; i += BLK_THREADS) reinterpret_cast<float4 *>(sE)[i] = blk_zero4();
    /* B warps: top-left output pixel (ty, tx) of the lane's quad (or pixel pair) in the warp's unit */
    uint32_t dwbase = 0; int tyx = 0x7fff << 16;
    {
        const int p = warp * 16 + 2 * g;
        int ty, tx; bool valid;
        if (QUAD) { const int np = (a.TH + 1) / 2 * a.TW, pc = min(p, np - 2); const int rp = pc / a.TW; tx = pc - rp * a.TW; ty = 2 * rp; valid = p < np; }
        else      { const int np = a.TH * a.TW, pc = min(p, np - 2); ty = pc / a.TW; tx = pc - ty * a.TW; valid = p < np; }
        dwbase = sE_addr + (uint32_t)(((ty * S) * HW + tx * S) * SEs + 4 * t) * 4;
        if (valid && warp < WS_BW) tyx = ty << 16 | tx;
    }
    const int nunits = QUAD ? ((a.TH + 1) / 2 * a.TW + 15) >> 4 : (a.TH * a.TW + 15) >> 4;
    const bool has_unit = warp < nunits;                                        /* nunits <= WS_BW (planner) */
    const uint32_t rowpitch = (uint32_t)HW * SEs * 4;
    sm100::tc_fence_before_sync();
    __syncthreads();
    sm100::tc_fence_after_sync();
    pdl_trigger(); pdl_wait();
    const f32x2 slope1_2 = sm100::f2_pack(a.slope1, a.slope1), sloped2 = sm100::f2_pack(a.sloped, a.sloped);

    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tq_addr = (uint32_t)((warp & 3) * 32) << 16;                 /* this warp's TMEM lane quarter */
    const uint32_t dcol0 = tmem_base + (uint32_t)a.nmt * 2 * KP;
    [[maybe_unused]] int tr_n = 0;
    const int my_tiles = (int)((a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int K = my_tiles * NC;                                                /* chunks this CTA walks, in order */
    auto tile_of = [&](int it) { return (long)blockIdx.x + (long)it * gridDim.x; };
    auto wslot = [&](int j, int c) { return resident ? c : (j & (WS_RING - 1)); };
    auto wpar  = [&](int j) { return resident ? 0u : (uint32_t)(j >> 2) & 1u; };

    auto load_x = [&](int it) {                                                 /* one thread */
        const BlkTile q = blk_tile<S>(a, tile_of(it));
        const int b = it % a.XB;
        sm100::mbar_arrive_expect_tx(full_x + b, x_bytes);
        sm100::tma_load_4d(sXB + b * a.xbuf_floats, &tmX, 0, q.ix0 + a.xo, q.iy0 + a.yo, q.n, full_x + b);
    };
    auto load_w = [&](int chunk, int slot) {                                    /* one thread */
        sm100::mbar_arrive_expect_tx(full_w + slot, w_bytes);
        bulk_load(sW_addr + (uint32_t)slot * w_bytes, a.wchunks + (long)chunk * off.total, w_bytes, full_w + slot);
    };
    /* expand GEMM of the chunk in weight slot `slot` for m-tile mt -> D[mt] (one thread) */
    auto issue_expand = [&](uint32_t slot, int mt, uint32_t buf) {
        constexpr uint32_t sub = 16 * GC * 128;                                 /* bytes of one [16*GC x 32] B sub-tile */
        constexpr uint32_t idesc = sm100::umma_idesc_tf32(128, 16 * GC);
        const uint32_t bh = sW_addr + slot * w_bytes, bl = bh + KC * sub;
        const uint64_t dbh = sm100::umma_desc_sw128(bh), dbl = sm100::umma_desc_sw128(bl);
        const uint32_t d = dcol0 + (uint32_t)(mt * 2 + buf) * 16 * GC;
        const uint32_t ahi = tmem_base + (uint32_t)mt * 2 * KP, alo = ahi + KP;
#pragma unroll
        for (int ks = 0; ks < KS1; ks++)                                        /* x_lo . W_hi */
            sm100::mma_tf32_ts(d, alo + 8 * ks, dbh + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < KS1; ks++)                                        /* x_hi . W_lo */
            sm100::mma_tf32_ts(d, ahi + 8 * ks, dbl + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, 1);
#pragma unroll
        for (int ks = 0; ks < KS1; ks++)                                        /* x_hi . W_hi */
            sm100::mma_tf32_ts(d, ahi + 8 * ks, dbh + (((ks >> 2) * sub + (ks & 3) * 32) >> 4), idesc, 1);
    };

    /* ------------------------------------------------------------------ A warps ------------------------------------------------------------------
       Every A warp's lane 0 issues the GEMMs of "its" m-tiles (aw, aw + 3) and commits to dfull (3 arrivals per chunk): one thread
       issuing all of them cost ~100 cycles per tcgen05.mma and was the longest part of the A side. */
    const int aw = warp - WS_BW;
    const int arow = aw * 32 + lane;                                            /* row of this thread inside an m-tile */
    constexpr int NMT_MAX = 4;                                                  /* planner: nmt <= 4 */
    int a_it = 0, a_c = 0, a_j = 0;                                             /* chunk the A warps expand next */
    int a_iy0 = 0, a_ix0 = 0; bool a_border = false;
    int m_it = 0, m_c = 0, m_j = 0;                                             /* chunk whose expand GEMM is issued next */
    /* x tile of tile `it` -> TMEM as the A operand, split hi/lo (thread = pixel = TMEM lane); every earlier GEMM has completed */
    auto split_x = [&](int it) {
        const int b = it % a.XB;
        sm100::mbar_wait(full_x + b, (uint32_t)(it / a.XB) & 1u);
        const float *sX = sXB + b * a.xbuf_floats;
        for (int mt = 0; mt < a.nmt; mt++) {
            const int p = mt * WS_MROWS + arow;
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
    };
    /* expand GEMM of chunk m_j into D[m_j & 1]: lane 0 of every A warp issues its m-tiles and commits.  Called after a CTA-wide
       barrier that follows the drain of that D buffer (stage A of chunk m_j - 2) and the x split of the chunk's tile. */
    auto issue_next = [&]() {
        if (lane == 0) {
            const int slot = wslot(m_j, m_c);
            sm100::mbar_wait(full_w + slot, wpar(m_j));
            sm100::tc_fence_after_sync();
            for (int mt = aw; mt < a.nmt; mt += BLK_WARPS - WS_BW) issue_expand((uint32_t)slot, mt, (uint32_t)m_j & 1u);
            sm100::tc_commit(dfull + (m_j & 1));
        }
        __syncwarp();
        if (++m_c == NC) { m_c = 0; m_it++; }
        m_j++;
    };
    /* stage A of chunk a_j: expand accumulators TMEM -> BN + act -> E[a_j & 1]; then get the chunk after it going */
    auto a_work = [&]() {
        const int j = a_j, c = a_c;
        if (c == 0) {
            const BlkTile q = blk_tile<S>(a, tile_of(a_it));
            a_iy0 = q.iy0; a_ix0 = q.ix0;
            a_border = q.iy0 < 0 || q.ix0 < 0 || q.iy0 + a.HH > a.H || q.ix0 + HW > a.W;
        }
        const int slot = wslot(j, c);
        sm100::mbar_wait(full_w + slot, wpar(j));
        const float *wc = sW + slot * off.total;
        const uint32_t eb = sE_addr + (uint32_t)(j & 1) * e_bytes;
        BTRACE(10);
        sm100::mbar_wait(dfull + (j & 1), (uint32_t)(j >> 1) & 1u);
        BTRACE(12);
        sm100::tc_fence_after_sync();
#pragma unroll
        for (int gr = 0; gr < GC; gr++) {
            f4p s1v[4], b1v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                s1v[i] = ld4p(wc + off.s1 + gr * 16 + 4 * i);
                b1v[i] = ld4p(wc + off.b1 + gr * 16 + 4 * i);
            }
            /* two m-tiles' accumulators are requested before the first is used, and nothing below branches: rows that map to no halo
               pixel store into a scratch row behind the E buffers, pixels outside the image are zeroed with a mask -- the eight
               float4 of a pair are independent instruction streams for the scheduler */
#pragma unroll
            for (int m0 = 0; m0 < NMT_MAX; m0 += 2) {
                if (m0 < a.nmt) {
                    uint32_t r[2][16];
                    sm100::tmem_ld16(dcol0 + tq_addr + (uint32_t)((m0 * 2 + (j & 1)) * GC + gr) * 16, r[0]);
                    if (m0 + 1 < a.nmt) sm100::tmem_ld16(dcol0 + tq_addr + (uint32_t)(((m0 + 1) * 2 + (j & 1)) * GC + gr) * 16, r[1]);
                    const int2 mp0 = sMap[m0 * WS_MROWS + arow], mp1 = sMap[(m0 + 1) * WS_MROWS + arow];   /* sMap has NMT_MAX * WS_MROWS entries */
                    uint32_t msk[2], dst[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int2 mp = u ? mp1 : mp0;
                        const int iy = a_iy0 + (mp.y & 0xffff), ix = a_ix0 + (mp.y >> 16);
                        const bool inside = !a_border || ((unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W);
                        msk[u] = inside ? 0xffffffffu : 0u;
                        dst[u] = (m0 + u < a.nmt && mp.x >= 0) ? eb + (uint32_t)mp.x : trash;
                    }
                    sm100::tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; u++)
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            f4p acc; acc.a = (f32x2)r[u][4 * i] | ((f32x2)r[u][4 * i + 1] << 32); acc.b = (f32x2)r[u][4 * i + 2] | ((f32x2)r[u][4 * i + 3] << 32);
                            float4 v = bn_act4p(acc, s1v[i], b1v[i], slope1_2);
                            v.x = __uint_as_float(__float_as_uint(v.x) & msk[u]); v.y = __uint_as_float(__float_as_uint(v.y) & msk[u]);
                            v.z = __uint_as_float(__float_as_uint(v.z) & msk[u]); v.w = __uint_as_float(__float_as_uint(v.w) & msk[u]);
                            sm100::sts128(dst[u] + (gr * 16 + 4 * i) * 4, v);
                        }
                }
            }
        }
        BTRACE(13);
        if (++a_c == NC) { a_c = 0; a_it++; }
        a_j++;
        /* the chunk whose GEMM is issued after the next barrier (m_j = j + 2) opens a tile: its x goes to TMEM now, once the GEMM
           of the previous tile's last chunk (j + 1, issued at the top of this step) has completed */
        if (m_j < K && m_c == 0) {
            sm100::mbar_wait(dfull + ((m_j - 1) & 1), (uint32_t)((m_j - 1) >> 1) & 1u);
            sm100::tc_fence_after_sync();
            split_x(m_it);
        }
        sm100::tc_fence_before_sync();
        BTRACE(14);
    };

    if (tid == WS_BW * 32) {                              /* lane 0 of the first A warp issues every TMA / bulk load */
        for (int it = 0; it < a.XB && it < my_tiles; it++) load_x(it);
        const int nw = resident ? NC : min(WS_RING, K);
        for (int j = 0; j < nw; j++) load_w(j % NC, j);
    }
    if (warp >= WS_BW) {
        split_x(0);
        sm100::tc_fence_before_sync();
        sm100::named_bar_sync(1, WS_ATHREADS);
        issue_next();                                     /* chunk 0 */
        if (K > 1) issue_next();                          /* chunk 1 (NC >= 2: same tile) */
        a_work();                                         /* chunk 0 -> E[0] */
    }

    /* ------------------------------------------------------------------ chunk loop ------------------------------------------------------------------ */
    float pacc[MTW][NT3][4];
    int it = 0, c = 0;                                                          /* tile / chunk of iteration k */
    int ring_c = WS_RING % max(NC, 1);                                          /* chunk id of the next ring load (sequence index k - 1 + WS_RING) */
    for (int k = 0; k < K; k++) {
        BTRACE(1);
        __syncthreads();                                  /* E[k & 1] is complete, B of chunk k - 1 is over: E[(k + 1) & 1], its weight slot and (at a tile start) the previous tile's x are free */
        BTRACE(11);
        if (tid == WS_BW * 32) {
            if (!resident && k >= 1 && k - 1 + WS_RING < K) { load_w(ring_c, (k - 1) & (WS_RING - 1)); if (++ring_c == NC) ring_c = 0; }
            if (c == 0 && it >= 1 && it - 1 + a.XB < my_tiles) load_x(it - 1 + a.XB);
        }
        if (warp >= WS_BW) {
            if (m_j < K) issue_next();                    /* chunk k + 2 */
            if (k + 1 < K) a_work();                      /* chunk k + 1 */
        } else {
            /* ---------------- stage B: depthwise 3x3 in registers -> projection GEMM ---------------- */
            if (c == 0) {
#pragma unroll
                for (int mi = 0; mi < MTW; mi++)
#pragma unroll
                    for (int nt = 0; nt < NT3; nt++)
#pragma unroll
                        for (int i = 0; i < 4; i++) pacc[mi][nt][i] = 0.f;
            }
            const int slot = wslot(k, c);
            sm100::mbar_wait(full_w + slot, wpar(k));
            const float *wc = sW + slot * off.total;
            const float *wl4 = wc + lane * 4, *wt4 = wc + 4 * t;
            const uint32_t ebase = dwbase + (uint32_t)(k & 1) * e_bytes;
            if (has_unit) {
#pragma unroll
                for (int grp = 0; grp < GC; grp++) {
                    f4p wd[9];
#pragma unroll
                    for (int i = 0; i < 9; i++) wd[i] = ld4p(wt4 + off.wd + (grp * 9 + i) * 16);
                    const f4p sd = ld4p(wt4 + off.sd + grp * 16), bd = ld4p(wt4 + off.bd + grp * 16);
                    uint32_t ah[MTW][2][4], al[MTW][2][4];
                    {
                        constexpr uint32_t px = SEs * 4;
                        constexpr int NR = QUAD ? S + 3 : 3, NCOL = S + 3;             /* input rows / columns the unit touches */
                        f4p d[2][2];                                                   /* [row of the quad][pixel] */
#pragma unroll
                        for (int h = 0; h < 2; h++) { d[h][0] = blk_zero4p(); d[h][1] = blk_zero4p(); }
#pragma unroll
                        for (int r = 0; r < NR; r++) {
                            const uint32_t row = ebase + grp * 64 + r * rowpitch;
                            f4p e[NCOL];
#pragma unroll
                            for (int i = 0; i < NCOL; i++) e[i] = lds128p(row + i * px);
#pragma unroll
                            for (int h = 0; h < (QUAD ? 2 : 1); h++) {
                                const int dy = r - h * S;                              /* tap row of this input row for quad row h */
                                if (dy >= 0 && dy < 3) {
                                    blk_fma4p(d[h][0], e[0], wd[dy * 3]); blk_fma4p(d[h][0], e[1], wd[dy * 3 + 1]); blk_fma4p(d[h][0], e[2], wd[dy * 3 + 2]);
                                    blk_fma4p(d[h][1], e[S], wd[dy * 3]); blk_fma4p(d[h][1], e[S + 1], wd[dy * 3 + 1]); blk_fma4p(d[h][1], e[S + 2], wd[dy * 3 + 2]);
                                }
                            }
                        }
#pragma unroll
                        for (int h = 0; h < (QUAD ? 2 : 1); h++) {
                            const float4 d0 = bn_act4p(d[h][0], sd, bd, sloped2), d1 = bn_act4p(d[h][1], sd, bd, sloped2);
                            split_tf32x2(d0.x, d1.x, ah[h][0][0], ah[h][0][1], al[h][0][0], al[h][0][1]);
                            split_tf32x2(d0.y, d1.y, ah[h][0][2], ah[h][0][3], al[h][0][2], al[h][0][3]);
                            split_tf32x2(d0.z, d1.z, ah[h][1][0], ah[h][1][1], al[h][1][0], al[h][1][1]);
                            split_tf32x2(d0.w, d1.w, ah[h][1][2], ah[h][1][3], al[h][1][2], al[h][1][3]);
                        }
                    }
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
            /* ---------------- block epilogue after the tile's last chunk: BN + act [+ shortcut from the resident x tile] -> y ---------------- */
            if (c == NC - 1) {
                const BlkTile q = blk_tile<S>(a, tile_of(it));
                const int b = it % a.XB;
                sm100::mbar_wait(full_x + b, (uint32_t)(it / a.XB) & 1u);       /* completed long ago (the A warps split this tile); makes the TMA data visible to this thread */
                const float *sX = sXB + b * a.xbuf_floats;
                const f32x2 slope3_2 = sm100::f2_pack(a.slope3, a.slope3), sloper_2 = sm100::f2_pack(a.slope_res, a.slope_res);
#pragma unroll
                for (int mi = 0; mi < MTW; mi++) {
                    const int ty = (tyx >> 16) + (QUAD ? mi : 0), tx = tyx & 0xffff;
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
        }
        if (warp < WS_BW && c == NC - 1) BTRACE(20);
        if (++c == NC) { c = 0; it++; }
    }
    sm100::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) { sm100::tc_fence_after_sync(); sm100::tmem_dealloc(tmem_base, a.tmem_cols); }
}

} // namespace ffb
