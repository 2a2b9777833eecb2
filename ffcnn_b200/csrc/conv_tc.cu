/*
 * conv_tc.cu -- dense convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
 *
 * Reference paths: convolution_generic + im2row (conv-v6.c:9-42), the explicit im2col + GEMM of conv-v2.c:7-33,63-87, and
 * convolution_pad0_fs1_stride1_all (conv-v6.c:46-91) for pointwise layers whose weights do not fit pw_tc.cu's
 * resident-weight plan (yolov3's 512 -> 1024, 1024 -> 256 ...):
 *     out[n][oy][ox][co] = act(s[co] * sum_{ky,kx,ci} W[co][ci][ky][kx] * in[n][oy*S - P + ky][ox*S - P + kx][ci] + b[co])
 * The reference materialises the unfolded patches (im2row: ow x K floats per output row; conv-v2: K x ow*oh) and then runs
 * dot products.  Here nothing is unfolded: the GEMM  D[M x N] = A[M x K] * W[N x K]^T  (M = output pixels, N = filters,
 * K = taps x channels) walks K as (tap, 32-channel chunk) blocks, and the A operand of a K block is fetched by ONE TMA tiled
 * load of the NHWC activation tensor at the tap's offset -- box [32 ch, TW, TH, TN] with element strides (1, S, S, 1), so
 * a stride-S conv reads every S-th pixel, and out-of-image coordinates (the conv's zero padding) arrive as zeros.  The
 * 128 rows of the box (TW*TH*TN output pixels, 128-byte rows, SWIZZLE_128B) ARE the K-major UMMA operand tile.
 *
 *   warp 0      TMA producer: per K block the A box + the matching [NS x 32] slices of W_hi and W_lo (weights stream
 *               through the same mbarrier ring; they are re-read from L2 by every tile, which is what lets any K x N fit)
 *   warp 1      MMA issuer: tcgen05.mma.kind::tf32 M=128 N=NS K=8, accumulators double buffered in tensor memory
 *   warps 2-9   3xTF32 split of each landed A block (hi in place, lo -> tensor memory, exactly as pw_tc.cu), then the
 *               epilogue of the previous tile: tcgen05.ld -> act(fma(acc, s, b)) -> 128-bit stores, clipped at the tensor edges
 *
 * Numerics: the rounded-split 3xTF32 of pw_tc.cu (fp32-equivalent; DESIGN.md "Numerics of the tensor-core paths").
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "conv_tc.h"
#include "pw_tc.h"
#include "ffb_internal.h"
#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int IG_BM = 128;
constexpr int IG_A_BYTES = IG_BM * 128;            /* one [128 px x 32 ch] fp32 A block */
constexpr int IG_EPI = 256;                        /* split + epilogue threads (8 warps: two per row) */
constexpr int IG_THREADS = 64 + IG_EPI;

struct IgArgs {
    int n, OH, OW, ic, fn, fs, S, P, ldo, coff, act;
    int TW, TH, TN, tiles_x, tiles_y, tiles_n, ntiles_m, nsl, NS, NP;
    int Kc, KB, nstages;                           /* 32-channel chunks per tap, K blocks = taps * Kc, ring slots */
    uint32_t tmem_cols, slot_bytes;
    float *out; const float *scale, *bias;         /* [NP] zero padded */
};

__device__ __forceinline__ float ig_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float ig_act(float v, float slope) { return v > 0.f ? v : v * slope; }

__global__ void __launch_bounds__(IG_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const IgArgs a)
{
    extern __shared__ uint8_t ig_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(ig_raw) + 1023) & ~uintptr_t(1023));
    const int S = a.nstages, NS = a.NS;
    const uint32_t b_bytes = (uint32_t)NS * 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * a.slot_bytes);
    uint64_t *full = bars, *empty = bars + S, *conv = bars + 2 * S, *tfull = bars + 3 * S, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(conv + s, IG_EPI); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, IG_EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_trigger(); pdl_wait();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t alo_col0 = 2 * NS;
    const long nwork = (long)a.ntiles_m * a.nsl;

    /* work item -> (m tile, N slice); m tile -> (frame group, tile row, tile column) */
    auto tile_origin = [&](long w, int &n0, int &oy0, int &ox0, int &slice) {
        slice = (int)(w % a.nsl); long mt = w / a.nsl;
        const int tx = (int)(mt % a.tiles_x); mt /= a.tiles_x;
        const int ty = (int)(mt % a.tiles_y); n0 = (int)(mt / a.tiles_y) * a.TN;
        oy0 = ty * a.TH; ox0 = tx * a.TW;
    };

    if (warp == 0) {
        /* ===================== TMA producer ===================== */
        if (elect_one()) {
            uint32_t q = 0;
            for (long w = blockIdx.x; w < nwork; w += gridDim.x) {
                int n0, oy0, ox0, slice; tile_origin(w, n0, oy0, ox0, slice);
                for (int kb = 0; kb < a.KB; kb++, q++) {
                    const int s = q % S; const uint32_t ph = (q / S) & 1;
                    const int tap = kb / a.Kc, kc = kb - tap * a.Kc, ky = tap / a.fs, kx = tap - ky * a.fs;
                    uint8_t *slot = smem + (size_t)s * a.slot_bytes;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_arrive_expect_tx(full + s, (uint32_t)IG_A_BYTES + 2 * b_bytes);
                    tma_load_4d(slot, &tmA, kc * 32, ox0 * a.S - a.P + kx, oy0 * a.S - a.P + ky, n0, full + s);
                    tma_load_2d(slot + IG_A_BYTES, &tmBh, kc * 32, tap * a.NP + slice * NS, full + s);
                    tma_load_2d(slot + IG_A_BYTES + b_bytes, &tmBl, kc * 32, tap * a.NP + slice * NS, full + s);
                }
            }
        }
    } else if (warp == 1) {
        /* ===================== MMA issuer ===================== */
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_tf32(IG_BM, NS);
            uint32_t q = 0; int it = 0;
            for (long w = blockIdx.x; w < nwork; w += gridDim.x, it++) {
                const int ab = it & 1; const uint32_t aph = (it >> 1) & 1;
                const uint32_t d = tmem_base + ab * NS;
                uint32_t accum = 0;
                for (int kb = 0; kb < a.KB; kb++, q++) {
                    const int s = q % S; const uint32_t ph = (q / S) & 1;
                    const int kc = kb % a.Kc;
                    mbar_wait(conv + s, ph);
                    if (kb == 0) mbar_wait(tempty + ab, aph ^ 1);
                    tc_fence_after_sync();
                    const uint32_t a_base = smem_u32(smem + (size_t)s * a.slot_bytes), bh = a_base + IG_A_BYTES, bl = bh + b_bytes;
                    const uint32_t alo = tmem_base + alo_col0 + s * 32;
                    const int nk = min(4, (a.ic - kc * 32 + 7) / 8);          /* k-steps (of 8 channels) that carry data in this block */
                    /* descriptors once per K block; a k-step advances the start-address field (bytes >> 4) by 32 B = 2 */
                    const uint64_t da = umma_desc_sw128(a_base), dbh = umma_desc_sw128(bh), dbl = umma_desc_sw128(bl);
                    if (nk == 4) {
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) mma_tf32_ts(d, alo + kk * 8, dbh + 2 * kk, idesc, kk ? 1u : accum);     /* A_lo . W_hi */
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) mma_tf32_ss(d, da + 2 * kk, dbl + 2 * kk, idesc, 1);                   /* A_hi . W_lo */
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) mma_tf32_ss(d, da + 2 * kk, dbh + 2 * kk, idesc, 1);                   /* A_hi . W_hi */
                    } else {
                        for (int kk = 0; kk < nk; kk++) mma_tf32_ts(d, alo + kk * 8, dbh + 2 * kk, idesc, kk ? 1u : accum);
                        for (int kk = 0; kk < nk; kk++) mma_tf32_ss(d, da + 2 * kk, dbl + 2 * kk, idesc, 1);
                        for (int kk = 0; kk < nk; kk++) mma_tf32_ss(d, da + 2 * kk, dbh + 2 * kk, idesc, 1);
                    }
                    accum = 1;
                    tc_commit(empty + s);
                }
                tc_commit(tfull + ab);
            }
        }
    } else {
        /* ===================== split + epilogue warps ===================== */
        const int qd = warp & 3, half = ((warp - 2) >> 2) & 1;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const float slope = a.act == 2 ? 0.1f : a.act == 1 ? 0.f : 1.f;

        auto epilogue = [&](long w, int it) {
            const int ab = it & 1; const uint32_t aph = (it >> 1) & 1;
            int n0, oy0, ox0, slice; tile_origin(w, n0, oy0, ox0, slice);
            const int dn = row / (a.TH * a.TW), rem = row - dn * a.TH * a.TW, dy = rem / a.TW, dx = rem - dy * a.TW;
            const bool valid = n0 + dn < a.n && oy0 + dy < a.OH && ox0 + dx < a.OW;
            float *op = a.out + (((long)(n0 + dn) * a.OH + oy0 + dy) * a.OW + ox0 + dx) * a.ldo + a.coff;
            mbar_wait(tfull + ab, aph);
            tc_fence_after_sync();
            for (int j = half; j < NS / 16; j += 2) {               /* 16-column blocks alternate between the two warps of a row */
                uint32_t r[16];
                tmem_ld16(tmem_base + lane_addr + ab * NS + j * 16, r);
                tmem_ld_wait();
                const int c0 = slice * NS + j * 16;
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int co = c0 + 4 * k;
                        if (co < a.fn) {                            /* columns >= fn inside the last group of 4 are the tensor's zero pad lanes */
                            const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.scale + co)), bi = __ldg(reinterpret_cast<const float4 *>(a.bias + co));
                            float4 v;
                            v.x = ig_act(fmaf(__uint_as_float(r[4 * k + 0]), sc.x, bi.x), slope);
                            v.y = ig_act(fmaf(__uint_as_float(r[4 * k + 1]), sc.y, bi.y), slope);
                            v.z = ig_act(fmaf(__uint_as_float(r[4 * k + 2]), sc.z, bi.z), slope);
                            v.w = ig_act(fmaf(__uint_as_float(r[4 * k + 3]), sc.w, bi.w), slope);
                            *reinterpret_cast<float4 *>(op + co) = v;
                        }
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(tempty + ab);
        };

        int it = 0, prev_it = -1; long prev_w = -1;
        for (long w = blockIdx.x; w < nwork; w += gridDim.x, it++) {
            for (int kb = 0; kb < a.KB; kb++) {
                const uint32_t q = (uint32_t)it * a.KB + kb;
                const int s = q % S; const uint32_t ph = (q / S) & 1;
                const int kc = kb % a.Kc;
                mbar_wait(full + s, ph);
                const uint32_t arow = smem_u32(smem + (size_t)s * a.slot_bytes + row * 128);
                const uint32_t alo = tmem_base + lane_addr + alo_col0 + s * 32;
                const int nu = min(4, (a.ic - kc * 32 + 7) / 8);
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int c2 = half + 2 * i;                     /* this warp's 8-channel units of the row */
                    if (c2 < nu) {
                        const uint32_t p0 = arow + (((2 * c2) ^ (row & 7)) << 4), p1 = arow + (((2 * c2 + 1) ^ (row & 7)) << 4);
                        const float4 x0 = lds128(p0), x1 = lds128(p1);
                        float4 h0, h1; uint32_t lo[8];
                        h0.x = ig_round(x0.x); h0.y = ig_round(x0.y); h0.z = ig_round(x0.z); h0.w = ig_round(x0.w);
                        h1.x = ig_round(x1.x); h1.y = ig_round(x1.y); h1.z = ig_round(x1.z); h1.w = ig_round(x1.w);
                        lo[0] = __float_as_uint(x0.x - h0.x); lo[1] = __float_as_uint(x0.y - h0.y); lo[2] = __float_as_uint(x0.z - h0.z); lo[3] = __float_as_uint(x0.w - h0.w);
                        lo[4] = __float_as_uint(x1.x - h1.x); lo[5] = __float_as_uint(x1.y - h1.y); lo[6] = __float_as_uint(x1.z - h1.z); lo[7] = __float_as_uint(x1.w - h1.w);
                        sts128(p0, h0); sts128(p1, h1);
                        tmem_st8(alo + c2 * 8, lo);
                    }
                }
                fence_proxy_async_smem();
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(conv + s);
            }
            if (prev_w >= 0) epilogue(prev_w, prev_it);
            prev_w = w; prev_it = it;
        }
        if (prev_w >= 0) epilogue(prev_w, prev_it);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem_base, a.tmem_cols); }
}

/* packed reference rows (ffcnn.c:218-234; within a row: ci-major, then ky, kx) -> tap-major tf32 hi / lo matrices
 * [taps][NP rows][icld floats], zero padded */
__global__ void k_ig_split_weights(const float *__restrict__ flt, int row, int fn, int ic, int fs, int NP, int icld,
                                   float *__restrict__ hi, float *__restrict__ lo, float *__restrict__ sc, float *__restrict__ bi)
{
    const int taps = fs * fs;
    const long total = (long)taps * NP * icld;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total + NP; i += (long)gridDim.x * blockDim.x) {
        if (i >= total) {
            const int o = (int)(i - total);
            sc[o] = o < fn ? flt[(long)o * row + row - 4] : 0.f;
            bi[o] = o < fn ? flt[(long)o * row + row - 3] : 0.f;
            continue;
        }
        const int ci = (int)(i % icld); long r = i / icld;
        const int co = (int)(r % NP), tap = (int)(r / NP);
        const float w = (co < fn && ci < ic) ? flt[(long)co * row + (long)ci * taps + tap] : 0.f;
        const float h = ig_round(w);
        hi[i] = h; lo[i] = ig_round(w - h);
    }
}

} // namespace

struct IgPlan {
    int ic, fn, fs, S, P, act;
    int nsl, NS, NP, Kc, KB, icld, nstages, num_sms;
    uint32_t tmem_cols, slot_bytes; size_t smem;
    float *d_hi, *d_lo, *d_scb;
    CUtensorMap tmBh, tmBl;
};

IgPlan *ig_plan_create(int ic, int fn, int fs, int stride, int pad, int groups, int act)
{
    if (groups != 1 || ic < 1 || fn < 8 || fs < 1 || fs > 11 || stride < 1 || stride > 8 || pad < 0) return nullptr;
    if ((long)fs * fs * ic < 32) return nullptr;                     /* tiny contractions: the generic kernel is as good */
    static const int off = getenv("FFCNN_NO_IGEMM") ? atoi(getenv("FFCNN_NO_IGEMM")) : 0;
    if (off) return nullptr;
    IgPlan *p = new IgPlan(); memset(p, 0, sizeof *p);
    p->ic = ic; p->fn = fn; p->fs = fs; p->S = stride; p->P = pad; p->act = act;
    const int N16 = (fn + 15) & ~15;
    p->nsl = (N16 + 127) / 128;                                       /* N slices of <= 128 filters: A block + W_hi + W_lo slices = 48 KB per ring slot */
    p->NS = ((N16 + p->nsl - 1) / p->nsl + 15) & ~15;
    p->NP = p->nsl * p->NS;
    p->Kc = (ic + 31) / 32; p->KB = fs * fs * p->Kc; p->icld = (ic + 3) & ~3;
    p->slot_bytes = (uint32_t)(IG_A_BYTES + 2 * p->NS * 128);
    p->nstages = std::max(2, std::min(6, (int)((220 * 1024) / p->slot_bytes)));
    p->smem = (size_t)p->nstages * p->slot_bytes + (3 * p->nstages + 4) * 8 + 16 + 1024;
    const int tmem = 2 * p->NS + p->nstages * 32;
    uint32_t c = 32; while ((int)c < tmem) c <<= 1;
    p->tmem_cols = c;
    if (c > 512 || p->smem > 227 * 1024) { delete p; return nullptr; }
    p->num_sms = ffb_num_sms();
    return p;
}

void ig_plan_destroy(IgPlan *p)
{
    if (!p) return;
    cudaFree(p->d_hi); cudaFree(p->d_lo); cudaFree(p->d_scb);
    delete p;
}

int ig_prepare(IgPlan *p, const float *d_packed, int row, cudaStream_t st)
{
    const size_t n = (size_t)p->fs * p->fs * p->NP * p->icld;
    if (!p->d_hi && (cudaMalloc(&p->d_hi, n * sizeof(float)) != cudaSuccess || cudaMalloc(&p->d_lo, n * sizeof(float)) != cudaSuccess ||
                     cudaMalloc(&p->d_scb, 2 * (size_t)p->NP * sizeof(float)) != cudaSuccess)) { ffb_set_error("conv_tc: cudaMalloc failed"); return -1; }
    k_ig_split_weights<<<(int)std::min<size_t>((n + p->NP + 255) / 256, 4096), 256, 0, st>>>(d_packed, row, p->fn, p->ic, p->fs, p->NP, p->icld, p->d_hi, p->d_lo, p->d_scb, p->d_scb + p->NP);
    if (cudaGetLastError() != cudaSuccess) { ffb_set_error("conv_tc: weight preparation launch failed"); return -1; }
    /* [taps * NP rows][ic] K-major, box = 32 channels x NS filters, 128-byte swizzle, OOB channels read as zero */
    const unsigned long long dims[2] = { (unsigned long long)p->ic, (unsigned long long)p->fs * p->fs * p->NP };
    const unsigned long long strides[1] = { (unsigned long long)p->icld * 4 };
    const unsigned box[2] = { 32u, (unsigned)p->NS };
    if (ffb_make_tensor_map_ex(&p->tmBh, p->d_hi, 2, dims, strides, box, nullptr, 1) != 0) return -1;
    if (ffb_make_tensor_map_ex(&p->tmBl, p->d_lo, 2, dims, strides, box, nullptr, 1) != 0) return -1;
    static ffb_smem_cfg cfg;
    return ffb_ensure_smem((const void *)k_conv_tc, 227 * 1024, &cfg);
}

bool ig_supports(const IgPlan *p, int ldi, int ldo, int coff, int ih, int iw)
{
    if (!p || ldi % 4 || ldo % 4 || coff % 4) return false;
    if (p->fn % 4 && (coff != 0 || ldo != ((p->fn + 3) & ~3))) return false;     /* the epilogue writes whole groups of 4 channels (pad lanes get zeros) */
    const int oh = (ih - p->fs + 2 * p->P) / p->S + 1, ow = (iw - p->fs + 2 * p->P) / p->S + 1;
    return oh >= 1 && ow >= 1;
}

int ig_run(IgPlan *p, const float *in, int ldi, float *out, int ldo, int coff, int n, int ih, int iw, cudaStream_t st)
{
    IgArgs a;
    a.n = n; a.OH = (ih - p->fs + 2 * p->P) / p->S + 1; a.OW = (iw - p->fs + 2 * p->P) / p->S + 1;
    a.ic = p->ic; a.fn = p->fn; a.fs = p->fs; a.S = p->S; a.P = p->P; a.ldo = ldo; a.coff = coff; a.act = p->act;
    /* m tile = TW x TH x TN output pixels (powers of two, product 128): the shape that wastes the fewest rows */
    double best = -1; a.TW = 128; a.TH = 1; a.TN = 1;
    for (int tw = 1; tw <= 128; tw *= 2)
        for (int th = 1; tw * th <= 128; th *= 2) {
            const int tn = 128 / (tw * th);
            if (tw * p->S > 256 || th * p->S > 256) continue;
            const double eff = (double)a.OW * a.OH * n / ((double)((a.OW + tw - 1) / tw * tw) * ((a.OH + th - 1) / th * th) * ((n + tn - 1) / tn * tn));
            const double score = eff + (tn == 1 ? 1e-3 : 0) + 1e-4 * tw / 128.0;     /* ties: whole frames, then wide rows */
            if (score > best) { best = score; a.TW = tw; a.TH = th; a.TN = tn; }
        }
    a.tiles_x = (a.OW + a.TW - 1) / a.TW; a.tiles_y = (a.OH + a.TH - 1) / a.TH; a.tiles_n = (n + a.TN - 1) / a.TN;
    a.ntiles_m = a.tiles_x * a.tiles_y * a.tiles_n;
    a.nsl = p->nsl; a.NS = p->NS; a.NP = p->NP; a.Kc = p->Kc; a.KB = p->KB; a.nstages = p->nstages;
    a.tmem_cols = p->tmem_cols; a.slot_bytes = p->slot_bytes;
    a.out = out; a.scale = p->d_scb; a.bias = p->d_scb + p->NP;
    CUtensorMap tmA;
    const unsigned long long dims[4] = { (unsigned long long)p->ic, (unsigned long long)iw, (unsigned long long)ih, (unsigned long long)n };
    const unsigned long long strides[3] = { (unsigned long long)ldi * 4, (unsigned long long)iw * ldi * 4, (unsigned long long)ih * iw * ldi * 4 };
    const unsigned box[4] = { 32u, (unsigned)(a.TW * p->S), (unsigned)(a.TH * p->S), (unsigned)a.TN };
    const unsigned estr[4] = { 1u, (unsigned)p->S, (unsigned)p->S, 1u };
    if (ffb_make_tensor_map_ex(&tmA, in, 4, dims, strides, box, estr, 1) != 0) return -1;
    const long nwork = (long)a.ntiles_m * a.nsl;
    const int grid = (int)std::max<long>(1, std::min<long>(nwork, p->num_sms));
    cudaError_t e = launch_pdl(k_conv_tc, dim3(grid), dim3(IG_THREADS), p->smem, st, tmA, p->tmBh, p->tmBl, a);
    if (e != cudaSuccess) { ffb_set_error("conv_tc launch failed: %s (grid %d smem %zu)", cudaGetErrorString(e), grid, p->smem); return -1; }
    return 0;
}
