/*
 * sm100.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) machinery the kernels use:
 * mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st / fences) and the
 * shared-memory / instruction descriptors of the 5th-gen tensor core.  Bit layouts follow the PTX ISA
 * (cross-checked against the descriptor unions in CUTLASS' cute/arch/mma_sm100_desc.hpp).
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

/* explicit shared-space 128-bit accesses on 32-bit shared addresses: pointers carved out of the dynamic smem blob lose
 * their address space and nvcc falls back to generic LD/ST (slower, and ordered against every other generic access) */
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

/* ---------------------------------------------------------------- packed fp32 pairs (FFMA2 / FMUL2 / FADD2)
 * sm_100 executes fma/mul/add.rn.f32x2 on a 64-bit register pair: two independent IEEE fp32 operations (bit-identical to
 * fmaf / * / +) in ONE issue slot.  The FMA pipe's peak is unchanged (tools/micro/ffma2.cu: 70.8 vs 73.2 TFLOP/s); what
 * they buy is issue slots, which is what the CUDA-core stages of the fused kernels run out of.  Not volatile: the compiler
 * may schedule, CSE and fold them; ptxas turns f2_pack(s, s) into a broadcast operand and constant-bank pairs into
 * uniform-register operands. */
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) { return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32); }
__device__ __forceinline__ float f2_lo(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

/* four channels as two packed pairs: (0, 1) and (2, 3) -- what one 128-bit load delivers */
struct f4p { f32x2 a, b; };
__device__ __forceinline__ f4p zero4p() { f4p r; r.a = 0ull; r.b = 0ull; return r; }
__device__ __forceinline__ f4p lds128p(uint32_t addr)
{
    f4p v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.a), "=l"(v.b) : "r"(addr));
    return v;
}
__device__ __forceinline__ f4p ldg128p(const float *p)              /* read-only global data */
{
    f4p v;
    asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(v.a), "=l"(v.b) : "l"(p));
    return v;
}
__device__ __forceinline__ void fma4p(f4p &acc, const f4p v, const f4p w) { acc.a = f2_fma(v.a, w.a, acc.a); acc.b = f2_fma(v.b, w.b, acc.b); }

/* ---------------------------------------------------------------- mbarrier */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}

/* ---------------------------------------------------------------- proxies / fences */
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

/* ---------------------------------------------------------------- TMA */
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m)
{
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
/* 2D tile load global -> shared, completion on an mbarrier (tx bytes) */
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *m, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *m, int x, int y, int z, int w, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar)) : "memory");
}
/* 2D tile store shared -> global (bulk async group) */
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *smem_src, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 :: "l"(reinterpret_cast<uint64_t>(m)), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_all()  { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

/* ---------------------------------------------------------------- TMEM */
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols)          /* whole warp; ncols power of 2 >= 32 */
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)               /* whole warp (the allocating one) */
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
/* make the mbarrier track completion of all tcgen05 ops issued so far by this thread (implies fence::before_thread_sync) */
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

/* D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate */
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
/* D[tmem] (+)= A[tmem] * B[smem desc] */
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

/* this thread's TMEM lane, 16 / 32 consecutive columns -> registers (warp-collective; lane quarter = warp_id % 4) */
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

/* ---------------------------------------------------------------- programmatic dependent launch (PDL) */
/* Every kernel of the forward pass is launched with programmatic stream serialization: its CTAs may start while the previous
 * kernel drains.  pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as resources free up; pdl_wait() blocks
 * until the PREVIOUS kernel has completed and flushed its writes -- it must precede the first access to activations. */
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait()    { asm volatile("griddepcontrol.wait;" ::: "memory"); }

/* ---------------------------------------------------------------- descriptors */
/* K-major operand tile in SWIZZLE_128B layout: rows of 128 bytes (32 tf32), 8-row groups 1024 B apart.
 * bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
 * [32,46) stride byte offset >> 4 (= 1024 >> 4), [46,48) version = 1, [61,64) layout type = 2 (SWIZZLE_128B). */
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
/* instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M x N tile:
 * [4,6) D format = 1 (f32), [7,10) A format = 2 (tf32), [10,13) B format = 2, [15] A major = 0 (K), [16] B major = 0 (K),
 * [17,23) N >> 3, [24,29) M >> 4 */
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

/* host: launch with the PDL attribute (falls back to a plain launch when pdl == 0) */
extern int g_ffb_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_ffb_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

} // namespace sm100
