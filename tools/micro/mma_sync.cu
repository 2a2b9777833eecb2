// microbenchmark: legacy mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16) next to FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 mma_sync.cu -o mma_sync
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE, int ACC> __global__ void k(float *out, int iters)
{
    float d[ACC][4]; unsigned a[4], b[2];
    for (int i = 0; i < ACC; i++) for (int j = 0; j < 4; j++) d[i][j] = 0.f;
    for (int j = 0; j < 4; j++) a[j] = 0x3f800000u + threadIdx.x + j;
    b[0] = 0x3f000000u + threadIdx.x; b[1] = 0x3e800000u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) { if (MODE == 0) mma_tf32(d[i], a, b); else mma_bf16(d[i], a, b); }
    }
    float acc = 0;
    for (int i = 0; i < ACC; i++) for (int j = 0; j < 4; j++) acc += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++)
        for (int warps = 4; warps <= 16; warps *= 2)
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0, 8><<<148, warps * 32>>>(d, iters); else k<1, 8><<<148, warps * 32>>>(d, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double flop = 148.0 * warps * 8.0 * iters * (mode == 0 ? 2048.0 : 4096.0);
                if (rep) printf("%s warps/SM=%2d: %.3f ms  %.1f TFLOP/s dense  (%.0f flop/clk/SM at 1.965 GHz)\n", mode ? "mma.sync bf16 m16n8k16" : "mma.sync tf32 m16n8k8 ",
                                warps, ms, flop / ms / 1e9, flop / ms / 1e9 * 1e12 / 148 / 1.965e9 / 1e3 * 1e-0);
            }
    return 0;
}
