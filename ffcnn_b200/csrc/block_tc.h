/* block_tc.h -- host interface of the all-tcgen05 fused inverted-residual block kernel (block_tc.cuh / block_tc.cu). */
#pragma once
#include <cuda_runtime.h>

struct Blk2Plan;

/* Same contract as blk_plan_create (block_mma.h): 1x1 expand (cin -> cexp) -> 3x3 depthwise stride `stride` pad 1 -> 1x1 project
 * (cexp -> cout) [+ x].  Returns NULL when the shape has no instantiated kernel or no tile fits (the caller falls back). */
Blk2Plan *blk2_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res);
void      blk2_plan_destroy(Blk2Plan *p);
int       blk2_prepare(Blk2Plan *p, const float *p1, const float *pd, const float *p3, cudaStream_t st);
int       blk2_run(Blk2Plan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st);
const char *blk2_describe(const Blk2Plan *p);
