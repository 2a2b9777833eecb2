/* block_mma.h -- host interface of the fused inverted-residual block kernel (block_mma.cuh / block_mma.cu). */
#pragma once
#include <cuda_runtime.h>

struct BlkPlan;

/* One fused block: 1x1 expand (cin -> cexp) -> 3x3 depthwise stride `stride` pad 1 -> 1x1 project (cexp -> cout) [+ x].
 * h, w: input spatial size.  act*: reference activation codes (utils.h:8-13).  res != 0: add the block input after the
 * projection (ffcnn.c:418-423) and apply act_res; needs stride 1 and cin == cout.
 * Returns NULL when this channel/stride combination has no instantiated kernel (the caller keeps the unfused layers). */
BlkPlan *blk_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res);
void     blk_plan_destroy(BlkPlan *p);
/* p1 / pd / p3: device pointers to the packed reference rows (ffcnn.c:218-234) of the three convs */
int      blk_prepare(BlkPlan *p, const float *p1, const float *pd, const float *p3, cudaStream_t st);
int      blk_run(BlkPlan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st);
const char *blk_describe(const BlkPlan *p);          /* "tile 16x32 gc1 mtw4 smem 93KB occ2" */
int      blk_uses_tcgen05(const BlkPlan *p);         /* the expand GEMM of this plan runs on tcgen05 (TMEM accumulators) */
