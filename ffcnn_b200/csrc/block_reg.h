/* block_reg.h -- host interface of the register-resident fused block kernel (block_reg.cuh / block_reg.cu). */
#pragma once
#include <cuda_runtime.h>

struct RegPlan;
/* Same contract as blk_plan_create (block_mma.h) for the low-channel blocks (expanded width <= 24).  h1 / hd / h3: HOST
 * pointers to the packed reference rows (ffcnn.c:218-234) of the expand, depthwise and projection convs: the weights
 * travel as a kernel parameter.  Returns NULL when the shape has no instantiated kernel. */
RegPlan *reg_plan_create(int cin, int cexp, int cout, int stride, int h, int w, int act1, int actd, int act3, int res, int act_res,
                         const float *h1, const float *hd, const float *h3);
void     reg_plan_destroy(RegPlan *p);
int      reg_run(RegPlan *p, const float *x, int ldx, float *y, int ldy, int n, cudaStream_t st);
const char *reg_describe(const RegPlan *p);
int      reg_launches(const RegPlan *p);          /* kernels one reg_run enqueues (3 for the sliced 24-channel block) */

/* The stem (3x3 s2 conv on the u8 frames, net_input fused) and the 8->8->4 block as one kernel (stem_block.cuh): `stemw` points to the
 * stem's StemW parameter block, y is the block's output tensor.  reg_stem_ok: the plan is that block and the frames qualify. */
int      reg_stem_ok(const RegPlan *p, int ih, int iw, int pitch, const void *frames);
int      reg_run_stem(RegPlan *p, const void *stemw, int act0, const unsigned char *frames, int pitch, float *y, int ldy, int n, int ih, int iw,
                      const float *mean, const float *norm, cudaStream_t st);
