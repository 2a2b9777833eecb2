/* conv_tc.h -- implicit-GEMM convolution on tcgen05 (conv_tc.cu): dense k x k convs (conv-v6.c:9-42, conv-v2.c:7-87) and
 * pointwise layers whose weights do not fit pw_tc.cu's resident-weight plan. */
#pragma once
#include <cuda_runtime.h>

struct IgPlan;
/* NULL when the shape is not supported (groups != 1, ic % 4, ...): the caller keeps its generic kernel */
IgPlan *ig_plan_create(int ic, int fn, int fs, int stride, int pad, int groups, int act);
void    ig_plan_destroy(IgPlan *p);
/* split the packed reference rows (ffcnn.c:218-234) into the tap-major tf32 hi/lo weight matrices */
int     ig_prepare(IgPlan *p, const float *d_packed, int row, cudaStream_t st);
/* run-time eligibility for a given tensor geometry (pixel strides must be multiples of 4 floats, coff of 4) */
bool    ig_supports(const IgPlan *p, int ldi, int ldo, int coff, int ih, int iw);
int     ig_run(IgPlan *p, const float *in, int ldi, float *out, int ldo, int coff, int n, int ih, int iw, cudaStream_t st);
