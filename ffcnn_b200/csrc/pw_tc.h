/*
 * pw_tc.h -- interface of the tcgen05 (5th-gen tensor core) pointwise-conv kernel, pw_tc.cu.
 * mode: 0 auto (3xTF32 where the layer is above the FFMA ridge), 2 force 3xTF32, 3 force 1xTF32.
 * pw_tc_plan_create returns NULL when the shape is not eligible (caller falls back to the FFMA kernel).
 */
#pragma once
#include <cuda_runtime.h>

struct PwTcPlan;
PwTcPlan   *pw_tc_plan_create(int K, int N, int act, int mode);
void        pw_tc_plan_destroy(PwTcPlan *p);
int         pw_tc_prepare(PwTcPlan *p, const float *d_packed, int row, cudaStream_t st);
/* res != NULL fuses the following shortcut layer: out = act2(conv(in) + res[m][0..N)), res row stride ldr floats */
int         pw_tc_run(PwTcPlan *p, const float *in, int ldi, float *out, int ldo, int coff, long M, cudaStream_t st,
                      const float *res, int ldr, int act2);
const char *pw_tc_mode_name(const PwTcPlan *p);
bool        pw_tc_supports(const PwTcPlan *p, int ldo, int coff);    /* false: this output geometry needs the FFMA kernel */

/* generic tiled tensor map over fp32 data: dims/box innermost first, strides_bytes[i] = byte stride of dimension i+1.
 * swizzle128 != 0 selects CU_TENSOR_MAP_SWIZZLE_128B.  Out-of-bounds elements read as zero. Returns 0 on success. */
#include <cuda.h>
int ffb_make_tensor_map(CUtensorMap *m, const void *base, int rank, const unsigned long long *dims,
                        const unsigned long long *strides_bytes, const unsigned *box, int swizzle128);
/* same with element strides (NULL = all 1): along a dimension with element stride s, box[i] = s * (elements wanted) */
int ffb_make_tensor_map_ex(CUtensorMap *m, const void *base, int rank, const unsigned long long *dims,
                           const unsigned long long *strides_bytes, const unsigned *box, const unsigned *elem_strides, int swizzle128);
