/* stem_common.cuh -- the pieces kernels.cuh (engine.cu) and stem_block.cuh (block_reg.cu) share: the activation of the reference
 * (utils.h:15-23) and the stem's kernel-parameter block. */
#pragma once
#include <cuda_runtime.h>

namespace ffb {

__device__ __forceinline__ float act_apply(float v, int act)
{
    return act == 2 ? (v > 0.f ? v : 0.1f * v) : act == 1 ? fmaxf(v, 0.f) : v;
}

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

struct alignas(16) StemW { float w[27 * 8]; float s[8]; float b[8]; };      /* w[(c*3+ky)*3+kx][oc] */

} // namespace ffb
