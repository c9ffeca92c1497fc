// kernels_common.cuh -- block-level reduction shared by the kernel groups
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

__device__ __forceinline__ uint32_t block_reduce_sum(uint32_t v, uint32_t* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0u;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xFFFFFFFFu, t, o);
    }
    return t;  // valid in thread 0
}
