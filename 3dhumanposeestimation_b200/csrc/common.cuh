// common.cuh -- shared helpers for the sm_100a kernels behind include/pose_b200.h
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/pose_b200.h"

#define POSE_API extern "C" __attribute__((visibility("default")))

namespace pose {

constexpr int kNumSMs = 148;  // B200

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? POSE_OK : (int)e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit store: outputs are written once and not re-read by the producing kernel
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace pose
