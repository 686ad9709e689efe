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

// cnn_train.cu: second stage of the BatchNorm statistics (fold of [parts, 2, C] partials -> mean / rstd, scale / shift,
// running statistics); also launched by the GEMM / convolution when the statistics come out of its epilogue (pose_bn_fuse)
int launch_bn_finalize_parts(const float *partials, int parts, long count, const float *gamma, const float *beta, float eps,
                             float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean,
                             float *running_var, cudaStream_t stream);

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

// Counter-based dropout mask shared by every kernel that applies or re-applies dropout: element `i` of the tensor that
// call `seed` drops is kept iff lowbias32(lo32(i) ^ key(seed, hi32(i))) >= p * 2^32, key = a 64-bit murmur finaliser of
// (seed, hi32(i)).  The key of hi32 == 0 (every tensor of this model) is computed once on the host (`DropSeed`), so the
// per-element cost is one 32-bit hash (2 multiplies) instead of a 64-bit one.  Backward regenerates the same mask.
__host__ __device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (uint32_t)x;
}
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
struct DropSeed {
    unsigned long long s;      // seed * golden ratio
    uint32_t key0;             // mix32(s): the key of elements below 2^32
    const uint32_t *step_key;  // device word XORed into every key, or null (pose_step_state_bind: the per-step part of the
                               // seed lives in device memory so that a captured CUDA graph of the step can be replayed)
};
// abi.cu: the bound per-step state (device pointer) or null
extern pose_step_state *g_step_state;
inline DropSeed make_drop_seed(unsigned long long seed) {
    DropSeed d;
    d.s = seed * 0x9E3779B97F4A7C15ULL;
    d.key0 = mix32(d.s);
    d.step_key = g_step_state ? &g_step_state->drop_key : nullptr;
    return d;
}
// key of the elements below 2^32 for this launch (one load per thread when a step state is bound)
__device__ __forceinline__ uint32_t drop_key0(const DropSeed &d) { return d.step_key ? (d.key0 ^ __ldg(d.step_key)) : d.key0; }
__device__ __forceinline__ uint32_t drop_key(const DropSeed &d, uint32_t hi) { return hi ? (mix32(d.s + hi) ^ (drop_key0(d) ^ d.key0)) : drop_key0(d); }
// element `lo` (< 2^32) under a key: branch-free, 2 multiplies -- the form the fused kernels use (their hosts reject
// tensors of 2^32 elements or more, so the key is always DropSeed::key0)
__device__ __forceinline__ bool drop_keep32(uint32_t key, uint32_t lo, uint32_t thresh) { return lowbias32(lo ^ key) >= thresh; }
__device__ __forceinline__ bool drop_keep(const DropSeed &d, uint64_t i, uint32_t thresh) {
    return drop_keep32(drop_key(d, (uint32_t)(i >> 32)), (uint32_t)i, thresh);
}
inline uint32_t drop_threshold(float p) { return (uint32_t)((double)p * 4294967296.0); }

// 16-byte asynchronous global -> shared copy (LDGSTS); `valid == false` writes zeros (padding) without touching memory
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 128-bit read-only global load that the compiler may not sink towards its first use: the row-streaming kernels issue a
// batch of these before touching any of the data (NVVM otherwise moves every __ldg next to its consumer, which leaves two
// 16-byte requests in flight per thread -- measured 4.4 TB/s on the BatchNorm passes)
__device__ __forceinline__ uint4 ldg_batch(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// streaming 128-bit store: outputs are written once and not re-read by the producing kernel
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace pose
