// rowpipe.cuh -- thread-private asynchronous row pipeline of the bandwidth-bound row-streaming kernels (BatchNorm passes,
// gate / pooling backward, LayerNorm): see row_stream below.
#pragma once
#include "common.cuh"

namespace pose {

// ---- thread-private asynchronous row pipeline ----------------------------------------------------------------------
// The row-streaming kernels below keep 32-48 per-channel values in registers, which leaves ptxas room for only TWO 16-byte
// loads in flight per thread (it sinks every further load next to its use: 25 KB in flight per SM = 4.4 TB/s measured on
// all BatchNorm passes).  Here every thread copies its own 16-byte chunks of the next kRowStages - 1 rows into its own
// shared-memory slots with cp.async (LDGSTS: no registers, no block synchronisation because nobody else reads the slot)
// and consumes them with one LDS.128 per tensor: bytes in flight no longer depend on the register file.
// Slot layout: [stage][tensor][thread] uint4 -> conflict-free 128-bit accesses.
#ifndef POSE_ROW_STAGES
#define POSE_ROW_STAGES 5
#endif
constexpr int kRowStages = POSE_ROW_STAGES;
static_assert(kRowStages >= 4, "the drained ring also holds the 64 B / thread scratch of rowmap_fold<2>");
extern __shared__ __align__(16) unsigned char g_rowpipe[];
static inline size_t rowpipe_bytes(int tensors, int threads) { return (size_t)kRowStages * tensors * 16 * threads; }
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// Streams rows r0, r0 + step, ... < M of NT row-major bf16 tensors (byte pitches `pitch`, this thread's chunk at byte
// offset `coff` of a row) through the ring and calls body(r, v[NT]) for each.
template <int NT, class F>
__device__ __forceinline__ void row_stream(const void *const (&base)[NT], const long (&pitch)[NT], long coff, long r0, long step,
                                           long M, F &&body) {
    const uint32_t tstride = blockDim.x * 16u, sstride = tstride * NT;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(g_rowpipe) + threadIdx.x * 16u;
    const int n = r0 < M ? (int)((M - r0 + step - 1) / step) : 0;
    const char *src[NT];
    long inc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        src[t] = (const char *)base[t] + r0 * pitch[t] + coff;
        inc[t] = step * pitch[t];
    }
#pragma unroll
    for (int s = 0; s < kRowStages - 1; ++s) {
        if (s < n) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                cp16(s0 + s * sstride + t * tstride, src[t]);
                src[t] += inc[t];
            }
        }
        cp_async_commit();
    }
    uint32_t rd = s0, wr = s0 + (kRowStages - 1) * sstride;
    const uint32_t end = s0 + kRowStages * sstride;
    long r = r0;
    for (int k = 0; k < n; ++k, r += step) {
        if (k + (kRowStages - 1) < n) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                cp16(wr + t * tstride, src[t]);
                src[t] += inc[t];
            }
        }
        cp_async_commit();
        cp_async_wait<kRowStages - 1>();
        uint4 v[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) v[t] = lds16(rd + t * tstride);
        body(r, v);
        rd += sstride; if (rd == end) rd = s0;
        wr += sstride; if (wr == end) wr = s0;
    }
}
// bf16 pair -> fp32 pair in two instructions (shift, mask) instead of the three of __bfloat1622float2's PRMT + shifts
__device__ __forceinline__ void up8q(const uint4 &p, float2 (&f)[4]) {
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u));
}

// the row-pipeline kernels take up to 5 stages x 2 tensors x 16 B x 384 threads = 60 KB of dynamic shared memory
template <class K>
static inline void rowpipe_optin(K kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowpipe_bytes(2, 384));
}

}  // namespace pose
