// gemm_tcgen05.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA, fused bias + activation epilogue.
//
//   C[M, N] = act(A[M, K] . W[N, K]^T + bias[N])         A, W bf16 row-major (K contiguous)
//
// This is the contraction behind every nn.Linear on the hot path (PoseRegressionHead:
// src/models/common.py:73-81, SE gates: src/models/cnn.py:16-18, ViT qkv/proj/fc1/fc2) and every 1x1
// convolution once activations are channels-last (NHWC makes a 1x1 conv exactly this GEMM with
// M = B*H*W: src/models/cnn.py:122-131).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D loads of A (128 x 64) and W (BN x 64) tiles into
//            a kStages-deep 128B-swizzled shared-memory ring, completion on `full` mbarriers
//   warp 1   TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) four
//            times per stage; tcgen05.commit releases the stage (`empty`) and finally signals `acc_full`
//   warps 2-5  epilogue: tcgen05.ld the fp32 accumulator (each warp owns its 32-lane TMEM quarter),
//            bias + activation, convert, store
// All mbarrier waits are bounded (trap instead of hanging the GPU if a descriptor is wrong).
#include <cuda.h>
#include <mutex>
#include "common.cuh"

namespace pose {

constexpr int kGemmThreads = 192;
constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;

// ------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 26)) __trap();  // never hang the device on a bad descriptor
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (8 rows x 128 B)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D = F32, A = B = BF16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case 1: return v > 0.f ? v : 0.f;                                   // relu
        case 2: return v / (1.0f + __expf(-v));                             // silu
        case 3: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // gelu (erf form, nn.GELU default)
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------ kernel
template <int BN, int kStages>
struct GemmSmem {
    static constexpr int kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = 256;
    static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // + alignment slack
};

template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N,
                    int K, const float *__restrict__ bias, int act, void *__restrict__ C, int ldc, int out_bf16) {
    using S = GemmSmem<BN, kStages>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024 B
    unsigned char *bars = smem + kStages * S::kStageBytes;
    uint64_t *full = (uint64_t *)bars;
    uint64_t *empty = full + kStages;
    uint64_t *acc_full = empty + kStages;
    uint32_t *tmem_slot = (uint32_t *)(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int num_kb = (K + BK - 1) / BK;
    constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;  // power of two >= 32

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(empty + s, ph ^ 1);
                unsigned char *sa = smem + s * S::kStageBytes, *sb = sa + S::kABytes;
                mbar_expect_tx(full + s, S::kStageBytes);
                tma_load_2d(sa, &map_a, full + s, kb * BK, m0);
                tma_load_2d(sb, &map_w, full + s, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (kb / kStages) & 1;
            mbar_wait(full + s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + s * S::kStageBytes), sb = sa + S::kABytes;
                const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sb);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advancing K by 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
                    tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                }
                tc_commit(empty + s);                      // frees the smem stage when these MMAs retire
                if (kb == num_kb - 1) tc_commit(acc_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int row = m0 + quarter * 32 + lane;
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
            const int col0 = n0 + c * 32;
            if (!row_ok || col0 >= N) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(r[j]);
                if (bias != nullptr && col0 + j < N) x += __ldg(bias + col0 + j);
                v[j] = apply_act(x, act);
            }
            if (out_bf16) {
                __nv_bfloat16 *dst = (__nv_bfloat16 *)C + (size_t)row * ldc + col0;
                if (col0 + 32 <= N && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                        uint4 pk = make_uint4(*(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
                        *(uint4 *)(dst + j) = pk;
                    }
                } else {
                    for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = __float2bfloat16_rn(v[j]);
                }
            } else {
                float *dst = (float *)C + (size_t)row * ldc + col0;
                if (col0 + 32 <= N && (((uintptr_t)dst) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *(float4 *)(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = v[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// fp32 -> bf16 (operand preparation for the GEMM); 8 elements per thread
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, long n) {
    const long n8 = n >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const float4 a = __ldg((const float4 *)in + i * 2), b = __ldg((const float4 *)in + i * 2 + 1);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        ((uint4 *)out)[i] = make_uint4(*(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
    }
    for (long i = n8 * 8 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}

// -------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = box_rows x 64 columns, 128 B swizzle, OOB reads give zeros
static int make_map_bf16(CUtensorMap *map, const void *ptr, long rows, long cols, long ld_elems, int box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

template <int BN, int kStages>
static int launch_gemm(const void *A, int lda, const void *W, int ldw, const float *bias, void *C, int ldc, int M,
                       int N, int K, int act, int out_bf16, cudaStream_t s) {
    using S = GemmSmem<BN, kStages>;
    CUtensorMap ma, mw;
    int e = make_map_bf16(&ma, A, M, K, lda, BM);
    if (e) return e;
    e = make_map_bf16(&mw, W, N, K, ldw, BN);
    if (e) return e;
    auto kern = gemm_bf16_tn_kernel<BN, kStages>;
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        if (ce != cudaSuccess) return (int)ce;
        configured = true;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    kern<<<grid, kGemmThreads, S::kTotal, s>>>(ma, mw, M, N, K, bias, act, C, ldc, out_bf16);
    return launch_status();
}

}  // namespace pose

POSE_API int pose_gemm_bf16(const void *A, int lda, const void *W, int ldw, const float *bias, void *C, int ldc,
                            int M, int N, int K, int act, int out_dtype, pose_stream_t stream) {
    using namespace pose;
    if (!A || !W || !C) return POSE_E_NULL;
    if (M <= 0 || N <= 0 || K <= 0) return POSE_E_SHAPE;
    if (lda < K || ldw < K || ldc < N) return POSE_E_SHAPE;
    if (lda % 8 || ldw % 8) return POSE_E_SHAPE;  // TMA: row pitch must be a multiple of 16 bytes
    if ((uintptr_t)A % 16 || (uintptr_t)W % 16) return POSE_E_ALIGN;
    if (act < 0 || act > 3 || (out_dtype != 0 && out_dtype != 1)) return POSE_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (N <= 32) return launch_gemm<32, 8>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, out_dtype, s);
    if (N <= 64) return launch_gemm<64, 8>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, out_dtype, s);
    return launch_gemm<128, 6>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, out_dtype, s);
}

POSE_API int pose_cast_f32_bf16(const float *in, void *out, long n, pose_stream_t stream) {
    using namespace pose;
    if (!in || !out) return POSE_E_NULL;
    if (n <= 0) return POSE_E_SHAPE;
    if ((uintptr_t)in % 16 || (uintptr_t)out % 16) return POSE_E_ALIGN;
    long blocks = (n / 8 + 255) / 256 + 1;
    int grid = (int)(blocks < (long)kNumSMs * 8 ? blocks : (long)kNumSMs * 8);
    cast_f32_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16 *)out, n);
    return launch_status();
}
