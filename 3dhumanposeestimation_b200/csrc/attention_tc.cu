// attention_tc.cu -- scaled dot-product attention, forward and backward, on the tcgen05 tensor cores with the score
// matrix resident in TMEM (reference: nn.MultiheadAttention in src/models/transformers.py:61-63, :98-106 and the timm
// ViT attention it wraps; sequences of this model: 257 / 273 tokens self-attention, 256 x 16 and 16 x 256 cross-attention;
// head_dim 64 | 48).
//
// Forward (one CTA per (sample, head), 128 threads, query tiles of 128 rows):
//   Q tile, K, V are staged in shared memory in the 128-byte-swizzled UMMA image (one row = one token's head slice); K
//   is the K-major B operand of S = Q K^T and the very same image of V is the MN-major B operand of O = P V.
//   S (128 x Nk fp32) is accumulated in TMEM; each thread owns one query row (tcgen05.ld 32x32b): row max, exp2,
//   row sum with no shuffles; P goes to shared memory as the K-major A operand (bf16); O accumulates in TMEM and is
//   normalised on the way out.  The log-sum-exp of every row is saved for the backward pass.
// Backward (probabilities rebuilt from the saved log-sum-exp, nothing of size Nq x Nk reaches HBM):
//   dq kernel : per (sample, head): for every query tile and key chunk of 128: S, dP = dO V^T in TMEM ->
//               dS = P (dP - D) scale -> shared memory -> dQ += dS K (K image read as an MN-major operand).
//   dkv kernel: per (sample, head, 128-key tile): for every query tile: S^T = K Q^T, dP^T = V dO^T in TMEM -> P^T, dS^T
//               -> shared memory -> dV += P^T dO, dK += dS^T Q (Q / dO images read as MN-major operands).
#include "tc_common.cuh"

namespace pose {

constexpr int kAttnThreads = 256;     // two warps per TMEM lane quarter: they split the column chunks of the elementwise phases
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {      // one MUFU.EX2 (exp2f without -use_fast_math is a multi-instruction sequence)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// rows of 128 B (one token's head slice, zero padded to 64 elements), 16-byte chunks XOR-swizzled by (row & 7):
// the SWIZZLE_128B image both the K-major and the MN-major UMMA descriptors read
template <int HD>
__device__ __forceinline__ void load_rows_sw128(unsigned char *smem, const __nv_bfloat16 *g, long ld, int rows_valid,
                                                int rows_total) {
    for (int i = threadIdx.x; i < rows_total * 8; i += kAttnThreads) {
        const int r = i >> 3, c = i & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows_valid && c < HD / 8) v = __ldg((const uint4 *)(g + (long)r * ld) + c);
        *(uint4 *)(smem + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
}

// 8 consecutive bf16 of row r, columns [col8 * 8, col8 * 8 + 8) of a K-major [128 x 64k] operand made of 16 KB blocks
__device__ __forceinline__ void store_chunk_kmajor(unsigned char *base, int r, int col8, uint4 v) {
    const int blk = col8 >> 3, c = col8 & 7;
    *(uint4 *)(base + blk * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *(uint32_t *)&h;
}

struct AttnSync {
    uint64_t *bar;       // [0]: MMA completion (tcgen05.commit); [1], [2]: TMA operand loads (complete_tx), used alternately
    uint32_t phase, lwait, lissue;
    __device__ __forceinline__ void commit_and_wait() {     // thread 0 commits the MMAs issued so far; everybody waits
        if (threadIdx.x == 0) tc_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
    }
    // Load group k lands on barrier 1 + (k & 1) with parity (k >> 1) & 1.  Two alternating barriers mean that the issuing
    // thread re-arms a barrier only two groups later -- with a __syncthreads in between -- so no thread can still be
    // polling the phase that is being re-armed (with ONE barrier a thread that had not yet observed phase P could miss it
    // once phase P + 1 completed as well: a parity alias, i.e. a hang).
    __device__ __forceinline__ uint64_t *load_bar() { return bar + 1 + (lissue++ & 1); }     // issuing thread only
    __device__ __forceinline__ void wait_loads() {
        mbar_wait(bar + 1 + (lwait & 1), (lwait >> 1) & 1);
        ++lwait;
    }
};

// Operand staging by TMA: every operand is a [B, T, cols] bf16 tensor (pitch ld, batch stride bs); one box = 64 columns
// (one head slice, 128 B: the SWIZZLE_128B row) x 32 tokens.  Rows past T and columns past `cols` are zero filled.
// Called by ONE thread; the bytes land on bar[1].
constexpr int kBoxRows = 32;
__device__ __forceinline__ void tma_rows(unsigned char *smem, const CUtensorMap *map, uint64_t *lbar, int col0, int row0, int b,
                                         int rows_padded) {
    for (int r = 0; r < rows_padded; r += kBoxRows) tma_load_3d(smem + r * 128, map, lbar, col0, row0 + r, b);
}
__device__ __forceinline__ int box_bytes(int rows) { return (rows + kBoxRows - 1) / kBoxRows * kBoxRows * 128; }

template <uint32_t TCOLS = 512>
__device__ __forceinline__ uint32_t attn_prologue(unsigned char *&smem, uint64_t *&bar) {
    extern __shared__ unsigned char smem_dyn[];
    smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t s_bar[3];
    __shared__ uint32_t s_tmem;
    bar = s_bar;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return s_tmem;
}

template <uint32_t TCOLS = 512>
__device__ __forceinline__ void attn_epilogue(uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// One token outside the tiles.  The backbone's sequences are 256 patches + 1 class token: 257 = 2 x 128 + 1 rows made a
// third query tile and a fifth key chunk that were all padding but one row / column (forward 88 us at 257 tokens against
// 62 us at 256).  With `AttnTail` the tensor-core tiles run on tokens 1 .. Nt - 1 (operand maps and output pointers are
// shifted by one token on the host) and token 0 is handled on the CUDA cores:
//   * as a KEY   inside every tile CTA: its score s_x = q_r . k_0 seeds the online softmax of row r (m = s_x, l = 1) and its
//                value row is added to the accumulator in the epilogue (O_r += p_x v_0) -- a rank-1 update;
//   * as a QUERY by ONE extra CTA per sample (blockIdx.z == 0, blockIdx.x == 0; all heads at once): a first version with one
//                CTA per (sample, head) spent ~9 us per CTA in three dependent global round trips while holding a 66 KB
//                slot -- 768 of them cost more (30 us) than the padding they replaced.  Here a warp owns a token row
//                (all heads: 1536 contiguous bytes, three 16-byte chunks per lane), and the 64 CTAs start first and run
//                beside the tiles.
struct AttnTail {
    const __nv_bfloat16 *q, *k, *v;    // ORIGINAL operand bases (token 0 of sample 0); null: no tail
    __nv_bfloat16 *o;                  // original output base
    float *lse;                        // original log-sum-exp base [B, heads, Nt] (may be null)
    long ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso;
    int Nt;                            // tokens in total (tiles cover Nt - 1)
};
constexpr int kTailBatch = 3;          // token rows in flight per warp (9 x 16 B per lane)

// the extra query row (token 0) of every head of sample b against all Nt keys.  256 threads; E = heads * 64 = 256 * CPL / 8.
// dynamic shared memory: s_q [E] | s_p [heads][Nt] | s_o [8][E]
template <int CPL>      // 16-byte chunks of a token row per lane (E = 256 * CPL elements)
__device__ void attn_tail_query_fwd(const AttnTail &x, unsigned char *smem, int b, int heads, float scale) {
    constexpr int E = 256 * CPL;
    float *s_q = (float *)smem, *s_p = s_q + E, *s_o = s_p + heads * x.Nt;
    __shared__ float s_max[32], s_sum[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < E; i += kAttnThreads) s_q[i] = __bfloat162float(x.q[b * x.bsq + i]) * scale * kLog2e;
    __syncthreads();
    // ---- scores: s[h][t] = q_h . k_t,h (log2 units) ----
    {
        float q[CPL][8];
#pragma unroll
        for (int j = 0; j < CPL; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) q[j][e] = s_q[(lane + 32 * j) * 8 + e];
        const __nv_bfloat16 *kb = x.k + b * x.bsk;
        for (int t0 = warp * kTailBatch; t0 < x.Nt; t0 += 8 * kTailBatch) {
            uint4 w[kTailBatch][CPL];
#pragma unroll
            for (int u = 0; u < kTailBatch; ++u)
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    w[u][j] = t0 + u < x.Nt ? __ldg((const uint4 *)(kb + (long)(t0 + u) * x.ldk) + lane + 32 * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < kTailBatch; ++u)
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const uint32_t ww[4] = {w[u][j].x, w[u][j].y, w[u][j].z, w[u][j].w};
                    float acc = 0.f;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc = fmaf(q[j][2 * e], __uint_as_float(ww[e] << 16), acc);
                        acc = fmaf(q[j][2 * e + 1], __uint_as_float(ww[e] & 0xffff0000u), acc);
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);       // the 8 lanes of a head (64 elements = 8 chunks)
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    if ((lane & 7) == 0 && t0 + u < x.Nt) s_p[((lane >> 3) + 4 * j) * x.Nt + t0 + u] = acc;
                }
        }
    }
    __syncthreads();
    // ---- softmax statistics per head (a warp per head), probabilities back into s_p ----
    for (int h = warp; h < heads; h += 8) {
        float *ph = s_p + h * x.Nt;
        float mx = -INFINITY;
        for (int t = lane; t < x.Nt; t += 32) mx = fmaxf(mx, ph[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int t = lane; t < x.Nt; t += 32) {
            const float pv = ex2_approx(ph[t] - mx);
            ph[t] = pv;
            sum += pv;
        }
        sum = warp_sum(sum);
        if (lane == 0) { s_max[h] = mx; s_sum[h] = sum; }
    }
    __syncthreads();
    // ---- o_h = p_h V_h: per-warp partial sums over its token rows ----
    {
        float acc[CPL][8];
#pragma unroll
        for (int j = 0; j < CPL; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
        const __nv_bfloat16 *vb = x.v + b * x.bsv;
        for (int t0 = warp * kTailBatch; t0 < x.Nt; t0 += 8 * kTailBatch) {
            uint4 w[kTailBatch][CPL];
#pragma unroll
            for (int u = 0; u < kTailBatch; ++u)
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    w[u][j] = t0 + u < x.Nt ? __ldg((const uint4 *)(vb + (long)(t0 + u) * x.ldv) + lane + 32 * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < kTailBatch; ++u)
                if (t0 + u < x.Nt)
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        const float pv = s_p[((lane >> 3) + 4 * j) * x.Nt + t0 + u];
                        const uint32_t ww[4] = {w[u][j].x, w[u][j].y, w[u][j].z, w[u][j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            acc[j][2 * e] = fmaf(pv, __uint_as_float(ww[e] << 16), acc[j][2 * e]);
                            acc[j][2 * e + 1] = fmaf(pv, __uint_as_float(ww[e] & 0xffff0000u), acc[j][2 * e + 1]);
                        }
                    }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) s_o[warp * E + (lane + 32 * j) * 8 + e] = acc[j][e];
    }
    __syncthreads();
    for (int c = tid; c < E; c += kAttnThreads) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_o[w * E + c];
        x.o[b * x.bso + c] = __float2bfloat16_rn(tot / s_sum[c >> 6]);
    }
    if (tid < heads && x.lse != nullptr) x.lse[((long)b * heads + tid) * x.Nt] = s_max[tid] * 0.6931471805599453f + __logf(s_sum[tid]);
}

// ---------------------------------------------------------------------------------------------------------
// Forward, one CTA per (sample, head): query tiles of 128 rows x key chunks of 64 streamed through double-buffered shared
// memory, ONLINE softmax (running row maximum / sum, the O accumulator in TMEM rescaled when the maximum moves), so the
// sequence length is unbounded (the 512 x 512 default of the reference: 1025 tokens) and the CTA needs only 128 TMEM
// columns and 65 KB of shared memory: several CTAs share an SM and cover each other's tensor / TMA round trips.  The loop
// is software pipelined over the flattened (query tile, key chunk) steps: O += P V of step t is issued together with
// S = Q K^T of step t + 1, whose operands arrive by TMA during the softmax of step t.
constexpr int kTilesPerCta = 1;      // query tiles of 128 rows per forward / dq CTA
constexpr int kFwdKC = 64;
template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 3)      // 80 registers, 67 KB, 128 TMEM columns: three CTAs per SM
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, __nv_bfloat16 *__restrict__ O, float *__restrict__ lse, int Nq, int Nk,
                   int Nkp, long ldo, long bso, float scale, uint32_t drop_thresh, float drop_scale,
                   const DropSeed drop_seed, const AttnTail tail, int lse_ld) {
    const bool has_tail = tail.q != nullptr;
    if (has_tail && blockIdx.z == 0) {                   // the z = 0 plane is scheduled first: one worker CTA per sample
        extern __shared__ unsigned char smem_tail[];
        if (blockIdx.x == 0) attn_tail_query_fwd<HD * 12 / 256>(tail, smem_tail, blockIdx.y, gridDim.x, scale);
        return;
    }
    const uint32_t dkey = drop_thresh ? drop_key0(drop_seed) : 0u;     // per-launch dropout key (+ bound step state)
    constexpr int KC = kFwdKC, KB = KC * 128;
    unsigned char *smem;
    uint64_t *bar;
    const uint32_t tmem = attn_prologue<128>(smem, bar);
    AttnSync sync{bar, 0, 0, 0};
    unsigned char *Qs = smem, *Ps = Qs + 16384;
    auto Kc = [&](int i) { return Ps + 16384 + i * 2 * KB; };       // double-buffered key / value chunks
    auto Vc = [&](int i) { return Ps + 16384 + KB + i * 2 * KB; };
    const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x;
    // blockIdx.z: this CTA's range of query tiles (one tile per CTA: key / value chunks are re-streamed per tile anyway, and
    // the finer granularity cuts the last-wave loss of 768 three-tile CTAs over 296-444 resident slots)
    const int q_begin = ((int)blockIdx.z - (has_tail ? 1 : 0)) * kTilesPerCta * 128, q_end = min(Nq, q_begin + kTilesPerCta * 128);
    const int warp = threadIdx.x >> 5, quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + (threadIdx.x & 31);              // this thread's query row (shared by its twin in the other group)
    __shared__ float red_m[2][128], red_s[2][128];
    __shared__ float s_vx[64];                           // tail: value row of the extra key
    __nv_bfloat16 *og = O + b * bso + (long)h * HD;
    const float sl2 = scale * kLog2e;
    if (has_tail && threadIdx.x < 64)
        s_vx[threadIdx.x] = threadIdx.x < HD ? __bfloat162float(tail.v[b * tail.bsv + (long)h * HD + threadIdx.x]) : 0.f;
    const uint32_t tS = tmem, tO = tmem + 64;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int nch = (Nkp + KC - 1) / KC;

    auto issue_s = [&](int buf, int n) {                // S = Q K^T for one key chunk   (thread 0)
        const uint32_t idesc = umma_idesc_bf16(128, n);
        const uint64_t da = umma_desc_k<128>(smem_u32(Qs)), db = umma_desc_k<128>(smem_u32(Kc(buf)));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tS, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
    };

    {   // prologue: first query tile, first key / value chunk; S(0)
        const int n0 = min(KC, Nkp);
        if (threadIdx.x == 0) {
            uint64_t *lb = sync.load_bar();
            mbar_expect_tx(lb, 16384 + 2 * box_bytes(n0));
            tma_rows(Qs, &mapQ, lb, h * HD, q_begin, b, 128);
            tma_rows(Kc(0), &mapK, lb, h * HD, 0, b, n0);
            tma_rows(Vc(0), &mapV, lb, h * HD, 0, b, n0);
        }
        sync.wait_loads();
        if (threadIdx.x == 0) issue_s(0, n0);
        sync.commit_and_wait();
    }

    int t = 0;
    for (int q0 = q_begin; q0 < q_end; q0 += 128) {
        const bool last_tile = q0 + 128 >= q_end;
        const bool live = q0 + quarter * 32 < Nq;        // warps whose 32 rows are all padding skip the softmax
        const uint32_t drop_row = (uint32_t)(((b * heads + h) * Nq + q0 + r) * Nk);   // mask index of (row, key 0); < 2^32 (host check)
        float m = -INFINITY, l = 0.f;                    // running row maximum (raw scores) and this thread's share of the row sum
        float sx = 0.f;
        if (has_tail) {
            // raw score of this thread's query row against the extra key: the Q tile is in shared memory (row r, 16-byte chunks
            // XOR-swizzled by r & 7), the key row comes straight from global memory (the same 128 bytes for every thread)
            const uint4 *kx = (const uint4 *)(tail.k + b * tail.bsk + (long)h * HD);
#pragma unroll
            for (int c8 = 0; c8 < HD / 8; ++c8) {
                const uint4 qv = *(const uint4 *)(Qs + r * 128 + ((c8 ^ (r & 7)) << 4)), kv = __ldg(kx + c8);
                const __nv_bfloat162 *qh = (const __nv_bfloat162 *)&qv, *kh = (const __nv_bfloat162 *)&kv;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a = __bfloat1622float2(qh[j]), bq = __bfloat1622float2(kh[j]);
                    sx = fmaf(a.x, bq.x, sx);
                    sx = fmaf(a.y, bq.y, sx);
                }
            }
            m = sx;                                       // the online softmax starts from this one key
            l = grp == 0 ? 1.0f : 0.f;
        }
        for (int c = 0; c < nch; ++c, ++t) {
            const int kc0 = c * KC, n = min(KC, Nkp - kc0);
            const bool last_c = c == nch - 1;
            const bool has_next = !(last_c && last_tile);
            const int cn = last_c ? 0 : c + 1, n_next = min(KC, Nkp - cn * KC);
            const int cur = t & 1;
            if (threadIdx.x == 0 && has_next) {           // operands of step t + 1 (their buffers were released by the last commit)
                uint64_t *lb = sync.load_bar();
                mbar_expect_tx(lb, 2 * box_bytes(n_next) + (last_c ? 16384 : 0));
                tma_rows(Kc(cur ^ 1), &mapK, lb, h * HD, cn * KC, b, n_next);
                tma_rows(Vc(cur ^ 1), &mapV, lb, h * HD, cn * KC, b, n_next);
                if (last_c) tma_rows(Qs, &mapQ, lb, h * HD, q0 + 128, b, 128);
            }
            // the two warps of a lane quarter take one 32-column half of the chunk each and meet through shared memory
            const bool mine = live && grp * 32 < n;
            uint32_t v[32];
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (mine) {
                tmem_ld32(tS + lane_off + grp * 32, v);
                const int lim = Nk - kc0 - grp * 32;     // columns of this half that are real keys
#pragma unroll
                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], j < lim ? __uint_as_float(v[j]) : -INFINITY);
            }
            red_m[grp][r] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            __syncthreads();
            const float m_new = fmaxf(m, fmaxf(red_m[0][r], red_m[1][r]));
            const float alpha = ex2_approx((m - m_new) * sl2);      // exp2(-inf) = 0 on the first chunk
            const float ms = m_new * sl2;
            m = m_new;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            if (mine) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float p[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = kc0 + grp * 32 + j8 * 8 + j;
                        p[j] = col < Nk ? ex2_approx(fmaf(__uint_as_float(v[j8 * 8 + j]), sl2, -ms)) : 0.f;
                        s4[j & 3] += p[j];
                        // attention-weight dropout (nn.MultiheadAttention(dropout=p)): the row still normalises by the full sum
                        if (drop_thresh) p[j] = drop_keep32(dkey, drop_row + col, drop_thresh) ? p[j] * drop_scale : 0.f;
                    }
                    const int col8 = grp * 4 + j8;
                    if (col8 * 8 < n)
                        store_chunk_kmajor(Ps, r, col8, make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]),
                                                                  pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7])));
                }
            }
            l = fmaf(l, alpha, (s4[0] + s4[1]) + (s4[2] + s4[3]));
            // the maximum moved: rescale this warp's 32-column half of the O accumulator (complete since the last commit)
            if (c > 0 && live && !__all_sync(0xffffffffu, alpha == 1.0f)) {
                uint32_t o[32];
                tmem_ld32(tO + lane_off + grp * 32, o);
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                tmem_st32(tO + lane_off + grp * 32, o);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (has_next) sync.wait_loads();
            if (threadIdx.x == 0) {
                tc_fence_after();
                constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
                for (int ks = 0; ks < n / 16; ++ks) {
                    const uint64_t da = umma_desc_k<128>(smem_u32(Ps)) + 2 * ks;
                    const uint64_t db = umma_desc_mn(smem_u32(Vc(cur) + ks * 2048), 1024);
                    tc_mma_f16(tO, da, db, idesc, (c | ks) ? 1u : 0u);           // O += P V
                }
                if (has_next) issue_s(cur ^ 1, n_next);
            }
            sync.commit_and_wait();
        }
        // end of the query tile: normalise and store; log-sum-exp for the backward pass
        red_s[grp][r] = l;
        __syncthreads();
        if (live) {                                      // warp-uniform: tcgen05.ld is a warp-collective instruction
            const float sum = red_s[0][r] + red_s[1][r];
            const float inv = 1.0f / sum;
            if (grp == 0 && lse != nullptr && q0 + r < Nq) lse[((long)b * heads + h) * lse_ld + q0 + r] = m * scale + __logf(sum);
            uint32_t v[32];
            const int c = grp;                           // each group stores one 32-column half of the output row
            tmem_ld32(tO + lane_off + c * 32, v);
            if (has_tail) {                              // + p_x v_0 (relative to the final row maximum, like the accumulator)
                const float px = ex2_approx((sx - m) * sl2);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(px, s_vx[c * 32 + j], __uint_as_float(v[j])));
            }
            if (q0 + r < Nq)
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    if (c * 32 + j8 * 8 >= HD) break;
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[j8 * 8]) * inv, __uint_as_float(v[j8 * 8 + 1]) * inv);
                    o.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]) * inv, __uint_as_float(v[j8 * 8 + 3]) * inv);
                    o.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]) * inv, __uint_as_float(v[j8 * 8 + 5]) * inv);
                    o.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]) * inv, __uint_as_float(v[j8 * 8 + 7]) * inv);
                    *((uint4 *)(og + (long)(q0 + r) * ldo) + c * 4 + j8) = o;
                }
        }
    }
    attn_epilogue<128>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// dQ (and D = rowsum(dO * O)) per (sample, head).  Keys stream through shared memory in double-buffered chunks of 64
// (256 TMEM columns, 81 KB of shared memory: TWO CTAs per SM), and the loop is software pipelined over the flattened
// (query tile, key chunk) steps: dQ += dS K of step t is issued together with S / dP of step t + 1, whose operands
// (next key chunk, and at a tile boundary the next Q / dO tile) arrive by TMA during the elementwise phase of step t.
constexpr int kDqKC = 64;
template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                      const __nv_bfloat16 *__restrict__ O, const __nv_bfloat16 *__restrict__ dO, const float *__restrict__ lse,
                      __nv_bfloat16 *__restrict__ dQ, float *__restrict__ Dout, int Nq, int Nk, int Nkp, long ldo, long lddo,
                      long lddq, long bso, long bsdo, long bsdq, float scale, uint32_t drop_thresh, float drop_scale,
                      const DropSeed drop_seed) {
    const uint32_t dkey = drop_thresh ? drop_key0(drop_seed) : 0u;     // per-launch dropout key (+ bound step state)
    constexpr int KC = kDqKC, KB = KC * 128;
    unsigned char *smem;
    uint64_t *bar;
    const uint32_t tmem = attn_prologue<256>(smem, bar);
    AttnSync sync{bar, 0, 0, 0};
    unsigned char *Qs = smem, *dOs = Qs + 16384, *dSs = dOs + 16384;
    auto Kc = [&](int i) { return dSs + 16384 + i * 2 * KB; };      // double-buffered key / value chunks
    auto Vc = [&](int i) { return dSs + 16384 + KB + i * 2 * KB; };
    const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x;
    // blockIdx.z: this CTA's range of query tiles (one tile per CTA: key / value chunks are re-streamed per tile anyway, and
    // the finer granularity cuts the last-wave loss of 768 three-tile CTAs over 296-444 resident slots)
    const int q_begin = blockIdx.z * kTilesPerCta * 128, q_end = min(Nq, q_begin + kTilesPerCta * 128);
    const int warp = threadIdx.x >> 5, quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + (threadIdx.x & 31);
    const __nv_bfloat16 *og = O + b * bso + (long)h * HD, *dog = dO + b * bsdo + (long)h * HD;
    __nv_bfloat16 *dqg = dQ + b * bsdq + (long)h * HD;
    const float sl2 = scale * kLog2e;
    const uint32_t tS = tmem, tP = tmem + 64, tQ = tmem + 128;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int nch = (Nkp + KC - 1) / KC;

    auto issue_s = [&](int buf, int n) {                   // S = Q K^T, dP = dO V^T for one key chunk   (thread 0)
        const uint32_t idesc = umma_idesc_bf16(128, n);
        const uint64_t dq = umma_desc_k<128>(smem_u32(Qs)), dd = umma_desc_k<128>(smem_u32(dOs));
        const uint64_t dk = umma_desc_k<128>(smem_u32(Kc(buf))), dv = umma_desc_k<128>(smem_u32(Vc(buf)));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tS, dq + 2 * k, dk + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tP, dd + 2 * k, dv + 2 * k, idesc, k ? 1u : 0u);
    };

    // prologue: first Q / dO tile, first key chunk; S(0), dP(0)
    {
        const int n0 = min(KC, Nkp);
        if (threadIdx.x == 0) {
            uint64_t *lb = sync.load_bar();
            mbar_expect_tx(lb, 2 * 16384 + 2 * box_bytes(n0));
            tma_rows(Qs, &mapQ, lb, h * HD, q_begin, b, 128);
            tma_rows(dOs, &mapdO, lb, h * HD, q_begin, b, 128);
            tma_rows(Kc(0), &mapK, lb, h * HD, 0, b, n0);
            tma_rows(Vc(0), &mapV, lb, h * HD, 0, b, n0);
        }
        sync.wait_loads();
        if (threadIdx.x == 0) issue_s(0, n0);
        sync.commit_and_wait();
    }

    int t = 0;
    for (int q0 = q_begin; q0 < q_end; q0 += 128) {
        const bool last_tile = q0 + 128 >= q_end;
        const bool row_ok = q0 + r < Nq;
        const bool live = q0 + quarter * 32 < Nq;
        const uint32_t drop_row = (uint32_t)(((b * heads + h) * Nq + q0 + r) * Nk);   // mask index of (row, key 0); < 2^32 (host check)
        // D = rowsum(dO * O), log-sum-exp (in exp2 units) of this thread's row
        float Dr = 0.f, l2 = 0.f;
        if (row_ok) {
            const uint4 *po = (const uint4 *)(og + (long)(q0 + r) * ldo), *pd = (const uint4 *)(dog + (long)(q0 + r) * lddo);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                const uint4 a = __ldg(po + c), d = __ldg(pd + c);
                const __nv_bfloat162 *ha = (const __nv_bfloat162 *)&a, *hd = (const __nv_bfloat162 *)&d;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fa = __bfloat1622float2(ha[j]), fd = __bfloat1622float2(hd[j]);
                    Dr = fmaf(fa.x, fd.x, fmaf(fa.y, fd.y, Dr));
                }
            }
            const long idx = ((long)b * heads + h) * Nq + q0 + r;
            if (grp == 0) Dout[idx] = Dr;
            l2 = lse[idx] * kLog2e;
        }
        for (int c = 0; c < nch; ++c, ++t) {
            const int kc0 = c * KC, n = min(KC, Nkp - kc0);
            const bool last_c = c == nch - 1;
            const bool has_next = !(last_c && last_tile);
            const int cn = last_c ? 0 : c + 1, n_next = min(KC, Nkp - cn * KC);
            const int cur = t & 1;
            if (threadIdx.x == 0 && has_next) {           // operands of step t + 1 (their buffers were released by the last commit)
                uint64_t *lb = sync.load_bar();
                mbar_expect_tx(lb, 2 * box_bytes(n_next) + (last_c ? 2 * 16384 : 0));
                tma_rows(Kc(cur ^ 1), &mapK, lb, h * HD, cn * KC, b, n_next);
                tma_rows(Vc(cur ^ 1), &mapV, lb, h * HD, cn * KC, b, n_next);
                if (last_c) {
                    tma_rows(Qs, &mapQ, lb, h * HD, q0 + 128, b, 128);
                    tma_rows(dOs, &mapdO, lb, h * HD, q0 + 128, b, 128);
                }
            }
            if (live && grp * 32 < n) {                   // the two warps of a lane quarter take one 32-column chunk each
                const int sc = grp;
                uint32_t s[32], p[32];
                tmem_ld32_nowait(tS + lane_off + sc * 32, s);
                tmem_ld32_nowait(tP + lane_off + sc * 32, p);
                tmem_wait_ld();
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float ds[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int col = kc0 + sc * 32 + j8 * 8 + j;
                        const float pr = (col < Nk && row_ok) ? ex2_approx(fmaf(__uint_as_float(s[j8 * 8 + j]), sl2, -l2)) : 0.f;
                        float dp = __uint_as_float(p[j8 * 8 + j]);
                        if (drop_thresh) dp = drop_keep32(dkey, drop_row + col, drop_thresh) ? dp * drop_scale : 0.f;
                        ds[j] = pr * (dp - Dr) * scale;
                    }
                    const int col8 = sc * 4 + j8;
                    if (col8 * 8 < n)
                        store_chunk_kmajor(dSs, r, col8, make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]),
                                                                    pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7])));
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (has_next) sync.wait_loads();
            if (threadIdx.x == 0) {
                tc_fence_after();
                constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
                for (int ks = 0; ks < n / 16; ++ks) {
                    const uint64_t da = umma_desc_k<128>(smem_u32(dSs)) + 2 * ks;
                    const uint64_t db = umma_desc_mn(smem_u32(Kc(cur) + ks * 2048), 1024);
                    tc_mma_f16(tQ, da, db, idesc, (c | ks) ? 1u : 0u);       // dQ += dS K
                }
                if (has_next) issue_s(cur ^ 1, n_next);
            }
            sync.commit_and_wait();
        }
        if (live) {
            uint32_t v[32];
            const int c = grp;
            tmem_ld32(tQ + lane_off + c * 32, v);
            if (row_ok)
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    if (c * 32 + j8 * 8 >= HD) break;
                    uint4 o;
                    o.x = pack_bf16x2(__uint_as_float(v[j8 * 8]), __uint_as_float(v[j8 * 8 + 1]));
                    o.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]), __uint_as_float(v[j8 * 8 + 3]));
                    o.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]), __uint_as_float(v[j8 * 8 + 5]));
                    o.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]), __uint_as_float(v[j8 * 8 + 7]));
                    *((uint4 *)(dqg + (long)(q0 + r) * lddq) + c * 4 + j8) = o;
                }
        }
    }
    attn_epilogue<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// dK, dV per (128-key tile, head, sample).  Query tiles of 64 columns keep the CTA at 256 TMEM columns and 97 KB of shared
// memory, so TWO CTAs share an SM: one's elementwise phase covers the other's tensor-core / TMA round trips, and the
// mostly-padding CTA of the last key tile (257 = 2 * 128 + 1 tokens) no longer holds an SM on its own.
constexpr int kDkvQT = 64;
template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                       const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapdO,
                       const float *__restrict__ lse, const float *__restrict__ Dg, __nv_bfloat16 *__restrict__ dK,
                       __nv_bfloat16 *__restrict__ dV, int Nq, int Nk, long lddk, long lddv, long bsdk, long bsdv, float scale,
                       uint32_t drop_thresh, float drop_scale, const DropSeed drop_seed) {
    const uint32_t dkey = drop_thresh ? drop_key0(drop_seed) : 0u;     // per-launch dropout key (+ bound step state)
    constexpr int QT = kDkvQT, QB = QT * 128;             // query rows per tile, bytes of one tile image
    unsigned char *smem;
    uint64_t *bar;
    const uint32_t tmem = attn_prologue<256>(smem, bar);
    AttnSync sync{bar, 0, 0, 0};
    unsigned char *Kt = smem, *Vt = Kt + 16384, *PT = Vt + 16384, *dST = PT + 16384;
    auto Qb = [&](int i) { return dST + 16384 + i * 2 * QB; };      // double-buffered query / dO tiles
    auto dOb = [&](int i) { return dST + 16384 + QB + i * 2 * QB; };
    __shared__ float lse_b[2][QT], D_b[2][QT];            // double buffered with the query / dO tiles
    const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z, heads = gridDim.y;
    const int warp = threadIdx.x >> 5, quarter = warp & 3, grp = warp >> 2;
    const int r = quarter * 32 + (threadIdx.x & 31);
    const float sl2 = scale * kLog2e;
    const uint32_t tS = tmem, tP = tmem + 64, tV = tmem + 128, tK = tmem + 192;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool warp_ok = k0 + quarter * 32 < Nk;

    auto load_stats = [&](int buf, int q0) {               // log-sum-exp (exp2 units) and D of a query tile
        if (threadIdx.x < QT) {
            const int t = threadIdx.x;
            const bool ok = q0 + t < Nq;
            const long idx = ((long)b * heads + h) * Nq + q0 + t;
            lse_b[buf][t] = ok ? lse[idx] * kLog2e : 0.f;
            D_b[buf][t] = ok ? Dg[idx] : 0.f;
        }
    };
    auto issue_st = [&](int buf, int np) {                 // S^T = K Q^T, dP^T = V dO^T   (thread 0)
        const uint32_t idesc = umma_idesc_bf16(128, np);
        const uint64_t dk = umma_desc_k<128>(smem_u32(Kt)), dv = umma_desc_k<128>(smem_u32(Vt));
        const uint64_t dq = umma_desc_k<128>(smem_u32(Qb(buf))), dd = umma_desc_k<128>(smem_u32(dOb(buf)));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tS, dk + 2 * k, dq + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tP, dv + 2 * k, dd + 2 * k, idesc, k ? 1u : 0u);
    };

    // prologue: this CTA's key / value tile, the first query / dO tile and its statistics; S^T(0), dP^T(0)
    if (threadIdx.x == 0) {
        uint64_t *lb = sync.load_bar();
        mbar_expect_tx(lb, 2 * 16384 + 2 * QB);
        tma_rows(Kt, &mapK, lb, h * HD, k0, b, 128);
        tma_rows(Vt, &mapV, lb, h * HD, k0, b, 128);
        tma_rows(Qb(0), &mapQ, lb, h * HD, 0, b, QT);
        tma_rows(dOb(0), &mapdO, lb, h * HD, 0, b, QT);
    }
    load_stats(0, 0);
    sync.wait_loads();
    if (threadIdx.x == 0) issue_st(0, (min(QT, Nq) + 15) / 16 * 16);
    sync.commit_and_wait();

    // steady state, ONE tensor-core round trip per query tile: dV / dK of tile i are issued together with S^T / dP^T of
    // tile i + 1, whose operands streamed in (TMA) during the elementwise phase
    for (int i = 0, q0 = 0; q0 < Nq; ++i, q0 += QT) {
        const int rows = min(QT, Nq - q0);
        const int np = (rows + 15) / 16 * 16;            // query columns of this tile, padded to the UMMA N / K granularity
        const bool has_next = q0 + QT < Nq;
        const int cur = i & 1;
        const uint32_t drop_col = (uint32_t)(((b * heads + h) * Nq + q0) * Nk + k0 + r);   // mask index of (query q0, this key); < 2^32 (host check)
        if (has_next) {
            if (threadIdx.x == 0) {
                uint64_t *lb = sync.load_bar();
                mbar_expect_tx(lb, 2 * QB);
                tma_rows(Qb(cur ^ 1), &mapQ, lb, h * HD, q0 + QT, b, QT);
                tma_rows(dOb(cur ^ 1), &mapdO, lb, h * HD, q0 + QT, b, QT);
            }
            load_stats(cur ^ 1, q0 + QT);
        }
        __syncthreads();                                  // lse / D of the current tile are visible (prologue or previous iteration)
        if (warp_ok && grp * 32 < np) {                   // the two warps of a lane quarter take one 32-column chunk each
            const float *lse_s = lse_b[cur], *D_s = D_b[cur];
            const int c = grp;
            uint32_t s[32], p[32];
            tmem_ld32_nowait(tS + lane_off + c * 32, s);
            tmem_ld32_nowait(tP + lane_off + c * 32, p);
            tmem_wait_ld();
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                float pv[8], ds[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = c * 32 + j8 * 8 + j;                  // query within the tile
                    const float pr = col < rows ? ex2_approx(fmaf(__uint_as_float(s[j8 * 8 + j]), sl2, -lse_s[col])) : 0.f;
                    float keep = 1.0f;
                    if (drop_thresh) keep = drop_keep32(dkey, drop_col + (unsigned)(col * Nk), drop_thresh) ? drop_scale : 0.f;
                    pv[j] = pr * keep;                                    // dV sees the dropped probabilities
                    ds[j] = pr * (__uint_as_float(p[j8 * 8 + j]) * keep - D_s[col]) * scale;
                }
                const int col8 = c * 4 + j8;
                if (col8 * 8 < np) {
                    store_chunk_kmajor(PT, r, col8, make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]),
                                                               pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7])));
                    store_chunk_kmajor(dST, r, col8, make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]),
                                                                pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7])));
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (has_next) sync.wait_loads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
            for (int ks = 0; ks < np / 16; ++ks) {
                const uint64_t dp = umma_desc_k<128>(smem_u32(PT)) + 2 * ks;
                const uint64_t ds = umma_desc_k<128>(smem_u32(dST)) + 2 * ks;
                const uint64_t bo = umma_desc_mn(smem_u32(dOb(cur) + ks * 2048), 1024);
                const uint64_t bq = umma_desc_mn(smem_u32(Qb(cur) + ks * 2048), 1024);
                tc_mma_f16(tV, dp, bo, idesc, (q0 | ks) ? 1u : 0u);      // dV += P^T dO
                tc_mma_f16(tK, ds, bq, idesc, (q0 | ks) ? 1u : 0u);      // dK += dS^T Q
            }
            if (has_next) issue_st(cur ^ 1, (min(QT, Nq - q0 - QT) + 15) / 16 * 16);
        }
        sync.commit_and_wait();      // P^T / dS^T and the S^T / dP^T accumulators are free again
    }
    if (warp_ok) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            __nv_bfloat16 *og = which == 0 ? dV + b * bsdv + (long)h * HD : dK + b * bsdk + (long)h * HD;
            const long ldo = which == 0 ? lddv : lddk;
            const uint32_t t = (which == 0 ? tV : tK) + lane_off;
            uint32_t v[32];
            {
                const int c = grp;
                tmem_ld32(t + c * 32, v);
                if (k0 + r < Nk)
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        if (c * 32 + j8 * 8 >= HD) break;
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[j8 * 8]), __uint_as_float(v[j8 * 8 + 1]));
                        o.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]), __uint_as_float(v[j8 * 8 + 3]));
                        o.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]), __uint_as_float(v[j8 * 8 + 5]));
                        o.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]), __uint_as_float(v[j8 * 8 + 7]));
                        *((uint4 *)(og + (long)(k0 + r) * ldo) + c * 4 + j8) = o;
                    }
            }
        }
    }
    attn_epilogue<256>(tmem);
}

constexpr int kFwdSmem = 2 * 16384 + 4 * kFwdKC * 128 + 1024;
constexpr int kDqSmem = 3 * 16384 + 4 * kDqKC * 128 + 1024;
constexpr int kDkvSmem = 4 * 16384 + 4 * kDkvQT * 128 + 1024;

// [B, T, cols] bf16 operand (pitch ld, batch stride bs, in elements); box = 64 columns x kBoxRows tokens, 128-byte swizzle
static int make_map_tokens(CUtensorMap *map, const void *ptr, int cols, int T, int B, long ld, long bs) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)kBoxRows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

// tail mode: self-attention over 128 m + 1 tokens of 12 heads x 64 (the tiles then carry no padding), no attention dropout
// (the mask index of the tiles would have to skip the extra column), enough (sample, head) pairs that the per-sample worker
// CTAs (~25 us) hide behind the tiles, and the worker's scratch fits the kernel's dynamic shared memory.
// POSE_ATTN_TAIL=0 switches it off (A/B measurements), 2 forces it at any batch (tests)
static bool attn_tail_ok(int B, int heads, int Nq, int Nk, int head_dim, uint32_t drop_thresh, int smem_bytes) {
    const char *env = getenv("POSE_ATTN_TAIL");          // read per call: the tests force the mode at small batch (2)
    const int mode = env ? atoi(env) : 1;
    const long scratch = 4L * (heads * head_dim + (long)heads * Nq + 8L * heads * head_dim);
    return mode != 0 && Nq == Nk && Nq > 128 && (Nq - 1) % 128 == 0 && head_dim == 64 && heads == 12 && drop_thresh == 0 &&
           (mode == 2 || (long)B * heads >= 2 * kNumSMs) && scratch <= smem_bytes;
}

template <class Kern>
static int set_smem(Kern kern, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? POSE_OK : (int)e;
}

}  // namespace pose

using namespace pose;

POSE_API int pose_attention_bf16(const void *Q, const void *K, const void *V, void *O, int B, int heads, int Nq, int Nk,
                                 int head_dim, long ldq, long ldk, long ldv, long ldo, long bsq, long bsk, long bsv, long bso,
                                 float scale, float *lse, float drop_p, uint64_t drop_seed, pose_stream_t stream) {
    if (!Q || !K || !V || !O) return POSE_E_NULL;
    if (drop_p < 0.f || drop_p >= 1.f) return POSE_E_SHAPE;
    const uint32_t dth = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
    const float dsc = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    if (B <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return POSE_E_SHAPE;
    if (head_dim != 48 && head_dim != 64) return POSE_E_UNSUPPORTED;
    if (drop_p > 0.f && (double)B * heads * Nq * Nk >= 4294967296.0) return POSE_E_UNSUPPORTED;   // 32-bit mask counter
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || bsq % 8 || bsk % 8 || bsv % 8 || bso % 8) return POSE_E_ALIGN;
    if ((uintptr_t)Q % 16 || (uintptr_t)K % 16 || (uintptr_t)V % 16 || (uintptr_t)O % 16) return POSE_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const int cols = heads * head_dim;
    if (ldq < cols || ldk < cols || ldv < cols || ldo < cols) return POSE_E_SHAPE;
    // one token outside the tiles (see AttnTail)
    AttnTail tail = {};
    const int lse_ld = Nq;
    if (attn_tail_ok(B, heads, Nq, Nk, head_dim, dth, kFwdSmem)) {
        tail = {(const __nv_bfloat16 *)Q, (const __nv_bfloat16 *)K, (const __nv_bfloat16 *)V, (__nv_bfloat16 *)O, lse,
                ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, Nq};
        Q = (const __nv_bfloat16 *)Q + ldq;
        K = (const __nv_bfloat16 *)K + ldk;
        V = (const __nv_bfloat16 *)V + ldv;
        O = (__nv_bfloat16 *)O + ldo;
        if (lse) lse += 1;
        Nq -= 1;
        Nk -= 1;
    }
    const int Nkp = (Nk + 15) / 16 * 16;
    const dim3 grid(heads, B, (Nq + kTilesPerCta * 128 - 1) / (kTilesPerCta * 128) + (tail.q ? 1 : 0));
    CUtensorMap mq, mk, mv;
    int e;
    if ((e = make_map_tokens(&mq, Q, cols, Nq, B, ldq, bsq))) return e;
    if ((e = make_map_tokens(&mk, K, cols, Nk, B, ldk, bsk))) return e;
    if ((e = make_map_tokens(&mv, V, cols, Nk, B, ldv, bsv))) return e;
#define FWD(HD_)                                                                                                       \
    if ((e = set_smem(attn_fwd_tc_kernel<HD_>, kFwdSmem))) return e;                                                    \
    attn_fwd_tc_kernel<HD_><<<grid, kAttnThreads, kFwdSmem, s>>>(mq, mk, mv, (__nv_bfloat16 *)O, lse, Nq, Nk, Nkp, ldo, bso, \
                                                                scale, dth, dsc, make_drop_seed(drop_seed), tail, lse_ld)
    if (head_dim == 64) { FWD(64); } else { FWD(48); }
#undef FWD
    return launch_status();
}

POSE_API int pose_attention_bwd_bf16(const void *Q, const void *K, const void *V, const void *O, const void *dO,
                                     const float *lse, void *dQ, void *dK, void *dV, float *Dws, int B, int heads, int Nq,
                                     int Nk, int head_dim, long ldq, long ldk, long ldv, long ldo, long lddo, long lddq,
                                     long lddk, long lddv, long bsq, long bsk, long bsv, long bso, long bsdo, long bsdq,
                                     long bsdk, long bsdv, float scale, float drop_p, uint64_t drop_seed, pose_stream_t stream) {
    if (!Q || !K || !V || !O || !dO || !lse || !dQ || !dK || !dV || !Dws) return POSE_E_NULL;
    if (drop_p < 0.f || drop_p >= 1.f) return POSE_E_SHAPE;
    const uint32_t dth = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
    const float dsc = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    if (B <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return POSE_E_SHAPE;
    if (head_dim != 48 && head_dim != 64) return POSE_E_UNSUPPORTED;
    if (drop_p > 0.f && (double)B * heads * Nq * Nk >= 4294967296.0) return POSE_E_UNSUPPORTED;   // 32-bit mask counter
    const long al[] = {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv};
    for (long a : al)
        if (a % 8) return POSE_E_ALIGN;
    const void *ptrs[] = {Q, K, V, O, dO, dQ, dK, dV};
    for (const void *q : ptrs)
        if ((uintptr_t)q % 16) return POSE_E_ALIGN;
    const int Nkp = (Nk + 15) / 16 * 16;
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 g1(heads, B, (Nq + kTilesPerCta * 128 - 1) / (kTilesPerCta * 128)), g2((Nk + 127) / 128, heads, B);
    const int cols = heads * head_dim;
    if (ldq < cols || ldk < cols || ldv < cols || ldo < cols || lddo < cols) return POSE_E_SHAPE;
    CUtensorMap mq, mk, mv, md;
    int e;
    if ((e = make_map_tokens(&mq, Q, cols, Nq, B, ldq, bsq))) return e;
    if ((e = make_map_tokens(&mk, K, cols, Nk, B, ldk, bsk))) return e;
    if ((e = make_map_tokens(&mv, V, cols, Nk, B, ldv, bsv))) return e;
    if ((e = make_map_tokens(&md, dO, cols, Nq, B, lddo, bsdo))) return e;
#define BWD(HD_)                                                                                                       \
    if ((e = set_smem(attn_bwd_dq_tc_kernel<HD_>, kDqSmem))) return e;                                                  \
    if ((e = set_smem(attn_bwd_dkv_tc_kernel<HD_>, kDkvSmem))) return e;                                                \
    attn_bwd_dq_tc_kernel<HD_><<<g1, kAttnThreads, kDqSmem, s>>>(mq, mk, mv, md, (const __nv_bfloat16 *)O,              \
                                                                (const __nv_bfloat16 *)dO, lse, (__nv_bfloat16 *)dQ, Dws, Nq, \
                                                                Nk, Nkp, ldo, lddo, lddq, bso, bsdo, bsdq, scale, dth, dsc, \
                                                                make_drop_seed(drop_seed));                                           \
    attn_bwd_dkv_tc_kernel<HD_><<<g2, kAttnThreads, kDkvSmem, s>>>(mq, mk, mv, md, lse, Dws, (__nv_bfloat16 *)dK,        \
                                                                  (__nv_bfloat16 *)dV, Nq, Nk, lddk, lddv, bsdk, bsdv, scale, \
                                                                  dth, dsc, make_drop_seed(drop_seed))
    if (head_dim == 64) { BWD(64); } else { BWD(48); }
#undef BWD
    return launch_status();
}
