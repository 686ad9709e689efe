// loss.cu -- fused forward + backward of the composite pose loss.
// Reference semantics: src/loss.py:57-85 (ComprehensivePoseLoss.forward), :29-47 (inter-joint
// distance term), :49-55 (absolute root term).  Gradient formula: SURVEY.md 8a row F.
//
// One warp owns one sample (J*3 <= 96 values live in shared memory); lane i < J produces the
// inter-joint term and gradient of joint i by walking the other joints, so no atomics are needed
// and the result is deterministic.  Block partial sums go to the workspace; the last block to
// finish (ticket counter) folds them in a fixed order in fp64 and writes the five scalars.
// HBM traffic: 2*J*12 B read + J*12 B written per sample (632 B for J = 17 with the scalars).
#include "common.cuh"

namespace pose {

constexpr int kLossWarps = 8;
constexpr int kLossMaxJ = 32;

struct LossWorkspaceHeader {
    unsigned int ticket;
    unsigned int pad[3];
};

__global__ void __launch_bounds__(kLossWarps * 32)
pose_loss_kernel(const float *__restrict__ pred, const float *__restrict__ gt, int B, int J, float w_mse, float w_l1,
                 float w_ij, float w_root, float *__restrict__ out5, float *__restrict__ grad, float grad_scale,
                 LossWorkspaceHeader *hdr, double *partials) {
    __shared__ float sp[kLossWarps][kLossMaxJ * 3];
    __shared__ float sg[kLossWarps][kLossMaxJ * 3];
    __shared__ float sgrad[kLossWarps][kLossMaxJ * 3];
    __shared__ float sblock[kLossWarps][4];
    __shared__ bool is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n3 = J * 3;
    const long n_el = (long)B * n3;
    const long n_pairs = (long)J * (J - 1) / 2;
    // d total / d pred coefficients (mean reductions folded in)
    const float c_mse = w_mse * 2.0f / (float)n_el * grad_scale;
    const float c_l1 = w_l1 / (float)n_el * grad_scale;
    const float c_ij = n_pairs ? w_ij / ((float)B * (float)n_pairs) * grad_scale : 0.0f;
    const float c_root = w_root / ((float)B * 3.0f) * grad_scale;

    float a_mse = 0.f, a_l1 = 0.f, a_ij = 0.f, a_root = 0.f;
    for (long b = (long)blockIdx.x * kLossWarps + warp; b < B; b += (long)gridDim.x * kLossWarps) {
        const float *p = pred + b * n3, *g = gt + b * n3;
        for (int i = lane; i < n3; i += 32) {
            sp[warp][i] = __ldg(p + i);
            sg[warp][i] = __ldg(g + i);
        }
        __syncwarp();
        // element-wise terms: MSE, L1, root
        float e_mse = 0.f, e_l1 = 0.f, e_root = 0.f;
        for (int i = lane; i < n3; i += 32) {
            float d = sp[warp][i] - sg[warp][i];
            float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
            e_mse += d * d;
            e_l1 += fabsf(d);
            float gr = c_mse * d + c_l1 * sgn;
            if (i < 3) {
                e_root += fabsf(d);
                gr += c_root * sgn;
            }
            sgrad[warp][i] = gr;
        }
        // pairwise term: lane i walks every other joint j
        float e_ij = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        if (lane < J) {
            const float pix = sp[warp][lane * 3], piy = sp[warp][lane * 3 + 1], piz = sp[warp][lane * 3 + 2];
            const float gix = sg[warp][lane * 3], giy = sg[warp][lane * 3 + 1], giz = sg[warp][lane * 3 + 2];
            for (int j = 0; j < J; ++j) {
                if (j == lane) continue;
                float dx = pix - sp[warp][j * 3], dy = piy - sp[warp][j * 3 + 1], dz = piz - sp[warp][j * 3 + 2];
                float ex = gix - sg[warp][j * 3], ey = giy - sg[warp][j * 3 + 1], ez = giz - sg[warp][j * 3 + 2];
                float dp = sqrtf(dx * dx + dy * dy + dz * dz);
                float dg = sqrtf(ex * ex + ey * ey + ez * ez);
                float e = dp - dg;
                if (j > lane) e_ij += fabsf(e);  // each unordered pair counted once (triu, offset 1)
                if (dp > 0.f) {                  // norm backward is 0 at coincident joints
                    float s = ((e > 0.f) ? c_ij : ((e < 0.f) ? -c_ij : 0.f)) / dp;
                    gx += s * dx;
                    gy += s * dy;
                    gz += s * dz;
                }
            }
        }
        __syncwarp();
        if (lane < J) {
            sgrad[warp][lane * 3] += gx;
            sgrad[warp][lane * 3 + 1] += gy;
            sgrad[warp][lane * 3 + 2] += gz;
        }
        __syncwarp();
        if (grad != nullptr)
            for (int i = lane; i < n3; i += 32) grad[b * n3 + i] = sgrad[warp][i];
        a_mse += warp_sum(e_mse);
        a_l1 += warp_sum(e_l1);
        a_ij += warp_sum(e_ij);
        a_root += warp_sum(e_root);
        __syncwarp();
    }
    if (lane == 0) {
        sblock[warp][0] = a_mse;
        sblock[warp][1] = a_l1;
        sblock[warp][2] = a_ij;
        sblock[warp][3] = a_root;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < kLossWarps; ++w) s += (double)sblock[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&hdr->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block: fixed-order fp64 fold of the per-block partials
    __shared__ double fin[4];
    if (warp < 4) {
        double s = 0.0;
        for (unsigned k = lane; k < gridDim.x; k += 32) s += __ldcg(partials + (size_t)k * 4 + warp);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) fin[warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mse = (float)(fin[0] / (double)n_el);
        float l1 = (float)(fin[1] / (double)n_el);
        float ij = n_pairs ? (float)(fin[2] / ((double)B * (double)n_pairs)) : __int_as_float(0x7fc00000);
        float root = (float)(fin[3] / ((double)B * 3.0));
        out5[0] = mse;
        out5[1] = l1;
        out5[2] = ij;
        out5[3] = root;
        out5[4] = w_mse * mse + w_l1 * l1 + w_ij * ij + w_root * root;
        hdr->ticket = 0;  // leave the workspace ready for the next call
    }
}

static int loss_grid(int B) {
    long blocks = ((long)B + kLossWarps - 1) / kLossWarps;
    long cap = (long)kNumSMs * 8;  // 8 resident 256-thread CTAs per SM
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pose

POSE_API size_t pose_loss_workspace_bytes(int B, int J) {
    (void)J;
    return sizeof(pose::LossWorkspaceHeader) + sizeof(double) * 4 * (size_t)pose::loss_grid(B);
}

POSE_API int pose_loss_fwd_bwd(const float *pred, const float *gt, int B, int J, const float *weights, float *out5,
                               float *grad, float grad_scale, void *workspace, size_t workspace_bytes,
                               pose_stream_t stream) {
    if (!pred || !gt || !weights || !out5 || !workspace) return POSE_E_NULL;
    if (B <= 0 || J <= 0 || J > pose::kLossMaxJ) return POSE_E_SHAPE;
    if (workspace_bytes < pose_loss_workspace_bytes(B, J)) return POSE_E_WORKSPACE;
    if ((uintptr_t)workspace % 16) return POSE_E_ALIGN;
    auto *hdr = (pose::LossWorkspaceHeader *)workspace;
    auto *partials = (double *)(hdr + 1);
    pose::pose_loss_kernel<<<pose::loss_grid(B), pose::kLossWarps * 32, 0, (cudaStream_t)stream>>>(
        pred, gt, B, J, weights[0], weights[1], weights[2], weights[3], out5, grad, grad_scale, hdr, partials);
    return pose::launch_status();
}
