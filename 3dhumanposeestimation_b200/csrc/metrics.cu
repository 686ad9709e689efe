// metrics.cu -- validation metrics on the device (reference: src/utils.py:55-165, called for every validation batch at
// src/train.py:249-254).  compute_pa_mpjpe is a per-sample Python loop with a 3x3 SVD in the reference; here one thread
// owns one sample: centre both 17-joint point sets, M = Pc^T Gc, closed-form 3x3 SVD (cyclic Jacobi on M^T M, double
// precision), reflection fix, scale, aligned error.  The reference applies the rotation as Pc @ (V U^T) -- the
// transpose of the optimal Procrustes rotation -- and that is reproduced, not corrected.  A second one-CTA kernel folds
// the per-sample values in a fixed order (deterministic means).
#include "common.cuh"

namespace pose {

__device__ __forceinline__ void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ void jacobi_eig3(double (&a)[3][3], double (&v)[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
}

__global__ void __launch_bounds__(128)
eval_metrics_kernel(const float *__restrict__ pred, const float *__restrict__ gt, int B, int J, float *__restrict__ per_sample) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *P = pred + (long)b * J * 3, *G = gt + (long)b * J * 3;
    double mp[3] = {0, 0, 0}, mg[3] = {0, 0, 0}, e_plain = 0;
    for (int j = 0; j < J; ++j) {
        float d2 = 0.f;
        for (int d = 0; d < 3; ++d) {
            mp[d] += P[j * 3 + d];
            mg[d] += G[j * 3 + d];
            const float df = P[j * 3 + d] - G[j * 3 + d];
            d2 += df * df;
        }
        e_plain += sqrtf(d2);
    }
    per_sample[b] = (float)(e_plain / J);                         // MPJPE of this sample
    for (int d = 0; d < 3; ++d) {
        mp[d] = (float)(mp[d] / J);
        mg[d] = (float)(mg[d] / J);
    }
    double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, var = 0;
    for (int j = 0; j < J; ++j)
        for (int r = 0; r < 3; ++r) {
            const double pc = (float)(P[j * 3 + r] - mp[r]);
            var += pc * pc;
            for (int c = 0; c < 3; ++c) M[r][c] += pc * (double)(float)(G[j * 3 + c] - mg[c]);
        }
    double A[3][3], V[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            A[r][c] = 0;
            for (int k = 0; k < 3; ++k) A[r][c] += M[k][r] * M[k][c];
        }
    jacobi_eig3(A, V);
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int k = i + 1; k < 3; ++k)
            if (A[idx[k]][idx[k]] > A[idx[i]][idx[i]]) {
                const int t = idx[i];
                idx[i] = idx[k];
                idx[k] = t;
            }
    double S[3], Vs[3][3], U[3][3];
    for (int i = 0; i < 3; ++i) {
        S[i] = sqrt(fmax(A[idx[i]][idx[i]], 0.0));
        for (int r = 0; r < 3; ++r) Vs[r][i] = V[r][idx[i]];
    }
    {
        const double c0[3] = {Vs[0][0], Vs[1][0], Vs[2][0]}, c1[3] = {Vs[0][1], Vs[1][1], Vs[2][1]};
        double vc[3];
        cross3(c0, c1, vc);
        if (vc[0] * Vs[0][2] + vc[1] * Vs[1][2] + vc[2] * Vs[2][2] < 0)
            for (int r = 0; r < 3; ++r) Vs[r][2] = -Vs[r][2];
    }
    const double tiny = 1e-12 * (S[0] > 0 ? S[0] : 1.0);
    int rank = 0;
    for (int i = 0; i < 3; ++i)
        if (S[i] > tiny) {
            for (int r = 0; r < 3; ++r) {
                double t = 0;
                for (int k = 0; k < 3; ++k) t += M[r][k] * Vs[k][i];
                U[r][i] = t / S[i];
            }
            rank = i + 1;
        }
    if (rank == 0) {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) U[r][c] = Vs[r][c];
    } else if (rank == 1) {
        const double u0[3] = {U[0][0], U[1][0], U[2][0]};
        double e[3] = {0, 0, 0}, u1[3], u2[3];
        const int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        e[m] = 1;
        cross3(u0, e, u1);
        const double n = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int r = 0; r < 3; ++r) u1[r] /= n;
        cross3(u0, u1, u2);
        for (int r = 0; r < 3; ++r) {
            U[r][1] = u1[r];
            U[r][2] = u2[r];
        }
    } else if (rank == 2) {
        const double u0[3] = {U[0][0], U[1][0], U[2][0]}, u1[3] = {U[0][1], U[1][1], U[2][1]};
        double u2[3];
        cross3(u0, u1, u2);
        for (int r = 0; r < 3; ++r) U[r][2] = u2[r];
    }
    double R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            R[r][c] = 0;
            for (int k = 0; k < 3; ++k) R[r][c] += Vs[r][k] * U[c][k];      // R = V U^T (utils.py:124)
        }
    double s_sum = S[0] + S[1] + S[2];
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    if (det < 0) {                                                        // utils.py:133-146
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R[r][c] -= 2.0 * Vs[r][2] * U[c][2];
        s_sum = S[0] + S[1] - S[2];
    }
    const double sc = var > 1e-9 ? s_sum / var : 1.0;
    double err = 0;
    for (int j = 0; j < J; ++j) {
        double d2 = 0;
        for (int c = 0; c < 3; ++c) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += (double)(float)(P[j * 3 + k] - mp[k]) * R[k][c];
            const double diff = sc * a + mg[c] - G[j * 3 + c];
            d2 += diff * diff;
        }
        err += sqrt(d2);
    }
    per_sample[B + b] = (float)(err / J);
}

__global__ void __launch_bounds__(256) eval_means_kernel(const float *__restrict__ per_sample, int B, float *__restrict__ means) {
    __shared__ double red[2][256];
    double a = 0, p = 0;
    for (int i = threadIdx.x; i < B; i += 256) {
        a += per_sample[i];
        p += per_sample[B + i];
    }
    red[0][threadIdx.x] = a;
    red[1][threadIdx.x] = p;
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0;
        for (int i = 0; i < 256; ++i) t += red[threadIdx.x][i];
        means[threadIdx.x] = (float)(t / B);
    }
}

}  // namespace pose

using namespace pose;

POSE_API int pose_eval_metrics(const float *pred, const float *gt, int B, int J, float *per_sample, float *means,
                               pose_stream_t stream) {
    if (!pred || !gt || !per_sample || !means) return POSE_E_NULL;
    if (B <= 0 || J <= 0) return POSE_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    eval_metrics_kernel<<<(B + 127) / 128, 128, 0, s>>>(pred, gt, B, J, per_sample);
    eval_means_kernel<<<1, 256, 0, s>>>(per_sample, B, means);
    return launch_status();
}
