// model.cu - GPMM instance reconstruction and per-sample mesh attributes.
//
//   launch_reconstruct     ModelFittingParameters.transformedMesh (ModelFittingParameters.scala:93-110)
//                          X = s (R (ref + mean + Q alpha - c) + c + t), batched over chains
//   launch_vertex_normals  TriangleMesh.vertexNormals (used at NonRigidIcpProposal.scala:100,120)
//   launch_gram            G = Q^T Q for the model.coefficients constant S
#include "icp_internal.h"
#include "icp_device.cuh"

namespace icp {

__global__ void k_scale_basis(int rows, int K, int Kp, const double *__restrict__ U, const double *__restrict__ var,
                              double *__restrict__ Q, double *__restrict__ QT) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)rows * Kp) return;
    int r = (int)(g / Kp), j = (int)(g % Kp);
    double v = j < K ? U[(size_t)r * K + j] * sqrt(var[j]) : 0.0;
    Q[(size_t)r * Kp + j] = v;
    QT[(size_t)j * rows + r] = v;
}

void launch_scale_basis(int rows, int K, int Kp, const double *d_U, const double *d_var, double *d_Q, double *d_QT,
                        cudaStream_t s) {
    long long total = (long long)rows * Kp;
    k_scale_basis<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(rows, K, Kp, d_U, d_var, d_Q, d_QT);
    ICP_CUDA(cudaGetLastError());
}

constexpr int kRecCH = 8;   // chains per CTA: every basis entry fetched from L2 serves 8 chains (at 4 the kernel sat on the L2 bandwidth cap)

// thread = vertex, CTA = 128 vertices x kRecCH chains; alpha tiles broadcast from shared memory,
// Q^T rows streamed coalesced from L2 (the basis is shared by every chain)
__global__ void __launch_bounds__(128) k_reconstruct(ModelDev m, int C, const double *__restrict__ theta,
                                                     double *__restrict__ X) {
    extern __shared__ double sm[];
    double *sa = sm;                      // [Kp][kRecCH]
    double *sp = sm + (size_t)m.Kp * kRecCH;  // [kRecCH][16]: R(9) s t(3) c(3)
    int c0 = blockIdx.y * kRecCH;
    int L = m.K + kTheta0;
    for (int e = threadIdx.x; e < m.Kp * kRecCH; e += blockDim.x) {
        int k = e / kRecCH, cc = e % kRecCH;
        sa[e] = (k < m.K && c0 + cc < C) ? theta[(size_t)(c0 + cc) * L + kTheta0 + k] : 0.0;
    }
    if (threadIdx.x < kRecCH && c0 + threadIdx.x < C) {
        const double *th = theta + (size_t)(c0 + threadIdx.x) * L;
        double *p = sp + threadIdx.x * 16;
        pose_matrix(th, p);
        p[9] = th[0];
        p[10] = th[1]; p[11] = th[2]; p[12] = th[3];
        p[13] = th[7]; p[14] = th[8]; p[15] = th[9];
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.N) return;
    double acc[3][kRecCH];
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int cc = 0; cc < kRecCH; cc++) acc[d][cc] = 0.0;
    size_t rows = (size_t)3 * m.N;
    const double *qt = m.QT + (size_t)3 * i;
#pragma unroll 4
    for (int k = 0; k < m.K; k++) {
        double q0 = __ldg(qt + k * rows), q1 = __ldg(qt + k * rows + 1), q2 = __ldg(qt + k * rows + 2);
#pragma unroll
        for (int cc = 0; cc < kRecCH; cc++) {
            double a = sa[k * kRecCH + cc];
            acc[0][cc] = fma(q0, a, acc[0][cc]);
            acc[1][cc] = fma(q1, a, acc[1][cc]);
            acc[2][cc] = fma(q2, a, acc[2][cc]);
        }
    }
    double r0 = m.ref[3 * i], r1 = m.ref[3 * i + 1], r2 = m.ref[3 * i + 2];
    double m0 = m.mean[3 * i], m1 = m.mean[3 * i + 1], m2 = m.mean[3 * i + 2];
#pragma unroll
    for (int cc = 0; cc < kRecCH; cc++) {
        if (c0 + cc >= C) break;
        const double *p = sp + cc * 16;
        double px = r0 + (m0 + acc[0][cc]) - p[13], py = r1 + (m1 + acc[1][cc]) - p[14], pz = r2 + (m2 + acc[2][cc]) - p[15];
        double *o = X + ((size_t)(c0 + cc) * m.N + i) * 3;
        o[0] = p[9] * ((p[0] * px + p[1] * py + p[2] * pz) + p[13] + p[10]);
        o[1] = p[9] * ((p[3] * px + p[4] * py + p[5] * pz) + p[14] + p[11]);
        o[2] = p[9] * ((p[6] * px + p[7] * py + p[8] * pz) + p[15] + p[12]);
    }
}

void launch_reconstruct(const ModelDev &m, int C, const double *d_theta, double *d_X, cudaStream_t s) {
    ProfScope _ps(ST_RECONSTRUCT, s);
    if (C <= 0) return;
    dim3 grid((m.N + 127) / 128, (C + kRecCH - 1) / kRecCH);
    size_t smem = sizeof(double) * ((size_t)m.Kp * kRecCH + kRecCH * 16);
    k_reconstruct<<<grid, 128, smem, s>>>(m, C, d_theta, d_X);
    ICP_CUDA(cudaGetLastError());
}

__global__ void k_vertex_normals(ModelDev m, int C, const double *__restrict__ X, double *__restrict__ out) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)C * m.N) return;
    int c = (int)(g / m.N), v = (int)(g % m.N);
    double nx, ny, nz;
    vertex_normal_dev(m, X + (size_t)c * m.N * 3, v, nx, ny, nz);
    out[3 * g] = nx; out[3 * g + 1] = ny; out[3 * g + 2] = nz;
}

void launch_vertex_normals(const ModelDev &m, int C, const double *d_X, double *d_normals, cudaStream_t s) {
    long long total = (long long)C * m.N;
    if (total <= 0) return;
    k_vertex_normals<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(m, C, d_X, d_normals);
    ICP_CUDA(cudaGetLastError());
}

// G[i][j] = sum_r Q[r][i] Q[r][j]; one-off at model creation
__global__ void k_gram(int rows, int Kp, const double *__restrict__ Q, double *__restrict__ G) {
    int i = blockIdx.x, j = threadIdx.x;
    if (j >= Kp) return;
    double acc = 0;
    for (int r = 0; r < rows; r++) acc = fma(Q[(size_t)r * Kp + i], Q[(size_t)r * Kp + j], acc);
    G[(size_t)i * Kp + j] = acc;
}

void launch_gram(const ModelDev &m, double *d_G, cudaStream_t s) {
    k_gram<<<m.Kp, ((m.Kp + 31) / 32) * 32, 0, s>>>(3 * m.N, m.Kp, m.Q, d_G);
    ICP_CUDA(cudaGetLastError());
}

}  // namespace icp
