// model.cu - GPMM instance reconstruction and per-sample mesh attributes.
//
//   launch_reconstruct     ModelFittingParameters.transformedMesh (ModelFittingParameters.scala:93-110)
//                          X = s (R (ref + mean + Q alpha - c) + c + t), batched over chains
//   launch_vertex_normals  TriangleMesh.vertexNormals (used at NonRigidIcpProposal.scala:100,120)
//   launch_gram            G = Q^T Q for the model.coefficients constant S
#include <algorithm>
#include <cstdlib>

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

// ---- tensor-pipe reconstruction ------------------------------------------------------------------------------------
// X[3N x C] = Q[3N x K] alpha[K x C] as FP64 DMMAs (mma.sync.m8n8k4.f64): a CTA owns 32 vertices (96 rows of Q) x 32 chains,
// a warp 8 vertices (three 8-row tiles) x the 32 chains (four 8-column tiles) = 12 accumulator tiles. A DMMA does the work
// of eight warp-wide DFMAs for one issue slot, and every basis entry fetched from L2 now serves 32 chains instead of 8.
// The k index is consumed in blocks of 8 in the order (0, 2, 4, 6), (1, 3, 5, 7), so that a lane's A operands of two
// consecutive DMMA steps are one 16-byte load of its Q row; the coefficient tile sits in shared memory with a row stride
// of 34 doubles, which makes those B-operand reads and the epilogue's reads of the staged tile conflict free.
// Epilogue: the accumulators go through shared memory so that one lane holds x, y, z of a vertex for the similarity
// transform (same arithmetic as k_reconstruct), 8 consecutive vertices of a chain per 8 lanes (192 contiguous bytes).
constexpr int kRmV = 96;        // vertices per CTA (three groups of 8 per warp share one staged coefficient tile); at N = 1622 and
                                // 2368 chains: 17 x 74 CTAs = 2.8 waves of 3 CTAs / SM (128 vertices: 2.2 waves, the third 17 % full)
constexpr int kRmC = 32;        // chains per CTA
constexpr int kRmLd = 34;       // row stride (doubles) of the coefficient tile and of the staged result

// RMV: vertices per CTA. kRmV for full batches; 32 (one group of 8 per warp) when the grid would not fill the GPU - a few chains
// are a chain of dependent kernel latencies, and three groups in sequence are three times the latency of one. The arithmetic of
// a vertex does not depend on RMV.
template <int RMV>
__global__ void __launch_bounds__(128) k_reconstruct_mma(ModelDev m, int C, const double *__restrict__ theta,
                                                         double *__restrict__ X) {
    extern __shared__ __align__(16) double sm[];
    const int Kp = m.Kp, L = m.K + kTheta0;
    double *sa = sm;                                 // [Kp][kRmLd] coefficients
    double *sp = sm + (size_t)Kp * kRmLd;            // [kRmC][16]: R(9) s t(3) c(3); then [4 warps][24][kRmLd] staged results
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.y * kRmC;
    // coefficient tile: a warp copies whole chains (consecutive lanes read consecutive coefficients), two chains = up to
    // eight independent loads in flight per lane (one load per store in a rolled loop spent a third of the kernel here)
    for (int cc = 2 * warp; cc < kRmC; cc += 8) {
        double v[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int k = lane + 32 * u;
                v[h][u] = (k < m.K && c0 + cc + h < C) ? __ldg(theta + (size_t)(c0 + cc + h) * L + kTheta0 + k) : 0.0;
            }
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int k = lane + 32 * u;
                if (k < Kp) sa[k * kRmLd + cc + h] = v[h][u];
            }
        for (int k = lane + 128; k < Kp; k += 32)   // ranks above 128
#pragma unroll
            for (int h = 0; h < 2; h++)
                sa[k * kRmLd + cc + h] = (k < m.K && c0 + cc + h < C) ? __ldg(theta + (size_t)(c0 + cc + h) * L + kTheta0 + k) : 0.0;
    }
    if (tid < kRmC && c0 + tid < C) {
        const double *th = theta + (size_t)(c0 + tid) * L;
        double *p = sp + tid * 16;
        pose_matrix(th, p);
        p[9] = th[0];
        p[10] = th[1]; p[11] = th[2]; p[12] = th[3];
        p[13] = th[7]; p[14] = th[8]; p[15] = th[9];
    }
    __syncthreads();
    const int g = lane >> 2, q = lane & 3;           // fragment row / column group
    const int last_row = 3 * m.N - 1;
    const double *sb = sa + (2 * q) * kRmLd + g;
    double *sx = sp + kRmC * 16 + (size_t)warp * 24 * kRmLd;   // [24 rows][kRmLd], private to the warp
    for (int grp = 0; grp < RMV / 32; grp++) {
        const int v0 = blockIdx.x * RMV + grp * 32 + warp * 8;
        if (v0 >= m.N) break;                        // warp-uniform
        const double2 *qa0, *qa1, *qa2;
        qa0 = reinterpret_cast<const double2 *>(m.Q + (size_t)min(3 * v0 + g, last_row) * Kp) + q;        // rows past the mesh are
        qa1 = reinterpret_cast<const double2 *>(m.Q + (size_t)min(3 * v0 + 8 + g, last_row) * Kp) + q;    // computed from a valid
        qa2 = reinterpret_cast<const double2 *>(m.Q + (size_t)min(3 * v0 + 16 + g, last_row) * Kp) + q;   // row and dropped
        double acc[3][4][2];
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int n = 0; n < 4; n++) acc[t][n][0] = acc[t][n][1] = 0.0;
        double2 a_next[3] = {__ldg(qa0), __ldg(qa1), __ldg(qa2)};
#pragma unroll 2
        for (int k8 = 0; k8 < Kp; k8 += 8) {
            double2 a[3];
#pragma unroll
            for (int t = 0; t < 3; t++) a[t] = a_next[t];
            if (k8 + 8 < Kp) {
                a_next[0] = __ldg(qa0 + (k8 + 8) / 2); a_next[1] = __ldg(qa1 + (k8 + 8) / 2); a_next[2] = __ldg(qa2 + (k8 + 8) / 2);
            }
            double b0[4], b1[4];
#pragma unroll
            for (int n = 0; n < 4; n++) {
                b0[n] = sb[(k8) * kRmLd + 8 * n];                 // coefficient k8 + 2 q of chain 8 n + g
                b1[n] = sb[(k8 + 1) * kRmLd + 8 * n];             // coefficient k8 + 2 q + 1
            }
#pragma unroll
            for (int t = 0; t < 3; t++)
#pragma unroll
                for (int n = 0; n < 4; n++) {
                    dmma_8x8x4(acc[t][n][0], acc[t][n][1], a[t].x, b0[n]);
                    dmma_8x8x4(acc[t][n][0], acc[t][n][1], a[t].y, b1[n]);
                }
        }
        __syncwarp();                                // the previous group's epilogue reads of sx are done
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int n = 0; n < 4; n++)
                *reinterpret_cast<double2 *>(sx + (8 * t + g) * kRmLd + 8 * n + 2 * q) = make_double2(acc[t][n][0], acc[t][n][1]);
        __syncwarp();
        const int vl = lane & 7;                     // vertex within the warp's 8
        const int i = v0 + vl;
        double r0 = 0, r1 = 0, r2 = 0, m0 = 0, m1 = 0, m2 = 0;
        if (i < m.N) {
            r0 = m.ref[3 * i]; r1 = m.ref[3 * i + 1]; r2 = m.ref[3 * i + 2];
            m0 = m.mean[3 * i]; m1 = m.mean[3 * i + 1]; m2 = m.mean[3 * i + 2];
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int cc = 4 * j + (lane >> 3);
            if (i < m.N && c0 + cc < C) {
                const double *p = sp + cc * 16;
                const double ax = sx[(3 * vl) * kRmLd + cc], ay = sx[(3 * vl + 1) * kRmLd + cc], az = sx[(3 * vl + 2) * kRmLd + cc];
                const double px = r0 + (m0 + ax) - p[13], py = r1 + (m1 + ay) - p[14], pz = r2 + (m2 + az) - p[15];
                double *o = X + ((size_t)(c0 + cc) * m.N + i) * 3;
                o[0] = p[9] * ((p[0] * px + p[1] * py + p[2] * pz) + p[13] + p[10]);
                o[1] = p[9] * ((p[3] * px + p[4] * py + p[5] * pz) + p[14] + p[11]);
                o[2] = p[9] * ((p[6] * px + p[7] * py + p[8] * pz) + p[15] + p[12]);
            }
        }
    }
}

void launch_reconstruct(const ModelDev &m, int C, const double *d_theta, double *d_X, cudaStream_t s) {
    ProfScope _ps(ST_RECONSTRUCT, s);
    if (C <= 0) return;
    static const bool no_mma = getenv("ICPCUDA_NO_DMMA") && getenv("ICPCUDA_NO_DMMA")[0] == '1';
    if (!no_mma) {   // every batch size takes the same kernel: a chain's mesh must not depend on how chains are batched or sharded
        dim3 grid((m.N + kRmV - 1) / kRmV, (C + kRmC - 1) / kRmC);
        const size_t smem = sizeof(double) * ((size_t)m.Kp * kRmLd + kRmC * 16 + (size_t)4 * 24 * kRmLd);
        if ((long long)grid.x * grid.y < 148) {   // latency-bound batch: three times the CTAs, a third of the work each
            grid.x = (m.N + 31) / 32;
            ICP_CUDA(cudaFuncSetAttribute(k_reconstruct_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_reconstruct_mma<32><<<grid, 128, smem, s>>>(m, C, d_theta, d_X);
        } else {
            ICP_CUDA(cudaFuncSetAttribute(k_reconstruct_mma<kRmV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_reconstruct_mma<kRmV><<<grid, 128, smem, s>>>(m, C, d_theta, d_X);
        }
        ICP_CUDA(cudaGetLastError());
        return;
    }
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
