// posterior.cu - the ICP proposal's GP posterior over sampled correspondences.
//
// Replaces (all paths relative to src/main/scala of the reference):
//   launch_observations      modelBased/targetBasedClosestPointsEstimation + surfaceNormalDependantNoise
//                            (api/sampling/proposals/NonRigidIcpProposal.scala:88-131,
//                             api/sampling/SurfaceNoiseHelpers.scala:32-60)
//   launch_posterior_build   interpolatedModel.posterior(obs) (NonRigidIcpProposal.scala:152), the
//                            M = I + sum Q_i^T Sigma_i^-1 Q_i and b = sum Q_i^T Sigma_i^-1 y_i of
//                            LowRankGaussianProcess.regression
//   launch_cholesky_solve    pinv(M) and mean_coeffs -> Cholesky M = L L^T, mu = M^-1 b
//   launch_propose           propose (NonRigidIcpProposal.scala:53-68) via Appendix A6
//   launch_log_transition    logTransitionProbability (NonRigidIcpProposal.scala:71-85) via Appendix A7
//
// Sigma_i^-1 = F_i^T F_i with F_i rows n/sd_n, t1/sd_t, t2/sd_t (orthonormal frame of the vertex
// normal), so M = I + A^T A with A = stack(F_i Q_i): a batched symmetric rank-3n update in FP64.
#include "icp_device.cuh"
#include "icp_internal.h"

namespace icp {

// ---------------------------------------------------------------------------------------------------
// observations
// ---------------------------------------------------------------------------------------------------
__global__ void k_observations(ObsArgs a, ObsDev o) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)a.C * o.n) return;
    int c = (int)(g / o.n), i = (int)(g % o.n);
    const ModelDev &m = a.m;
    const double *th = a.theta + (size_t)c * (m.K + kTheta0);
    const double *Xc = a.X + (size_t)c * m.N * 3;
    int id;
    double tx, ty, tz;
    bool drop = false;
    if (a.prm.direction == ICP_TARGET_SAMPLING) {
        id = a.near_vid[g];                                         // :118 findClosestPoint on the current mesh
        tx = a.tp[3 * i]; ty = a.tp[3 * i + 1]; tz = a.tp[3 * i + 2];
        if (id < 0) drop = true;
        else if (a.prm.boundary_aware && m.boundary[id]) drop = true;  // :119,124
    } else {
        id = a.ids[i];
        tx = a.cp[3 * g]; ty = a.cp[3 * g + 1]; tz = a.cp[3 * g + 2];  // :97 closest point on the target
        if (a.prm.boundary_aware && a.cp_on_boundary && a.cp_on_boundary[g]) drop = true;  // :99,104
    }
    double *F = o.F + 9 * g, *y = o.y + 3 * g;
    if (drop) {
        o.vid[g] = -1;
        for (int k = 0; k < 9; k++) F[k] = 0.0;
        y[0] = y[1] = y[2] = 0.0;
        return;
    }
    atomicAdd(&o.nobs[c], 1);
    o.vid[g] = id;
    double f[9];
    if (a.iso) {
        double w = 1.0 / sqrt(a.iso_sigma2);
        for (int k = 0; k < 9; k++) f[k] = 0.0;
        f[0] = f[4] = f[8] = w;
    } else {
        double nx, ny, nz;
        vertex_normal_dev(m, Xc, id, nx, ny, nz);                   // :100,120 currentMesh.vertexNormals.atPoint(id)
        // SurfaceNoiseHelpers.scala:39: normalize again
        double nn = sqrt(nx * nx + ny * ny + nz * nz);
        nx /= nn; ny /= nn; nz /= nn;
        // :44-48 candidate = n x e_x; (inverted) fallback test; n x e_y otherwise
        double c0 = 0.0, c1 = nz, c2 = -ny;
        double t1x, t1y, t1z;
        if (c0 * c0 + c1 * c1 + c2 * c2 < 0.0001) { t1x = c0; t1y = c1; t1z = c2; }
        else { t1x = -nz; t1y = 0.0; t1z = nx; }
        double tn = sqrt(t1x * t1x + t1y * t1y + t1z * t1z);
        t1x /= tn; t1y /= tn; t1z /= tn;                             // 0/0 = NaN exactly where the reference yields NaN
        double t2x = ny * t1z - nz * t1y, t2y = nz * t1x - nx * t1z, t2z = nx * t1y - ny * t1x;
        double t2n = sqrt(t2x * t2x + t2y * t2y + t2z * t2z);
        t2x /= t2n; t2y /= t2n; t2z /= t2n;
        double wn = 1.0 / a.prm.noise_along_normal, wt = 1.0 / a.prm.tangential_noise;
        f[0] = nx * wn; f[1] = ny * wn; f[2] = nz * wn;
        f[3] = t1x * wt; f[4] = t1y * wt; f[5] = t1z * wt;
        f[6] = t2x * wt; f[7] = t2y * wt; f[8] = t2z * wt;
    }
    double R[9];
    pose_matrix(th, R);
    double ix, iy, iz;
    inverse_pose(th, R, tx, ty, tz, ix, iy, iz);                    // :108,129 inversePoseTransform(targetPoint)
    double y0 = (ix - m.ref[3 * id]) - m.mean[3 * id], y1 = (iy - m.ref[3 * id + 1]) - m.mean[3 * id + 1],
           y2 = (iz - m.ref[3 * id + 2]) - m.mean[3 * id + 2];
    for (int k = 0; k < 9; k++) F[k] = f[k];
    y[0] = f[0] * y0 + f[1] * y1 + f[2] * y2;
    y[1] = f[3] * y0 + f[4] * y1 + f[5] * y2;
    y[2] = f[6] * y0 + f[7] * y1 + f[8] * y2;
}

void launch_observations(const ObsArgs &a, const ObsDev &o, cudaStream_t s) {
    ProfScope _ps(ST_OBSERVATIONS, s);
    long long total = (long long)a.C * o.n;
    ICP_CUDA(cudaMemsetAsync(o.nobs, 0, sizeof(int) * a.C, s));
    if (total <= 0) return;
    k_observations<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a, o);
    ICP_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// M = I + A^T A, b = A^T y   (one CTA per chain, FP64)
// ---------------------------------------------------------------------------------------------------
constexpr int kPbThreads = 256;
constexpr int kPbObs = 8;  // observations (x3 rows) staged per chunk

__device__ __forceinline__ void tri_index(int idx, int &ti, int &tj) {
    // idx -> (ti, tj) with tj <= ti, row-major over the lower triangle
    int t = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((t + 1) * (t + 2) / 2 <= idx) t++;
    while (t * (t + 1) / 2 > idx) t--;
    ti = t;
    tj = idx - t * (t + 1) / 2;
}

template <int TILES>  // 4x4 register tiles per thread
__global__ void __launch_bounds__(kPbThreads) k_posterior_build(ModelDev m, ObsDev o, double *__restrict__ M,
                                                                double *__restrict__ bvec) {
    extern __shared__ double sm[];
    const int Kp = m.Kp, nt = Kp / 4, ntiles = nt * (nt + 1) / 2;
    double *sA = sm;                              // [3 kPbObs][Kp]
    double *sy = sm + (size_t)3 * kPbObs * Kp;    // [3 kPbObs]
    int c = blockIdx.x;
    const int *vid = o.vid + (size_t)c * o.n;
    const double *F = o.F + (size_t)c * o.n * 9;
    const double *y = o.y + (size_t)c * o.n * 3;

    int ti[TILES], tj[TILES];
    bool on[TILES];
    double acc[TILES][4][4];
#pragma unroll
    for (int t = 0; t < TILES; t++) {
        int idx = threadIdx.x + t * kPbThreads;
        on[t] = idx < ntiles;
        tri_index(on[t] ? idx : 0, ti[t], tj[t]);
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[t][r][q] = 0.0;
    }
    double bacc = 0.0;

    for (int base = 0; base < o.n; base += kPbObs) {
        int nobs = min(kPbObs, o.n - base);
        // stage A rows of this chunk: A[3 o + r][j] = sum_d F[o][r][d] Q[3 vid + d][j]
        for (int e = threadIdx.x; e < kPbObs * Kp; e += kPbThreads) {
            int ob = e / Kp, j = e % Kp;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (ob < nobs) {
                int v = vid[base + ob];
                if (v >= 0) {
                    const double *f = F + (size_t)(base + ob) * 9;
                    const double *q = m.Q + (size_t)3 * v * Kp + j;
                    double q0 = __ldg(q), q1 = __ldg(q + Kp), q2 = __ldg(q + 2 * Kp);
                    a0 = f[0] * q0 + f[1] * q1 + f[2] * q2;
                    a1 = f[3] * q0 + f[4] * q1 + f[5] * q2;
                    a2 = f[6] * q0 + f[7] * q1 + f[8] * q2;
                }
            }
            sA[(3 * ob) * Kp + j] = a0;
            sA[(3 * ob + 1) * Kp + j] = a1;
            sA[(3 * ob + 2) * Kp + j] = a2;
        }
        if (threadIdx.x < 3 * kPbObs) {
            int ob = threadIdx.x / 3;
            sy[threadIdx.x] = (ob < nobs && vid[base + ob] >= 0) ? y[(size_t)(base + ob) * 3 + threadIdx.x % 3] : 0.0;
        }
        __syncthreads();
        const int rows = 3 * kPbObs;
#pragma unroll
        for (int t = 0; t < TILES; t++) {
            if (!on[t]) continue;
            const double *pa = sA + 4 * ti[t], *pb = sA + 4 * tj[t];
#pragma unroll 4
            for (int r = 0; r < rows; r++) {
                double2 a01 = *reinterpret_cast<const double2 *>(pa + r * Kp), a23 = *reinterpret_cast<const double2 *>(pa + r * Kp + 2);
                double2 b01 = *reinterpret_cast<const double2 *>(pb + r * Kp), b23 = *reinterpret_cast<const double2 *>(pb + r * Kp + 2);
                double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[t][p][q] = fma(av[p], bv[q], acc[t][p][q]);
            }
        }
        if (threadIdx.x < Kp) {
            for (int r = 0; r < rows; r++) bacc = fma(sA[r * Kp + threadIdx.x], sy[r], bacc);
        }
        __syncthreads();
    }
    double *Mc = M + (size_t)c * Kp * Kp;
#pragma unroll
    for (int t = 0; t < TILES; t++) {
        if (!on[t]) continue;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int i = 4 * ti[t] + p, j = 4 * tj[t] + q;
                double v = acc[t][p][q] + (i == j ? 1.0 : 0.0);
                Mc[(size_t)i * Kp + j] = v;
                if (ti[t] != tj[t]) Mc[(size_t)j * Kp + i] = v;
            }
    }
    if (threadIdx.x < Kp) bvec[(size_t)c * Kp + threadIdx.x] = bacc;
}

void launch_posterior_build(const ModelDev &m, int C, const ObsDev &o, double *d_M, double *d_b, cudaStream_t s) {
    ProfScope _ps(ST_POSTERIOR_BUILD, s);
    if (C <= 0) return;
    int nt = m.Kp / 4, ntiles = nt * (nt + 1) / 2;
    size_t smem = sizeof(double) * ((size_t)3 * kPbObs * m.Kp + 3 * kPbObs);
    ICP_REQUIRE(m.Kp <= kPbThreads, "rank too large for the posterior build kernel (K <= 256)");
    int tiles = (ntiles + kPbThreads - 1) / kPbThreads;
    if (tiles <= 1) k_posterior_build<1><<<C, kPbThreads, smem, s>>>(m, o, d_M, d_b);
    else if (tiles <= 2) k_posterior_build<2><<<C, kPbThreads, smem, s>>>(m, o, d_M, d_b);
    else if (tiles <= 4) k_posterior_build<4><<<C, kPbThreads, smem, s>>>(m, o, d_M, d_b);
    else if (tiles <= 8) k_posterior_build<8><<<C, kPbThreads, smem, s>>>(m, o, d_M, d_b);
    else throw ArgError{"rank too large for the posterior build kernel"};
    ICP_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// Cholesky M = L L^T with the right-hand side carried as an extra row (forward solve for free),
// then back substitution L^T mu = y. One CTA per matrix, matrix resident in shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kChThreads = 256;

__global__ void __launch_bounds__(kChThreads) k_cholesky_solve(int K, int Kp, const double *__restrict__ M,
                                                               const double *__restrict__ bvec, double *__restrict__ L,
                                                               double *__restrict__ mu, const int *__restrict__ out_slot,
                                                               int *__restrict__ status) {
    extern __shared__ double sm[];
    const int ld = Kp + 1;        // odd stride: column accesses are bank-conflict free
    double *A = sm;               // [Kp + 1][ld]: rows 0..Kp-1 lower triangle of M, row Kp = b
    double *x = sm + (size_t)(Kp + 1) * ld;  // [Kp] solution
    __shared__ int bad;
    int c = blockIdx.x;
    const double *Mc = M + (size_t)c * Kp * Kp;
    if (threadIdx.x == 0) bad = 0;
    for (int e = threadIdx.x; e < Kp * Kp; e += kChThreads) {
        int i = e / Kp, j = e % Kp;
        A[i * ld + j] = Mc[e];
    }
    for (int j = threadIdx.x; j < Kp; j += kChThreads) A[Kp * ld + j] = bvec[(size_t)c * Kp + j];
    __syncthreads();
    const int R = Kp + 1;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int j = 0; j < Kp; j++) {
        double ajj = A[j * ld + j];
        if (!(ajj > 0.0)) { if (threadIdx.x == 0) bad = 1; }
        double inv = 1.0 / sqrt(ajj);
        __syncthreads();  // everyone has read the pivot before it is overwritten
        for (int i = j + threadIdx.x; i < R; i += kChThreads) A[i * ld + j] *= inv;  // i == j: a/sqrt(a) = sqrt(a)
        __syncthreads();
        // trailing update of the lower triangle (and of the b row)
        for (int i = j + 1 + ty; i < R; i += 16) {
            double lij = A[i * ld + j];
            int kmax = i < Kp ? i : Kp - 1;
            for (int k = j + 1 + tx; k <= kmax; k += 16) A[i * ld + k] = fma(-lij, A[k * ld + j], A[i * ld + k]);
        }
        __syncthreads();
    }
    // A row Kp now holds y = L^-1 b. Back substitution L^T x = y, column oriented.
    for (int j = threadIdx.x; j < Kp; j += kChThreads) x[j] = A[Kp * ld + j];
    __syncthreads();
    for (int i = Kp - 1; i >= 0; i--) {
        double xi = x[i] / A[i * ld + i];
        __syncthreads();
        if (threadIdx.x == 0) x[i] = xi;
        for (int k = threadIdx.x; k < i; k += kChThreads) x[k] = fma(-A[i * ld + k], xi, x[k]);
        __syncthreads();
    }
    int oc = out_slot ? out_slot[c] : c;
    double *Lc = L + (size_t)oc * Kp * Kp;
    bool isbad = bad != 0;
    for (int e = threadIdx.x; e < Kp * Kp; e += kChThreads) {
        int i = e / Kp, j = e % Kp;
        double v = j <= i ? A[i * ld + j] : 0.0;
        Lc[e] = isbad ? NAN : v;
    }
    for (int j = threadIdx.x; j < Kp; j += kChThreads) mu[(size_t)oc * Kp + j] = isbad ? NAN : x[j];
    if (threadIdx.x == 0 && status) status[c] = isbad ? 1 : 0;
    (void)K;
}

void launch_cholesky_solve(int C, int K, int Kp, const double *d_M, const double *d_b, double *d_L, double *d_mu,
                           const int *d_out_slot, int *d_status, cudaStream_t s) {
    ProfScope _ps(ST_CHOLESKY, s);
    if (C <= 0) return;
    size_t smem = sizeof(double) * ((size_t)(Kp + 1) * (Kp + 1) + Kp);
    ICP_REQUIRE(smem <= 227 * 1024, "rank too large for the shared-memory Cholesky (K <= 160)");
    ICP_CUDA(cudaFuncSetAttribute(k_cholesky_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cholesky_solve<<<C, kChThreads, smem, s>>>(K, Kp, d_M, d_b, d_L, d_mu, d_out_slot, d_status);
    ICP_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------
// propose / log-transition
// ---------------------------------------------------------------------------------------------------
// w = L^-T z by a single warp (registers + shuffles), L rows read from global memory (L2 resident).
// Requires Kp <= 32 * 8.
__device__ void warp_backsolve_LT(const double *__restrict__ Lc, int Kp, const double *z_sm, double *w_sm) {
    int lane = threadIdx.x & 31;
    double zr[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { int k = lane + 32 * q; zr[q] = k < Kp ? z_sm[k] : 0.0; }
    for (int i = Kp - 1; i >= 0; i--) {
        int owner = i & 31, slot = i >> 5;
        double zi = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) if (q == slot) zi = zr[q];
        double wi = __shfl_sync(0xffffffffu, zi, owner) / __ldg(&Lc[(size_t)i * Kp + i]);
        if (lane == 0) w_sm[i] = wi;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int k = lane + 32 * q;
            if (k < i) zr[q] = fma(-__ldg(&Lc[(size_t)i * Kp + k]), wi, zr[q]);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_propose(ModelDev m, double step, const double *__restrict__ theta,
                                                 const double *__restrict__ z, const double *__restrict__ L,
                                                 const double *__restrict__ mu, const int *__restrict__ slot,
                                                 double *__restrict__ theta_out) {
    extern __shared__ double sm[];
    const int Kp = m.Kp, K = m.K, Lt = K + kTheta0;
    double *sz = sm, *sw = sm + Kp;
    int c = blockIdx.x;
    int sl = slot ? slot[c] : c;
    const double *Lc = L + (size_t)sl * Kp * Kp, *muc = mu + (size_t)sl * Kp;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) sz[k] = k < K ? z[(size_t)c * K + k] : 0.0;
    __syncthreads();
    if (threadIdx.x < 32) warp_backsolve_LT(Lc, Kp, sz, sw);
    __syncthreads();
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) sz[k] = muc[k] + sw[k];   // v = mu + W z
    __syncthreads();
    const double *th = theta + (size_t)c * Lt;
    double *to = theta_out + (size_t)c * Lt;
    for (int j = threadIdx.x; j < Lt; j += blockDim.x) {
        if (j < kTheta0) { to[j] = th[j]; continue; }
        int jj = j - kTheta0;
        double acc = 0.0;  // (S v)_jj, S symmetric: read column-wise for coalescing
        for (int k = 0; k < Kp; k++) acc = fma(__ldg(&m.S[(size_t)k * Kp + jj]), sz[k], acc);
        to[j] = th[j] + (acc - th[j]) * step;                                       // :61-62
    }
}

void launch_propose(const ModelDev &m, int C, double step, const double *d_theta, const double *d_z,
                    const double *d_L, const double *d_mu, const int *d_slot, double *d_theta_out, cudaStream_t s) {
    if (C <= 0) return;
    ICP_REQUIRE(m.Kp <= 256, "rank too large for the propose kernel (K <= 256)");
    k_propose<<<C, 128, sizeof(double) * 2 * m.Kp, s>>>(m, step, d_theta, d_z, d_L, d_mu, d_slot, d_theta_out);
    ICP_CUDA(cudaGetLastError());
}

// |L^T d|^2 with d in shared memory; all threads of the block participate; red: >= 33 doubles
__device__ double block_quad_LT(const double *__restrict__ Lc, int Kp, const double *d_sm, double *red) {
    double part = 0.0;
    for (int j = threadIdx.x; j < Kp; j += blockDim.x) {
        double v = 0.0;
        for (int i = j; i < Kp; i++) v = fma(__ldg(&Lc[(size_t)i * Kp + j]), d_sm[i], v);
        part = fma(v, v, part);
    }
    return block_sum(part, red);
}

__global__ void __launch_bounds__(128) k_log_transition(int K, int Kp, double step, const double *__restrict__ from,
                                                        const double *__restrict__ to, const double *__restrict__ L,
                                                        const double *__restrict__ mu, const int *__restrict__ slot,
                                                        double *__restrict__ out) {
    extern __shared__ double sm[];
    double *sd = sm, *red = sm + Kp;
    __shared__ int differs;
    int c = blockIdx.x, Lt = K + kTheta0;
    int sl = slot ? slot[c] : c;
    const double *f = from + (size_t)c * Lt, *t = to + (size_t)c * Lt;
    if (threadIdx.x == 0) differs = 0;
    __syncthreads();
    if (threadIdx.x < kTheta0 && !(f[threadIdx.x] == t[threadIdx.x])) differs = 1;   // :72 only alpha may change
    for (int k = threadIdx.x; k < Kp; k += blockDim.x)
        sd[k] = k < K ? (f[kTheta0 + k] + ((t[kTheta0 + k] - f[kTheta0 + k]) / step)) - mu[(size_t)sl * Kp + k] : 0.0;  // :79
    __syncthreads();
    double q = block_quad_LT(L + (size_t)sl * Kp * Kp, Kp, sd, red);
    if (threadIdx.x == 0) out[c] = differs ? -INFINITY : -0.5 * (K * ICP_LOG_2PI + q);  // :83
}

void launch_log_transition(int C, int K, int Kp, double step, const double *d_from, const double *d_to,
                           const double *d_L, const double *d_mu, const int *d_slot, double *d_out, cudaStream_t s) {
    if (C <= 0) return;
    k_log_transition<<<C, 128, sizeof(double) * (Kp + 40), s>>>(K, Kp, step, d_from, d_to, d_L, d_mu, d_slot, d_out);
    ICP_CUDA(cudaGetLastError());
}

}  // namespace icp
