// gpmm.cu - GPMM construction from analytic kernels: the step before the hot path.
//
//   icp_gpmm_kernel_matrix    the matrix-valued kernel of apps/femur/CreateGPModel.scala:68-83 evaluated on two point sets:
//                             k(x, y) = sum_t scale_t exp(-|x - y|^2 / sigma_t^2) A_t   (Scalismo GaussianKernel3D(sigma) * scale,
//                             DiagonalKernel3D -> A = I, the anisotropic base kernel -> A = baseMatrix)
//   icp_gpmm_nystrom_extend   LowRankGaussianProcess.approximateGPNystrom (CreateGPModel.scala:86): the eigenfunctions of the
//                             m-point kernel matrix extended to all N model points,
//                             phi_i(x) = sqrt(m) / w_i * k(x, X_m) v_i,   lambda_i = w_i / m
//
//   icp_gpmm_eigen_psd        the symmetric eigenproblem of the m-point kernel matrix between the two: one-sided Jacobi
//                             (Hestenes), one launch per round of the round-robin pair schedule
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "icp_internal.h"
#include "icp_device.cuh"

namespace icp {

constexpr int kMaxKernelTerms = 8;
struct KernelTerms {
    int n;
    double scale[kMaxKernelTerms], inv_s2[kMaxKernelTerms], A[kMaxKernelTerms][9];
};

__device__ __forceinline__ void kernel_block(const KernelTerms &kt, double dx, double dy, double dz, double (&b)[9]) {
    const double d2 = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int k = 0; k < 9; k++) b[k] = 0.0;
    for (int t = 0; t < kt.n; t++) {
        const double g = kt.scale[t] * exp(-d2 * kt.inv_s2[t]);
#pragma unroll
        for (int k = 0; k < 9; k++) b[k] = fma(g, kt.A[t][k], b[k]);
    }
}

struct GaussSpec {
    KernelTerms kt;
    __device__ __forceinline__ void block(int, int, double px, double py, double pz, double qx, double qy, double qz, double (&b)[9]) const {
        kernel_block(kt, px - qx, py - qy, pz - qz, b);
    }
};

// ---- the face kernel of apps/bfm/FaceKernel.scala --------------------------------------------------------------------
// SpatiallyVaryingMultiscaleKernel (:26-55): k(x, y) = sum_l scale_l w_l(x) w_l(y) B3(2^level_l x, 2^level_l y) I_3 with the
// order-3 B-spline kernel of Scalismo's BSplineKernel[_3D](order = 3, scale = 0) [S-recall]:
//     B3(a, b) = prod_d sum_k beta3(a_d - k) beta3(b_d - k),   k over the integers with both factors non-zero,
// and the symmetrisation about the plane x = 0 of FaceKernel (:58-100):
//     k_face(x, y) = 0.7 (I k(x, y) + Ibar k(x, ybar)) + 0.3 k(x, y),   ybar = (-y_x, y_y, y_z), Ibar = diag(-1, 1, 1).
// The region weights w_l come from the face mask (FaceMask.computeSmoothedRegions, data that is not part of the reference
// checkout): the caller passes them per level and point, for y also at the mirrored points.
constexpr int kMaxFaceLevels = 8;
struct FaceSpec {
    int n, nx, ny;
    double mul[kMaxFaceLevels], scale[kMaxFaceLevels];     // 2^level, LevelWithScale.scale
    double sym, plain;
    const double *wx, *wy, *wyb;                           // [n][nx], [n][ny], [n][ny] or null (= 1)
    static __device__ __forceinline__ double beta3(double t) {
        t = fabs(t);
        if (t >= 2.0) return 0.0;
        if (t >= 1.0) { const double u = 2.0 - t; return u * u * u * (1.0 / 6.0); }
        return 2.0 / 3.0 - t * t + 0.5 * t * t * t;
    }
    static __device__ __forceinline__ double lattice_sum(double a, double b) {
        const int kl = (int)ceil(fmax(a, b) - 2.0), ku = (int)floor(fmin(a, b) + 2.0);
        double s = 0.0;
        for (int k = kl; k <= ku; k++) s += beta3(a - (double)k) * beta3(b - (double)k);
        return s;
    }
    __device__ __forceinline__ void block(int i, int j, double px, double py, double pz, double qx, double qy, double qz, double (&b)[9]) const {
        double k = 0.0, kb = 0.0;
        for (int l = 0; l < n; l++) {
            const double c = mul[l], wi = wx ? wx[(size_t)l * nx + i] : 1.0;
            const double syz = lattice_sum(c * py, c * qy) * lattice_sum(c * pz, c * qz) * scale[l] * wi;
            k = fma(syz * (wy ? wy[(size_t)l * ny + j] : 1.0), lattice_sum(c * px, c * qx), k);
            if (sym != 0.0) kb = fma(syz * (wyb ? wyb[(size_t)l * ny + j] : 1.0), lattice_sum(c * px, -c * qx), kb);
        }
#pragma unroll
        for (int e = 0; e < 9; e++) b[e] = 0.0;
        b[0] = sym * (k - kb) + plain * k;
        b[4] = b[8] = sym * (k + kb) + plain * k;
    }
};

// thread / (x_i, y_j) pair: writes the 3 x 3 block (j fastest: a warp writes three runs of 768 contiguous bytes)
template <class KS>
__global__ void __launch_bounds__(256) k_kernel_matrix(KS kt, int nx, const double *__restrict__ x, int ny,
                                                       const double *__restrict__ y, double *__restrict__ out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)nx * ny) return;
    const int i = (int)(g / ny), j = (int)(g % ny);
    double b[9];
    kt.block(i, j, x[3 * i], x[3 * i + 1], x[3 * i + 2], y[3 * j], y[3 * j + 1], y[3 * j + 2], b);
    const size_t ld = (size_t)3 * ny;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) out[((size_t)3 * i + r) * ld + 3 * j + c] = b[3 * r + c];
}

// Nystrom extension. CTA = 8 model points x all `rank` columns, 128 threads: thread (v, cg) owns rows 3 v .. 3 v + 2 and the
// columns cg + 16 u. Per chunk of 16 Nystrom points the CTA first evaluates its 8 x 16 kernel blocks (one per thread, the
// exp() is the expensive part and is done once), stages the 48 matching rows of V, then accumulates.
constexpr int kNyV = 8, kNyY = 16, kNyCols = 14;   // 16 x 14 = 224 columns at most
template <class KS>
__global__ void __launch_bounds__(128) k_nystrom_extend(KS kt, int N, const double *__restrict__ pts, int m,
                                                        const double *__restrict__ nys, int rank, const double *__restrict__ V,
                                                        const double *__restrict__ w, double *__restrict__ basis) {
    extern __shared__ double sm[];
    double *sblk = sm;                           // [kNyV][kNyY][9]
    double *sV = sm + kNyV * kNyY * 9;           // [3 kNyY][rank]
    const int tid = threadIdx.x, v = tid >> 4, cg = tid & 15;
    const int i = blockIdx.x * kNyV + v;
    double acc[3][kNyCols];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int u = 0; u < kNyCols; u++) acc[r][u] = 0.0;
    const int ic = i < N ? i : N - 1;
    const double px = pts[3 * ic], py = pts[3 * ic + 1], pz = pts[3 * ic + 2];
    for (int y0 = 0; y0 < m; y0 += kNyY) {
        {   // this thread's block: model point v, Nystrom point y0 + cg
            const int j = y0 + cg;
            double b[9];
            if (j < m) kt.block(ic, j, px, py, pz, nys[3 * j], nys[3 * j + 1], nys[3 * j + 2], b);
            else {
#pragma unroll
                for (int k = 0; k < 9; k++) b[k] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < 9; k++) sblk[(v * kNyY + cg) * 9 + k] = b[k];
        }
        for (int e = tid; e < 3 * kNyY * rank; e += blockDim.x) {
            const int row = e / rank, col = e - row * rank;
            sV[e] = 3 * y0 + row < 3 * m ? __ldg(V + (size_t)(3 * y0 + row) * rank + col) : 0.0;
        }
        __syncthreads();
#pragma unroll 2
        for (int yy = 0; yy < kNyY; yy++) {
            double b[9];
#pragma unroll
            for (int k = 0; k < 9; k++) b[k] = sblk[(v * kNyY + yy) * 9 + k];
#pragma unroll
            for (int u = 0; u < kNyCols; u++) {
                const int col = cg + 16 * u;
                if (col < rank) {
                    const double v0 = sV[(3 * yy) * rank + col], v1 = sV[(3 * yy + 1) * rank + col], v2 = sV[(3 * yy + 2) * rank + col];
#pragma unroll
                    for (int r = 0; r < 3; r++) acc[r][u] = fma(b[3 * r], v0, fma(b[3 * r + 1], v1, fma(b[3 * r + 2], v2, acc[r][u])));
                }
            }
        }
        __syncthreads();
    }
    if (i >= N) return;
    const double sq = sqrt((double)m);
#pragma unroll
    for (int u = 0; u < kNyCols; u++) {
        const int col = cg + 16 * u;
        if (col < rank) {
            const double s = sq / w[col];
#pragma unroll
            for (int r = 0; r < 3; r++) basis[((size_t)3 * i + r) * rank + col] = acc[r][u] * s;
        }
    }
}

// ---- symmetric positive semi-definite eigen-decomposition: one-sided Jacobi (Hestenes) ----------------------------------
// G = A V is kept column by column (A symmetric: its rows are its columns); a rotation of the column pair (i, j) makes
// g_i and g_j orthogonal. At convergence the columns of V are the eigenvectors and |g_k| the eigenvalues (A is PSD).
// One launch = one round of the round-robin tournament (circle method): n / 2 disjoint pairs, one CTA each.
__global__ void __launch_bounds__(128) k_jacobi_round(int n, int npad, int round, double *__restrict__ G, double *__restrict__ V,
                                                      double tol, double floor2, int *__restrict__ rotated) {
    const int k = blockIdx.x, np1 = npad - 1;
    int i, j;
    if (k == 0) { i = np1; j = round; }
    else { i = (round + k) % np1; j = (round - k + np1) % np1; }
    if (i > j) { const int t = i; i = j; j = t; }
    if (j >= n) return;   // padding player
    double *gi = G + (size_t)i * n, *gj = G + (size_t)j * n, *vi = V + (size_t)i * n, *vj = V + (size_t)j * n;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double x = gi[r], y = gj[r];
        a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
    }
    __shared__ double red[40];
    __shared__ double rot[2];
    a = block_sum(a, red);
    b = block_sum(b, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
        double cs = 1.0, sn = 0.0;
        // columns whose norm has sunk below the rounding noise of the large ones are null vectors: never rotated
        if (a > floor2 && b > floor2 && fabs(c) > tol * sqrt(a * b)) {
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cs = 1.0 / sqrt(1.0 + t * t);
            sn = cs * t;
            *rotated = 1;
        }
        rot[0] = cs; rot[1] = sn;
    }
    __syncthreads();
    const double cs = rot[0], sn = rot[1];
    if (sn == 0.0) return;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const double x = gi[r], y = gj[r];
        gi[r] = cs * x - sn * y; gj[r] = sn * x + cs * y;
        const double p = vi[r], q = vj[r];
        vi[r] = cs * p - sn * q; vj[r] = sn * p + cs * q;
    }
}

__global__ void k_identity(int n, double *__restrict__ V) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < (long long)n * n) V[g] = (g / n == g % n) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(128) k_column_norms(int n, const double *__restrict__ G, double *__restrict__ w) {
    __shared__ double red[40];
    const double *g = G + (size_t)blockIdx.x * n;
    double a = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) a = fma(g[r], g[r], a);
    a = block_sum(a, red);
    if (threadIdx.x == 0) w[blockIdx.x] = sqrt(a);
}

// d_A: n x n symmetric PSD (destroyed: becomes G). d_V: n x n, column k = eigenvector k. d_w: n eigenvalues (unsorted).
static int jacobi_eigen_psd(int n, double *d_A, double *d_V, double *d_w, double frob, cudaStream_t s) {
    const int npad = (n + 1) & ~1;
    const double eps = 2.220446049250313e-16, tol = std::max(1e-12, 32.0 * n * eps)   /* the orthogonality a sweep of n rotations per column can hold */, floor2 = (n * eps * frob) * (n * eps * frob);
    DevBuf<int> flag;
    flag.alloc(1);
    k_identity<<<(unsigned)(((long long)n * n + 255) / 256), 256, 0, s>>>(n, d_V);
    ICP_CUDA(cudaGetLastError());
    // one sweep = npad - 1 dependent launches: captured once in a CUDA graph and replayed per sweep (launch-bound otherwise)
    auto enqueue_sweep = [&]() {
        ICP_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
        for (int round = 0; round < npad - 1; round++) k_jacobi_round<<<npad / 2, 128, 0, s>>>(n, npad, round, d_A, d_V, tol, floor2, flag.p);
        ICP_CUDA(cudaGetLastError());
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        bool ok = true;
        try { enqueue_sweep(); } catch (...) { ok = false; }
        if (cudaStreamEndCapture(s, &graph) != cudaSuccess || !ok || !graph) { if (graph) cudaGraphDestroy(graph); graph = nullptr; }
        if (graph && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
    }
    cudaGetLastError();
    int sweeps = 0;
    for (; sweeps < 40; sweeps++) {
        if (exec) ICP_CUDA(cudaGraphLaunch(exec, s));
        else enqueue_sweep();
        int h = 0;
        ICP_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        if (!h) break;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    k_column_norms<<<n, 128, 0, s>>>(n, d_A, d_w);
    ICP_CUDA(cudaGetLastError());
    return sweeps;
}

static KernelTerms pack_terms(const icp_kernel_term *terms, int n_terms) {
    ICP_REQUIRE(terms != nullptr && n_terms >= 1 && n_terms <= kMaxKernelTerms, "between 1 and 8 kernel terms");
    KernelTerms kt;
    kt.n = n_terms;
    for (int t = 0; t < n_terms; t++) {
        ICP_REQUIRE(terms[t].sigma > 0.0, "kernel sigma must be positive");
        kt.scale[t] = terms[t].scale;
        kt.inv_s2[t] = 1.0 / (terms[t].sigma * terms[t].sigma);
        for (int k = 0; k < 9; k++) kt.A[t][k] = terms[t].A[k];
    }
    return kt;
}

}  // namespace icp

using namespace icp;

template <class KS>
static void run_kernel_matrix(icp_ctx ctx, const KS &ks, int nx, const double *d_x, int ny, const double *d_y, double *out) {
    cudaStream_t s = ctx->stream;
    const size_t total = (size_t)9 * nx * ny;
    DevBuf<double> dout;
    dout.alloc(total);
    const long long pairs = (long long)nx * ny;
    k_kernel_matrix<KS><<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(ks, nx, d_x, ny, d_y, dout.p);
    ICP_CUDA(cudaGetLastError());
    ICP_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * total, cudaMemcpyDeviceToHost, s));
    ICP_CUDA(cudaStreamSynchronize(s));
}

template <class KS>
static void run_nystrom_extend(icp_ctx ctx, const KS &ks, int N, const double *d_pts, int m, const double *d_nys, int rank, const double *V,
                               const double *w, double *basis, double *variance) {
    ICP_REQUIRE(rank >= 1 && rank <= 16 * kNyCols && rank <= 3 * m, "rank must be in [1, min(224, 3 m)]");
    for (int k = 0; k < rank; k++) ICP_REQUIRE(w[k] > 0.0, "eigenvalues of the kernel matrix must be positive");
    cudaStream_t s = ctx->stream;
    DevBuf<double> dV, dw, dB;
    dV.upload(V, (size_t)3 * m * rank, s);
    dw.upload(w, rank, s);
    dB.alloc((size_t)3 * N * rank);
    const size_t smem = sizeof(double) * ((size_t)kNyV * kNyY * 9 + (size_t)3 * kNyY * rank);
    ICP_CUDA(cudaFuncSetAttribute(k_nystrom_extend<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nystrom_extend<KS><<<(N + kNyV - 1) / kNyV, 128, smem, s>>>(ks, N, d_pts, m, d_nys, rank, dV.p, dw.p, dB.p);
    ICP_CUDA(cudaGetLastError());
    ICP_CUDA(cudaMemcpyAsync(basis, dB.p, sizeof(double) * (size_t)3 * N * rank, cudaMemcpyDeviceToHost, s));
    ICP_CUDA(cudaStreamSynchronize(s));
    if (variance)
        for (int k = 0; k < rank; k++) variance[k] = w[k] / m;
}

static FaceSpec pack_face(const icp_face_kernel *k, int nx, int ny) {
    ICP_REQUIRE(k != nullptr && k->n_levels >= 1 && k->n_levels <= kMaxFaceLevels, "between 1 and 8 kernel levels");
    ICP_REQUIRE(std::isfinite(k->symmetric_weight) && std::isfinite(k->plain_weight), "kernel weights must be finite");
    FaceSpec f{};
    f.n = k->n_levels; f.nx = nx; f.ny = ny; f.sym = k->symmetric_weight; f.plain = k->plain_weight;
    for (int l = 0; l < k->n_levels; l++) {
        ICP_REQUIRE(k->level[l] >= -60 && k->level[l] <= 60 && std::isfinite(k->scale[l]), "bad kernel level");
        f.mul[l] = std::ldexp(1.0, k->level[l]);
        f.scale[l] = k->scale[l];
    }
    return f;
}

extern "C" int32_t icp_gpmm_kernel_matrix(icp_ctx ctx, int32_t nx, const double *x, int32_t ny, const double *y,
                                          const icp_kernel_term *terms, int32_t n_terms, double *out) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(nx >= 1 && ny >= 1 && x && y && out, "bad point sets");
        const GaussSpec ks{pack_terms(terms, n_terms)};
        DevBuf<double> dx, dy;
        dx.upload(x, (size_t)3 * nx, _ctx->stream);
        dy.upload(y, (size_t)3 * ny, _ctx->stream);
        run_kernel_matrix(_ctx, ks, nx, dx.p, ny, dy.p, out);
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_gpmm_nystrom_extend(icp_ctx ctx, int32_t N, const double *pts, int32_t m, const double *nys_pts,
                                           const icp_kernel_term *terms, int32_t n_terms, int32_t rank, const double *V,
                                           const double *w, double *basis, double *variance) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(N >= 1 && m >= 1 && pts && nys_pts && V && w && basis, "bad argument");
        const GaussSpec ks{pack_terms(terms, n_terms)};
        DevBuf<double> dp, dn;
        dp.upload(pts, (size_t)3 * N, _ctx->stream);
        dn.upload(nys_pts, (size_t)3 * m, _ctx->stream);
        run_nystrom_extend(_ctx, ks, N, dp.p, m, dn.p, rank, V, w, basis, variance);
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_gpmm_face_kernel_matrix(icp_ctx ctx, int32_t nx, const double *x, const double *wx, int32_t ny, const double *y,
                                               const double *wy, const double *wy_mirror, const icp_face_kernel *kernel, double *out) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(nx >= 1 && ny >= 1 && x && y && out, "bad point sets");
        FaceSpec ks = pack_face(kernel, nx, ny);
        cudaStream_t s = _ctx->stream;
        DevBuf<double> dx, dy, dwx, dwy, dwb;
        dx.upload(x, (size_t)3 * nx, s);
        dy.upload(y, (size_t)3 * ny, s);
        if (wx) { dwx.upload(wx, (size_t)ks.n * nx, s); ks.wx = dwx.p; }
        if (wy) { dwy.upload(wy, (size_t)ks.n * ny, s); ks.wy = dwy.p; }
        if (wy_mirror) { dwb.upload(wy_mirror, (size_t)ks.n * ny, s); ks.wyb = dwb.p; }
        run_kernel_matrix(_ctx, ks, nx, dx.p, ny, dy.p, out);
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_gpmm_face_nystrom_extend(icp_ctx ctx, int32_t N, const double *pts, const double *w_pts, int32_t m,
                                                const double *nys_pts, const double *w_nys, const double *w_nys_mirror,
                                                const icp_face_kernel *kernel, int32_t rank, const double *V, const double *w, double *basis,
                                                double *variance) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(N >= 1 && m >= 1 && pts && nys_pts && V && w && basis, "bad argument");
        FaceSpec ks = pack_face(kernel, N, m);
        cudaStream_t s = _ctx->stream;
        DevBuf<double> dp, dn, dwx, dwy, dwb;
        dp.upload(pts, (size_t)3 * N, s);
        dn.upload(nys_pts, (size_t)3 * m, s);
        if (w_pts) { dwx.upload(w_pts, (size_t)ks.n * N, s); ks.wx = dwx.p; }
        if (w_nys) { dwy.upload(w_nys, (size_t)ks.n * m, s); ks.wy = dwy.p; }
        if (w_nys_mirror) { dwb.upload(w_nys_mirror, (size_t)ks.n * m, s); ks.wyb = dwb.p; }
        run_nystrom_extend(_ctx, ks, N, dp.p, m, dn.p, rank, V, w, basis, variance);
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

// Leading eigenpairs of a symmetric positive semi-definite matrix (the Nystrom kernel matrix) on the device.
extern "C" int32_t icp_gpmm_eigen_psd(icp_ctx ctx, int32_t n, const double *A, int32_t n_top, double *w, double *V) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(n >= 1 && A && w && V && n_top >= 1 && n_top <= n, "bad argument");
        cudaStream_t s = _ctx->stream;
        DevBuf<double> dA, dV, dw;
        dA.upload(A, (size_t)n * n, s);
        dV.alloc((size_t)n * n);
        dw.alloc(n);
        double frob = 0.0;
        for (size_t e = 0; e < (size_t)n * n; e++) frob += A[e] * A[e];
        const int sweeps = jacobi_eigen_psd(n, dA.p, dV.p, dw.p, std::sqrt(frob), s);
        if (getenv("ICPCUDA_VERBOSE")) fprintf(stderr, "icp_gpmm_eigen_psd: n = %d, %d sweeps\n", n, sweeps + 1);
        std::vector<double> hw(n), hV((size_t)n * n);
        ICP_CUDA(cudaMemcpyAsync(hw.data(), dw.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaMemcpyAsync(hV.data(), dV.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        std::vector<int> order(n);
        for (int k = 0; k < n; k++) order[k] = k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hw[a] > hw[b]; });
        for (int k = 0; k < n_top; k++) {
            const double *col = hV.data() + (size_t)order[k] * n;
            int big = 0;
            for (int r = 1; r < n; r++)
                if (fabs(col[r]) > fabs(col[big])) big = r;
            const double sgn = col[big] < 0.0 ? -1.0 : 1.0;   // sign convention: the largest-magnitude entry is positive
            w[k] = hw[order[k]];
            for (int r = 0; r < n; r++) V[(size_t)r * n_top + k] = sgn * col[r];
        }
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}
