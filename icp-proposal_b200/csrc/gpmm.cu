// gpmm.cu - GPMM construction from analytic kernels: the step before the hot path.
//
//   icp_gpmm_kernel_matrix    the matrix-valued kernel of apps/femur/CreateGPModel.scala:68-83 evaluated on two point sets:
//                             k(x, y) = sum_t scale_t exp(-|x - y|^2 / sigma_t^2) A_t   (Scalismo GaussianKernel3D(sigma) * scale,
//                             DiagonalKernel3D -> A = I, the anisotropic base kernel -> A = baseMatrix)
//   icp_gpmm_nystrom_extend   LowRankGaussianProcess.approximateGPNystrom (CreateGPModel.scala:86): the eigenfunctions of the
//                             m-point kernel matrix extended to all N model points,
//                             phi_i(x) = sqrt(m) / w_i * k(x, X_m) v_i,   lambda_i = w_i / m
//
// The 3m x 3m symmetric eigenproblem between the two calls is left to the host (LAPACK), where the reference has it too
// (Breeze); everything that scales with the number of model points runs here.
#include <algorithm>

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

// thread / (x_i, y_j) pair: writes the 3 x 3 block (j fastest: a warp writes three runs of 768 contiguous bytes)
__global__ void __launch_bounds__(256) k_kernel_matrix(KernelTerms kt, int nx, const double *__restrict__ x, int ny,
                                                       const double *__restrict__ y, double *__restrict__ out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)nx * ny) return;
    const int i = (int)(g / ny), j = (int)(g % ny);
    double b[9];
    kernel_block(kt, x[3 * i] - y[3 * j], x[3 * i + 1] - y[3 * j + 1], x[3 * i + 2] - y[3 * j + 2], b);
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
__global__ void __launch_bounds__(128) k_nystrom_extend(KernelTerms kt, int N, const double *__restrict__ pts, int m,
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
            if (j < m) kernel_block(kt, px - nys[3 * j], py - nys[3 * j + 1], pz - nys[3 * j + 2], b);
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

extern "C" int32_t icp_gpmm_kernel_matrix(icp_ctx ctx, int32_t nx, const double *x, int32_t ny, const double *y,
                                          const icp_kernel_term *terms, int32_t n_terms, double *out) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(nx >= 1 && ny >= 1 && x && y && out, "bad point sets");
        const KernelTerms kt = pack_terms(terms, n_terms);
        cudaStream_t s = _ctx->stream;
        DevBuf<double> dx, dy, dout;
        dx.upload(x, (size_t)3 * nx, s);
        dy.upload(y, (size_t)3 * ny, s);
        const size_t total = (size_t)9 * nx * ny;
        dout.alloc(total);
        const long long pairs = (long long)nx * ny;
        k_kernel_matrix<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(kt, nx, dx.p, ny, dy.p, dout.p);
        ICP_CUDA(cudaGetLastError());
        ICP_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * total, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
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
        ICP_REQUIRE(rank >= 1 && rank <= 16 * kNyCols && rank <= 3 * m, "rank must be in [1, min(224, 3 m)]");
        for (int k = 0; k < rank; k++) ICP_REQUIRE(w[k] > 0.0, "eigenvalues of the kernel matrix must be positive");
        const KernelTerms kt = pack_terms(terms, n_terms);
        cudaStream_t s = _ctx->stream;
        DevBuf<double> dp, dn, dV, dw, dB;
        dp.upload(pts, (size_t)3 * N, s);
        dn.upload(nys_pts, (size_t)3 * m, s);
        dV.upload(V, (size_t)3 * m * rank, s);
        dw.upload(w, rank, s);
        dB.alloc((size_t)3 * N * rank);
        const size_t smem = sizeof(double) * ((size_t)kNyV * kNyY * 9 + (size_t)3 * kNyY * rank);
        ICP_CUDA(cudaFuncSetAttribute(k_nystrom_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nystrom_extend<<<(N + kNyV - 1) / kNyV, 128, smem, s>>>(kt, N, dp.p, m, dn.p, rank, dV.p, dw.p, dB.p);
        ICP_CUDA(cudaGetLastError());
        ICP_CUDA(cudaMemcpyAsync(basis, dB.p, sizeof(double) * (size_t)3 * N * rank, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        if (variance)
            for (int k = 0; k < rank; k++) variance[k] = w[k] / m;
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}
