// tc_i8.cu - 5th-generation tensor cores (tcgen05, INT8 operands, INT32 accumulators in TMEM) for the rank update of
// the ICP posterior, M = I + A^T A (interpolatedModel.posterior, NonRigidIcpProposal.scala:152).
//
// The path is FP64 in the reference and tcgen05 has no FP64 kind, so the FP64 product is EMULATED exactly enough for the
// 1e-5 contract by split-integer arithmetic (Ozaki scheme): every column of A is scaled by a model-level bound into
// (-1/2, 1/2) and written as four balanced base-255 digits d_0..d_3 in [-127, 127] (int8 operands),
//     A[r][j] / s_j = sum_k d_k[r][j] 255^-(k+1) + O(255^-4),
// the digit matrices are multiplied on the tensor cores with exact INT32 accumulation (606 rows x 127^2 x 4 products
// < 2^31), digit pairs of equal weight k + l share one accumulator, and the four accumulators are recombined in FP64:
//     (A^T A)[i][j] = s_i s_j sum_t 255^-(t+2) ACC_t[i][j],   ACC_t = sum_{k+l=t} D_k^T D_l,  t = 0..3.
// Measured on the reference's femur GPMM-100 (tools/ozaki_study.py): max relative error of M 2e-9, of the posterior mean
// 1.4e-8 - three orders inside the 1e-5 contract, but not the 1e-9 the FP64 tensor-pipe path (DMMA) delivers, so this
// path is selected explicitly (ICPCUDA_RANK_UPDATE=int8).
//
// Both MMA operands are the SAME shared-memory image of a digit matrix (rows = the MMA's K dimension, columns = M resp. N),
// i.e. MN-major operands, which tcgen05 accepts for 8-bit types: no transposition anywhere.
//   canonical MN-major layout, no swizzle: 16 consecutive columns of one row = one 16-byte line; 8 consecutive rows = one
//   128-byte core matrix; core matrices tile the columns with stride SBO and the 8-row groups with stride LBO.
#include <algorithm>
#include <vector>

#include "icp_device.cuh"
#include "icp_internal.h"

namespace icp {

// ---- PTX wrappers (CUDA 12.9, sm_100a) ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell): start address, leading (K-direction) and stride
// (MN-direction) byte offsets, all in units of 16 bytes
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor of kind::i8: S32 accumulator, signed 8-bit A and B, both MN-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 consecutive accumulator columns of this thread's TMEM lane (warp w of a warpgroup reads lanes 32 (w % 4) ..)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of the 16-byte line (row r, column chunk ch) inside a digit image: blocks of 32 rows (one MMA K step) of
// 4 KB, inside a block [row group of 8][chunk][row in group][16 bytes]  ->  SBO = 128, LBO = 1024
constexpr int kI8Cols = 128;          // columns of a digit image (MMA M; the first 112 are the MMA N)
constexpr int kI8N = 112;
constexpr int kI8BlockBytes = 32 * kI8Cols;
__device__ __forceinline__ uint32_t i8_line_offset(int r, int ch) {
    return (uint32_t)((r >> 5) * kI8BlockBytes + ((r & 31) >> 3) * 1024 + ch * 128 + (r & 7) * 16);
}

// ---- micro-benchmark / self-test: D = A^T A of one int8 matrix (rows x 128), `iters` passes over the same image ---------
__global__ void __launch_bounds__(128) k_i8_gram_test(int rows, const int8_t *__restrict__ A, int *__restrict__ D, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&s_tmem, 128);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    const int nblk = rows / 32;
    for (int e = tid; e < rows * (kI8Cols / 16); e += blockDim.x) {
        const int r = e / (kI8Cols / 16), ch = e % (kI8Cols / 16);
        *reinterpret_cast<uint4 *>(smem + i8_line_offset(r, ch)) = *reinterpret_cast<const uint4 *>(A + (size_t)r * kI8Cols + ch * 16);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_i8(128, kI8N);
        for (int it = 0; it < iters; it++)
            for (int b = 0; b < nblk; b++) {
                const uint64_t desc = umma_desc_noswizzle(smem_u32(smem) + b * kI8BlockBytes, 1024, 128);
                umma_i8(tmem, desc, desc, idesc, (it | b) ? 1u : 0u);
            }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (blockIdx.x == 0 && D) {
        for (int c0 = 0; c0 < kI8N; c0 += 16) {
            int v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
#pragma unroll
            for (int k = 0; k < 16; k++) D[(size_t)(32 * warp + lane) * kI8Cols + c0 + k] = v[k];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace icp

using namespace icp;

extern "C" int32_t icp_debug_i8_gram(icp_ctx ctx, int32_t rows, const int8_t *A, int32_t *D, int32_t iters, int32_t ctas,
                                     double *ms) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(ctx && A && rows >= 32 && rows % 32 == 0 && rows <= 1600 && iters >= 1 && ctas >= 1, "bad arguments");
        CtxLock lock(ctx);
        cudaStream_t s = ctx->stream;
        DevBuf<int8_t> dA;
        DevBuf<int> dD;
        dA.upload(A, (size_t)rows * kI8Cols, s);
        dD.alloc((size_t)kI8Cols * kI8Cols);
        ICP_CUDA(cudaMemsetAsync(dD.p, 0, sizeof(int) * kI8Cols * kI8Cols, s));
        const size_t smem = (size_t)rows * kI8Cols + 1024;
        ICP_CUDA(cudaFuncSetAttribute(k_i8_gram_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        ICP_CUDA(cudaEventCreate(&e0));
        ICP_CUDA(cudaEventCreate(&e1));
        k_i8_gram_test<<<ctas, 128, smem, s>>>(rows, dA.p, dD.p, 1);       // warm-up + the checked result
        ICP_CUDA(cudaGetLastError());
        if (D) ICP_CUDA(cudaMemcpyAsync(D, dD.p, sizeof(int) * kI8Cols * kI8Cols, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        ICP_CUDA(cudaEventRecord(e0, s));
        k_i8_gram_test<<<ctas, 128, smem, s>>>(rows, dA.p, nullptr, iters);
        ICP_CUDA(cudaGetLastError());
        ICP_CUDA(cudaEventRecord(e1, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        float t = 0;
        ICP_CUDA(cudaEventElapsedTime(&t, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (ms) *ms = t;
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

// =======================================================================================================================
// Rank update of the ICP posterior on the INT8 tensor cores:  Mp = block-packed lower triangle of I + A^T A (+ Gs term),
// b = A^T y~.  One CTA per chain-posterior, 512 threads, one CTA per SM (the four INT32 accumulators take all 512 TMEM
// columns). Stage = 32 observations (RPO rows each): thread (observation slot, 16-column chunk, half) gathers its 8 columns
// of the observation's three basis rows, whitens them in FP64 (the same arithmetic as the DMMA kernel's producers), adds
// its share of b, cuts every value into four base-255 digits and stores them into the stage's four digit images. One thread
// then issues the ten digit-pair MMAs per 32-row K step; tcgen05.commit releases the stage through an mbarrier, so the
// producers fill the other stage while the tensor core works. The epilogue reads the accumulators back with tcgen05.ld,
// recombines them in FP64 (Horner in 1/255) and writes the blocks k_cholesky_packed consumes.
// =======================================================================================================================
namespace icp {

constexpr int kRuThreads = 512;
constexpr int kRuObs = 32;                 // observations per stage
constexpr double kRuBase = 255.0;

template <int RPO>
__global__ void __launch_bounds__(kRuThreads, 1) k_rank_update_i8(ModelDev m, ObsDev o, const double *__restrict__ col_scale,
                                                                  double row_scale, const double *__restrict__ Gs,
                                                                  double gs_scale, double *__restrict__ Mp,
                                                                  double *__restrict__ bvec) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kStageRows = kRuObs * RPO;                    // 96 or 32
    constexpr int kImageBytes = kStageRows * kI8Cols;           // one digit image of a stage
    constexpr int kStageBytes = 4 * kImageBytes;
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_poison;
    const int Kp = m.Kp, c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *s_inv = reinterpret_cast<double *>(smem + 2 * kStageBytes);   // [128] 1 / s_j (0 beyond Kp)
    double *s_scl = s_inv + 128;                                           // [128] s_j
    double *s_b = s_scl + 128;                                             // [kRuObs][Kp] per-slot partial sums of b
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        s_poison = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = tid; j < 128; j += kRuThreads) {
        const double s = j < Kp ? col_scale[j] : 0.0;
        s_scl[j] = s;
        s_inv[j] = s > 0.0 ? 1.0 / s : 0.0;
    }
    // the images start as zeros: rows past the last observation and columns >= Kp are never written
    for (int e = tid; e < 2 * kStageBytes / 16; e += kRuThreads) reinterpret_cast<uint4 *>(smem)[e] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int nrows = o.nrows ? min(o.nrows[c], o.n) : o.n;    // observation slots in use
    const int nstage = (nrows + kRuObs - 1) / kRuObs;
    const int half = tid & 1, ch = (tid >> 1) & 7, os = tid >> 4;
    const int j0 = 16 * ch + 8 * half;                          // this thread's 8 columns
    const bool cols_live = j0 < Kp;
    const int *vid = o.vid + (size_t)c * o.n;
    const double *F = o.F + (size_t)c * o.n * 9, *yt = o.y + (size_t)c * o.n * 3;
    double bacc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) bacc[k] = 0.0;
    double inv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) inv[k] = cols_live ? s_inv[j0 + k] : 0.0;
    // frames travel one stage ahead in registers (they come from DRAM)
    double fn[12];
    int vn = -1;
    auto frame = [&](int st) {
        const int gi = st * kRuObs + os;
        vn = (st < nstage && gi < nrows) ? __ldg(&vid[gi]) : -1;
        if (vn >= 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) fn[k] = __ldg(F + (size_t)gi * 9 + k);
#pragma unroll
            for (int k = 0; k < 3; k++) fn[9 + k] = __ldg(yt + (size_t)gi * 3 + k);
        }
    };
    frame(0);
    const uint32_t idesc = umma_idesc_i8(128, kI8N);
    bool poison = false;
    for (int st = 0; st < nstage; st++) {
        const int buf = st & 1;
        unsigned char *stage = smem + buf * kStageBytes;
        // the MMAs that read this buffer two stages ago must have finished
        if (st >= 2) mbar_wait(&bar[buf], ((st >> 1) - 1) & 1);
        const int v = vn;
        double f[12];
#pragma unroll
        for (int k = 0; k < 12; k++) f[k] = fn[k];
        frame(st + 1);
        if (cols_live) {
            unsigned long long dig[RPO][4];
#pragma unroll
            for (int d = 0; d < RPO; d++)
#pragma unroll
                for (int q = 0; q < 4; q++) dig[d][q] = 0ull;
            if (v >= 0) {
                const double2 *q0 = reinterpret_cast<const double2 *>(m.Q + (size_t)3 * v * Kp + j0);
                const double2 *q1 = reinterpret_cast<const double2 *>(m.Q + ((size_t)3 * v + 1) * Kp + j0);
                const double2 *q2 = reinterpret_cast<const double2 *>(m.Q + ((size_t)3 * v + 2) * Kp + j0);
                double qa[8], qb[8], qc[8];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double2 a = __ldg(q0 + k), b = __ldg(q1 + k), cc = __ldg(q2 + k);
                    qa[2 * k] = a.x; qa[2 * k + 1] = a.y; qb[2 * k] = b.x; qb[2 * k + 1] = b.y; qc[2 * k] = cc.x; qc[2 * k + 1] = cc.y;
                }
                if (RPO == 1) {
                    // b += Q_i^T (F^T F y): w = F^T (F y)
                    const double w0 = f[0] * f[9] + f[3] * f[10] + f[6] * f[11], w1 = f[1] * f[9] + f[4] * f[10] + f[7] * f[11],
                                 w2 = f[2] * f[9] + f[5] * f[10] + f[8] * f[11];
#pragma unroll
                    for (int k = 0; k < 8; k++) bacc[k] = fma(w0, qa[k], fma(w1, qb[k], fma(w2, qc[k], bacc[k])));
                }
#pragma unroll
                for (int d = 0; d < RPO; d++) {
                    const double f0 = RPO == 1 ? f[0] * row_scale : f[3 * d], f1 = RPO == 1 ? f[1] * row_scale : f[3 * d + 1],
                                 f2 = RPO == 1 ? f[2] * row_scale : f[3 * d + 2];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const double a = f0 * qa[k] + f1 * qb[k] + f2 * qc[k];       // the DMMA producers' arithmetic
                        if (RPO == 3) bacc[k] = fma(f[9 + d], a, bacc[k]);
                        double t = a * inv[k];
                        if (!(fabs(t) < 0.5)) { poison = poison || (t != 0.0); t = 0.0; }   // out of the scaled range or NaN
                        // four balanced base-255 digits: two in FP64, the remainder (|r| <= 1/2, 2^-16 needed) in FP32
                        // (a remainder of exactly +-1/2 would round to +-128: clamped to +-127, the next digit absorbs it)
                        const int i0 = __double2int_rn(t * kRuBase);
                        double r = fma(t, kRuBase, -(double)i0);
                        const int i1 = max(-127, min(127, __double2int_rn(r * kRuBase)));
                        r = fma(r, kRuBase, -(double)i1);
                        float rf = (float)r;
                        const int i2 = max(-127, min(127, __float2int_rn(rf * 255.f)));
                        rf = fmaf(rf, 255.f, -(float)i2);
                        const int i3 = max(-127, min(127, __float2int_rn(rf * 255.f)));
                        dig[d][0] |= (unsigned long long)(unsigned char)i0 << (8 * k);
                        dig[d][1] |= (unsigned long long)(unsigned char)i1 << (8 * k);
                        dig[d][2] |= (unsigned long long)(unsigned char)i2 << (8 * k);
                        dig[d][3] |= (unsigned long long)(unsigned char)i3 << (8 * k);
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < RPO; d++) {
                const uint32_t off = i8_line_offset(os * RPO + d, ch) + 8 * half;
#pragma unroll
                for (int q = 0; q < 4; q++) *reinterpret_cast<unsigned long long *>(stage + q * kImageBytes + off) = dig[d][q];
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t sbase = smem_u32(stage);
#pragma unroll 1
            for (int ks = 0; ks < RPO; ks++) {
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int k = 0; k <= t; k++) {
                        const uint64_t da = umma_desc_noswizzle(sbase + k * kImageBytes + ks * kI8BlockBytes, 1024, 128);
                        const uint64_t db = umma_desc_noswizzle(sbase + (t - k) * kImageBytes + ks * kI8BlockBytes, 1024, 128);
                        umma_i8(tmem + 128 * t, da, db, idesc, (st | ks | k) ? 1u : 0u);
                    }
            }
            umma_commit(&bar[buf]);
        }
    }
    // b: per-slot partial sums, reduced in slot order (deterministic)
    if (cols_live) {
#pragma unroll
        for (int k = 0; k < 8; k++) s_b[os * Kp + j0 + k] = bacc[k];
    }
    if (poison) atomicOr(&s_poison, 1);
    // every issued MMA has completed once the last commit of each buffer has arrived
    if (nstage >= 1) mbar_wait(&bar[(nstage - 1) & 1], ((nstage - 1) >> 1) & 1);
    if (nstage >= 2) mbar_wait(&bar[(nstage - 2) & 1], ((nstage - 2) >> 1) & 1);
    tc_fence_after();
    __syncthreads();
    const bool bad = s_poison != 0;
    for (int j = tid; j < Kp; j += kRuThreads) {
        double s = 0.0;
        for (int sl = 0; sl < kRuObs; sl++) s += s_b[sl * Kp + j];
        bvec[(size_t)c * Kp + j] = bad ? NAN : s;
    }
    // epilogue: thread = (row i = TMEM lane, column chunks cg, cg + 4)
    {
        const int i = 32 * (warp & 3) + lane, cg = warp >> 2, NB = Kp >> 3, ntri = NB * (NB + 1) / 2;
        double *Mc = Mp + (size_t)c * ntri * 64;
        const double si = i < Kp ? s_scl[i] : 0.0;
        const double w2 = 1.0 / (kRuBase * kRuBase), cinv = 1.0 / kRuBase;
        for (int chn = cg; chn < 7; chn += 4) {
            const int jc = 16 * chn;
            if (nstage == 0 || i >= Kp || jc >= Kp || (jc >> 3) > (i >> 3)) {
                // (all lanes of a warp must still issue the TMEM loads together: the branch is evaluated per lane below)
            }
            int a0[16], a1[16], a2[16], a3[16];
            const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + jc;
            if (nstage > 0) { tmem_ld16(ta, a0); tmem_ld16(ta + 128, a1); tmem_ld16(ta + 256, a2); tmem_ld16(ta + 384, a3); }
            else {
#pragma unroll
                for (int k = 0; k < 16; k++) a0[k] = a1[k] = a2[k] = a3[k] = 0;
            }
            if (i >= Kp) continue;
            const int bi = i >> 3;
#pragma unroll
            for (int hb = 0; hb < 2; hb++) {
                const int bj = (jc >> 3) + hb;
                if (bj > bi || 8 * bj >= Kp) continue;
                double *dst = Mc + (size_t)((bi * (bi + 1) >> 1) + bj) * 64 + (i & 7) * 8;
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    double v[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int kk = 8 * hb + k + u, j = jc + kk;
                        const double acc = ((((double)a3[kk] * cinv + (double)a2[kk]) * cinv + (double)a1[kk]) * cinv + (double)a0[kk]) * w2;
                        double val = si * s_scl[j] * acc + (i == j ? 1.0 : 0.0);
                        if (RPO == 1) val = fma(__ldg(Gs + (size_t)i * Kp + j), gs_scale, val);
                        v[u] = bad ? NAN : val;
                    }
                    *reinterpret_cast<double2 *>(dst + k) = make_double2(v[0], v[1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// false: outside the kernel's domain (nothing launched) - the caller takes the FP64 tensor-pipe path
bool launch_rank_update_i8(const ModelDev &m, int C, const ObsDev &o, const GramFast *gf, const double *d_col_scale,
                           double *d_Mp, double *d_b, cudaStream_t s) {
    if (C <= 0 || m.Kp > kI8N || (m.Kp & 7)) return false;
    ProfScope _ps(ST_POSTERIOR_BUILD, s);
    const size_t tail = sizeof(double) * (256 + (size_t)kRuObs * m.Kp) + 1024;
    if (gf) {
        const size_t smem = (size_t)2 * 4 * kRuObs * kI8Cols + tail;
        // one CTA per SM (all 512 TMEM columns): ask for more than half of the shared memory so that no second CTA is
        // scheduled only to wait for the allocation
        const size_t ask = std::max<size_t>(smem, 116 * 1024);
        ICP_CUDA(cudaFuncSetAttribute(k_rank_update_i8<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ask));
        k_rank_update_i8<1><<<C, kRuThreads, ask, s>>>(m, o, d_col_scale, gf->row_scale, gf->Gs, gf->gs_scale, d_Mp, d_b);
    } else {
        const size_t smem = (size_t)2 * 4 * kRuObs * 3 * kI8Cols + tail;
        ICP_CUDA(cudaFuncSetAttribute(k_rank_update_i8<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_rank_update_i8<3><<<C, kRuThreads, smem, s>>>(m, o, d_col_scale, 1.0, nullptr, 0.0, d_Mp, d_b);
    }
    ICP_CUDA(cudaGetLastError());
    return true;
}

}  // namespace icp
