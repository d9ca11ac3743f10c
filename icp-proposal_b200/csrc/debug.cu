// debug.cu - profiling support: per-stage CUDA-event timers and FP64 peak micro-benchmarks
// (the roofline denominators bench.py reports for the FP64 kernels; MEASURED_PEAKS.json only
// carries HBM bandwidth and bf16 tensor throughput).
#include <algorithm>

#include "icp_internal.h"

namespace icp {

thread_local Profiler *g_prof = nullptr;

ProfScope::ProfScope(int stage, cudaStream_t stream) : s(stream) {
    if (!g_prof) return;
    Profiler::Rec r;
    r.stage = stage;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, s);
    idx = (int)g_prof->recs.size();
    g_prof->recs.push_back(r);
}
ProfScope::~ProfScope() {
    if (idx >= 0 && g_prof) cudaEventRecord(g_prof->recs[idx].e1, s);
}
void Profiler::collect() {
    for (auto &r : recs) {
        float ms_ = 0;
        if (cudaEventElapsedTime(&ms_, r.e0, r.e1) == cudaSuccess) { ms[r.stage] += ms_; launches[r.stage]++; }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    recs.clear();
}

__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters) {
    double a[8], x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// L2 -> SM read bandwidth: every thread streams 16-byte vectors of a working set that stays resident in L2 (a few MB),
// L1 bypassed (ld.global.cg), four independent loads in flight per thread. The denominator of the closest-point
// traversal's roofline: its tree and triangles (~0.5 MB) are cache resident, so HBM bandwidth does not bound it.
__global__ void __launch_bounds__(256) k_l2_read(const float4 *__restrict__ buf, unsigned long long n_vec, int iters,
                                                 float4 *__restrict__ sink) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x,
                             stride = (unsigned long long)gridDim.x * blockDim.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; it++) {
        for (unsigned long long i = tid; i + 3 * stride < n_vec; i += 4 * stride) {
            const float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride),
                         d = __ldcg(buf + i + 3 * stride);
            acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
            acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
        }
    }
    if (acc.x == 12345.678f) sink[tid] = acc;
}

}  // namespace icp

using namespace icp;

extern "C" int32_t icp_debug_l2_bandwidth(icp_ctx ctx, int64_t working_set_bytes, double *gbps) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(ctx && gbps && working_set_bytes >= (1 << 20), "bad argument");
        CtxLock lock(ctx);
        cudaStream_t s = ctx->stream;
        const int threads = 256;
        unsigned long long n_vec = (unsigned long long)working_set_bytes / 16;
        int blocks = ctx->sm_count * 8;
        while (blocks > ctx->sm_count && 4ull * threads * blocks > n_vec) blocks -= ctx->sm_count;
        const unsigned long long stride = (unsigned long long)threads * blocks;
        n_vec = n_vec / (4 * stride) * (4 * stride);          // whole rounds: every thread issues the same number of loads
        ICP_REQUIRE(n_vec > 0, "working set smaller than one round of the grid");
        DevBuf<float4> buf, sink;
        buf.alloc(n_vec); sink.alloc(stride);
        ICP_CUDA(cudaMemsetAsync(buf.p, 0, sizeof(float4) * n_vec, s));
        cudaEvent_t e0, e1;
        ICP_CUDA(cudaEventCreate(&e0));
        ICP_CUDA(cudaEventCreate(&e1));
        const int iters = (int)std::max<unsigned long long>(8, (4ull << 30) / (n_vec * 16));   // ~4 GB of reads per launch
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            ICP_CUDA(cudaEventRecord(e0, s));
            k_l2_read<<<blocks, threads, 0, s>>>(buf.p, n_vec, iters, sink.p);
            ICP_CUDA(cudaGetLastError());
            ICP_CUDA(cudaEventRecord(e1, s));
            ICP_CUDA(cudaStreamSynchronize(s));
            float ms = 0;
            ICP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double g = (double)n_vec * 16.0 * iters / (ms * 1e-3) / 1e9;
            if (rep > 0 && g > best) best = g;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *gbps = best;
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_debug_fp64_peak(icp_ctx ctx, double out[2]) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(ctx && out, "null argument");
        CtxLock lock(ctx);
        cudaStream_t s = ctx->stream;
        DevBuf<double> d;
        d.alloc(4);
        cudaEvent_t e0, e1;
        ICP_CUDA(cudaEventCreate(&e0));
        ICP_CUDA(cudaEventCreate(&e1));
        const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
        for (int which = 0; which < 2; which++) {
            double best = 0;
            for (int rep = 0; rep < 4; rep++) {
                ICP_CUDA(cudaEventRecord(e0, s));
                if (which == 0) k_dfma_peak<<<blocks, threads, 0, s>>>(d.p, iters);
                else k_dmma_peak<<<blocks, threads, 0, s>>>(d.p, iters);
                ICP_CUDA(cudaGetLastError());
                ICP_CUDA(cudaEventRecord(e1, s));
                ICP_CUDA(cudaStreamSynchronize(s));
                float ms = 0;
                ICP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
                double flops = which == 0 ? 2.0 * 8 * iters * (double)blocks * threads
                                          : 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
                double tf = flops / (ms * 1e-3) / 1e12;
                if (rep > 0 && tf > best) best = tf;
            }
            out[which] = best;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_debug_time_closest_point(icp_target t, int64_t nq, const double *q_dev, int32_t *tri_dev,
                                                double *cp_dev, double *d2_dev, int32_t iters, double *ms) {
    icp_ctx _ctx = t ? t->ctx : nullptr;
    try {
        ICP_REQUIRE(_ctx && q_dev && ms && nq > 0 && iters > 0, "bad argument");
        CtxLock lock(_ctx);
        cudaStream_t s = _ctx->stream;
        NearestArgs a;
        a.bvh = &t->tri_bvh; a.prim_data = t->tri_data.p; a.wide = t->wide(); a.sm_count = t->ctx->sm_count; a.nq = nq; a.q = q_dev;
        a.out_prim = tri_dev; a.out_cp = cp_dev; a.out_d2 = d2_dev;
        launch_nearest_sorted(a, t->qsort, t->lo, t->hi, s);
        cudaEvent_t e0, e1;
        ICP_CUDA(cudaEventCreate(&e0));
        ICP_CUDA(cudaEventCreate(&e1));
        ICP_CUDA(cudaEventRecord(e0, s));
        for (int i = 0; i < iters; i++) launch_nearest_sorted(a, t->qsort, t->lo, t->hi, s);
        ICP_CUDA(cudaEventRecord(e1, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        float t_ms = 0;
        ICP_CUDA(cudaEventElapsedTime(&t_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms = t_ms / iters;
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

// DMMA issue-rate experiment (tools/dmma_sweep.py): throughput as a function of the number of resident warps and of
// independent accumulators per warp. out: TFLOP/s.
template <int NACC>
__global__ void k_dmma_sweep(double *out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

extern "C" int32_t icp_debug_dmma_sweep(icp_ctx ctx, int32_t warps_per_cta, int32_t ctas_per_sm, int32_t nacc, double *tflops) {
    icp_ctx _ctx = ctx;
    try {
        ICP_REQUIRE(ctx && tflops, "null argument");
        CtxLock lock(ctx);
        cudaStream_t s = ctx->stream;
        DevBuf<double> d;
        d.alloc(4);
        cudaEvent_t e0, e1;
        ICP_CUDA(cudaEventCreate(&e0));
        ICP_CUDA(cudaEventCreate(&e1));
        const int iters = 8192, blocks = ctx->sm_count * ctas_per_sm, threads = warps_per_cta * 32;
        // dynamic shared memory pins the number of resident CTAs per SM
        size_t smem = (size_t)(200 * 1024) / ctas_per_sm;
        double best = 0;
        for (int rep = 0; rep < 3; rep++) {
            ICP_CUDA(cudaEventRecord(e0, s));
            if (nacc == 1) { ICP_CUDA(cudaFuncSetAttribute(k_dmma_sweep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_dmma_sweep<1><<<blocks, threads, smem, s>>>(d.p, iters); }
            else if (nacc == 2) { ICP_CUDA(cudaFuncSetAttribute(k_dmma_sweep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_dmma_sweep<2><<<blocks, threads, smem, s>>>(d.p, iters); }
            else if (nacc == 4) { ICP_CUDA(cudaFuncSetAttribute(k_dmma_sweep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_dmma_sweep<4><<<blocks, threads, smem, s>>>(d.p, iters); }
            else { ICP_CUDA(cudaFuncSetAttribute(k_dmma_sweep<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_dmma_sweep<8><<<blocks, threads, smem, s>>>(d.p, iters); nacc = 8; }
            ICP_CUDA(cudaGetLastError());
            ICP_CUDA(cudaEventRecord(e1, s));
            ICP_CUDA(cudaStreamSynchronize(s));
            float ms = 0;
            ICP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            double tf = 2.0 * 256 * nacc * iters * (double)blocks * warps_per_cta / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *tflops = best;
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}
