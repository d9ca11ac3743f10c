// metrics.cu - registration quality measures between transformedMesh(theta) and the target, batched over chains.
//
// Replaces (paths relative to src/main/scala of the reference):
//   api/other/RegistrationComparison.scala:24-49   MeshMetrics.avgDistance / hausdorffDistance and the boundary-aware
//                                                   average + maximum, printed every acceptInfoPrintInterval steps for the
//                                                   best sample so far (api/sampling/SamplingRegistration.scala:75-82)
//   apps/femur/StdIcpVsChainICPrandomInitComparisonAll.scala:43-47   MeshMetrics.diceCoefficient(best, target)
// All of them are the closest-point / nearest-vertex queries of the hot path over every vertex (or over sample points).
#include <algorithm>
#include <cmath>

#include "icp_device.cuh"
#include "icp_internal.h"

using namespace icp;

namespace icp {

// per chain: {avg, hausdorff, avg_boundary_aware, max_boundary_aware}
__global__ void __launch_bounds__(128) k_metrics(int N, int Nt, const double *__restrict__ d2_m2t,
                                                 const uint8_t *__restrict__ skip, const double *__restrict__ d2_t2m,
                                                 double *__restrict__ out) {
    __shared__ double red[4][128];
    int c = blockIdx.x;
    double s = 0, mx = 0, sb = 0, cb = 0, mb = -INFINITY, mt = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double d = sqrt(d2_m2t[(size_t)c * N + i]);
        s += d; mx = fmax(mx, d);
        if (!(skip && skip[(size_t)c * N + i])) { sb += d; cb += 1; mb = fmax(mb, d); }
    }
    for (int i = threadIdx.x; i < Nt; i += blockDim.x) mt = fmax(mt, sqrt(d2_t2m[(size_t)c * Nt + i]));
    red[0][threadIdx.x] = s; red[1][threadIdx.x] = fmax(mx, mt); red[2][threadIdx.x] = sb; red[3][threadIdx.x] = cb;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + o]);
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
            red[3][threadIdx.x] += red[3][threadIdx.x + o];
        }
        __syncthreads();
    }
    // an empty filtered list: the reference's `filteredDists.sum / size` is NaN and `.max` throws (:42); both NaN here
    double avg = red[0][0] / N, hd = red[1][0], avgb = red[2][0] / red[3][0];
    const bool empty = red[3][0] == 0.0;
    __syncthreads();
    red[0][threadIdx.x] = mb;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[4 * c] = avg; out[4 * c + 1] = hd; out[4 * c + 2] = avgb; out[4 * c + 3] = empty ? NAN : red[0][0]; }
}

void registration_metrics_device(icp_model m, icp_target t, int C, const double *d_theta, double *d_out, MetricsWork &w,
                                 cudaStream_t s) {
    if (C <= 0) return;
    const int N = m->N, Nt = t->Nt;
    w.X.ensure((size_t)C * N * 3); w.d2a.ensure((size_t)C * N); w.cpa.ensure((size_t)C * N * 3); w.d2b.ensure((size_t)C * Nt);
    launch_reconstruct(m->dev(), C, d_theta, w.X.p, s);
    // every model vertex against the target surface
    NearestArgs a;
    a.bvh = &t->tri_bvh; a.prim_data = t->tri_data.p; a.wide = t->wide(); a.sm_count = t->ctx->sm_count; a.C = C; a.nq = N; a.q = w.X.p; a.q_per_chain = 1;
    a.out_d2 = w.d2a.p; a.out_cp = w.cpa.p;
    launch_nearest(a, s);
    const uint8_t *skip_p = nullptr;
    if (t->has_boundary) {
        // :35-37 nearest target vertex of the closest point, dropped when it lies on the target's boundary
        w.prim.ensure((size_t)C * N); w.skip.ensure((size_t)C * N);
        NearestArgs v;
        v.bvh = &t->vert_bvh; v.prim_data = t->vert_data.p; v.C = C; v.nq = N; v.q = w.cpa.p; v.q_per_chain = 1; v.out_prim = w.prim.p;
        launch_nearest(v, s);
        launch_lookup_flags((int64_t)C * N, w.prim.p, t->boundary.p, Nt, w.skip.p, s);
        skip_p = w.skip.p;
    }
    // every target vertex against the model surfaces (second half of hausdorffDistance)
    bvh_refit(m->tri_bvh, C, w.X.p, N, m->tris.p, s);
    NearestArgs b;
    b.bvh = &m->tri_bvh; b.X = w.X.p; b.tris = m->tris.p; b.N = N; b.C = C; b.nq = Nt; b.q = t->verts.p; b.out_d2 = w.d2b.p;
    launch_nearest(b, s);
    k_metrics<<<C, 128, 0, s>>>(N, Nt, w.d2a.p, skip_p, w.d2b.p, d_out);
    ICP_CUDA(cudaGetLastError());
}

// ---- Dice coefficient ------------------------------------------------------------------------------------------------
// Scalismo MeshMetrics.diceCoefficient [S-recall]: n uniform samples in the union of the two bounding boxes; a sample is
// inside a mesh when  vertexNormal(v) . (v - p) > 0  for the mesh vertex v nearest to p (toBinaryImage);
// dice = 2 |A and B| / (|A| + |B|). The reference draws the samples from an unseeded global RNG (10 000 of them), so the
// caller supplies unit-cube samples (or a Philox seed) and they are scaled into each chain's evaluation region here.
__global__ void __launch_bounds__(256) k_dice_points(int N, const double *__restrict__ X, double tlx, double tly, double tlz,
                                                     double thx, double thy, double thz, int n,
                                                     const double *__restrict__ unit, unsigned long long seed,
                                                     double *__restrict__ pts) {
    __shared__ double red[6][256];
    const int c = blockIdx.x, tid = threadIdx.x;
    const double *Xc = X + (size_t)c * N * 3;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int v = tid; v < N; v += blockDim.x)
#pragma unroll
        for (int d = 0; d < 3; d++) { const double x = Xc[3 * v + d]; lo[d] = fmin(lo[d], x); hi[d] = fmax(hi[d], x); }
#pragma unroll
    for (int d = 0; d < 3; d++) { red[d][tid] = lo[d]; red[3 + d][tid] = hi[d]; }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
#pragma unroll
            for (int d = 0; d < 3; d++) {
                red[d][tid] = fmin(red[d][tid], red[d][tid + o]);
                red[3 + d][tid] = fmax(red[3 + d][tid], red[3 + d][tid + o]);
            }
        __syncthreads();
    }
    const double l0 = fmin(red[0][0], tlx), l1 = fmin(red[1][0], tly), l2 = fmin(red[2][0], tlz);
    const double e0 = fmax(red[3][0], thx) - l0, e1 = fmax(red[4][0], thy) - l1, e2 = fmax(red[5][0], thz) - l2;
    for (int i = tid; i < n; i += blockDim.x) {
        double u0, u1, u2;
        if (unit) { u0 = unit[3 * i]; u1 = unit[3 * i + 1]; u2 = unit[3 * i + 2]; }
        else {
            // the same sample set for every chain (keyed by the sample index only): chains are compared on equal terms
            const uint4 r0 = chain_philox(seed, (unsigned long long)i, 0u, 0u), r1 = chain_philox(seed, (unsigned long long)i, 0u, 1u);
            u0 = u53(r0.x, r0.y); u1 = u53(r0.z, r0.w); u2 = u53(r1.x, r1.y);
        }
        double *p = pts + ((size_t)c * n + i) * 3;
        p[0] = l0 + u0 * e0; p[1] = l1 + u1 * e1; p[2] = l2 + u2 * e2;
    }
}

__global__ void __launch_bounds__(128) k_dice_count(ModelDev m, const double *__restrict__ X, int n,
                                                    const double *__restrict__ pts, const int *__restrict__ vid_model,
                                                    const int *__restrict__ vid_target, const double *__restrict__ tverts,
                                                    const double *__restrict__ tnormals, double *__restrict__ out) {
    __shared__ double red[40];
    const int c = blockIdx.x;
    const double *Xc = X + (size_t)c * m.N * 3;
    double na = 0.0, nb = 0.0, nab = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double *p = pts + ((size_t)c * n + i) * 3;
        const int va = vid_model[(size_t)c * n + i], vb = vid_target[(size_t)c * n + i];
        double nx, ny, nz;
        vertex_normal_dev(m, Xc, va, nx, ny, nz);
        const bool ina = nx * (Xc[3 * va] - p[0]) + ny * (Xc[3 * va + 1] - p[1]) + nz * (Xc[3 * va + 2] - p[2]) > 0.0;
        const bool inb = tnormals[3 * vb] * (tverts[3 * vb] - p[0]) + tnormals[3 * vb + 1] * (tverts[3 * vb + 1] - p[1]) +
                         tnormals[3 * vb + 2] * (tverts[3 * vb + 2] - p[2]) > 0.0;
        na += ina ? 1.0 : 0.0; nb += inb ? 1.0 : 0.0; nab += (ina && inb) ? 1.0 : 0.0;
    }
    na = block_sum(na, red);
    __syncthreads();
    nb = block_sum(nb, red);
    __syncthreads();
    nab = block_sum(nab, red);
    if (threadIdx.x == 0) out[c] = 2.0 * nab / (na + nb);
}

}  // namespace icp

extern "C" int32_t icp_registration_metrics(icp_model m, icp_target t, int32_t C, const double *theta, double *out) {
    icp_ctx _ctx = m ? m->ctx : nullptr;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(t && t->ctx == m->ctx, "bad target");
        if (C == 0) return ICP_OK;
        ICP_REQUIRE(C > 0 && theta != nullptr && out != nullptr, "bad arguments");
        cudaStream_t s = _ctx->stream;
        m->s_theta.upload(theta, (size_t)C * (m->K + kTheta0), s);
        MetricsWork w;
        DevBuf<double> dout;
        dout.alloc((size_t)4 * C);
        registration_metrics_device(m, t, C, m->s_theta.p, dout.p, w, s);
        ICP_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * 4 * C, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}

extern "C" int32_t icp_dice_coefficient(icp_model m, icp_target t, int32_t C, const double *theta, int32_t n_samples,
                                        const double *unit_samples, uint64_t seed, double *out) {
    icp_ctx _ctx = m ? m->ctx : nullptr;
    try {
        ICP_REQUIRE(_ctx != nullptr, "null handle");
        CtxLock lock(_ctx);
        ICP_REQUIRE(t && t->ctx == m->ctx, "bad target");
        if (C == 0) return ICP_OK;
        ICP_REQUIRE(C > 0 && theta != nullptr && out != nullptr && n_samples >= 1, "bad arguments");
        cudaStream_t s = _ctx->stream;
        const int N = m->N, n = n_samples;
        m->s_theta.upload(theta, (size_t)C * (m->K + kTheta0), s);
        DevBuf<double> X, unit, pts, dout;
        DevBuf<int> va, vb;
        X.alloc((size_t)C * N * 3); pts.alloc((size_t)C * n * 3); dout.alloc(C);
        va.alloc((size_t)C * n); vb.alloc((size_t)C * n);
        if (unit_samples) {
            for (size_t i = 0; i < (size_t)3 * n; i++)
                ICP_REQUIRE(unit_samples[i] >= 0.0 && unit_samples[i] <= 1.0, "unit_samples must lie in [0, 1]");
            unit.upload(unit_samples, (size_t)3 * n, s);
        }
        launch_reconstruct(m->dev(), C, m->s_theta.p, X.p, s);
        k_dice_points<<<C, 256, 0, s>>>(N, X.p, t->lo[0], t->lo[1], t->lo[2], t->hi[0], t->hi[1], t->hi[2], n,
                                        unit_samples ? unit.p : nullptr, seed, pts.p);
        ICP_CUDA(cudaGetLastError());
        nearest_model_vertex(m, C, X.p, n, pts.p, 1, nullptr, va.p, s);
        NearestArgs v;
        v.bvh = &t->vert_bvh; v.prim_data = t->vert_data.p; v.C = C; v.nq = n; v.q = pts.p; v.q_per_chain = 1; v.out_prim = vb.p;
        launch_nearest(v, s);
        k_dice_count<<<C, 128, 0, s>>>(m->dev(), X.p, n, pts.p, va.p, vb.p, t->verts.p, t->vnormals.p, dout.p);
        ICP_CUDA(cudaGetLastError());
        ICP_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * C, cudaMemcpyDeviceToHost, s));
        ICP_CUDA(cudaStreamSynchronize(s));
        return ICP_OK;
    } catch (...) {
        return translate_exception(_ctx);
    }
}
