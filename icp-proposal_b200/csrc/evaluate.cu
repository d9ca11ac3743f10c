// evaluate.cu - likelihood evaluators as fused warp-shuffle reductions over closest-point distances.
//
// Replaces (paths relative to src/main/scala/api/sampling of the reference):
//   evaluators/IndependentPointDistanceEvaluator.scala:40-66   sum of Gaussian.logPdf(distance)
//   evaluators/HausdorffDistanceEvaluator.scala:31-35          Exponential.logPdf(max distance, both ways)
//   evaluators/CollectiveAverageHausdorffDistanceBoundaryAwareEvaluator.scala:41-78
//   evaluators/ModelPriorEvaluator.scala:28-30                 N(0, I_K).logpdf(alpha)
//   ProductEvaluators.scala:44-53                              product = prior + distance
// Breeze densities: Gaussian(mu, sd).logPdf(x) = -(x-mu)^2/(2 sd^2) - ln(sd sqrt(2 pi)),
// Exponential(r).logPdf(x) = ln r - r x.
#include "icp_device.cuh"
#include "icp_internal.h"

namespace icp {

__device__ __forceinline__ double gauss_logpdf(double x, double mu, double sd) {
    return -((x - mu) * (x - mu)) / (2.0 * sd * sd) - log(sd * sqrt(2.0 * 3.14159265358979323846));
}

__device__ __forceinline__ double block_max(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = l < nw ? red[l] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (l == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// one direction of one chain: sum of log-densities, sum / count / max of kept distances
struct DirStats {
    double sum_logpdf, sum_d, cnt, max_d;
};

__device__ DirStats reduce_direction(int n, const double *__restrict__ d2, const uint8_t *__restrict__ skip,
                                     double g_mean, double g_sd, bool want_logpdf, double *red) {
    double sl = 0.0, sd = 0.0, cn = 0.0, mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (skip && skip[i]) continue;
        double d = sqrt(d2[i]);
        if (want_logpdf) sl += gauss_logpdf(d, g_mean, g_sd);
        sd += d;
        cn += 1.0;
        mx = fmax(mx, d);
        if (d != d) mx = d;  // NaN propagates
    }
    DirStats r;
    r.sum_logpdf = want_logpdf ? block_sum(sl, red) : 0.0;
    r.sum_d = block_sum(sd, red);
    r.cnt = block_sum(cn, red);
    double nanflag = block_sum((mx != mx) ? 1.0 : 0.0, red);
    r.max_d = block_max((mx != mx) ? -INFINITY : mx, red);
    if (nanflag > 0.0) r.max_d = NAN;
    return r;
}

__global__ void __launch_bounds__(128) k_eval_reduce(EvalReduceArgs a) {
    __shared__ double red[40];
    int c = blockIdx.x, K = a.K;
    const double *th = a.theta + (size_t)c * (K + kTheta0);
    double prior = 0.0;
    if (a.prm.use_prior) {
        double s = 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) s = fma(th[kTheta0 + k], th[kTheta0 + k], s);
        s = block_sum(s, red);
        prior = -0.5 * (K * ICP_LOG_2PI + s);
    }
    double dist = 0.0;
    int st = ICP_OK;
    const int kind = a.prm.kind, mode = a.prm.mode;
    if (kind != ICP_EVAL_ACCEPT_ALL) {
        bool use_m = kind == ICP_EVAL_HAUSDORFF || mode != ICP_TARGET_TO_MODEL;
        bool use_t = kind == ICP_EVAL_HAUSDORFF || mode != ICP_MODEL_TO_TARGET;
        DirStats sm = {0, 0, 0, -INFINITY}, stt = {0, 0, 0, -INFINITY};
        bool lp = kind == ICP_EVAL_INDEPENDENT;
        if (use_m)
            sm = reduce_direction(a.n_m2t, a.d2_m2t + (size_t)c * a.n_m2t,
                                  a.skip_m2t ? a.skip_m2t + (size_t)c * a.n_m2t : nullptr, a.prm.p0, a.prm.p1, lp, red);
        if (use_t)
            stt = reduce_direction(a.n_t2m, a.d2_t2m + (size_t)c * a.n_t2m,
                                   a.skip_t2m ? a.skip_t2m + (size_t)c * a.n_t2m : nullptr, a.prm.p0, a.prm.p1, lp, red);
        if (kind == ICP_EVAL_INDEPENDENT) {
            dist = mode == ICP_MODEL_TO_TARGET ? sm.sum_logpdf
                   : mode == ICP_TARGET_TO_MODEL ? stt.sum_logpdf
                                                 : 0.5 * sm.sum_logpdf + 0.5 * stt.sum_logpdf;
        } else if (kind == ICP_EVAL_HAUSDORFF) {
            double hd = fmax(sm.max_d, stt.max_d);
            if (sm.max_d != sm.max_d || stt.max_d != stt.max_d) hd = NAN;
            dist = log(a.prm.p0) - a.prm.p0 * hd;
        } else {  // collective average + max, boundary hits dropped
            double avg, mx;
            if ((use_m && sm.cnt == 0.0) || (use_t && stt.cnt == 0.0)) st = ICP_ERR_EMPTY_SET;
            double am = sm.sum_d / sm.cnt, at = stt.sum_d / stt.cnt;
            if (mode == ICP_MODEL_TO_TARGET) { avg = am; mx = sm.max_d; }
            else if (mode == ICP_TARGET_TO_MODEL) { avg = at; mx = stt.max_d; }
            else { avg = 0.5 * am + 0.5 * at; mx = fmax(sm.max_d, stt.max_d); if (sm.max_d != sm.max_d || stt.max_d != stt.max_d) mx = NAN; }
            dist = st == ICP_OK ? gauss_logpdf(avg, a.prm.p0, a.prm.p1) + (log(a.prm.p2) - a.prm.p2 * mx) : NAN;
        }
    }
    if (threadIdx.x == 0) {
        a.values[3 * c] = prior + dist;
        a.values[3 * c + 1] = prior;
        a.values[3 * c + 2] = dist;
        if (a.status) a.status[c] = st;
    }
}

void launch_eval_reduce(const EvalReduceArgs &a, cudaStream_t s) {
    ProfScope _ps(ST_EVAL_REDUCE, s);
    if (a.C <= 0) return;
    k_eval_reduce<<<a.C, 128, 0, s>>>(a);
    ICP_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(128) k_prior(int K, const double *__restrict__ theta, double *__restrict__ out) {
    __shared__ double red[40];
    int c = blockIdx.x;
    const double *th = theta + (size_t)c * (K + kTheta0);
    double s = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) s = fma(th[kTheta0 + k], th[kTheta0 + k], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[c] = -0.5 * (K * ICP_LOG_2PI + s);
}

void launch_prior(int C, int K, const double *d_theta, double *d_out, cudaStream_t s) {
    if (C <= 0) return;
    k_prior<<<C, 128, 0, s>>>(K, d_theta, d_out);
    ICP_CUDA(cudaGetLastError());
}

}  // namespace icp
