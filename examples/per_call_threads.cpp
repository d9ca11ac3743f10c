// per_call_threads.cpp - the reference's threading against the plain C ABI: ONE target-sampling proposal, ONE model-sampling
// proposal and ONE evaluator shared by T host threads (apps/femur/RunMHRandomInitComparison.scala:59-86 shares its proposal
// mixture and evaluator between ten fitting threads the same way), every thread driving its own Metropolis-Hastings walk
// through the per-call entry points: icp_propose, icp_log_transition both ways for both ICP components, icp_eval_log_value.
// Checks that every walk ends where the same walk ends when it runs alone (serially), and reports MH steps/s for 1 and T
// threads. No GIL here: this is what JVM threads would see.
//
//   g++ -std=c++17 -O2 -pthread -I include examples/per_call_threads.cpp -L icp-proposal_b200 -licpcuda -o per_call_threads
//   ./per_call_threads model.bin n_threads steps_per_thread        (model.bin: the layout tests/test_gpu_cpp_host.py writes)
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "icpcuda.h"

template <class T>
static std::vector<T> rd(std::ifstream &f, size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * sizeof(T)));
    if (!f) throw std::runtime_error("short read");
    return v;
}

static icp_ctx g_ctx = nullptr;
static void ck(int32_t rc, const char *what) {
    if (rc == ICP_OK) return;
    char buf[512] = {0};
    icp_last_error(g_ctx, buf, sizeof buf);
    throw std::runtime_error(std::string(what) + ": " + buf);
}

// the host's random numbers (the reference draws them from scalismo.utils.Random on the JVM side): SplitMix64 + Box-Muller
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uniform() { return (double)((next() >> 11) + 1) / 9007199254740993.0; }
    double normal() { return std::sqrt(-2.0 * std::log(uniform())) * std::cos(6.283185307179586 * uniform()); }
};

struct Walk {
    std::vector<double> theta;   // final state
    int accepted = 0;
};

static Walk walk(icp_proposal pt, icp_proposal pm, icp_evaluator ev, int K, const std::vector<double> &theta0, uint64_t seed, int steps) {
    const int Lt = K + 10;
    Rng rng(seed);
    Walk w;
    w.theta = theta0;
    std::vector<double> prop(Lt), z(K);
    double v_cur[3], v_prop[3], lt[4];
    ck(icp_eval_log_value(ev, 1, w.theta.data(), v_cur, nullptr), "icp_eval_log_value");
    for (int s = 0; s < steps; s++) {
        for (double &x : z) x = rng.normal();
        icp_proposal gen = rng.uniform() < 0.5 ? pt : pm;      // MixtureProposal: which component proposes
        ck(icp_propose(gen, 1, w.theta.data(), z.data(), prop.data()), "icp_propose");
        // MixtureProposal.logTransitionProbability sums over every component, both directions
        ck(icp_log_transition(pt, 1, w.theta.data(), prop.data(), &lt[0]), "icp_log_transition");
        ck(icp_log_transition(pt, 1, prop.data(), w.theta.data(), &lt[1]), "icp_log_transition");
        ck(icp_log_transition(pm, 1, w.theta.data(), prop.data(), &lt[2]), "icp_log_transition");
        ck(icp_log_transition(pm, 1, prop.data(), w.theta.data(), &lt[3]), "icp_log_transition");
        ck(icp_eval_log_value(ev, 1, prop.data(), v_prop, nullptr), "icp_eval_log_value");
        auto lse = [](double a, double b) { double m = a > b ? a : b; return m + std::log(0.5 * std::exp(a - m) + 0.5 * std::exp(b - m)); };
        const double a = v_prop[0] - v_cur[0] - (lse(lt[0], lt[2]) - lse(lt[1], lt[3]));
        if (a > 0.0 || rng.uniform() < std::exp(a)) { w.theta = prop; std::memcpy(v_cur, v_prop, sizeof v_cur); w.accepted++; }
    }
    return w;
}

int main(int argc, char **argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s model.bin n_threads steps_per_thread\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        auto hdr = rd<int32_t>(f, 8);
        const int N = hdr[0], T = hdr[1], K = hdr[2], Nt = hdr[3], Tt = hdr[4], n_ids = hdr[5], n_tp = hdr[6], n_eval = hdr[7];
        auto ref = rd<double>(f, 3 * (size_t)N), basis = rd<double>(f, 3 * (size_t)N * K), var = rd<double>(f, K);
        auto tv = rd<double>(f, 3 * (size_t)Nt), tp = rd<double>(f, 3 * (size_t)n_tp);
        auto cells = rd<int32_t>(f, 3 * (size_t)T), tcells = rd<int32_t>(f, 3 * (size_t)Tt), ids = rd<int32_t>(f, n_ids), eids = rd<int32_t>(f, n_eval);
        const int n_threads = std::atoi(argv[2]), steps = std::atoi(argv[3]);

        ck(icp_ctx_create(0, &g_ctx), "icp_ctx_create");
        icp_model model; icp_target target; icp_proposal pt, pm; icp_evaluator ev;
        ck(icp_model_create(g_ctx, N, T, K, ref.data(), nullptr, basis.data(), var.data(), cells.data(), &model), "icp_model_create");
        ck(icp_target_create(g_ctx, Nt, Tt, tv.data(), tcells.data(), &target), "icp_target_create");
        icp_proposal_params pp{0.1, 10.0, 5.0, ICP_TARGET_SAMPLING, 1, ICP_FACTOR_CHOLESKY, ICP_RANK_UPDATE_FP64};
        ck(icp_proposal_create(model, target, &pp, ids.data(), n_ids, tp.data(), n_tp, &pt), "icp_proposal_create");
        pp.direction = ICP_MODEL_SAMPLING;
        ck(icp_proposal_create(model, target, &pp, ids.data(), n_ids, tp.data(), n_tp, &pm), "icp_proposal_create");
        icp_evaluator_params ep{ICP_EVAL_INDEPENDENT, ICP_MODEL_TO_TARGET, 1, 0, 0.0, 2.0, 0.0};
        ck(icp_evaluator_create(model, target, &ep, eids.data(), n_eval, tp.data(), n_tp, &ev), "icp_evaluator_create");

        // every walk starts from the mean shape (the first calls of all threads collide on one cache key)
        std::vector<double> theta0(K + 10, 0.0);
        theta0[0] = 1.0;
        for (int i = 0; i < N; i++) for (int d = 0; d < 3; d++) theta0[7 + d] += ref[3 * i + d] / N;

        using clk = std::chrono::steady_clock;
        walk(pt, pm, ev, K, theta0, 999, 5);     // warm-up
        std::vector<Walk> alone(n_threads), together(n_threads);
        auto t0 = clk::now();
        for (int t = 0; t < n_threads; t++) alone[t] = walk(pt, pm, ev, K, theta0, 1000 + t, steps);
        const double serial_s = std::chrono::duration<double>(clk::now() - t0).count();
        ck(icp_proposal_clear_cache(pt), "icp_proposal_clear_cache");
        ck(icp_proposal_clear_cache(pm), "icp_proposal_clear_cache");
        std::vector<std::string> errors(n_threads);
        for (int round = 0; round < 2; round++) {   // the second round runs on the warm call slots (graphs captured)
            std::vector<std::thread> th;
            t0 = clk::now();
            for (int t = 0; t < n_threads; t++)
                th.emplace_back([&, t] {
                    try { together[t] = walk(pt, pm, ev, K, theta0, 1000 + t, steps); } catch (const std::exception &e) { errors[t] = e.what(); }
                });
            for (auto &x : th) x.join();
        }
        const double threaded_s = std::chrono::duration<double>(clk::now() - t0).count();
        int mismatches = 0, failed = 0, accepted = 0;
        for (int t = 0; t < n_threads; t++) {
            if (!errors[t].empty()) { failed++; std::fprintf(stderr, "thread %d: %s\n", t, errors[t].c_str()); continue; }
            if (together[t].accepted != alone[t].accepted || std::memcmp(together[t].theta.data(), alone[t].theta.data(), sizeof(double) * (K + 10)) != 0) mismatches++;
            accepted += alone[t].accepted;
        }
        std::printf("{\"threads\": %d, \"steps_per_thread\": %d, \"failed_threads\": %d, \"walks_that_differ_from_the_serial_run\": %d, "
                    "\"accepted_total\": %d, \"one_thread_steps_per_s\": %.1f, \"threads_steps_per_s\": %.1f}\n",
                    n_threads, steps, failed, mismatches, accepted, n_threads * steps / serial_s, n_threads * steps / threaded_s);
        icp_evaluator_destroy(ev); icp_proposal_destroy(pm); icp_proposal_destroy(pt); icp_target_destroy(target); icp_model_destroy(model);
        icp_ctx_destroy(g_ctx);
        return (failed || mismatches) ? 1 : 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
