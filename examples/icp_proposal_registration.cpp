// C++ counterpart of apps/femur/IcpProposalRegistration.scala:50-103 on top of the host-side mirror
// (icp-proposal_b200/host/icp_host.hpp) and libicpcuda.so.
//
//   icp_proposal_registration <model.bin> <n_samples> [log.json]
//
// model.bin (written by tests/test_gpu_cpp_host.py): int32 N, T, K, Nt, Tt, n_ids, n_tp, n_eval; then doubles ref[3N],
// basis[3N K], variance[K], target[3Nt], target_points[3 n_tp]; then int32 cells[3T], target_cells[3Tt], ids[n_ids],
// eval_ids[n_eval]. Prints one JSON line with the outcome.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "icp_host.hpp"

using namespace icp_host;

template <class T>
static std::vector<T> rd(std::ifstream &f, size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * sizeof(T)));
    if (!f) throw std::runtime_error("short read");
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s model.bin n_samples [log.json]\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        auto hdr = rd<int32_t>(f, 8);
        const int N = hdr[0], T = hdr[1], K = hdr[2], Nt = hdr[3], Tt = hdr[4], n_ids = hdr[5], n_tp = hdr[6], n_eval = hdr[7];
        auto ref = rd<double>(f, 3 * (size_t)N), basis = rd<double>(f, 3 * (size_t)N * K), var = rd<double>(f, K);
        auto tv = rd<double>(f, 3 * (size_t)Nt), tp = rd<double>(f, 3 * (size_t)n_tp);
        auto cells = rd<int32_t>(f, 3 * (size_t)T), tcells = rd<int32_t>(f, 3 * (size_t)Tt), ids = rd<int32_t>(f, n_ids), eids = rd<int32_t>(f, n_eval);
        const int numOfSamples = std::atoi(argv[2]);

        Context ctx(0);
        StatisticalMeshModel model(ctx, ref, cells, basis, var);
        TriangleMesh3D target(ctx, tv, tcells);

        // MixedProposalDistributions.mixedProposalICP(..., ModelAndTargetSampling, tangentialNoise = 10, noiseAlongNormal = 5, stepLength = 0.1)
        auto icpT = std::make_shared<NonRigidIcpProposal>(model, target, 0.1, 10.0, 5.0, ids, tp, IcpProjectionDirection::TargetSampling, true,
                                                          "IcpProposal-TargetSampling-0.1Step");
        auto icpM = std::make_shared<NonRigidIcpProposal>(model, target, 0.1, 10.0, 5.0, ids, tp, IcpProjectionDirection::ModelSampling, true,
                                                          "IcpProposal-ModelSampling-0.1Step");
        auto proposalICP = std::make_shared<MixtureProposal>(std::vector<std::pair<double, std::shared_ptr<ProposalGeneratorWithTransition>>>{{0.5, icpT}, {0.5, icpM}});
        auto proposalRND = std::make_shared<MixtureProposal>(std::vector<std::pair<double, std::shared_ptr<ProposalGeneratorWithTransition>>>{
            {0.5, std::make_shared<RandomShapeUpdateProposal>(model, 0.1, "RandomShape-0.1")}});
        MixtureProposal proposal({{0.90, proposalICP}, {0.10, proposalRND}});

        // ProductEvaluators.proximityAndIndependent(model, target, ModelToTargetEvaluation, uncertainty = 2.0, 4 * rank points)
        auto distance = IndependentPointDistanceEvaluator(model, target, 0.0, 2.0, EvaluationMode::ModelToTargetEvaluation, eids, tp);
        auto prior = std::make_shared<ModelPriorEvaluator>(model);
        ProductEvaluator product({prior, distance});
        auto productDev = makeProductEvaluator(model, target, distance->params(), eids, tp);

        ModelFittingParameters theta0 = model.initialParameters();
        const double p0 = product.logValue(theta0);

        // (1) the reference's driver structure: Scalismo-style MH over the per-call classes
        MetropolisHastings chain(proposal, product);
        ModelFittingParameters theta = theta0;
        int acc = 0;
        for (int i = 0; i < 20; i++) { bool ok = false; theta = chain.next(theta, &ok); acc += ok; }
        const double p20 = product.logValue(theta);

        // (2) SamplingRegistration.runfitting on the fused device runner
        SamplingRegistration reg(model, target);
        auto res = reg.runfitting(*productDev, "distance", proposal, numOfSamples, theta0, argc > 3 ? argv[3] : "");
        const double pbest = product.logValue(res.best);

        // (3) apps/femur/PosteriorVariabilityToMeshColor.scala:45-52: every 10th logged state after a burn-in of 20 -> shapes ->
        // per-vertex variability maps, reconstructed and reduced on the device
        std::vector<ModelFittingParameters> picked;
        for (auto &e : LogHelper::samplesFromLog(res.log, 10, numOfSamples, 20)) picked.push_back(LogHelper::sampleToModelParameters(*e.first));
        double meanTotal = 0.0, meanNormal = 0.0;
        if (picked.size() > 1) {
            auto maps = PosteriorVariability::statistics(model, picked, true);
            for (double v : maps.total) meanTotal += v / maps.total.size();
            for (double v : maps.normal) meanNormal += v / maps.normal.size();
        }

        // (4) apps/femur/CreateGPModel.scala:70-86 in small: isotropic two-scale Gaussian kernel, Nystrom on every 20th reference
        // point, 8 basis functions - kernel matrix, eigen-decomposition and extension all on the device
        std::vector<double> nys;
        for (int i = 0; i < N; i += 20) nys.insert(nys.end(), ref.begin() + 3 * i, ref.begin() + 3 * i + 3);
        auto gpKernel = MatrixValuedKernel::gaussian(40.0) * 5.0 + MatrixValuedKernel::gaussian(10.0) * 3.0;
        auto lowRank = LowRankGaussianProcess::approximateGPNystrom(ctx, gpKernel, ref, nys, 8);
        double varSum = 0.0, basisAbs = 0.0;
        for (double v : lowRank.variance) varSum += v;
        for (double v : lowRank.basis) basisAbs += std::fabs(v);

        std::printf("{\"gp_variance_sum\": %.12g, \"gp_basis_abs_sum\": %.12g, \"K\": %d, \"product_initial\": %.12g, \"host_mh_steps\": 20, \"host_mh_accepted\": %d, \"product_after_host_mh\": %.12g, "
                    "\"fused_steps\": %d, \"fused_accepted\": %lld, \"fused_best_product\": %.12g, \"fused_best_product_recomputed\": %.12g, "
                    "\"fused_best_generated_by\": \"%s\", \"variability_samples\": %d, \"mean_total_variance\": %.12g, "
                    "\"mean_normal_variance\": %.12g}\n",
                    varSum, basisAbs, K, p0, acc, p20, res.steps, (long long)res.accepted, res.bestProduct, pbest, res.best.generatedBy.c_str(),
                    (int)picked.size(), meanTotal, meanNormal);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
