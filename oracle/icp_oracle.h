/*
 * icp_oracle.h - CPU restatement (plain C, FP64) of the icp-proposal hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (libicpcuda.so, the host-side mirror in
 * icp-proposal_b200/) may include, link or call this file. It is the checker for tests/, for
 * __graft_entry__.smoke() and for bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (unibas-gravis/icp-proposal, Scala on Scalismo 0.90.0) ships no
 * tests or golden vectors for this path and cannot run here (no JVM, Scalismo jars not vendored;
 * SURVEY.md section 8c). This oracle follows the reference's own call sites line by line and the
 * Scalismo 0.90.0 / Breeze semantics recalled in SURVEY.md Appendix A ([S-recall]).
 *
 * Paths are relative to /root/reference/src/main/scala.
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- geometry ------------------------------------------------------------------------------- */
typedef struct orc_mesh orc_mesh;

/* Scalismo TriangleMesh3D with its lazily built helpers: bounding-volume tree over triangles
 * (operations.closestPointOnSurface), KD-tree over vertices (pointSet.findClosestPoint), boundary
 * table (operations.pointIsOnBoundary) and vertex normals. Appendix A10-A13. */
orc_mesh *orc_mesh_create(int nv, const double *verts, int nt, const int32_t *tris);
void orc_mesh_free(orc_mesh *m);

/* exact closest point, brute force over all triangles; ties -> lowest triangle index.
 * feature: 0 vertex, 1 edge, 2 face interior. */
void orc_closest_point_brute(int nv, const double *verts, int nt, const int32_t *tris, int nq,
                             const double *q, int32_t *tri, int32_t *feat, double *cp, double *d2);
/* same through the tree (nearest-child-first descent, sphere-distance pruning) */
void orc_mesh_closest_point(const orc_mesh *m, int nq, const double *q, int32_t *tri, int32_t *feat,
                            double *cp, double *d2);
/* squared distance from q to one given triangle (used by tests to certify equidistant ties) */
double orc_point_triangle_d2(const double *q, const double *a, const double *b, const double *c,
                             double *cp, int32_t *feat);
void orc_closest_vertex_brute(int nv, const double *verts, int nq, const double *q, int32_t *id, double *d2);
void orc_mesh_closest_vertex(const orc_mesh *m, int nq, const double *q, int32_t *id, double *d2);
void orc_mesh_boundary_flags(const orc_mesh *m, uint8_t *flags);   /* nv flags */
void orc_mesh_vertex_normals(const orc_mesh *m, double *normals);  /* nv x 3 */

/* ---- model ---------------------------------------------------------------------------------- */
typedef struct orc_model orc_model;

/* StatisticalMeshModel: reference mesh, mean deformation (3N), unscaled basis U (3N x K row-major),
 * variances lambda (K). Q = U diag(sqrt(lambda)). */
orc_model *orc_model_create(int N, int T, int K, const double *ref, const double *mean_def,
                            const double *U, const double *variance, const int32_t *tris);
void orc_model_free(orc_model *m);
int orc_model_rank(const orc_model *m);

/* theta = [s, tx,ty,tz, phi,theta,psi, cx,cy,cz, alpha_0..alpha_{K-1}]
 * (api/sampling/ModelFittingParameters.scala:64). ModelFittingParameters.transformedMesh (:108-110) */
void orc_transformed_mesh(const orc_model *m, const double *theta, double *xyz);
void orc_pose_matrix(const double *theta, double R[9]);            /* Rz(phi) Ry(theta) Rx(psi) */
/* SurfaceNoiseHelpers.surfaceNormalDependantNoise (api/sampling/SurfaceNoiseHelpers.scala:32-60) */
void orc_surface_noise_cov(const double normal[3], double sd_normal, double sd_tangent, double cov[9]);

/* ---- ICP proposal --------------------------------------------------------------------------- */
typedef struct orc_proposal orc_proposal;
enum { ORC_MODEL_SAMPLING = 0, ORC_TARGET_SAMPLING = 1 };

/* NonRigidIcpProposal (api/sampling/proposals/NonRigidIcpProposal.scala:30-51). The id list and
 * the target point list are what model.decimate / target.operations.decimate returned (:45-46). */
orc_proposal *orc_proposal_create(const orc_model *model, const orc_mesh *target, double step_length,
                                  double tangential_noise, double noise_along_normal, int direction,
                                  int boundary_aware, int n_ids, const int32_t *ids, int n_tp,
                                  const double *target_points);
void orc_proposal_free(orc_proposal *p);

/* icpPosterior (:88-153) + LowRankGaussianProcess.regression. Returns the number of observations
 * that survived the boundary filter. mu (K), M (K x K), Minv (K x K) may be NULL. Optional
 * obs_ids (n), obs_y (3n), obs_cov (9n) expose the observation list. */
int orc_icp_posterior(const orc_proposal *p, const double *theta, double *mu, double *M, double *Minv,
                      int32_t *obs_ids, double *obs_y, double *obs_cov);
/* propose (:53-68) in the reference's own structure: posterior -> SVD of D Minv D -> rotated basis
 * on all N reference points -> sampled field -> model.coefficients (full-mesh regression, 1e-5). */
void orc_propose(const orc_proposal *p, const double *theta, const double *z, double *theta_out);
/* logTransitionProbability (:71-85), reference structure (discretised posterior + regression). */
double orc_log_transition(const orc_proposal *p, const double *from, const double *to);
/* the same two functions through the closed forms of SURVEY.md Appendix A6/A7 (what an optimised
 * CPU implementation would do; used for the "optimised CPU" column and to bound the shortcut error) */
void orc_propose_closed_form(const orc_proposal *p, const double *theta, const double *z, double *theta_out);
double orc_log_transition_closed_form(const orc_proposal *p, const double *from, const double *to);

/* model.posterior(corr, sigma2).mean coefficients + re-projection step of the deterministic ICP
 * (api/other/IcpBasedSurfaceFitting.scala:55-92): one iteration, alpha -> alpha'. direction as above;
 * ids / target_points are the sample lists drawn once (:51-53). */
void orc_std_icp_iteration(const orc_model *model, const orc_mesh *target, int direction, int n_ids,
                           const int32_t *ids, int n_tp, const double *target_points, double sigma2,
                           double step_length, const double *alpha, double *alpha_out);

/* the same iteration under a rigid transform: theta = [s, t, rot, centre, alpha] (:61 model.transform(currentTrans)) */
void orc_std_icp_iteration_theta(const orc_model *model, const orc_mesh *target, int direction, int n_ids,
                                 const int32_t *ids, int n_tp, const double *target_points, double sigma2,
                                 double step_length, const double *theta, double *alpha_out);

/* ---- evaluators ----------------------------------------------------------------------------- */
enum { ORC_MODEL_TO_TARGET = 0, ORC_TARGET_TO_MODEL = 1, ORC_SYMMETRIC = 2 };

/* IndependentPointDistanceEvaluator (evaluators/IndependentPointDistanceEvaluator.scala:27-67)
 * with breeze Gaussian(mean, sd) */
double orc_eval_independent(const orc_model *model, const orc_mesh *target, int mode, double g_mean,
                            double g_sd, int n_ids, const int32_t *ids, int n_tp,
                            const double *target_points, const double *theta);
/* HausdorffDistanceEvaluator (evaluators/HausdorffDistanceEvaluator.scala:25-36), Exponential(rate) */
double orc_eval_hausdorff(const orc_model *model, const orc_mesh *target, double rate, const double *theta);
/* CollectiveAverageHausdorffDistanceBoundaryAwareEvaluator (:27-79). *status = 1 when a filtered
 * distance list is empty (the reference throws, :51,63). avg_max (2) optional. */
double orc_eval_collective(const orc_model *model, const orc_mesh *target, int mode, double avg_mean,
                           double avg_sd, double max_rate, int n_ids, const int32_t *ids, int n_tp,
                           const double *target_points, const double *theta, int *status, double *avg_max);
/* ModelPriorEvaluator (evaluators/ModelPriorEvaluator.scala:24-31) */
double orc_eval_prior(int K, const double *theta);
/* RandomShapeUpdateProposal.logTransitionProbability (proposals/RandomShapeUpdateProposal.scala:38-45) */
double orc_random_walk_log_transition(int K, double sd, const double *from, const double *to);
/* PoseProposals.scala:47-62 / :81-89: kind 0 = rotation axis (0 roll/phi,1 pitch/theta,2 yaw/psi),
 * kind 1 = translation axis */
double orc_pose_log_transition(int K, int kind, int axis, double sd, const double *from, const double *to);

/* ---- registration quality measures ------------------------------------------------------------ */
/* api/other/RegistrationComparison.scala:24-49: {avg, hausdorff, boundary-aware avg, boundary-aware max} */
void orc_registration_metrics(const orc_model *model, const orc_mesh *target, const double *theta, double out[4]);
/* MeshMetrics.diceCoefficient over n unit-cube samples (apps/femur/StdIcpVsChainICPrandomInitComparisonAll.scala:46) */
double orc_dice_coefficient(const orc_model *model, const orc_mesh *target, const double *theta, int n, const double *unit);

/* ---- Metropolis-Hastings chain (Scalismo MetropolisHastings.next + MixtureProposal) ---------- */
enum { ORC_PROP_ICP = 0, ORC_PROP_RANDOM_SHAPE = 1, ORC_PROP_ROTATION = 2, ORC_PROP_TRANSLATION = 3 };
enum { ORC_EVAL_ACCEPT_ALL = 0, ORC_EVAL_INDEPENDENT = 1, ORC_EVAL_HAUSDORFF = 2, ORC_EVAL_COLLECTIVE = 3 };

typedef struct {
    int kind;                 /* ORC_PROP_* */
    double weight;            /* effective (flattened) mixture weight */
    const orc_proposal *icp;  /* kind == ICP */
    double sd;                /* random walk / pose std-dev */
    int axis;                 /* pose axis */
} orc_component;

typedef struct {
    const orc_model *model;
    const orc_mesh *target;
    int n_components;
    const orc_component *components;
    int use_prior;            /* ProductEvaluator(prior, distance) vs distance only */
    int eval_kind;            /* ORC_EVAL_* */
    int eval_mode;            /* ORC_MODEL_TO_TARGET ... */
    double p0, p1, p2;        /* independent: mean, sd | hausdorff: rate | collective: mean, sd, rate */
    int n_ids; const int32_t *ids;
    int n_tp; const double *target_points;
    int closed_form;          /* 0: reference structure, 1: Appendix-A closed forms */
} orc_chain_desc;

/* One chain, n_steps sequential MH steps from theta0. Randomness is supplied by the caller:
 * u_comp[step] picks the mixture component, z[step*K..] are the standard normals of the proposal,
 * u_acc[step] is the acceptance uniform. Per step the log receives:
 * comp[step], accepted[step], logv[step*3 + {product, prior, distance}] of the state that is
 * current AFTER the step (accepted proposal or retained state, JSONAcceptRejectLogger.scala:93-106)
 * and theta_log[step*(K+10)..] likewise. Returns the number of accepted steps. */
int orc_chain_run(const orc_chain_desc *d, const double *theta0, int n_steps, const double *u_comp,
                  const double *z, const double *u_acc, int32_t *comp, uint8_t *accepted, double *logv,
                  double *theta_log);

/* Switches the dense linear algebra (products, symmetric eigen-decomposition) to an OpenBLAS / LAPACKE library opened from
 * `path` (NULL switches back to the built-in loops). Returns 0 on success. Used by bench.py's timed CPU arm only. */
int orc_use_blas(const char *path);
int orc_blas_enabled(void);

/* Philox4x32-10, the counter-based generator the device chain runner uses (bit-exact integer check) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
