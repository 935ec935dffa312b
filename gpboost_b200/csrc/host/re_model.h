// Host-side REModel of the B200 build: the object behind the GPB_* C API (include/gpboost_b200_c_api.h).
//
// Mirrors the reference facade GPBoost::REModel (include/GPBoost/re_model.h:28-595, src/GPBoost/re_model.cpp)
// for the hot-path configurations of SURVEY §8: argument meaning, defaults, parameter transformations,
// optimiser driver and error behaviour follow the reference; every numeric pass runs on the device engine
// (include/gpboost_b200_dev.h). Configurations outside the hot path are rejected with the reference's
// error channel (exception -> -1 + LGBM_GetLastError) — there is no CPU fallback.
#ifndef GPB200_RE_MODEL_H_
#define GPB200_RE_MODEL_H_
#include <cstdint>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../../include/gpboost_b200_dev.h"
#include "lbfgs.h"

namespace gpb200 {

class REModel {
 public:
  // argument list = GPB_CreateREModel (include/LightGBM/c_api.h:1359-1391)
  REModel(int32_t num_data, const int32_t* cluster_ids_data, const char* re_group_data, int32_t num_re_group,
          const double* re_group_rand_coef_data, const int32_t* ind_effect_group_rand_coef, int32_t num_re_group_rand_coef,
          const int* drop_intercept_group_rand_effect, int32_t num_gp, const double* gp_coords_data, int dim_gp_coords,
          const double* gp_rand_coef_data, int32_t num_gp_rand_coef, const char* cov_fct, double cov_fct_shape,
          const char* gp_approx, double cov_fct_taper_range, double cov_fct_taper_shape, int num_neighbors,
          const char* vecchia_ordering, int num_ind_points, double cover_tree_radius, const char* ind_points_selection,
          const char* likelihood, double likelihood_additional_param, const char* matrix_inversion_method, int seed,
          int num_parallel_threads, bool GPU_use, bool has_weights, const double* weights, double likelihood_learning_rate);
  ~REModel();
  REModel(const REModel&) = delete;
  REModel& operator=(const REModel&) = delete;

  // GPB_SetOptimConfig (c_api.h:1437-1467): only the fields the hot path consumes are stored
  void SetOptimConfig(const double* init_cov_pars, double lr, int max_iter, double delta_rel_conv, bool trace,
                      const char* optimizer, const char* convergence_criterion, int m_lbfgs,
                      const int* estimate_cov_par_index);

  // iterative-method settings of GPB_SetOptimConfig consumed by the Laplace-Vecchia path (re_model_template.h:860-900)
  void SetIterativeConfig(int cg_max_num_it, int cg_max_num_it_tridiag, double cg_delta_conv, int num_rand_vec_trace,
                          const char* cg_preconditioner_type, int seed_rand_vec_trace, double delta_conv_mode_finding);
  // REModel::OptimCovPar (re_model.cpp:483-541)
  void OptimCovPar(const double* y_data, const double* fixed_effects, bool called_in_GPBoost_algorithm,
                   bool reuse_learning_rates_from_previous_call);
  // REModel::EvalNegLogLikelihood (re_model.cpp:752-794); cov_pars on the ORIGINAL scale or nullptr
  void EvalNegLogLikelihood(const double* y_data, const double* cov_pars, double* negll, const double* fixed_effects);
  // REModel::CalcGradient (re_model.cpp:809): y <- Psi^-1 y / sigma^2 at the current covariance parameters
  void CalcGradient(double* y, const double* fixed_effects, bool calc_cov_factor);
  // Device-resident forms used by the boosting loop (GBDT::Boosting -> objective -> REModel, gbdt.cpp:194,
  // regression_objective.hpp:164-165): y_dev is a device pointer to n doubles in original order, complete on the
  // caller's side (the caller synchronised its stream). No host copy of the response is made.
  void OptimCovParDevice(const double* y_dev, bool called_in_GPBoost_algorithm, bool reuse_learning_rates_from_previous_call);
  // response_is_current: y_dev is the vector the last OptimCovParDevice call installed (the boosting objective calls both on the same
  // F - y, regression_objective.hpp:164-165): it is not installed again, and the factor of the optimiser's last gradient pass is reused
  void CalcGradientDevice(double* y_dev, bool response_is_current = false);
  bool DevicePathReady() const;
  // REModel::NewtonUpdateLeafValues (re_model.cpp:1298-1310 -> re_model_template.h:4982-5063), device-resident: leaf ids and the
  // gradient of the tree that was just grown stay in HBM; only the L x L system comes to the host. After CalcGradient*.
  void NewtonUpdateLeafValuesDevice(const int32_t* leaf_of_row_dev, int num_leaves, const double* grad_dev, double* leaf_values);
  // GPB_SetPredictionData (c_api.h:1601-1613): prediction locations / neighbour count kept for later Predict calls
  void SetPredictionData(int32_t num_data_pred, const double* gp_coords_data_pred, const char* vecchia_pred_type, int num_neighbors_pred);
  // REModel::Predict (re_model.cpp:1081-1215) for the Gaussian Vecchia model: out_predict = mean (num_data_pred),
  // followed by the predictive variances when predict_var. gp_coords_data_pred column-major like every matrix of the API.
  void Predict(const double* y_obs, int32_t num_data_pred, double* out_predict, bool predict_cov_mat, bool predict_var,
               bool predict_response, const double* gp_coords_data_pred, const double* cov_pars_pred, bool use_saved_data,
               const double* fixed_effects);
  // GPB_GetCovPar / GPB_GetInitCovPar (original scale)
  void GetCovPar(double* out, bool calc_std_dev) const;
  void GetInitCovPar(double* out) const;
  int GetNumIt() const { return num_it_; }
  int NumCovPars() const { return num_cov_pars_; }
  int NumData() const { return num_data_; }
  double CurrentNegLogLikelihood() const { return neg_log_likelihood_; }
  const std::string& LikelihoodName() const { return likelihood_; }
  const std::string& OptimizerCovPars() const { return optimizer_; }
  int64_t NumLikelihoodEvals() const { return num_ll_evals_; }
  gpbdev_vecchia_t Engine() const { return engine_; }
  bool IsGrouped() const { return grouped_ != nullptr; }
  // transformed <-> original scale (cov_fcts.h:485-623)
  void TransformCovPars(const double* orig, double* trans) const;
  void TransformBackCovPars(const double* trans, double* orig) const;

 private:
  void InitializeCovParsIfNotDefined(const double* y_data, const double* fixed_effects);
  void FindInitCovPar(const double* y_data, const double* fixed_effects, double* init_trans);
  void SetY(const double* y_data, const double* fixed_effects);
  void SetYDevice(const double* y_dev);
  void OptimCovParCore(bool called_in_GPBoost_algorithm, bool reuse_learning_rates_from_previous_call);
  void OptimCovParLaplace(const double* y_data, const double* fixed_effects);
  // one device pass at transformed (var, range); fills sums_
  void DevicePass(double var, double range, int mode);
  double NegLLFromSums(double sigma2) const;

  int32_t num_data_ = 0;
  int dim_ = 0;
  int num_neighbors_ = 20;
  int num_neighbors_pred_ = 40;          // 2 x num_neighbors (re_model_template.h:299)
  std::vector<double> coords_pred_saved_;  // np x d row-major (GPB_SetPredictionData)
  int32_t num_data_pred_saved_ = 0;
  bool y_has_been_set_ = false;
  int cov_id_ = 0;
  std::string cov_fct_, gp_approx_, vecchia_ordering_, likelihood_;
  double shape_ = 0.;
  int num_cov_pars_ = 3;
  std::mt19937 rng_;
  std::vector<int32_t> perm_;            // ordered position -> original index (data_indices_per_cluster_)
  std::vector<double> coords_ordered_;   // n x d row-major
  gpbdev_vecchia_t engine_ = nullptr;
  gpbdev_grouped_t grouped_ = nullptr;   // single-level grouped random effect backend
  gpbdev_dense_t dense_ = nullptr;       // exact GP backend, gp_approx = "none" 
  void DensePass(double var, double range, bool with_grad = false);
  // non-Gaussian likelihood (bernoulli_logit) with a latent Vecchia GP: Laplace approximation on the device
  bool gauss_ = true;
  bool device_collective_ = false;  // the engine all-reduces its results itself (NCCL on its stream)
  void EvalLaplace(const double* y_data, const double* cov_pars, double* negll, const double* fixed_effects);
  void EnsureProbes();
  double TransformRange(double range) const;
  int cg_max_num_it_ = 1000, cg_max_num_it_tridiag_ = 1000, num_rand_vec_trace_ = 50, seed_rand_vec_trace_ = 1;
  double cg_delta_conv_ = 1e-2, delta_conv_mode_finding_ = 1e-8;
  uint64_t cg_generator_counter_ = 0;   // likelihoods.h:17395
  int probes_t_ = 0, probes_seed_ = -1;
  double laplace_out_[6];
 public:
  const double* LaplaceInfo() const { return laplace_out_; }
  void GetLaplaceMode(double* out) const;
  void EvalLaplaceWithGradient(const double* y_data, const double* cov_pars, const double* fixed_effects, double* negll, double* grad2);
 private:
  int num_groups_ = 0;
  double gsums_[5];
  void CreateGroupedBackend(const char* re_group_data);
  void GroupedPass(double var_ratio);

  // state (all covariance parameters kept on the TRANSFORMED scale like REModel::cov_pars_)
  std::vector<double> cov_pars_, init_cov_pars_;
  bool cov_pars_initialized_ = false, init_cov_pars_provided_ = false;
  double neg_log_likelihood_ = 0.;
  int num_it_ = 0;
  int64_t num_ll_evals_ = 0;
  double sums_[GPBDEV_NUM_SUMS];
  std::vector<double> work_;

  // optimiser settings (defaults: re_model_template.h:8277-8347, :5851)
  std::string optimizer_ = "lbfgs";
  std::string convergence_criterion_ = "relative_change_in_log_likelihood";
  int max_iter_ = 1000;
  double delta_rel_conv_ = 1e-6;
  double lr_cov_init_ = 1.;
  int m_lbfgs_ = 6;
  bool trace_ = false;
  std::vector<int> estimate_cov_par_index_;
  LbfgsMemory lbfgs_mem_;
  bool cov_pars_estimated_once_ = false;
};

}  // namespace gpb200
#endif  // GPB200_RE_MODEL_H_
