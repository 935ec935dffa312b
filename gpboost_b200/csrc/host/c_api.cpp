// GPB_* C API of the B200 build (include/gpboost_b200_c_api.h): thin, exception-to-error-code boundary
// like the reference's API_BEGIN/API_END (src/LightGBM/c_api.cpp:45-59).
#include "../../../include/gpboost_b200_c_api.h"

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "re_model.h"
#include "runtime.h"

namespace gpb200 {
Runtime& GetRuntime() {
  static Runtime rt;
  return rt;
}
}  // namespace gpb200

namespace {
thread_local char g_last_error[512] = "Everything is fine";  // c_api.h:1837-1849
void SetLastError(const char* msg) { std::snprintf(g_last_error, sizeof(g_last_error), "%s", msg); }
inline gpb200::REModel* M(REModelHandle h) {
  if (h == nullptr) throw std::runtime_error("REModel handle is null");
  return reinterpret_cast<gpb200::REModel*>(h);
}
int CopyString(const std::string& s, char* out_str, int* num_char) {
  *num_char = (int)s.size() + 1;
  std::memcpy(out_str, s.c_str(), s.size() + 1);
  return 0;
}
}  // namespace

#define API_BEGIN() try {
#define API_END()                                                          \
  }                                                                        \
  catch (std::exception & ex) { SetLastError(ex.what()); return -1; }      \
  catch (std::string & ex) { SetLastError(ex.c_str()); return -1; }        \
  catch (...) { SetLastError("unknown exception"); return -1; }            \
  return 0;

extern "C" {

const char* LGBM_GetLastError(void) { return g_last_error; }

int GPB_CreateREModel(int32_t num_data, const int32_t* cluster_ids_data, const char* re_group_data, int32_t num_re_group,
                      const double* re_group_rand_coef_data, const int32_t* ind_effect_group_rand_coef,
                      int32_t num_re_group_rand_coef, const int* drop_intercept_group_rand_effect, int32_t num_gp,
                      const double* gp_coords_data, const int dim_gp_coords, const double* gp_rand_coef_data,
                      int32_t num_gp_rand_coef, const char* cov_fct, double cov_fct_shape, const char* gp_approx,
                      double cov_fct_taper_range, double cov_fct_taper_shape, int num_neighbors, const char* vecchia_ordering,
                      int num_ind_points, double cover_tree_radius, const char* ind_points_selection, const char* likelihood,
                      double likelihood_additional_param, const char* matrix_inversion_method, int seed,
                      int num_parallel_threads, bool GPU_use, bool has_weights, const double* weights,
                      double likelihood_learning_rate, REModelHandle* out) {
  API_BEGIN();
  *out = new gpb200::REModel(num_data, cluster_ids_data, re_group_data, num_re_group, re_group_rand_coef_data,
                             ind_effect_group_rand_coef, num_re_group_rand_coef, drop_intercept_group_rand_effect, num_gp,
                             gp_coords_data, dim_gp_coords, gp_rand_coef_data, num_gp_rand_coef, cov_fct, cov_fct_shape,
                             gp_approx, cov_fct_taper_range, cov_fct_taper_shape, num_neighbors, vecchia_ordering,
                             num_ind_points, cover_tree_radius, ind_points_selection, likelihood,
                             likelihood_additional_param, matrix_inversion_method, seed, num_parallel_threads, GPU_use,
                             has_weights, weights, likelihood_learning_rate);
  API_END();
}

int GPB_REModelFree(REModelHandle handle) {
  API_BEGIN();
  delete reinterpret_cast<gpb200::REModel*>(handle);
  API_END();
}

int GPB_SetOptimConfig(REModelHandle handle, double* init_cov_pars, double lr, double /*acc_rate_cov*/, int max_iter,
                       double delta_rel_conv, bool /*use_nesterov_acc*/, int /*nesterov_schedule_version*/, bool trace,
                       const char* optimizer, int /*momentum_offset*/, const char* convergence_criterion,
                       int /*num_covariates*/, double* /*init_coef*/, double /*lr_coef*/, double /*acc_rate_coef*/,
                       const char* /*optimizer_coef*/, int /*cg_max_num_it*/, int /*cg_max_num_it_tridiag*/,
                       double /*cg_delta_conv*/, int /*num_rand_vec_trace*/, bool /*reuse_rand_vec_trace*/,
                       const char* /*cg_preconditioner_type*/, int /*seed_rand_vec_trace*/, int /*piv_chol_rank*/,
                       double* /*init_aux_pars*/, bool /*estimate_aux_pars*/, bool /*init_coef_aux_pars_from_iid_model*/,
                       const int* estimate_cov_par_index, int m_lbfgs, double /*delta_conv_mode_finding*/) {
  API_BEGIN();
  M(handle)->SetOptimConfig(init_cov_pars, lr, max_iter, delta_rel_conv, trace, optimizer, convergence_criterion, m_lbfgs,
                            estimate_cov_par_index);
  API_END();
}

int GPB_OptimCovPar(REModelHandle handle, const double* y_data, const double* fixed_effects) {
  API_BEGIN();
  M(handle)->OptimCovPar(y_data, fixed_effects, false, false);
  API_END();
}

int GPB_EvalNegLogLikelihood(REModelHandle handle, const double* y_data, double* cov_pars, const double* fixed_effects,
                             double* negll) {
  API_BEGIN();
  M(handle)->EvalNegLogLikelihood(y_data, cov_pars, negll, fixed_effects);
  API_END();
}

int GPB_GetCurrentNegLogLikelihood(REModelHandle handle, double* negll) {
  API_BEGIN();
  *negll = M(handle)->CurrentNegLogLikelihood();
  API_END();
}

int GPB_GetCovPar(REModelHandle handle, double* optim_cov_pars, bool calc_std_dev) {
  API_BEGIN();
  M(handle)->GetCovPar(optim_cov_pars, calc_std_dev);
  API_END();
}

int GPB_GetInitCovPar(REModelHandle handle, double* init_cov_pars) {
  API_BEGIN();
  M(handle)->GetInitCovPar(init_cov_pars);
  API_END();
}

int GPB_GetNumIt(REModelHandle handle, int* num_it) {
  API_BEGIN();
  *num_it = M(handle)->GetNumIt();
  API_END();
}

int GPB_HasStdCylBesselK(int* has_bessel) {
  API_BEGIN();
  *has_bessel = 0;  // general-smoothness Matern is outside the hot path
  API_END();
}

int GPB_GetLikelihoodName(REModelHandle handle, char* out_str, int* num_char) {
  API_BEGIN();
  CopyString(M(handle)->LikelihoodName(), out_str, num_char);
  API_END();
}

int GPB_GetOptimizerCovPars(REModelHandle handle, char* out_str, int* num_char) {
  API_BEGIN();
  CopyString(M(handle)->OptimizerCovPars(), out_str, num_char);
  API_END();
}

int GPB_CanCalculateStandardErrorsCovPars(REModelHandle handle, int* out) {
  API_BEGIN();
  M(handle);
  *out = 0;
  API_END();
}

int GPB200_SetDevice(int device) {
  API_BEGIN();
  if (device < 0 || device >= gpbdev_device_count())
    throw std::runtime_error("GPB200_SetDevice: no CUDA device " + std::to_string(device));
  gpb200::GetRuntime().device = device;
  API_END();
}

int GPB200_SetCollective(int rank, int world_size, void* allreduce_sum) {
  API_BEGIN();
  if (world_size < 1 || rank < 0 || rank >= world_size) throw std::runtime_error("GPB200_SetCollective: bad rank / world_size");
  if (world_size > 1 && allreduce_sum == nullptr) throw std::runtime_error("GPB200_SetCollective: world_size > 1 needs an all-reduce function");
  gpb200::Runtime& rt = gpb200::GetRuntime();
  rt.rank = rank;
  rt.world_size = world_size;
  rt.allreduce_sum = reinterpret_cast<gpb200::AllReduceSumFn>(allreduce_sum);
  API_END();
}

int GPB200_CalcGradient(REModelHandle handle, double* y_inout) {
  API_BEGIN();
  M(handle)->CalcGradient(y_inout, nullptr, true);
  API_END();
}

int GPB200_GetNumLikelihoodEvals(REModelHandle handle, int64_t* out) {
  API_BEGIN();
  *out = M(handle)->NumLikelihoodEvals();
  API_END();
}

int GPB200_GetDeviceEngine(REModelHandle handle, void** out) {
  API_BEGIN();
  *out = M(handle)->Engine();
  API_END();
}

}  // extern "C"
