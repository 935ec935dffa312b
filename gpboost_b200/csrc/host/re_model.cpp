// Host-side REModel: configuration parsing, Vecchia ordering, parameter transformations, initial values,
// the L-BFGS driver and the likelihood algebra around the device sums. See re_model.h.
#include "re_model.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <unordered_map>

#include "runtime.h"

namespace gpb200 {

namespace {

[[noreturn]] void Fatal(const std::string& msg) { throw std::runtime_error(msg); }

void DevCheck(int rc) {
  if (rc != 0) Fatal(std::string(gpbdev_last_error()));
}
void DenseCheck(int rc) {
  if (rc != 0) Fatal(std::string(gpbdev_dense_last_error()));
}
void GrpCheck(int rc) {
  if (rc != 0) Fatal(std::string(gpbdev_grouped_last_error()));
}

bool NearlyEqual(double a, double b) { return std::fabs(a - b) < 1e-10 * std::max({1.0, std::fabs(a), std::fabs(b)}); }

// likelihood aliases: include/GPBoost/likelihoods.h:10255-10280
std::string ParseLikelihoodAlias(const std::string& l) {
  if (l == "regression") return "gaussian";
  if (l == "binary" || l == "binary_logit") return "bernoulli_logit";
  return l;
}

}  // namespace

REModel::REModel(int32_t num_data, const int32_t* cluster_ids_data, const char* re_group_data, int32_t num_re_group,
                 const double* /*re_group_rand_coef_data*/, const int32_t* /*ind_effect_group_rand_coef*/,
                 int32_t num_re_group_rand_coef, const int* /*drop_intercept_group_rand_effect*/, int32_t num_gp,
                 const double* gp_coords_data, int dim_gp_coords, const double* /*gp_rand_coef_data*/, int32_t num_gp_rand_coef,
                 const char* cov_fct, double cov_fct_shape, const char* gp_approx, double /*cov_fct_taper_range*/,
                 double /*cov_fct_taper_shape*/, int num_neighbors, const char* vecchia_ordering, int /*num_ind_points*/,
                 double /*cover_tree_radius*/, const char* /*ind_points_selection*/, const char* likelihood,
                 double /*likelihood_additional_param*/, const char* matrix_inversion_method, int seed,
                 int /*num_parallel_threads*/, bool /*GPU_use*/, bool has_weights, const double* /*weights*/,
                 double likelihood_learning_rate) {
  // ---- checks in the order of REModelTemplate's constructor (re_model_template.h:102-472)
  if (!(num_data > 0)) Fatal("Check failed: num_data > 0");
  if (!(seed >= 0)) Fatal("Check failed: seed >= 0");
  num_data_ = num_data;
  rng_ = std::mt19937((uint32_t)seed);  // rng_ = RNG_t(seed), re_model_template.h:160
  likelihood_ = likelihood == nullptr ? "gaussian" : ParseLikelihoodAlias(likelihood);
  if (likelihood_ != "gaussian" && likelihood_ != "bernoulli_logit")
    Fatal("Likelihood '" + likelihood_ + "' is not supported by the B200 engine yet (hot path: 'gaussian', 'bernoulli_logit')");
  gauss_ = likelihood_ == "gaussian";
  if (!gauss_) {
    // SetDefaultMatrixInversionMethod / UseIterativeByDefault (re_model_template.h:7093-7105, 7431-7444): "default" resolves
    // to "iterative" for a non-Gaussian likelihood with a Vecchia approximation; only that variant runs on the device
    const std::string mim = matrix_inversion_method == nullptr ? "default" : std::string(matrix_inversion_method);
    if (mim != "default" && mim != "iterative")
      Fatal("matrix_inversion_method = '" + mim + "' is not supported for likelihood '" + likelihood_ +
            "' by the B200 engine (use 'iterative', the reference's default for this model)");
    if (num_re_group > 0 || num_gp != 1 || gp_approx == nullptr || std::string(gp_approx) != "vecchia")
      Fatal("Likelihood '" + likelihood_ + "' is only supported with a single GP and gp_approx = 'vecchia' by the B200 engine");
  }
  gp_approx_ = gp_approx == nullptr ? "none" : std::string(gp_approx);
  if (cluster_ids_data != nullptr) {
    for (int32_t i = 1; i < num_data; ++i)
      if (cluster_ids_data[i] != cluster_ids_data[0])
        Fatal("Multiple independent realizations ('cluster_ids') are not supported by the B200 engine yet");
  }
  if (num_re_group_rand_coef > 0) Fatal("Grouped random coefficients are not supported by the B200 engine yet");
  if (num_re_group > 0) {
    if (num_gp > 0) Fatal("Combined grouped random effects and Gaussian processes are not supported by the B200 engine yet");
    if (num_re_group != 1) Fatal("Only a single level of grouped random effects is supported by the B200 engine yet (hot path: config 3)");
    if (has_weights) Fatal("'weights' are not supported by the B200 engine yet");
    if (re_group_data == nullptr) Fatal("Check failed: re_group_data != nullptr");
    num_cov_pars_ = 2;  // error variance, group variance
    CreateGroupedBackend(re_group_data);
    estimate_cov_par_index_.assign(num_cov_pars_, 1);
    std::memset(sums_, 0, sizeof(sums_));
    return;
  }
  if (num_gp != 1) Fatal("num_gp can only be either 0 or 1 in the current implementation");
  if (num_gp_rand_coef > 0) Fatal("GP random coefficients are not supported by the B200 engine");
  if (has_weights) Fatal("'weights' are not supported by the B200 engine yet");
  if (!(likelihood_learning_rate > 0.)) Fatal("Check failed: likelihood_learning_rate > 0.");
  if (!(dim_gp_coords > 0)) Fatal("Check failed: dim_gp_coords > 0");
  if (gp_coords_data == nullptr) Fatal("Check failed: gp_coords_data != nullptr");
  if (cov_fct == nullptr) Fatal("Check failed: cov_fct != nullptr");
  dim_ = dim_gp_coords;
  // ---- covariance function (aliases cov_fcts.h:3170-3200; closed forms :2100-2164)
  cov_fct_ = cov_fct;
  shape_ = cov_fct_shape;
  if (cov_fct_ == "exponential" || cov_fct_ == "Matern") { cov_fct_ = "matern"; shape_ = 0.5; }
  if (cov_fct_ == "Gaussian") cov_fct_ = "gaussian";
  if (cov_fct_ == "matern") {
    if (NearlyEqual(shape_, 0.5)) cov_id_ = GPBDEV_COV_EXPONENTIAL;
    else if (NearlyEqual(shape_, 1.5)) cov_id_ = GPBDEV_COV_MATERN15;
    else if (NearlyEqual(shape_, 2.5)) cov_id_ = GPBDEV_COV_MATERN25;
    else Fatal("Only Matern smoothness 0.5, 1.5 and 2.5 are supported by the B200 engine (found " + std::to_string(shape_) + ")");
  } else if (cov_fct_ == "gaussian") {
    cov_id_ = GPBDEV_COV_GAUSSIAN;
  } else {
    Fatal("Covariance of type '" + cov_fct_ + "' is not supported by the B200 engine.");
  }
  num_cov_pars_ = gauss_ ? 3 : 2;  // (nugget,) marginal variance, range
  // ---- GP approximation
  std::memset(laplace_out_, 0, sizeof(laplace_out_));
  if (gp_approx_ == "none") {  // exact GP: dense Gram + Cholesky on the device, original observation order
    perm_.resize(num_data_);
    std::iota(perm_.begin(), perm_.end(), 0);
    coords_ordered_.resize((size_t)num_data_ * dim_);
    for (int32_t i = 0; i < num_data_; ++i)
      for (int k = 0; k < dim_; ++k) coords_ordered_[(size_t)i * dim_ + k] = gp_coords_data[(size_t)k * num_data_ + i];
    DenseCheck(gpbdev_dense_create(&dense_, GetRuntime().device, num_data_, dim_, coords_ordered_.data()));
    estimate_cov_par_index_.assign(num_cov_pars_, 1);
    std::memset(sums_, 0, sizeof(sums_));
    return;
  }
  if (gp_approx_ != "vecchia")
    Fatal("GP approximation '" + gp_approx_ + "' is currently not supported by the B200 engine (hot path: 'vecchia', 'none')");
  num_neighbors_ = num_neighbors > 0 ? num_neighbors : 20;  // re_model_template.h:288-294
  num_neighbors_pred_ = 2 * num_neighbors_;                  // re_model_template.h:299
  vecchia_ordering_ = vecchia_ordering == nullptr ? "none" : std::string(vecchia_ordering);
  if (vecchia_ordering_ != "none" && vecchia_ordering_ != "random")
    Fatal("Ordering of type '" + vecchia_ordering_ + "' is not supported for the Veccia approximation ");
  if (num_neighbors_ > num_data_ - 1) num_neighbors_ = std::max(num_data_ - 1, 1);  // Vecchia_utils.cpp:755-758
  // ---- ordering: data_indices_per_cluster = 0..n-1, shuffled with rng_ for "random" (Vecchia_utils.cpp:1129-1131)
  perm_.resize(num_data_);
  {
    std::vector<int> idx(num_data_);
    std::iota(idx.begin(), idx.end(), 0);
    if (vecchia_ordering_ == "random") std::shuffle(idx.begin(), idx.end(), rng_);
    for (int32_t i = 0; i < num_data_; ++i) perm_[i] = idx[i];
  }
  // coordinates arrive column-major (gp_coords_data[j*num_data+i], Vecchia_utils.cpp:1133-1137); keep them
  // row-major in Vecchia order for the device
  coords_ordered_.resize((size_t)num_data_ * dim_);
  for (int32_t i = 0; i < num_data_; ++i)
    for (int k = 0; k < dim_; ++k) coords_ordered_[(size_t)i * dim_ + k] = gp_coords_data[(size_t)k * num_data_ + perm_[i]];
  // ---- device state (+ device neighbour search)
  const Runtime& rt = GetRuntime();
  int64_t rb = 0, re = num_data_;
  if (rt.world_size > 1 && gauss_) {  // non-Gaussian: every rank holds the whole factor, the SLQ probe columns are sharded
    const int64_t chunk = (num_data_ + rt.world_size - 1) / rt.world_size;
    rb = std::min<int64_t>(num_data_, chunk * rt.rank);
    re = std::min<int64_t>(num_data_, rb + chunk);
  }
  DevCheck(gpbdev_vecchia_create(&engine_, rt.device, num_data_, dim_, num_neighbors_, coords_ordered_.data(), perm_.data(),
                                 nullptr, rb, re));
  if (rt.world_size > 1 && gauss_ && rt.allreduce_dev != nullptr) {
    // native collective (GPB200_NcclInit): the engine sums its shard results over the ranks on its own stream
    DevCheck(gpbdev_vecchia_set_allreduce(engine_, rt.allreduce_dev, rt.allreduce_ctx));
    device_collective_ = true;
  }
  estimate_cov_par_index_.assign(num_cov_pars_, 1);
  std::memset(sums_, 0, sizeof(sums_));
}

REModel::~REModel() {
  if (engine_) gpbdev_vecchia_free(engine_);
  if (grouped_) gpbdev_grouped_free(grouped_);
  if (dense_) gpbdev_dense_free(dense_);
}

void REModel::DensePass(double var, double range, bool with_grad) {
  double o[3];
  DenseCheck(gpbdev_dense_eval(dense_, cov_id_, var, range, o));
  sums_[GPBDEV_SUM_QUAD] = o[0];    // y' Psi^-1 y = ||L^-1 y||^2 (re_model_template.h:10002)
  sums_[GPBDEV_SUM_LOGDET] = o[1];  // 2 sum log L_ii (:3127)
  sums_[GPBDEV_SUM_NBAD] = o[2];
  ++num_ll_evals_;
  if (with_grad) {
    // re_model_template.h:2018-2039: grad_k = -alpha' dPsi_k alpha / (2 sigma^2) + tr(Psi^-1 dPsi_k) / 2, alpha = Psi^-1 y
    double g[4];
    DenseCheck(gpbdev_dense_grad(dense_, g));
    sums_[GPBDEV_SUM_UKU0] = 0.; sums_[GPBDEV_SUM_UKU1] = 0.;
    sums_[GPBDEV_SUM_TR0] = g[0]; sums_[GPBDEV_SUM_TR1] = g[1];
    sums_[GPBDEV_SUM_UDU0] = g[2]; sums_[GPBDEV_SUM_UDU1] = g[3];
  }
}

// group labels arrive as num_data NUL-terminated strings (c_api.h:1325, ConvertCharToStringGroupLevels); Z is kept as an
// int32 group index per observation (RECompGroup, re_comp.h:228-360)
void REModel::CreateGroupedBackend(const char* re_group_data) {
  std::unordered_map<std::string, int32_t> level_of;
  std::vector<int32_t> gidx(num_data_);
  const char* p = re_group_data;
  for (int32_t i = 0; i < num_data_; ++i) {
    std::string lab(p);
    p += lab.size() + 1;
    auto it = level_of.find(lab);
    if (it == level_of.end()) it = level_of.emplace(lab, (int32_t)level_of.size()).first;
    gidx[i] = it->second;
  }
  num_groups_ = (int)level_of.size();
  GrpCheck(gpbdev_grouped_create(&grouped_, GetRuntime().device, num_data_, gidx.data(), num_groups_));
}

void REModel::GroupedPass(double var_ratio) {
  GrpCheck(gpbdev_grouped_eval(grouped_, var_ratio, gsums_));
  sums_[GPBDEV_SUM_QUAD] = gsums_[0] - gsums_[1];  // y'Psi^-1 y (Woodbury, re_model_template.h:9966-9985)
  sums_[GPBDEV_SUM_LOGDET] = gsums_[2];            // log|Psi| (:3029-3031)
  ++num_ll_evals_;
}

// cov_fcts.h:485-552
void REModel::TransformCovPars(const double* orig, double* trans) const {
  const double s2 = orig[0];
  trans[0] = s2;
  trans[1] = orig[1] / s2;
  if (grouped_) return;  // RECompGroup: only the variance is rescaled (re_comp.h:300-310)
  if (!(orig[2] > 0.)) Fatal("Check failed: pars[1] > 0.");
  switch (cov_id_) {
    case GPBDEV_COV_EXPONENTIAL: trans[2] = 1. / orig[2]; break;
    case GPBDEV_COV_MATERN15: trans[2] = std::sqrt(3.) / orig[2]; break;
    case GPBDEV_COV_MATERN25: trans[2] = std::sqrt(5.) / orig[2]; break;
    default: trans[2] = 1. / (orig[2] * orig[2]); break;
  }
}

// cov_fcts.h:560-623
void REModel::TransformBackCovPars(const double* trans, double* orig) const {
  if (!gauss_) {  // no nugget: [sigma_1^2, range]
    orig[0] = trans[0];
    switch (cov_id_) {
      case GPBDEV_COV_EXPONENTIAL: orig[1] = 1. / trans[1]; break;
      case GPBDEV_COV_MATERN15: orig[1] = std::sqrt(3.) / trans[1]; break;
      case GPBDEV_COV_MATERN25: orig[1] = std::sqrt(5.) / trans[1]; break;
      default: orig[1] = 1. / std::sqrt(trans[1]); break;
    }
    return;
  }
  const double s2 = trans[0];
  orig[0] = s2;
  orig[1] = s2 * trans[1];
  if (grouped_) return;
  switch (cov_id_) {
    case GPBDEV_COV_EXPONENTIAL: orig[2] = 1. / trans[2]; break;
    case GPBDEV_COV_MATERN15: orig[2] = std::sqrt(3.) / trans[2]; break;
    case GPBDEV_COV_MATERN25: orig[2] = std::sqrt(5.) / trans[2]; break;
    default: orig[2] = 1. / std::sqrt(trans[2]); break;
  }
}

void REModel::SetOptimConfig(const double* init_cov_pars, double lr, int max_iter, double delta_rel_conv, bool trace,
                             const char* optimizer, const char* convergence_criterion, int m_lbfgs,
                             const int* estimate_cov_par_index) {
  // re_model.cpp:SetOptimConfig / re_model_template.h:790-960: -999 and nullptr mean "keep the default"
  if (init_cov_pars != nullptr) {
    for (int i = 0; i < num_cov_pars_; ++i)
      if (!(init_cov_pars[i] > 0.) || std::isnan(init_cov_pars[i]) || std::isinf(init_cov_pars[i]))
        Fatal("Found negative, zero, NaN or Inf values in 'init_cov_pars'");
    init_cov_pars_.assign(num_cov_pars_, 0.);
    if (gauss_) TransformCovPars(init_cov_pars, init_cov_pars_.data());
    else { init_cov_pars_[0] = init_cov_pars[0]; init_cov_pars_[1] = TransformRange(init_cov_pars[1]); }
    cov_pars_ = init_cov_pars_;
    init_cov_pars_provided_ = true;
    cov_pars_initialized_ = true;
  }
  if (lr > 0.) lr_cov_init_ = lr;
  else if (!NearlyEqual(lr, -999.)) Fatal("lr_cov is not > 0");
  if (max_iter >= 0) max_iter_ = max_iter;
  if (delta_rel_conv > 0.) delta_rel_conv_ = delta_rel_conv;
  else if (!NearlyEqual(delta_rel_conv, -999.)) Fatal("delta_rel_conv is not > 0");
  trace_ = trace;
  if (optimizer != nullptr && std::string(optimizer) != "") {
    optimizer_ = optimizer;
    if (optimizer_ != "lbfgs")
      Fatal("Optimizer option '" + optimizer_ + "' is not supported for covariance parameters by the B200 engine (use 'lbfgs')");
  }
  if (convergence_criterion != nullptr && std::string(convergence_criterion) != "" &&
      std::string(convergence_criterion) != "default") {
    convergence_criterion_ = convergence_criterion;
    if (convergence_criterion_ != "relative_change_in_log_likelihood" && convergence_criterion_ != "relative_change_in_parameters")
      Fatal("Convergence criterion '" + convergence_criterion_ + "' is not supported.");
  }
  if (m_lbfgs > 0) m_lbfgs_ = m_lbfgs;
  if (estimate_cov_par_index != nullptr && estimate_cov_par_index[0] >= 0) {
    for (int i = 0; i < num_cov_pars_; ++i) {
      estimate_cov_par_index_[i] = estimate_cov_par_index[i];
      if (estimate_cov_par_index[i] <= 0) Fatal("Holding covariance parameters fixed ('estimate_cov_par_index') is not supported by the B200 engine yet");
    }
  }
}

// re_model_template.h:4849-4968 (Gaussian branch) + cov_fcts.h:1422-1690 (median-distance range heuristic)
void REModel::FindInitCovPar(const double* y_data, const double* fixed_effects, double* init_trans) {
  const int n = num_data_;
  double mean = 0., var = 0.;
  for (int i = 0; i < n; ++i) mean += fixed_effects ? y_data[i] - fixed_effects[i] : y_data[i];
  mean /= n;
  for (int i = 0; i < n; ++i) {
    const double r = (fixed_effects ? y_data[i] - fixed_effects[i] : y_data[i]) - mean;
    var += r * r;
  }
  var /= (n - 1);
  init_trans[0] = var / 2;  // nugget
  init_trans[1] = 1.;       // marginal variance on the transformed scale (init_marg_var / num_comps_total_)
  if (grouped_) return;     // RECompGroup::FindInitCovPar: pars[0] = marginal variance only (re_comp.h:411-417)
  // range: median pairwise distance on (a sub-sample of) the ORDERED coordinates, sampled with the model's rng_
  const int kMaxPoints = 1000;
  const int ns = n > kMaxPoints ? kMaxPoints : n;
  std::vector<int> sample(ns);
  if (ns < n) {
    std::uniform_int_distribution<> dis(0, n - 1);
    for (int i = 0; i < ns; ++i) sample[i] = dis(rng_);
  } else {
    std::iota(sample.begin(), sample.end(), 0);
  }
  std::vector<double> dists;
  dists.reserve((size_t)ns * (ns - 1) / 2);
  for (int i = 0; i < ns - 1; ++i)
    for (int j = i + 1; j < ns; ++j) {
      double s = 0.;
      for (int k = 0; k < dim_; ++k) {
        const double t = coords_ordered_[(size_t)sample[i] * dim_ + k] - coords_ordered_[(size_t)sample[j] * dim_ + k];
        s += t * t;
      }
      dists.push_back(std::sqrt(s));
    }
  if (dists.empty()) Fatal("Cannot find an initial value for the range parameter");
  const size_t pos_med = dists.size() / 2;
  std::nth_element(dists.begin(), dists.begin() + pos_med, dists.end());
  double med = dists[pos_med];
  if (dists.size() % 2 == 0) {
    std::nth_element(dists.begin(), dists.begin() + pos_med - 1, dists.end());
    med = (med + dists[pos_med - 1]) / 2.;
  }
  if (med < 1e-10) med = std::accumulate(dists.begin(), dists.end(), 0.) / dists.size();
  if (med < 1e-10)
    Fatal("Cannot find an initial value for the range parameter since both the median and the average distances among coordinates are zero ");
  if (cov_fct_ == "matern") {
    if (shape_ <= 1.) init_trans[2] = 2. * 3. / med;
    else if (shape_ <= 2.) init_trans[2] = 2. * 4.7 / med;
    else init_trans[2] = 2. * 5.9 / med;
  } else {
    init_trans[2] = 3. / std::pow(med / 2., 2.);
  }
}

// re_model.cpp:1312-1334
void REModel::InitializeCovParsIfNotDefined(const double* y_data, const double* fixed_effects) {
  if (cov_pars_initialized_) return;
  if (init_cov_pars_provided_) {
    cov_pars_ = init_cov_pars_;
  } else {
    cov_pars_.assign(num_cov_pars_, 0.);
    FindInitCovPar(y_data, fixed_effects, cov_pars_.data());
    init_cov_pars_ = cov_pars_;
  }
  cov_pars_initialized_ = true;
}

void REModel::SetY(const double* y_data, const double* fixed_effects) {
  y_has_been_set_ = true;
  const double* src = y_data;
  if (fixed_effects != nullptr) {  // y - fixed_effects (re_model_template.h:2907-2917)
    work_.resize(num_data_);
    for (int32_t i = 0; i < num_data_; ++i) work_[i] = y_data[i] - fixed_effects[i];
    src = work_.data();
  }
  if (grouped_) GrpCheck(gpbdev_grouped_set_y(grouped_, src));
  else if (dense_) DenseCheck(gpbdev_dense_set_y(dense_, src));
  else DevCheck(gpbdev_vecchia_set_y(engine_, src));
}

void REModel::DevicePass(double var, double range, int mode) {
  DevCheck(gpbdev_vecchia_eval(engine_, cov_id_, var, range, mode, sums_));
  const Runtime& rt = GetRuntime();
  if (rt.world_size > 1 && !device_collective_) {
    if (rt.allreduce_sum == nullptr) Fatal("world_size > 1 but no all-reduce callback was registered (GPB200_SetCollective)");
    rt.allreduce_sum(sums_, GPBDEV_NUM_SUMS);
  }
  ++num_ll_evals_;
  if (sums_[GPBDEV_SUM_NBAD] > 0.) {
    // Vecchia_utils.cpp:1685-1698: warning for Gaussian likelihoods; the likelihood becomes NaN/Inf and the
    // line search backs off
  }
}

// re_model_template.h:3132
double REModel::NegLLFromSums(double sigma2) const {
  const double kLog2Pi = std::log(2. * M_PI);
  return sums_[GPBDEV_SUM_QUAD] / 2. / sigma2 + sums_[GPBDEV_SUM_LOGDET] / 2. + num_data_ / 2. * (std::log(sigma2) + kLog2Pi);
}

double REModel::TransformRange(double range) const {
  if (!(range > 0.)) Fatal("Check failed: pars[1] > 0.");
  switch (cov_id_) {
    case GPBDEV_COV_EXPONENTIAL: return 1. / range;
    case GPBDEV_COV_MATERN15: return std::sqrt(3.) / range;
    case GPBDEV_COV_MATERN25: return std::sqrt(5.) / range;
    default: return 1. / (range * range);
  }
}

void REModel::SetIterativeConfig(int cg_max_num_it, int cg_max_num_it_tridiag, double cg_delta_conv, int num_rand_vec_trace,
                                 const char* cg_preconditioner_type, int seed_rand_vec_trace, double delta_conv_mode_finding) {
  // re_model_template.h:860-900: non-positive / -999 values keep the defaults
  if (cg_max_num_it > 0) cg_max_num_it_ = cg_max_num_it;
  if (cg_max_num_it_tridiag > 0) cg_max_num_it_tridiag_ = cg_max_num_it_tridiag;
  if (cg_delta_conv > 0.) cg_delta_conv_ = cg_delta_conv;
  if (num_rand_vec_trace > 0) num_rand_vec_trace_ = num_rand_vec_trace;
  if (seed_rand_vec_trace >= 0) seed_rand_vec_trace_ = seed_rand_vec_trace;
  if (delta_conv_mode_finding > 0.) delta_conv_mode_finding_ = delta_conv_mode_finding;
  if (!gauss_ && cg_preconditioner_type != nullptr && std::string(cg_preconditioner_type) != "") {
    std::string pc(cg_preconditioner_type);
    if (pc == "Sigma_inv_plus_BtWB" || pc == "vadu" || pc == "VADU") pc = "vadu";  // ParsePreconditionerAlias
    if (pc != "vadu")
      Fatal("Preconditioner type '" + pc + "' is not supported by the B200 engine (hot path: 'vadu', the reference's default)");
  }
}

// GenRandVecNormalParallel (src/GPBoost/CG_utils.cpp:978-994): column col_i of the probe matrix is drawn from
// mt19937(seed_seq{seed, run_id lo, run_id hi, col_i}) with std::normal_distribution — the same standard-library calls,
// so the stochastic Lanczos quadrature sees the reference's probe vectors. Rows index the latent process in Vecchia order.
void REModel::EnsureProbes() {
  if (probes_t_ == num_rand_vec_trace_ && probes_seed_ == seed_rand_vec_trace_) return;  // reuse_rand_vec_trace
  const int t = num_rand_vec_trace_;
  // multi-GPU: rank r owns the probe columns r, r + world, r + 2 world, ... (the columns' generators are independent)
  const Runtime& rt = GetRuntime();
  std::vector<int> cols;
  for (int col = rt.rank; col < t; col += rt.world_size) cols.push_back(col);
  if (cols.empty()) Fatal("num_rand_vec_trace is smaller than the number of ranks");
  const int tl = (int)cols.size();
  std::vector<double> probes((size_t)num_data_ * tl);
  const uint64_t run_id = cg_generator_counter_;
  const uint32_t b32 = static_cast<uint32_t>(seed_rand_vec_trace_);
#pragma omp parallel for schedule(static) num_threads(16)
  for (int lc = 0; lc < tl; ++lc) {
    std::normal_distribution<double> ndist(0.0, 1.0);
    std::seed_seq seq{b32, static_cast<uint32_t>(run_id), static_cast<uint32_t>(run_id >> 32), static_cast<uint32_t>(cols[lc])};
    std::mt19937 gen(seq);
    double* dst = probes.data() + (size_t)lc * num_data_;
    for (int32_t row = 0; row < num_data_; ++row) dst[row] = ndist(gen);
  }
  ++cg_generator_counter_;
  DevCheck(gpbdev_vecchia_laplace_set_probes(engine_, probes.data(), tl));
  if (rt.world_size > 1) {
    if (rt.allreduce_sum == nullptr) Fatal("world_size > 1 but no all-reduce callback was registered (GPB200_SetCollective)");
    DevCheck(gpbdev_vecchia_laplace_set_collective(engine_, rt.allreduce_sum, t));
  }
  probes_t_ = t;
  probes_seed_ = seed_rand_vec_trace_;
}

// EvalLaplaceApproxNegLogLikelihood (re_model_template.h:3175-3214): mode re-initialised to 0, factor at cov_pars, then
// FindModePostRandEffCalcMLLVecchia on the device
void REModel::EvalLaplace(const double* y_data, const double* cov_pars, double* negll, const double* fixed_effects) {
  double var, range_t;
  if (cov_pars == nullptr) {
    if (!cov_pars_initialized_) Fatal("Check failed: cov_pars != nullptr");
    var = cov_pars_[0]; range_t = cov_pars_[1];
  } else {
    for (int i = 0; i < num_cov_pars_; ++i)
      if (!(cov_pars[i] > 0.)) Fatal("Covariance parameters must be positive");
    var = cov_pars[0]; range_t = TransformRange(cov_pars[1]);
  }
  if (y_data != nullptr) {
    for (int32_t i = 0; i < num_data_; ++i)  // CheckY, likelihoods.h:1321-1329
      if (std::fabs(y_data[i]) >= 1e-10 && !NearlyEqual(y_data[i], 1.))
        Fatal("The response variable ('y') needs to be 0 or 1 for likelihood = '" + likelihood_ + "' ");
    DevCheck(gpbdev_vecchia_set_y(engine_, y_data));
  }
  EnsureProbes();
  const double cfg[8] = {1000., delta_conv_mode_finding_, 20., (double)cg_max_num_it_, (double)cg_max_num_it_tridiag_,
                         cg_delta_conv_, 1., 1e-4};  // likelihoods.h:17316-17332
  DevCheck(gpbdev_vecchia_laplace_eval(engine_, cov_id_, var, range_t, fixed_effects, cfg, laplace_out_));
  ++num_ll_evals_;
  *negll = laplace_out_[0];
  neg_log_likelihood_ = *negll;
}

// Laplace-approximated negative log-likelihood AND its gradient w.r.t. (log variance, log range) at cov_pars (original scale):
// what EvalLLforLBFGSpp / CalcGradPars hand to the optimiser for a non-Gaussian likelihood (re_model_template.h:2055-2100).
void REModel::EvalLaplaceWithGradient(const double* y_data, const double* cov_pars, const double* fixed_effects, double* negll, double* grad2) {
  if (gauss_ || engine_ == nullptr) Fatal("EvalLaplaceWithGradient: only for non-Gaussian likelihoods with a Vecchia GP");
  DevCheck(gpbdev_vecchia_laplace_keep_solutions(engine_, 1));
  EvalLaplace(y_data, cov_pars, negll, fixed_effects);
  const double var = cov_pars ? cov_pars[0] : cov_pars_[0];
  const double range_t = cov_pars ? TransformRange(cov_pars[1]) : cov_pars_[1];
  const double cfg[8] = {1000., delta_conv_mode_finding_, 20., (double)cg_max_num_it_, (double)cg_max_num_it_tridiag_,
                         cg_delta_conv_, 1., 1e-4};
  double out[3] = {0., 0., 0.};
  DevCheck(gpbdev_vecchia_laplace_grad(engine_, cov_id_, var, range_t, cfg, out));
  DevCheck(gpbdev_vecchia_laplace_keep_solutions(engine_, 0));
  grad2[0] = out[0]; grad2[1] = out[1];
}

void REModel::GetLaplaceMode(double* out) const {
  if (gauss_ || engine_ == nullptr) Fatal("The posterior mode is only available for non-Gaussian likelihoods");
  DevCheck(gpbdev_vecchia_laplace_get_mode(engine_, out));
}

void REModel::EvalNegLogLikelihood(const double* y_data, const double* cov_pars, double* negll, const double* fixed_effects) {
  if (!gauss_) { EvalLaplace(y_data, cov_pars, negll, fixed_effects); return; }
  double trans[3] = {0., 0., 1.};
  if (cov_pars == nullptr) {
    if (y_data != nullptr) InitializeCovParsIfNotDefined(y_data, fixed_effects);
    if (!cov_pars_initialized_) Fatal("Check failed: cov_pars_initialized_");
    for (int i = 0; i < num_cov_pars_; ++i) trans[i] = cov_pars_[i];
  } else {
    for (int i = 0; i < num_cov_pars_; ++i)
      if (!(cov_pars[i] > 0.)) Fatal("Covariance parameters must be positive");
    TransformCovPars(cov_pars, trans);
  }
  if (fixed_effects != nullptr && y_data == nullptr) Fatal("EvalNegLogLikelihoodGauss: 'y_data' cannot nullptr when 'fixed_effects' is provided ");
  if (y_data != nullptr) SetY(y_data, fixed_effects);
  if (grouped_) GroupedPass(trans[1]);
  else if (dense_) DensePass(trans[1], trans[2]);
  else DevicePass(trans[1], trans[2], GPBDEV_MODE_NLL);
  *negll = NegLLFromSums(trans[0]);
  // Gaussian data: the value goes to the caller only; GPB_GetCurrentNegLogLikelihood keeps reporting the last optimisation's value
  // (the reference passes `negll` by reference here and stores neg_log_likelihood_ in CalcCovFactorOrModeAndNegLL, i.e. during fits:
  // re_model.cpp:752-794, re_model_template.h:2832-2851)
}

void REModel::SetPredictionData(int32_t num_data_pred, const double* gp_coords_data_pred, const char* vecchia_pred_type, int num_neighbors_pred) {
  if (engine_ == nullptr) Fatal("Prediction data can only be set for a Vecchia GP model in the B200 build");
  if (vecchia_pred_type != nullptr && vecchia_pred_type[0] != '\0') {
    const std::string t(vecchia_pred_type);
    if (gauss_ && t != "order_obs_first_cond_obs_only")
      Fatal("Prediction type '" + t + "' is not supported by the B200 engine (supported: 'order_obs_first_cond_obs_only', the reference's default)");
    if (!gauss_) Fatal("Prediction is not supported for likelihood '" + likelihood_ + "' by the B200 engine yet");
  }
  if (num_neighbors_pred > 0) num_neighbors_pred_ = num_neighbors_pred;
  if (gp_coords_data_pred != nullptr) {
    if (!(num_data_pred > 0)) Fatal("Check failed: num_data_pred > 0");
    coords_pred_saved_.resize((size_t)num_data_pred * dim_);
    for (int32_t i = 0; i < num_data_pred; ++i)
      for (int k = 0; k < dim_; ++k) coords_pred_saved_[(size_t)i * dim_ + k] = gp_coords_data_pred[(size_t)k * num_data_pred + i];
    num_data_pred_saved_ = num_data_pred;
  }
}

void REModel::Predict(const double* y_obs, int32_t num_data_pred, double* out_predict, bool predict_cov_mat, bool predict_var,
                      bool predict_response, const double* gp_coords_data_pred, const double* cov_pars_pred, bool use_saved_data,
                      const double* fixed_effects) {
  if (engine_ == nullptr) Fatal("GPB_PredictREModel: only the Vecchia GP model predicts on the device in the B200 build (grouped / exact models: not yet)");
  if (!gauss_) Fatal("Prediction is not supported for likelihood '" + likelihood_ + "' by the B200 engine yet");
  if (predict_cov_mat) Fatal("Predictive covariance matrices are not supported by the B200 engine (predict_var gives the variances)");
  if (out_predict == nullptr) Fatal("Check failed: out_predict != nullptr");
  double trans[3] = {0., 0., 1.};
  if (cov_pars_pred != nullptr) {
    for (int i = 0; i < num_cov_pars_; ++i)
      if (!(cov_pars_pred[i] > 0.)) Fatal("Covariance parameters must be positive");
    TransformCovPars(cov_pars_pred, trans);
  } else {
    if (!cov_pars_initialized_) Fatal("Covariance parameters have not been estimated or are not given.");  // re_model.cpp:1119-1121
    for (int i = 0; i < num_cov_pars_; ++i) trans[i] = cov_pars_[i];
  }
  const double* cp = nullptr;
  std::vector<double> rowmajor;
  if (use_saved_data) {
    if (num_data_pred_saved_ <= 0) Fatal("No data has been set for making predictions. Call set_prediction_data first");
    if (num_data_pred != num_data_pred_saved_) Fatal("Check failed: num_data_pred == num_data_pred_ (saved prediction data)");
    cp = coords_pred_saved_.data();
  } else {
    if (gp_coords_data_pred == nullptr) Fatal("Check failed: gp_coords_data_pred != nullptr");
    if (!(num_data_pred > 0)) Fatal("Check failed: num_data_pred > 0");
    rowmajor.resize((size_t)num_data_pred * dim_);
    for (int32_t i = 0; i < num_data_pred; ++i)
      for (int k = 0; k < dim_; ++k) rowmajor[(size_t)i * dim_ + k] = gp_coords_data_pred[(size_t)k * num_data_pred + i];
    cp = rowmajor.data();
  }
  (void)fixed_effects;  // ignored for Gaussian data (c_api.h:1640 ff.)
  if (y_obs != nullptr) SetY(y_obs, nullptr);
  else if (!y_has_been_set_) Fatal("Response variable data is not available for making predictions (pass y or fit the model first)");
  std::vector<double> dvar((size_t)num_data_pred);
  DevCheck(gpbdev_vecchia_predict(engine_, cov_id_, trans[1], trans[2], cp, num_data_pred, num_neighbors_pred_, out_predict, dvar.data()));
  if (predict_var) {
    // back to the original scale (re_model_template.h:3427 ff.): sigma^2 D_p, plus the error variance when the response is predicted
    const double add = predict_response ? 1. : 0.;
    for (int32_t i = 0; i < num_data_pred; ++i) out_predict[(size_t)num_data_pred + i] = trans[0] * (dvar[(size_t)i] + add);
  }
}

void REModel::OptimCovPar(const double* y_data, const double* fixed_effects, bool called_in_GPBoost_algorithm,
                          bool reuse_learning_rates_from_previous_call) {
  if (!gauss_) {
    // L-BFGS on log(cov_pars) with the device gradient of the Laplace-approximated likelihood (tests/test_laplace_gpu.py: gradient and
    // fits against the reference's goldens on the B200; tests/test_laplace_oracle_pinned.py: the same driver fed by the oracle)
    OptimCovParLaplace(y_data, fixed_effects);
    return;
  }
  if (y_data == nullptr) Fatal("Check failed: y_data != nullptr");
  for (int32_t i = 0; i < num_data_; ++i)
    if (std::isnan(y_data[i]) || std::isinf(y_data[i])) Fatal("NaN or Inf in response variable / label ");
  InitializeCovParsIfNotDefined(y_data, fixed_effects);
  SetY(y_data, fixed_effects);
  OptimCovParCore(called_in_GPBoost_algorithm, reuse_learning_rates_from_previous_call);
}

// GPB_OptimCovPar for a non-Gaussian likelihood with a latent Vecchia GP: L-BFGS on log(cov_pars) (original scale, no nugget),
// objective = Laplace-approximated negative log-likelihood, gradient = CalcGradNegMargLikelihoodLaplaceApproxVecchia
// (re_model_template.h:972-1800 with EvalLLforLBFGSpp, optim_utils.h:244-340). Initial values: marginal variance 1, range from the
// coordinates (FindInitCovPar, re_model_template.h:4901-4925).
void REModel::OptimCovParLaplace(const double* y_data, const double* fixed_effects) {
  if (y_data == nullptr) Fatal("Check failed: y_data != nullptr");
  if (!cov_pars_initialized_) {
    if (init_cov_pars_provided_) {
      cov_pars_ = init_cov_pars_;
    } else {
      double tmp[3] = {0., 0., 0.};
      FindInitCovPar(y_data, fixed_effects, tmp);  // tmp[2] = transformed range; the variance part is the Gaussian rule, unused here
      cov_pars_.assign(2, 0.);
      cov_pars_[0] = 1.;
      cov_pars_[1] = tmp[2];
      init_cov_pars_ = cov_pars_;
    }
    cov_pars_initialized_ = true;
  }
  num_it_ = max_iter_;
  if (max_iter_ <= 0) return;
  double orig[2];
  TransformBackCovPars(cov_pars_.data(), orig);
  bool have_cached = false;
  std::vector<double> cached_x, cached_grad(2, 0.);
  double cached_f = 0.;
  LbfgsObjective objective = [&](const std::vector<double>& x, std::vector<double>* grad, bool) -> double {
    if (grad != nullptr && have_cached && cached_x == x) { *grad = cached_grad; return cached_f; }
    const double cp[2] = {std::exp(x[0]), std::exp(x[1])};
    double f = 0.;
    if (grad != nullptr) {
      double g[2];
      EvalLaplaceWithGradient(y_data, cp, fixed_effects, &f, g);
      grad->assign(g, g + 2);
      cached_x = x; cached_grad = *grad; cached_f = f; have_cached = true;
    } else {
      EvalLaplace(y_data, cp, &f, fixed_effects);
    }
    return f;
  };
  LbfgsMaxStep max_step = [&](const std::vector<double>& neg_dir) {
    double mx = 0.;
    for (double v : neg_dir) mx = std::max(mx, std::fabs(v));
    return std::log(100.) / mx;
  };
  LbfgsHook hook = [](bool) {};
  LbfgsParams par;
  par.max_iterations = max_iter_;
  par.delta = delta_rel_conv_;
  par.m = m_lbfgs_;
  par.initial_step_factor = lr_cov_init_;
  std::vector<double> x = {std::log(orig[0]), std::log(orig[1])};
  double fx = 0.;
  LbfgsMemory mem;
  num_it_ = lbfgs_minimize(objective, max_step, hook, par, &x, &fx, &mem, false);
  cov_pars_[0] = std::exp(x[0]);
  cov_pars_[1] = TransformRange(std::exp(x[1]));
  for (double v : cov_pars_)
    if (std::isnan(v) || std::isinf(v)) Fatal("NaN or Inf occurred in covariance parameter optimization using 'lbfgs'");
  neg_log_likelihood_ = fx;
  cov_pars_estimated_once_ = true;
}

bool REModel::DevicePathReady() const {
  // initial covariance parameters come from the sample variance of the response (FindInitCovPar): the first call takes the
  // host-pointer form; the dense backend and an injected host collective have no device-resident entry
  const Runtime& rt = GetRuntime();
  return gauss_ && cov_pars_initialized_ && dense_ == nullptr && !(rt.world_size > 1 && !device_collective_);
}

void REModel::SetYDevice(const double* y_dev) {
  y_has_been_set_ = true;
  if (grouped_) GrpCheck(gpbdev_grouped_set_y_device(grouped_, y_dev));
  else DevCheck(gpbdev_vecchia_set_y_device(engine_, y_dev));
}

void REModel::OptimCovParDevice(const double* y_dev, bool called_in_GPBoost_algorithm, bool reuse_learning_rates_from_previous_call) {
  if (y_dev == nullptr) Fatal("Check failed: y_data != nullptr");
  if (!DevicePathReady()) Fatal("OptimCovParDevice: no device-resident path for this model state (use OptimCovPar)");
  SetYDevice(y_dev);
  OptimCovParCore(called_in_GPBoost_algorithm, reuse_learning_rates_from_previous_call);
}

void REModel::OptimCovParCore(bool called_in_GPBoost_algorithm, bool reuse_learning_rates_from_previous_call) {
  num_it_ = max_iter_;
  if (max_iter_ <= 0) return;
  const bool reuse_mem = reuse_learning_rates_from_previous_call && called_in_GPBoost_algorithm && cov_pars_estimated_once_;
  // optimisation variables: log of the transformed (variance ratio, range); the error variance is profiled out
  // (optim_utils.h:244-340, re_model_template.h:1082-1084, :2640-2650)
  double sigma2 = cov_pars_[0], sigma2_lag1 = cov_pars_[0];
  bool have_cached_grad = false;
  std::vector<double> cached_x, cached_grad;
  double cached_f = 0.;
  auto grad_from_sums = [&](double s2, std::vector<double>* g) {
    for (int k = 0; k < 2; ++k)  // re_model_template.h:2002-2004
      (*g)[k] = (sums_[GPBDEV_SUM_UKU0 + k] - 0.5 * sums_[GPBDEV_SUM_UDU0 + k]) / s2 + 0.5 * sums_[GPBDEV_SUM_TR0 + k];
  };
  LbfgsObjective objective_grouped = [&](const std::vector<double>& x, std::vector<double>* grad, bool) -> double {
    GroupedPass(std::exp(x[0]));
    sigma2 = sums_[GPBDEV_SUM_QUAD] / num_data_;  // ProfileOutSigma2
    // d negll / d log v at the profiled sigma^2: -(d yPy)/(2 sigma^2) ... + (1/2) d log|Psi|
    if (grad != nullptr) (*grad)[0] = -gsums_[3] / (2. * sigma2) + 0.5 * gsums_[4];
    return NegLLFromSums(sigma2);
  };
  LbfgsObjective objective = [&](const std::vector<double>& x, std::vector<double>* grad, bool speculative) -> double {
    if (grad != nullptr && have_cached_grad && cached_x == x) {  // gradient right after an accepted first trial
      *grad = cached_grad;
      return cached_f;
    }
    const double var = std::exp(x[0]), range = std::exp(x[1]);
    // the dense gradient costs two more n^3/3 passes: no speculative gradient with the first line-search trial
    const bool with_grad = grad != nullptr || (speculative && !dense_);
    if (dense_) DensePass(var, range, with_grad);
    else DevicePass(var, range, with_grad ? GPBDEV_MODE_GRAD : GPBDEV_MODE_NLL);
    sigma2 = sums_[GPBDEV_SUM_QUAD] / num_data_;  // ProfileOutSigma2
    const double f = NegLLFromSums(sigma2);
    have_cached_grad = false;
    if (with_grad) {
      cached_grad.assign(2, 0.);
      grad_from_sums(sigma2, &cached_grad);
      cached_x = x; cached_f = f; have_cached_grad = true;
      if (grad != nullptr) *grad = cached_grad;
    }
    return f;
  };
  LbfgsMaxStep max_step = [&](const std::vector<double>& neg_dir) {  // re_model_template.h:5413-5421
    double mx = 0.;
    for (double v : neg_dir) mx = std::max(mx, std::fabs(v));
    return std::log(100.) / mx;
  };
  LbfgsHook hook = [&](bool commit) {  // Set/ResetProfiledOutVariables (re_model_template.h:2798-2823)
    if (commit) sigma2_lag1 = sigma2; else sigma2 = sigma2_lag1;
  };
  LbfgsParams par;
  par.max_iterations = max_iter_;
  par.delta = delta_rel_conv_;
  par.m = m_lbfgs_;
  par.initial_step_factor = lr_cov_init_;
  std::vector<double> x = {std::log(cov_pars_[1])};
  if (!grouped_) x.push_back(std::log(cov_pars_[2]));
  double fx = 0.;
  num_it_ = lbfgs_minimize(grouped_ ? objective_grouped : objective, max_step, hook, par, &x, &fx, &lbfgs_mem_, reuse_mem);
  cov_pars_[0] = sigma2;
  cov_pars_[1] = std::exp(x[0]);
  if (!grouped_) cov_pars_[2] = std::exp(x[1]);
  for (double v : cov_pars_)
    if (std::isnan(v) || std::isinf(v)) Fatal("NaN or Inf occurred in covariance parameter optimization using 'lbfgs'");
  neg_log_likelihood_ = fx;
  cov_pars_estimated_once_ = true;
}

void REModel::CalcGradient(double* y, const double* fixed_effects, bool /*calc_cov_factor*/) {
  if (!gauss_) Fatal("CalcGradient for likelihood '" + likelihood_ + "' is not built on the device yet");
  if (y == nullptr) Fatal("Check failed: y != nullptr");
  InitializeCovParsIfNotDefined(y, fixed_effects);
  // re_model_template.h:3298-3321: SetY(y); y_aux = Psi^-1 y / sigma^2; written back on y.
  // The factor is always recomputed here: the device keeps no B between calls unless a STORE pass ran.
  if (grouped_) {
    GrpCheck(gpbdev_grouped_set_y(grouped_, y));
    GrpCheck(gpbdev_grouped_yaux(grouped_, cov_pars_[1], 1. / cov_pars_[0], y));
    return;
  }
  if (dense_) {
    DenseCheck(gpbdev_dense_set_y(dense_, y));
    DensePass(cov_pars_[1], cov_pars_[2]);
    DenseCheck(gpbdev_dense_yaux(dense_, 1. / cov_pars_[0], y));
    return;
  }
  DevCheck(gpbdev_vecchia_set_y(engine_, y));
  DevicePass(cov_pars_[1], cov_pars_[2], GPBDEV_MODE_STORE);
  DevCheck(gpbdev_vecchia_yaux(engine_, y));
  const Runtime& rt = GetRuntime();
  if (rt.world_size > 1 && !device_collective_) rt.allreduce_sum(y, num_data_);
  const double inv_s2 = 1. / cov_pars_[0];
  for (int32_t i = 0; i < num_data_; ++i) y[i] *= inv_s2;
}

void REModel::CalcGradientDevice(double* y_dev, bool response_is_current) {
  if (!gauss_) Fatal("CalcGradient for likelihood '" + likelihood_ + "' is not built on the device yet");
  if (y_dev == nullptr) Fatal("Check failed: y != nullptr");
  if (!DevicePathReady()) Fatal("CalcGradientDevice: no device-resident path for this model state (use CalcGradient)");
  if (grouped_) {
    if (!response_is_current) GrpCheck(gpbdev_grouped_set_y_device(grouped_, y_dev));
    GrpCheck(gpbdev_grouped_yaux_device(grouped_, cov_pars_[1], 1. / cov_pars_[0], y_dev));
    return;
  }
  // (with the response still installed, a store request at the parameters of the optimiser's last gradient pass costs no launch:
  // that pass wrote A, D^-1 and u already — gpbdev_vecchia_eval)
  if (!response_is_current) DevCheck(gpbdev_vecchia_set_y_device(engine_, y_dev));
  DevicePass(cov_pars_[1], cov_pars_[2], GPBDEV_MODE_STORE);
  DevCheck(gpbdev_vecchia_yaux_device(engine_, y_dev, 1. / cov_pars_[0]));
}

void REModel::NewtonUpdateLeafValuesDevice(const int32_t* leaf_of_row_dev, int num_leaves, const double* grad_dev, double* leaf_values) {
  if (!gauss_) Fatal("Newton updates for leaf values is only supported for Gaussian data");  // re_model_template.h:4986-4988
  if (engine_ == nullptr) Fatal("Newton updates for leaf values are only built for the Vecchia GP model in the B200 engine");
  const int L = num_leaves;
  std::vector<double> M((size_t)L * L), rhs((size_t)L);
  DevCheck(gpbdev_vecchia_newton_system(engine_, leaf_of_row_dev, L, grad_dev, M.data(), rhs.data()));
  // new leaf values = (H^T Psi^-1 H)^-1 (-sigma^2 H^T g)   (:5057-5061: HTYAux *= marg_variance; llt().solve) — dense Cholesky, L <= 256
  for (int j = 0; j < L; ++j) {
    double d = M[(size_t)j * L + j];
    for (int k = 0; k < j; ++k) d -= M[(size_t)j * L + k] * M[(size_t)j * L + k];
    if (!(d > 0.)) Fatal("NewtonUpdateLeafValues: H^T Psi^-1 H is not positive definite");
    d = std::sqrt(d);
    M[(size_t)j * L + j] = d;
    for (int i = j + 1; i < L; ++i) {
      double v = M[(size_t)i * L + j];
      for (int k = 0; k < j; ++k) v -= M[(size_t)i * L + k] * M[(size_t)j * L + k];
      M[(size_t)i * L + j] = v / d;
    }
  }
  for (int i = 0; i < L; ++i) {
    double v = -cov_pars_[0] * rhs[i];
    for (int k = 0; k < i; ++k) v -= M[(size_t)i * L + k] * leaf_values[k];
    leaf_values[i] = v / M[(size_t)i * L + i];
  }
  for (int i = L - 1; i >= 0; --i) {
    double v = leaf_values[i];
    for (int k = i + 1; k < L; ++k) v -= M[(size_t)k * L + i] * leaf_values[k];
    leaf_values[i] = v / M[(size_t)i * L + i];
  }
}

void REModel::GetCovPar(double* out, bool calc_std_dev) const {
  if (!cov_pars_initialized_) Fatal("Covariance parameters have not been estimated or set");
  if (calc_std_dev) Fatal("Standard deviations of covariance parameters are not available in the B200 engine");
  TransformBackCovPars(cov_pars_.data(), out);
}

void REModel::GetInitCovPar(double* out) const {
  if (init_cov_pars_.empty()) {
    for (int i = 0; i < num_cov_pars_; ++i) out[i] = -1.;  // re_model.cpp: not yet determined
    return;
  }
  TransformBackCovPars(init_cov_pars_.data(), out);
}

}  // namespace gpb200
