/*
 * TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by, or called from the product
 * library (gpboost_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Plain-C restatement of the reference's CPU algorithm for the Vecchia-approximated Gaussian-process
 * hot path (fabsig/GPBoost @ c93fa49). Parity is PINNED: tests/test_oracle_pinned.py checks this file
 * against (i) the golden negative log-likelihoods hard-coded in the reference's own R tests
 * (R-package/tests/testthat/test_GPModel_gaussian_process.R:86-120,1145-1149) and (ii) outputs of the
 * unmodified reference library built by oracle/Makefile.ref (oracle/_ref/lib_gpboost.so).
 *
 * Each function cites the reference file:line it restates.  All arithmetic is fp64; indices are int32.
 * Row i of B is stored dense as (n x m): Bneg[i*m + k] = A_i[k] = -B[i, nn[i*m+k]]  (B = I - A),
 * rows with fewer than m neighbours (i < m) are padded with nn = -1, A = 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* covariance function ids shared with the product's C-ABI (include/gpboost_b200_dev.h) */
enum { ORC_COV_EXPONENTIAL = 0, ORC_COV_MATERN15 = 1, ORC_COV_MATERN25 = 2, ORC_COV_GAUSSIAN = 3 };

/* ---- covariance closed forms: include/GPBoost/cov_fcts.h:2100-2118 (Matern 0.5/1.5/2.5), :2154 (Gaussian).
 * 'var' and 'range' are the TRANSFORMED parameters (cov_fcts.h:485-552): var = sigma1^2/sigma^2,
 * range = 1/rho | sqrt(3)/rho | sqrt(5)/rho | 1/rho^2. */
static double orc_cov(int type, double dist, double var, double range) {
  double rd;
  switch (type) {
    case ORC_COV_EXPONENTIAL: return var * exp(-range * dist);
    case ORC_COV_MATERN15: rd = range * dist; return var * (1. + rd) * exp(-rd);
    case ORC_COV_MATERN25: rd = range * dist; return var * (1. + rd + rd * rd / 3.) * exp(-rd);
    default: return var * exp(-range * dist * dist);
  }
}

/* ---- d Sigma / d log(range_transformed) on the transformed scale (transf_scale = true):
 * constants cov_fcts.h:2183-2203 (DetermineConstantsForGradient), element formulas :2535-2563.
 * 'sigma' is the covariance value at this distance (needed by the 0.5 and Gaussian forms). */
static double orc_cov_grad_range(int type, double dist, double var, double range, double sigma) {
  double cm, rd;
  switch (type) {
    case ORC_COV_EXPONENTIAL: cm = -1. * range; return cm * dist * sigma;
    case ORC_COV_MATERN15: cm = -1. * var * range * range; return cm * dist * dist * exp(-range * dist);
    case ORC_COV_MATERN25:
      cm = -1. * var * range * range; rd = range * dist;
      return cm / 3. * dist * dist * (1. + rd) * exp(-rd);
    default: cm = -1. * range; return cm * dist * dist * sigma;
  }
}

/* Euclidean distance, sequential (non-FMA) accumulation as Eigen's row-expression redux does:
 * Vecchia_utils.cpp:798,966 ((a-b).lpNorm<2>()) and :1065 (squaredNorm). */
static double orc_sqdist(const double* coords, int n, int d, int a, int b) {
  double s = 0.;
  for (int k = 0; k < d; ++k) {
    double t = coords[(size_t)k * n + a] - coords[(size_t)k * n + b];
    s += t * t;
  }
  return s;
}

/* ---- insertion sort keeping (a,b) ascending in a: include/GPBoost/utils.h:250-262 */
static void orc_sort_increasing(double* a, int* b, int n) {
  for (int j = 1; j <= n - 1; ++j) {
    int k = j;
    while (k > 0 && a[k] < a[k - 1]) {
      double v = a[k]; int l = b[k];
      a[k] = a[k - 1]; b[k] = b[k - 1];
      a[k - 1] = v; b[k - 1] = l;
      --k;
    }
  }
}

/* std::sort-based index sort of the reference (utils.h:230-238), restated in shuffle_oracle.cpp */
void orc_sort_indices(const double* v, int n, int32_t* idx_out);

/* ---- Vecchia neighbour search among previously ordered points ("nearest" selection):
 * src/GPBoost/Vecchia_utils.cpp:733-985 (driver) and :1029-1093 (find_nearest_neighbors_fast_internal).
 * coords: column-major n x d (already in Vecchia order). nn: n x m int32, -1 padded.
 * The coordinate sums are sorted by the same std::sort call as the reference (orc_sort_indices), because the
 * order of equal sums decides which of several equidistant candidates is met first (lattice data). */
void orc_knn_vecchia(const double* coords, int n, int d, int m, int32_t* nn) {
  for (size_t t = 0; t < (size_t)n * m; ++t) nn[t] = -1;
  double* csum = (double*)malloc(sizeof(double) * n);
  int32_t* sort_sum = (int32_t*)malloc(sizeof(int32_t) * n);
  int* sort_inv = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) {
    double s = 0.;
    for (int k = 0; k < d; ++k) s += coords[(size_t)k * n + i]; /* :778 row sum */
    csum[i] = s;
  }
  orc_sort_indices(csum, n, sort_sum);
  for (int i = 0; i < n; ++i) sort_inv[sort_sum[i]] = i;
  int end_search_at = n - 2; /* :752-754 */
  /* :788-813 the first m+1 points condition on all predecessors, in index order */
  for (int i = 1; i < n && i <= m; ++i)
    for (int j = 0; j < i; ++j) nn[(size_t)i * m + j] = j;
#pragma omp parallel
  {
    double* sq = (double*)malloc(sizeof(double) * m);
#pragma omp for schedule(dynamic, 256)
    for (int i = m + 1; i < n; ++i) { /* :885-900 */
      int32_t* nb = nn + (size_t)i * m;
      for (int j = 0; j < m; ++j) sq[j] = INFINITY; /* :1040-1043 */
      int down = 1, up = 1, up_i = sort_inv[i], down_i = sort_inv[i];
      while (up || down) { /* :1049-1092 */
        if (down_i == 0) down = 0;
        if (up_i == n - 1) up = 0;
        if (down) {
          --down_i;
          int c = sort_sum[down_i];
          if (c < i && c <= end_search_at) {
            double dd = csum[c] - csum[i], smd = dd * dd; /* std::pow(x,2) == x*x */
            if (smd > d * sq[m - 1]) down = 0;
            else {
              double sed = orc_sqdist(coords, n, d, c, i);
              if (sed < sq[m - 1]) { sq[m - 1] = sed; nb[m - 1] = c; orc_sort_increasing(sq, nb, m); }
            }
          }
        }
        if (up) {
          ++up_i;
          int c = sort_sum[up_i];
          if (c < i && c <= end_search_at) {
            double dd = csum[c] - csum[i], smd = dd * dd;
            if (smd > d * sq[m - 1]) up = 0;
            else {
              double sed = orc_sqdist(coords, n, d, c, i);
              if (sed < sq[m - 1]) { sq[m - 1] = sed; nb[m - 1] = c; orc_sort_increasing(sq, nb, m); }
            }
          }
        }
      }
    }
    free(sq);
  }
  free(csum); free(sort_sum); free(sort_inv);
}

/* dense Cholesky (lower, in place, row-major ld=m) + solve; stands in for Eigen::LLT at
 * Vecchia_utils.cpp:1617-1618 (same algorithm class; rounding differs at the 1e-16 level). */
static int orc_chol(double* a, int m) {
  for (int j = 0; j < m; ++j) {
    double s = a[j * m + j];
    for (int k = 0; k < j; ++k) s -= a[j * m + k] * a[j * m + k];
    if (!(s > 0.)) return -1;
    double l = sqrt(s);
    a[j * m + j] = l;
    for (int i = j + 1; i < m; ++i) {
      double t = a[i * m + j];
      for (int k = 0; k < j; ++k) t -= a[i * m + k] * a[j * m + k];
      a[i * m + j] = t / l;
    }
  }
  return 0;
}
static void orc_chol_solve(const double* l, int m, double* b) {
  for (int i = 0; i < m; ++i) { /* L z = b */
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= l[i * m + k] * b[k];
    b[i] = t / l[i * m + i];
  }
  for (int i = m - 1; i >= 0; --i) { /* L^T x = z */
    double t = b[i];
    for (int k = i + 1; k < m; ++k) t -= l[k * m + i] * b[k];
    b[i] = t / l[i * m + i];
  }
}

/* ---- Vecchia factor B = I - A, D^-1 and (optionally) their derivatives w.r.t. the log of the
 * transformed covariance parameters (marginal variance, range), Gaussian likelihood, transformed scale
 * (nugget = 1): src/GPBoost/Vecchia_utils.cpp:1367-1699; covariance blocks via
 * include/GPBoost/re_comp.h:1476-1503 -> cov_fcts.h:635-755 (symmetric block: diagonal = var, :737),
 * gradients cov_fcts.h:1073-1218 (ind_par==0: dSigma = Sigma :1088-1093; diagonal of range gradient = 0 :1179).
 * pars = {var, range} transformed. Outputs: A (n x m), Dinv (n); if calc_grad: Agrad (2 x n x m) holding
 * dA = -dB, Dgrad (2 x n). Returns number of non-positive D (Vecchia_utils.cpp:1685-1698). */
int orc_vecchia_factor(const double* coords, int n, int d, int m, const int32_t* nn, int cov_type,
                       const double* pars, double* A, double* Dinv, int calc_grad, double* Agrad,
                       double* Dgrad) {
  const double var = pars[0], range = pars[1];
  int bad = 0;
#pragma omp parallel
  {
    double* S = (double*)malloc(sizeof(double) * m * m);   /* cov among neighbours (+ nugget) */
    double* G = (double*)malloc(sizeof(double) * m * m);   /* d/d log range of the same block */
    double* G0 = (double*)malloc(sizeof(double) * m * m);  /* d/d log var */
    double* s1 = (double*)malloc(sizeof(double) * m);      /* cov obs - neighbours */
    double* g1 = (double*)malloc(sizeof(double) * m);
    double* a = (double*)malloc(sizeof(double) * m);
    double* t1 = (double*)malloc(sizeof(double) * m);
    double* t2 = (double*)malloc(sizeof(double) * m);
#pragma omp for schedule(static) reduction(+ : bad)
    for (int i = 0; i < n; ++i) {
      const int32_t* nb = nn + (size_t)i * m;
      int q = 0;
      while (q < m && nb[q] >= 0) ++q;
      double Di = 1. + var; /* :1411 identity (nugget 1) + :1555-1563 marginal variance */
      for (int k = 0; k < m; ++k) A[(size_t)i * m + k] = 0.;
      if (calc_grad) {
        for (int k = 0; k < m; ++k) { Agrad[(size_t)i * m + k] = 0.; Agrad[((size_t)n + i) * m + k] = 0.; }
        Dgrad[i] = var; /* :1571 transf_scale: dD/dlog var starts at d_comp_j */
        Dgrad[(size_t)n + i] = 0.;
      }
      if (q > 0) {
        for (int j = 0; j < q; ++j) {
          double dist = sqrt(orc_sqdist(coords, n, d, nb[j], i));
          s1[j] = orc_cov(cov_type, dist, var, range);
          if (calc_grad) g1[j] = orc_cov_grad_range(cov_type, dist, var, range, s1[j]);
          for (int k = j; k < q; ++k) {
            if (k == j) { S[j * q + j] = var; if (calc_grad) { G[j * q + j] = 0.; G0[j * q + j] = var; } }
            else {
              double djk = sqrt(orc_sqdist(coords, n, d, nb[j], nb[k]));
              double c = orc_cov(cov_type, djk, var, range);
              S[j * q + k] = S[k * q + j] = c;
              if (calc_grad) {
                G0[j * q + k] = G0[k * q + j] = c;
                G[j * q + k] = G[k * q + j] = orc_cov_grad_range(cov_type, djk, var, range, c);
              }
            }
          }
        }
        for (int j = 0; j < q; ++j) S[j * q + j] += 1.; /* :1601 nugget on transformed scale */
        if (orc_chol(S, q) != 0) { ++bad; continue; }
        memcpy(a, s1, sizeof(double) * q);
        orc_chol_solve(S, q, a); /* :1618 A_i */
        double dot = 0.;
        for (int j = 0; j < q; ++j) { A[(size_t)i * m + j] = a[j]; dot += a[j] * s1[j]; }
        Di -= dot; /* :1623 */
        if (calc_grad) { /* :1636-1652, A_grad = Sigma^-1 dSigma_iN - Sigma^-1 dSigma_NN A_i */
          for (int p = 0; p < 2; ++p) {
            const double* GG = p == 0 ? G0 : G;
            const double* gg = p == 0 ? s1 : g1;
            for (int j = 0; j < q; ++j) {
              double t = 0.;
              for (int k = 0; k < q; ++k) t += GG[j * q + k] * a[k];
              t1[j] = gg[j]; t2[j] = t;
            }
            orc_chol_solve(S, q, t1);
            orc_chol_solve(S, q, t2);
            double d1 = 0., d2 = 0.;
            for (int j = 0; j < q; ++j) {
              double ag = t1[j] - t2[j];
              Agrad[((size_t)p * n + i) * m + j] = ag;
              d1 += ag * s1[j]; d2 += a[j] * gg[j];
            }
            if (p == 0) Dgrad[i] -= (d1 + d2);           /* :1646 */
            else Dgrad[(size_t)n + i] = -(d1 + d2);      /* :1650 */
          }
        }
      }
      if (!(Di > 0.)) ++bad;
      Dinv[i] = 1. / Di; /* :1682 */
    }
    free(S); free(G); free(G0); free(s1); free(g1); free(a); free(t1); free(t2);
  }
  return bad;
}

/* Vecchia factor for a NON-Gaussian likelihood (latent GP, no nugget): Vecchia_utils.cpp:1415-1417 (D starts at 0),
 * :1555-1563 (+ marginal variance), :1607-1609 (neighbour block diagonal *= JITTER_MULT_VECCHIA = 1 + 1e-10, utils.h:38),
 * :1618-1623, :1682. pars = {var, range} with var on the ORIGINAL scale (sigma2 = 1) and range transformed. */
int orc_vecchia_factor_latent(const double* coords, int n, int d, int m, const int32_t* nn, int cov_type,
                              const double* pars, double* A, double* Dinv) {
  const double var = pars[0], range = pars[1];
  int bad = 0;
#pragma omp parallel
  {
    double* S = (double*)malloc(sizeof(double) * m * m);
    double* s1 = (double*)malloc(sizeof(double) * m);
    double* a = (double*)malloc(sizeof(double) * m);
#pragma omp for schedule(static) reduction(+ : bad)
    for (int i = 0; i < n; ++i) {
      const int32_t* nb = nn + (size_t)i * m;
      int q = 0;
      while (q < m && nb[q] >= 0) ++q;
      double Di = var;
      for (int k = 0; k < m; ++k) A[(size_t)i * m + k] = 0.;
      if (q > 0) {
        for (int j = 0; j < q; ++j) {
          s1[j] = orc_cov(cov_type, sqrt(orc_sqdist(coords, n, d, nb[j], i)), var, range);
          S[j * q + j] = var * (1. + 1e-10);
          for (int k = j + 1; k < q; ++k)
            S[j * q + k] = S[k * q + j] = orc_cov(cov_type, sqrt(orc_sqdist(coords, n, d, nb[j], nb[k])), var, range);
        }
        if (orc_chol(S, q) != 0) { ++bad; continue; }
        memcpy(a, s1, sizeof(double) * q);
        orc_chol_solve(S, q, a);
        double dot = 0.;
        for (int j = 0; j < q; ++j) { A[(size_t)i * m + j] = a[j]; dot += a[j] * s1[j]; }
        Di -= dot;
      }
      if (!(Di > 0.)) ++bad;
      Dinv[i] = 1. / Di;
    }
    free(S); free(s1); free(a);
  }
  return bad;
}

/* u = B y  (B = I - A): (B y)_i = y_i - sum_k A[i,k] y[nn[i,k]] */
/* Latent factor with the derivative w.r.t. the log of the (transformed) range parameter: what CalcCovFactorGradientVecchia
 * (src/GPBoost/Vecchia_utils.cpp:1620-1652) leaves in B_grad[1] = -Agrad and D_grad[1] for a non-Gaussian likelihood with one GP.
 * (The marginal-variance derivative is not formed there — exclude_marg_var_grad — because dSigma^-1/dlog(var) = -Sigma^-1.)
 * The gradient block of the neighbours has a zero diagonal (cov_fcts.h:1179), so the jitter does not enter it. */
int orc_vecchia_factor_latent_grad(const double* coords, int n, int d, int m, const int32_t* nn, int cov_type,
                                   const double* pars, double* A, double* Dinv, double* Agrad, double* Dgrad) {
  const double var = pars[0], range = pars[1];
  int bad = 0;
#pragma omp parallel
  {
    double* S = (double*)malloc(sizeof(double) * m * m);
    double* G = (double*)malloc(sizeof(double) * m * m);
    double* s1 = (double*)malloc(sizeof(double) * m);
    double* g1 = (double*)malloc(sizeof(double) * m);
    double* a = (double*)malloc(sizeof(double) * m);
    double* t1 = (double*)malloc(sizeof(double) * m);
    double* t2 = (double*)malloc(sizeof(double) * m);
#pragma omp for schedule(static) reduction(+ : bad)
    for (int i = 0; i < n; ++i) {
      const int32_t* nb = nn + (size_t)i * m;
      int q = 0;
      while (q < m && nb[q] >= 0) ++q;
      double Di = var;
      for (int k = 0; k < m; ++k) { A[(size_t)i * m + k] = 0.; Agrad[(size_t)i * m + k] = 0.; }
      Dgrad[i] = 0.;
      if (q > 0) {
        for (int j = 0; j < q; ++j) {
          double dist = sqrt(orc_sqdist(coords, n, d, nb[j], i));
          s1[j] = orc_cov(cov_type, dist, var, range);
          g1[j] = orc_cov_grad_range(cov_type, dist, var, range, s1[j]);
          S[j * q + j] = var * (1. + 1e-10);
          G[j * q + j] = 0.;
          for (int k = j + 1; k < q; ++k) {
            double djk = sqrt(orc_sqdist(coords, n, d, nb[j], nb[k]));
            double c = orc_cov(cov_type, djk, var, range);
            S[j * q + k] = S[k * q + j] = c;
            G[j * q + k] = G[k * q + j] = orc_cov_grad_range(cov_type, djk, var, range, c);
          }
        }
        if (orc_chol(S, q) != 0) { ++bad; continue; }
        memcpy(a, s1, sizeof(double) * q);
        orc_chol_solve(S, q, a);
        double dot = 0.;
        for (int j = 0; j < q; ++j) { A[(size_t)i * m + j] = a[j]; dot += a[j] * s1[j]; }
        Di -= dot;
        for (int j = 0; j < q; ++j) {
          double t = 0.;
          for (int k = 0; k < q; ++k) t += G[j * q + k] * a[k];
          t1[j] = g1[j]; t2[j] = t;
        }
        orc_chol_solve(S, q, t1);
        orc_chol_solve(S, q, t2);
        double d1 = 0., d2 = 0.;
        for (int j = 0; j < q; ++j) {
          double ag = t1[j] - t2[j];
          Agrad[(size_t)i * m + j] = ag;
          d1 += ag * s1[j]; d2 += a[j] * g1[j];
        }
        Dgrad[i] = -(d1 + d2); /* :1650 */
      }
      if (!(Di > 0.)) ++bad;
      Dinv[i] = 1. / Di;
    }
    free(S); free(G); free(s1); free(g1); free(a); free(t1); free(t2);
  }
  return bad;
}

void orc_apply_B(int n, int m, const int32_t* nn, const double* A, const double* y, double* u) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double s = y[i];
    for (int k = 0; k < m; ++k) { int j = nn[(size_t)i * m + k]; if (j >= 0) s -= A[(size_t)i * m + k] * y[j]; }
    u[i] = s;
  }
}
/* w = B^T v */
void orc_apply_Bt(int n, int m, const int32_t* nn, const double* A, const double* v, double* w) {
  for (int i = 0; i < n; ++i) w[i] = v[i];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < m; ++k) { int j = nn[(size_t)i * m + k]; if (j >= 0) w[j] -= A[(size_t)i * m + k] * v[i]; }
}

/* ---- Gaussian negative log-likelihood with the Vecchia factor.
 * y^T Psi^-1 y = (By)^T D^-1 (By): re_model_template.h:9957-9964; log|Psi| = -sum log D^-1_ii :2947;
 * negll = yPy/(2 sigma2) + log|Psi|/2 + n/2 (log sigma2 + log 2pi) :3132.
 * out[0]=negll, out[1]=yTPsiInvy, out[2]=log_det_Psi */
void orc_vecchia_nll(int n, int m, const int32_t* nn, const double* A, const double* Dinv,
                     const double* y, double sigma2, double* out) {
  double* u = (double*)malloc(sizeof(double) * n);
  orc_apply_B(n, m, nn, A, y, u);
  double q = 0., ld = 0.;
  for (int i = 0; i < n; ++i) { q += u[i] * u[i] * Dinv[i]; ld -= log(Dinv[i]); }
  out[1] = q; out[2] = ld;
  out[0] = q / 2. / sigma2 + ld / 2. + n / 2. * (log(sigma2) + log(2 * M_PI));
  free(u);
}

/* y_aux = Psi^-1 y = B^T D^-1 B y : re_model_template.h:9772 */
void orc_vecchia_yaux(int n, int m, const int32_t* nn, const double* A, const double* Dinv,
                      const double* y, double* yaux) {
  double* u = (double*)malloc(sizeof(double) * n);
  orc_apply_B(n, m, nn, A, y, u);
  for (int i = 0; i < n; ++i) u[i] *= Dinv[i];
  orc_apply_Bt(n, m, nn, A, u, yaux);
  free(u);
}

/* ---- gradient of the negll w.r.t. log(transformed var), log(transformed range) with the error variance
 * profiled out / not included (include_error_var = false): re_model_template.h:1988-2010
 *   u = D^-1 B y ; u_k = dB_k y ; grad_k = (u_k.u - 0.5 u^T dD_k u)/sigma2 + 0.5 sum_i D^-1_ii dD_k,ii */
void orc_vecchia_grad(int n, int m, const int32_t* nn, const double* A, const double* Dinv,
                      const double* Agrad, const double* Dgrad, const double* y, double sigma2,
                      double* grad) {
  double* u = (double*)malloc(sizeof(double) * n);
  orc_apply_B(n, m, nn, A, y, u);
  for (int i = 0; i < n; ++i) u[i] *= Dinv[i];
  for (int p = 0; p < 2; ++p) {
    double uku = 0., udu = 0., tr = 0.;
    for (int i = 0; i < n; ++i) {
      double uk = 0.; /* (dB y)_i = - sum dA[i,k] y[nn] ; diagonal of dB is 0 (:1432) */
      for (int k = 0; k < m; ++k) { int j = nn[(size_t)i * m + k]; if (j >= 0) uk -= Agrad[((size_t)p * n + i) * m + k] * y[j]; }
      uku += uk * u[i];
      udu += u[i] * Dgrad[(size_t)p * n + i] * u[i];
      tr += Dinv[i] * Dgrad[(size_t)p * n + i];
    }
    grad[p] = (uku - 0.5 * udu) / sigma2 + 0.5 * tr;
  }
  free(u);
}

/* ---- exact (dense) Gaussian GP negll, config C1: Psi = I + Sigma (re_model_template.h:9273), Cholesky :6492,
 * log|Psi| = 2 sum log L_ii :3127, quadratic form via two triangular solves :9894, formula :3132.
 * coords column-major n x d; pars transformed {var, range}. out as in orc_vecchia_nll. Returns 0 on success. */
int orc_dense_nll(const double* coords, int n, int d, int cov_type, const double* pars, const double* y,
                  double sigma2, double* out) {
  double* P = (double*)malloc(sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; ++i) {
    P[(size_t)i * n + i] = pars[0] + 1.;
    for (int j = 0; j < i; ++j)
      P[(size_t)i * n + j] = P[(size_t)j * n + i] =
          orc_cov(cov_type, sqrt(orc_sqdist(coords, n, d, i, j)), pars[0], pars[1]);
  }
  if (orc_chol(P, n) != 0) { free(P); return -1; }
  double* z = (double*)malloc(sizeof(double) * n);
  double ld = 0., q = 0.;
  for (int i = 0; i < n; ++i) {
    double t = y[i];
    for (int k = 0; k < i; ++k) t -= P[(size_t)i * n + k] * z[k];
    z[i] = t / P[(size_t)i * n + i];
    q += z[i] * z[i];
    ld += 2. * log(P[(size_t)i * n + i]);
  }
  out[1] = q; out[2] = ld;
  out[0] = q / 2. / sigma2 + ld / 2. + n / 2. * (log(sigma2) + log(2 * M_PI));
  free(P); free(z);
  return 0;
}
