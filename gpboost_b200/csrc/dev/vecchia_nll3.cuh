// Likelihood / factor-store pass of the Vecchia factor, two observations per warp (the layout of vecchia_nll2.cuh), with the pair
// covariances of the NEXT pair of observations computed inside the factorisation of the current one.
//
// Why: in vecchia_nll2_kernel every one of the 31 elimination steps costs about the latency of its pivot chain (DFMA -> shuffle ->
// MUFU.RSQ64H -> four dependent FP64 levels -> DMUL) whatever its share of rank-1 updates is (profiles/r02_ncu_nll2_steps.txt: ~3300
// stall samples per step for 8..51 FP64 instructions): the steps behind column 15 leave the FP64 pipe idle, and three resident warps
// per scheduler cannot fill it. The 30 covariance rounds of an observation are the opposite: independent 22-instruction FP64 chains.
// Here two rounds of the next pair ride in each of the steps 15..29 of the current pair:
//   * the matrix of the pair being factorised lives in the lower triangle of the shared 32 x 33 buffer (column c, rows >= c — the
//     columns of L replace it in place), the covariances of the next pair go to the strict UPPER triangle of the same buffer (pair
//     (a, b), a < b, at column b / row a), which the factorisation never touches once its column stores skip the rows above the diagonal;
//   * at the top of an iteration a lane reads its two matrix rows straight out of the upper triangle (row r is contiguous there),
//     after the diagonal and the response row have been put in place;
//   * the gather runs three pairs ahead: neighbour indices of pair it+3 and responses / coordinates of pair it+2 are requested in
//     step 15 of pair it, the points of pair it+1 are in shared memory from the top of the iteration.
// Same arithmetic per entry as vecchia_nll2_kernel (same covariance function, same elimination order): the sums agree with it to
// the last bit for equal grids. Modes: MODE_NLL, MODE_STORE (the gradient pass keeps its range-derivative values in registers across
// the factorisation and stays in vecchia_nll2_kernel).
// Reference: CalcCovFactorGradientVecchia, src/GPBoost/Vecchia_utils.cpp:1461-1684; re_model_template.h:9957-9964, :2947.
#pragma once
#include "vecchia_nll2.cuh"

namespace gpb {

template <int COV, int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, GPB_NLL2_BLOCKS) vecchia_nll3_kernel(const FactorArgs p) {
  static_assert(MODE == MODE_NLL || MODE == MODE_STORE, "modes: NLL, STORE");
  constexpr bool SOLVE = MODE == MODE_STORE;
  constexpr int MT = 30, P = 31, NR = 30;
  constexpr int K0 = 15;  // first elimination step that carries rounds of the next pair (two per step: 2 * (MT - K0) = NR)
  static_assert(2 * (MT - K0) == NR, "rounds per step");
  extern __shared__ __align__(16) double smem_raw[];
  const int lane = threadIdx.x & 31, hl = lane & 15, hh = lane >> 4, wib = threadIdx.x >> 5;
  const int hbase = lane & 16;
  double* S = smem_raw + (size_t)wib * (2 * (kNll2Half + 64)) + (size_t)hh * kNll2Half;
  double* pts = smem_raw + (size_t)wib * (2 * (kNll2Half + 64)) + 2 * kNll2Half + (size_t)hh * 64;
  const int64_t gwarp = (int64_t)blockIdx.x * kWarpsPerBlock + wib, nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int m = p.m;
  const double var = p.var;
  const double sc = COV == COV_GAUSSIAN ? sqrt(p.range) : p.range;
  const double logvar = log(var);
  const double tab = kExp2Tab32[lane];
  double acc0 = 0., acc1 = 0., acc2 = 0.;

  auto slot_src = [&](int64_t ii, int s) -> int {
    const int qq = ii < m ? (int)ii : m;
    if (s < qq) return p.nn[ii * m + s];
    return s == MT ? (int)ii : -1;
  };
  auto row_of = [&](int64_t itx) -> int64_t { return p.row_begin + 2 * itx + hh; };
  auto fetch_idx = [&](int64_t ii, bool act, int& a_lo, int& a_hi) {
    a_lo = -1; a_hi = -1;
    if (act) {
      a_lo = slot_src(ii, hl);
      a_hi = hl + 16 <= MT ? slot_src(ii, hl + 16) : -1;
    }
  };
  auto fetch_data = [&](int a_lo, int a_hi, double2& k_lo, double2& k_hi, double& v_lo, double& v_hi) {
    v_lo = 0.; v_hi = 0.;
    if (a_lo >= 0) { v_lo = p.y[a_lo]; k_lo = *reinterpret_cast<const double2*>(p.coords + (int64_t)a_lo * 2); }
    if (a_hi >= 0) { v_hi = p.y[a_hi]; k_hi = *reinterpret_cast<const double2*>(p.coords + (int64_t)a_hi * 2); }
  };
  // covariance of round r (offset t = r / 2 + 1, own point pi = hl + 16 (r & 1)) for the pair whose scaled points are in pts
  auto pair_value = [&](int r, const double2& mlo, const double2& mhi) -> double {
    const int t = (r >> 1) + 1;
    const bool odd = (r & 1) != 0;
    const int pi = hl + (odd ? 16 : 0);
    int o = pi + t;
    if (o >= P) o -= P;
    if (pi >= P) o = 0;
    const double2 po = *reinterpret_cast<const double2*>(pts + o * 2);
    const double2 me = odd ? mhi : mlo;
    const double dx = me.x - po.x, dy = me.y - po.y;
    const double d2s = fma(dy, dy, fma(dx, dx, 1e-300));
    double g = 0.;
    return cov_eval_scaled<COV, false>(d2s, var, logvar, tab, g);
  };
  // ... and its place in the upper triangle: pair (a, b), a < b -> column b, row a; zero for pairs with a dummy slot; the one padded
  // pair (lane 15, odd rounds) goes to the padding row 32
  auto pair_store = [&](int r, double v, bool rlo, bool rhi, unsigned rmask, bool fullw) {
    const int t = (r >> 1) + 1;
    const bool odd = (r & 1) != 0;
    const int pi = hl + (odd ? 16 : 0);
    const bool valid = pi < P;
    int o = pi + t;
    if (o >= P) o -= P;
    if (!valid) o = 0;
    const bool keep = ((fullw ? 1u : 0u) | ((odd ? (unsigned)rhi : (unsigned)rlo) & (rmask >> o) & 1u)) != 0u;
    S[valid ? max(pi, o) * kLd + min(pi, o) : hl * kLd + 32] = keep ? v : 0.;
  };

  // ---- prologue: pair 0 completely (gather, points, rounds), data of pair 1, indices of pair 2
  int64_t it = gwarp;
  int64_t i = row_of(it);
  bool active = i < p.row_end;
  bool real_lo, real_hi;
  double yl, yh;
  {
    int s_lo, s_hi;
    double2 c_lo = make_double2(0., 0.), c_hi = c_lo;
    fetch_idx(i, active, s_lo, s_hi);
    fetch_data(s_lo, s_hi, c_lo, c_hi, yl, yh);
    real_lo = s_lo >= 0; real_hi = s_hi >= 0;
    const double2 my_lo = make_double2(c_lo.x * sc, c_lo.y * sc), my_hi = make_double2(c_hi.x * sc, c_hi.y * sc);
    if (real_lo) *reinterpret_cast<double2*>(pts + hl * 2) = my_lo;
    if (real_hi) *reinterpret_cast<double2*>(pts + (hl + 16) * 2) = my_hi;
    const unsigned blo = __ballot_sync(0xffffffffu, real_lo), bhi = __ballot_sync(0xffffffffu, real_hi);
    const unsigned rmask = ((blo >> hbase) & 0xffffu) | (((bhi >> hbase) & 0xffffu) << 16);
    const int q = i < m ? (int)i : m;
    const bool fullw = (__all_sync(0xffffffffu, (q == MT) || !active)) != 0;
    __syncwarp();
    constexpr int G = 6;
#pragma unroll
    for (int r0 = 0; r0 < NR; r0 += G) {
      double val[G];
#pragma unroll
      for (int j = 0; j < G; ++j) val[j] = pair_value(r0 + j, my_lo, my_hi);
#pragma unroll
      for (int j = 0; j < G; ++j) pair_store(r0 + j, val[j], real_lo, real_hi, rmask, fullw);
    }
  }
  int64_t i1 = row_of(it + nwarps);
  bool active1 = i1 < p.row_end;
  int s1_lo, s1_hi;
  double2 c1_lo = make_double2(0., 0.), c1_hi = c1_lo;
  double y1_lo, y1_hi;
  fetch_idx(i1, active1, s1_lo, s1_hi);
  fetch_data(s1_lo, s1_hi, c1_lo, c1_hi, y1_lo, y1_hi);
  bool real1_lo = s1_lo >= 0, real1_hi = s1_hi >= 0;
  int64_t i2 = row_of(it + 2 * nwarps);
  bool active2 = i2 < p.row_end;
  int s2_lo, s2_hi;
  fetch_idx(i2, active2, s2_lo, s2_hi);
  __syncwarp();  // the rounds above are done with pts; the upper triangle is complete

  for (; __any_sync(0xffffffffu, active); ) {
    const int q = i < m ? (int)i : m;
    const bool was_active = active;
    // ---- this pair: diagonal (variance + nugget, 1 for dummy slots: Vecchia_utils.cpp:1601 / :1411, :1563) and response row go next
    // to the covariances that the previous iteration left in the upper triangle
    const double dlo = real_lo ? p.diag_nb : 1.;
    const double dhi = real_hi ? (hl + 16 == MT ? p.diag_obs : p.diag_nb) : 1.;
    S[hl * kLd + hl] = dlo;
    S[31 * kLd + hl] = yl;
    if (hl + 16 <= MT) {
      S[(hl + 16) * kLd + hl + 16] = dhi;
      S[31 * kLd + hl + 16] = yh;
    }
    // ---- next pair: scaled points into shared memory, dummy-slot masks
    const double2 nx_lo = make_double2(c1_lo.x * sc, c1_lo.y * sc), nx_hi = make_double2(c1_hi.x * sc, c1_hi.y * sc);
    if (real1_lo) *reinterpret_cast<double2*>(pts + hl * 2) = nx_lo;
    if (real1_hi) *reinterpret_cast<double2*>(pts + (hl + 16) * 2) = nx_hi;
    const unsigned blo = __ballot_sync(0xffffffffu, real1_lo), bhi = __ballot_sync(0xffffffffu, real1_hi);
    const unsigned rmask1 = ((blo >> hbase) & 0xffffu) | (((bhi >> hbase) & 0xffffu) << 16);
    const int q1 = i1 < m ? (int)i1 : m;
    const bool full1 = (__all_sync(0xffffffffu, (q1 == MT) || !active1)) != 0;
    __syncwarp();

    // ---- my two rows of the lower triangle -> registers: row r of the matrix is row r of the upper-triangle store (entries right of
    // the diagonal: don't-care values); lane 15's second row is the response row
    double lo[16], hi[MT + 1];
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = S[hl * kLd + c];
#pragma unroll
    for (int c = 0; c <= MT; ++c) hi[c] = S[(hl + 16) * kLd + c];
    __syncwarp();

    // ---- right-looking Cholesky with look-ahead, pivots 0..MT (see vecchia_nll2_kernel for the pivot chain)
    double Di, lk_lo, lk_hi;
    double dg_lo = dlo, dg_hi = dhi;
    {
      const double d0 = __shfl_sync(0xffffffffu, lo[0], hbase);
      const double r0 = rsqrt_fast(d0);
      lk_lo = lo[0] * r0; lk_hi = hi[0] * r0;
      S[hl] = lk_lo; S[hl + 16] = lk_hi;
      Di = d0;
    }
    __syncwarp();
    // gather registers of the pair after the next one (requested in step K0)
    double2 c2_lo = make_double2(0., 0.), c2_hi = c2_lo;
    double y2_lo = 0., y2_hi = 0.;
    int s3_lo = -1, s3_hi = -1;
    int64_t i3 = 0;
    bool active3 = false;
#pragma unroll
    for (int k = 0; k < MT; ++k) {
      dg_lo = fma(-lk_lo, lk_lo, dg_lo);
      dg_hi = fma(-lk_hi, lk_hi, dg_hi);
      const double dn = (k + 1 < 16) ? __shfl_sync(0xffffffffu, dg_lo, hbase + k + 1) : __shfl_sync(0xffffffffu, dg_hi, hbase + k + 1 - 16);
      if (k == K0) {  // responses / coordinates of pair it+2 (its indices arrived an iteration ago), indices of pair it+3
        fetch_data(s2_lo, s2_hi, c2_lo, c2_hi, y2_lo, y2_hi);
        i3 = row_of(it + 3 * nwarps);
        active3 = i3 < p.row_end;
        fetch_idx(i3, active3, s3_lo, s3_hi);
      }
      // two covariance rounds of the next pair ride in this step
      double v0 = 0., v1 = 0.;
      if (k >= K0) {
        v0 = pair_value(2 * (k - K0), nx_lo, nx_hi);
        v1 = pair_value(2 * (k - K0) + 1, nx_lo, nx_hi);
      }
      {
        const double m1 = S[k * kLd + k + 1];
        hi[k + 1] -= lk_hi * m1;
        if (k + 1 < 16) lo[k + 1] -= lk_lo * m1;
      }
      if (k + 1 == MT) Di = dn;
      const double rn = rsqrt_fast(dn);
      const double lk1_lo = (k + 1 < 16) ? lo[k + 1] * rn : 0.;
      const double lk1_hi = hi[k + 1] * rn;
#pragma unroll
      for (int c = k + 2; c <= MT; c += 2) {
        const double2 l2 = *reinterpret_cast<const double2*>(&S[k * kLd + c]);
        hi[c] -= lk_hi * l2.x;
        if (c + 1 <= MT) hi[c + 1] -= lk_hi * l2.y;
        if (c < 16) lo[c] -= lk_lo * l2.x;
        if (c + 1 < 16) lo[c + 1] -= lk_lo * l2.y;
      }
      // stores last (vecchia_nll2_kernel); rows above the diagonal are skipped: those slots belong to the next pair
      if (k + 1 < 16 && hl >= k + 1) S[(k + 1) * kLd + hl] = lk1_lo;
      if (k + 1 <= 16 || hl + 16 >= k + 1) S[(k + 1) * kLd + hl + 16] = lk1_hi;
      if (k >= K0) {
        pair_store(2 * (k - K0), v0, real1_lo, real1_hi, rmask1, full1);
        pair_store(2 * (k - K0) + 1, v1, real1_lo, real1_hi, rmask1, full1);
      }
      __syncwarp();
      lk_lo = lk1_lo; lk_hi = lk1_hi;
    }
    // lane hl = 15 (row 31) wrote L[31][30] = (By)_i / sqrt(D_i) into column 30
    const double r_over_sd = S[MT * kLd + (MT + 1)];
    if (hl == 0 && was_active) {
      acc0 += r_over_sd * r_over_sd;
      acc1 += log(Di);
      acc2 += !(Di > 0.) ? 1. : 0.;
    }
    if (SOLVE) {
      // ---- back substitution L_NN^T x = L[30][.] (-> A_i): lane hl owns unknowns hl, hl + 16 (as vecchia_nll2_kernel<MODE_STORE>)
      const bool has_hi = hl + 16 < MT;
      double xa_lo = S[hl * kLd + MT];
      double xa_hi = has_hi ? S[(hl + 16) * kLd + MT] : 0.;
      const double dinv_lo = 1. / S[hl * kLd + hl];
      const double dinv_hi = has_hi ? 1. / S[(hl + 16) * kLd + hl + 16] : 0.;
#pragma unroll
      for (int r = MT - 1; r >= 0; --r) {
        double fa;
        if (r < 16) {
          fa = __shfl_sync(0xffffffffu, xa_lo * dinv_lo, hbase + r);
          if (hl == r) xa_lo = fa;
        } else {
          fa = __shfl_sync(0xffffffffu, xa_hi * dinv_hi, hbase + r - 16);
          if (hl + 16 == r) xa_hi = fa;
        }
        const double l_lo = hl < r ? S[hl * kLd + r] : 0.;  // L[r][hl]
        xa_lo -= l_lo * fa;
        if (r > 16) {
          const double l_hi = (has_hi && hl + 16 < r) ? S[(hl + 16) * kLd + r] : 0.;  // L[r][hl+16]
          xa_hi -= l_hi * fa;
        }
      }
      if (was_active) {
        if (hl < m) p.A[i * m + hl] = hl < q ? xa_lo : 0.;
        if (hl + 16 < m) p.A[i * m + hl + 16] = hl + 16 < q ? xa_hi : 0.;
        if (hl == 0) { const double Dinv_i = 1. / Di; p.Dinv[i] = Dinv_i; p.w[i] = r_over_sd * sqrt(Di) * Dinv_i; }
      }
    }
    __syncwarp();
    // rotate the pipeline
    it += nwarps;
    i = i1; active = active1; real_lo = real1_lo; real_hi = real1_hi; yl = y1_lo; yh = y1_hi;
    i1 = i2; active1 = active2; c1_lo = c2_lo; c1_hi = c2_hi; y1_lo = y2_lo; y1_hi = y2_hi; real1_lo = s2_lo >= 0; real1_hi = s2_hi >= 0;
    i2 = i3; active2 = active3; s2_lo = s3_lo; s2_hi = s3_hi;
  }
  if (hl == 0) {
    double* out = p.partials + (size_t)(gwarp * 2 + hh) * kNumAcc;
    out[0] = acc0; out[1] = acc1; out[2] = acc2;
#pragma unroll
    for (int k = 3; k < kNumAcc; ++k) out[k] = 0.;
  }
}

}  // namespace gpb
