/*
 * TEST INFRASTRUCTURE ONLY (oracle). Plain-C restatement of the reference's serial histogram tree learner for
 * dense uint8 bins, numerical features without missing values, constant hessian (L2 / GP objective):
 *   SerialTreeLearner::Train            src/LightGBM/treelearner/serial_tree_learner.cpp:159-209
 *   BeforeTrain / BeforeFindBestSplit   :253-322   (root sums, depth and 2*min_data_in_leaf checks, smaller/larger leaf)
 *   ConstructHistograms                 :351-373 -> Dataset::ConstructHistogramsInner io/dataset.cpp:1143-1245
 *                                        -> DenseBin::ConstructHistogramInner io/dense_bin.hpp:98-141 (col-wise: one
 *                                        sequential pass over the leaf's rows per feature; count -> hessian :1223-1226)
 *   FindBestSplitsFromHistograms        :375-455  (parent - smaller subtraction, feature_histogram.hpp:79-83)
 *   FeatureHistogram::FindBestThreshold feature_histogram.hpp:85-95, BeforeNumercal :97-113,
 *   FindBestThresholdSequentially<REVERSE> :858-960, result :1057-1083, GetLeafGain :826-835,
 *   CalculateSplittedLeafOutput :743-764, SplitInfo::operator> split_info.hpp:126-153
 *   SplitInner / DataPartition::Split   serial_tree_learner.cpp:565-681, data_partition.hpp:101-120 (stable partition)
 *   Tree::Split                         include/LightGBM/tree.h:533-575
 * Pinned by tests/test_tree_oracle_pinned.py against trees grown by the unmodified reference library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_KEPS ((double)1e-15f)   /* kEpsilon, include/LightGBM/meta.h:54 (float literal widened to score_t=double) */

typedef struct {
  int num_leaves;
  int min_data_in_leaf;
  double min_sum_hessian_in_leaf;
  double lambda_l2;
  double min_gain_to_split;
  int max_depth;
} orc_tree_config;

typedef struct {
  double gain;
  int feature;
  int threshold;
  int left_count, right_count;
  double left_output, right_output;
  double left_sum_gradient, left_sum_hessian, right_sum_gradient, right_sum_hessian;
} orc_split;

static double orc_leaf_gain(double g, double h, double l2) { return (g * g) / (h + l2); }
static double orc_leaf_output(double g, double h, double l2) { return -g / (h + l2); }

/* SplitInfo::operator> : larger gain wins; on equal gain the smaller feature index wins (feature -1 = int max) */
static int orc_split_better(const orc_split* a, const orc_split* b) {
  if (a->gain != b->gain) return a->gain > b->gain;
  int fa = a->feature == -1 ? 2147483647 : a->feature, fb = b->feature == -1 ? 2147483647 : b->feature;
  return fa < fb;
}

/* hist: num_bin x 2 doubles (grad, hess) */
static void orc_find_best_threshold(const double* hist, int num_bin, double sum_gradient, double sum_hessian_in, int num_data,
                                    const orc_tree_config* cfg, orc_split* out, int* is_splittable) {
  out->gain = -INFINITY;
  out->feature = -1;
  const double sum_hessian = sum_hessian_in + 2 * ORC_KEPS;                                   /* :93 */
  const double min_gain_shift = orc_leaf_gain(sum_gradient, sum_hessian, cfg->lambda_l2) + cfg->min_gain_to_split; /* :102-112 */
  double best_sum_left_gradient = NAN, best_sum_left_hessian = NAN, best_gain = -INFINITY;
  int best_left_count = 0, best_threshold = num_bin;
  const double cnt_factor = num_data / sum_hessian;
  double sum_right_gradient = 0., sum_right_hessian = ORC_KEPS;
  int right_count = 0, splittable = 0;
  for (int t = num_bin - 1; t >= 1; --t) {                                                     /* :885-887, absolute bins */
    const double grad = hist[2 * t], hess = hist[2 * t + 1];
    const int cnt = (int)(hess * cnt_factor + 0.5f);                                          /* RoundInt, common.h:920 */
    sum_right_gradient += grad;
    sum_right_hessian += hess;
    right_count += cnt;
    if (right_count < cfg->min_data_in_leaf || sum_right_hessian < cfg->min_sum_hessian_in_leaf) continue;
    const int left_count = num_data - right_count;
    if (left_count < cfg->min_data_in_leaf) break;
    const double sum_left_hessian = sum_hessian - sum_right_hessian;
    if (sum_left_hessian < cfg->min_sum_hessian_in_leaf) break;
    const double sum_left_gradient = sum_gradient - sum_right_gradient;
    const double current_gain = orc_leaf_gain(sum_left_gradient, sum_left_hessian, cfg->lambda_l2) +
                                orc_leaf_gain(sum_right_gradient, sum_right_hessian, cfg->lambda_l2);
    if (current_gain <= min_gain_shift) continue;
    splittable = 1;
    if (current_gain > best_gain) {
      best_left_count = left_count;
      best_sum_left_gradient = sum_left_gradient;
      best_sum_left_hessian = sum_left_hessian;
      best_threshold = t - 1;
      best_gain = current_gain;
    }
  }
  *is_splittable = splittable;                                                                 /* is_splittable_, :100,:949 */
  if (splittable && best_gain > out->gain + min_gain_shift) {                                  /* :1057-1083 */
    out->threshold = best_threshold;
    out->left_output = orc_leaf_output(best_sum_left_gradient, best_sum_left_hessian, cfg->lambda_l2);
    out->left_count = best_left_count;
    out->left_sum_gradient = best_sum_left_gradient;
    out->left_sum_hessian = best_sum_left_hessian - ORC_KEPS;
    out->right_output = orc_leaf_output(sum_gradient - best_sum_left_gradient, sum_hessian - best_sum_left_hessian, cfg->lambda_l2);
    out->right_count = num_data - best_left_count;
    out->right_sum_gradient = sum_gradient - best_sum_left_gradient;
    out->right_sum_hessian = sum_hessian - best_sum_left_hessian - ORC_KEPS;
    out->gain = best_gain - min_gain_shift;
  }
}

/* one sequential pass per feature over the leaf's rows (col-wise DenseBin path) */
static void orc_construct_hist(const uint8_t* bins, int n, int F, const int* num_bin, const int32_t* idx, int cnt, const double* grad,
                               double hess_const, double* hist /* F x 256 x 2 */) {
  memset(hist, 0, sizeof(double) * (size_t)F * 512);
  for (int f = 0; f < F; ++f) {
    double* h = hist + (size_t)f * 512;
    const uint8_t* col = bins + (size_t)f * n; /* feature-major */
    for (int j = 0; j < cnt; ++j) {
      const int r = idx[j];
      const int b = col[r];
      h[2 * b] += grad[r];
      h[2 * b + 1] += 1.0; /* count; scaled below */
    }
    for (int b = 0; b < num_bin[f]; ++b) h[2 * b + 1] *= hess_const; /* dataset.cpp:1223-1226 */
  }
}

/*
 * Grow one tree. bins: feature-major F x n uint8. Outputs (arrays sized num_leaves): per internal node
 * split_feature, threshold_bin, left_child, right_child, split_gain(float), internal_count; per leaf leaf_value, leaf_count;
 * leaf_of_row (n). Returns the number of leaves.
 */
int orc_tree_train(const uint8_t* bins, int n, int F, const int* num_bin, const double* grad, double hess_const,
                   const orc_tree_config* cfg, int* split_feature, int* threshold_bin, int* left_child, int* right_child,
                   float* split_gain, double* leaf_value, int* leaf_count, int32_t* leaf_of_row) {
  const int L = cfg->num_leaves;
  int32_t* indices = (int32_t*)malloc(sizeof(int32_t) * n);
  int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * n);
  int* leaf_begin = (int*)calloc(L, sizeof(int));
  int* leaf_cnt = (int*)calloc(L, sizeof(int));
  int* leaf_depth = (int*)calloc(L, sizeof(int));
  int* leaf_parent = (int*)malloc(sizeof(int) * L);
  double* leaf_sum_g = (double*)calloc(L, sizeof(double));
  double* leaf_sum_h = (double*)calloc(L, sizeof(double));
  orc_split* best = (orc_split*)malloc(sizeof(orc_split) * L);
  double** hist_of = (double**)calloc(L, sizeof(double*));
  /* per leaf & feature: FeatureHistogram::is_splittable_ (a feature that produced no admissible threshold in the parent is
   * not examined in the children: serial_tree_learner.cpp:329-336) */
  char* splittable = (char*)malloc((size_t)L * F);
  memset(splittable, 1, (size_t)L * F);
  for (int i = 0; i < n; ++i) indices[i] = i;
  for (int l = 0; l < L; ++l) { best[l].gain = -INFINITY; best[l].feature = -1; leaf_parent[l] = -1; }
  leaf_cnt[0] = n;
  /* root sums: LeafSplits::Init leaf_splits.hpp:70-83 */
  double sg = 0.;
  for (int i = 0; i < n; ++i) sg += grad[i];
  leaf_sum_g[0] = sg;
  leaf_sum_h[0] = hess_const * n;
  int num_leaves = 1;
  int left_leaf = 0, right_leaf = -1;
  leaf_value[0] = 0.;
  leaf_count[0] = n;
  for (int split = 0; split < L - 1; ++split) {
    /* BeforeFindBestSplit */
    int do_find = 1;
    if (cfg->max_depth > 0 && leaf_depth[left_leaf] >= cfg->max_depth) {
      best[left_leaf].gain = -INFINITY;
      if (right_leaf >= 0) best[right_leaf].gain = -INFINITY;
      do_find = 0;
    }
    if (do_find) {
      const int nl = leaf_cnt[left_leaf], nr = right_leaf >= 0 ? leaf_cnt[right_leaf] : 0;
      if (nr < cfg->min_data_in_leaf * 2 && nl < cfg->min_data_in_leaf * 2) {
        best[left_leaf].gain = -INFINITY;
        if (right_leaf >= 0) best[right_leaf].gain = -INFINITY;
        do_find = 0;
      }
    }
    if (do_find) {
      int smaller, larger;
      double* parent_hist = NULL;
      if (right_leaf < 0) { smaller = left_leaf; larger = -1; }
      else if (leaf_cnt[left_leaf] < leaf_cnt[right_leaf]) { smaller = left_leaf; larger = right_leaf; parent_hist = hist_of[left_leaf]; hist_of[left_leaf] = NULL; }
      else { smaller = right_leaf; larger = left_leaf; parent_hist = hist_of[left_leaf]; hist_of[left_leaf] = NULL; }
      /* children inherit the parent's flags (the larger leaf re-uses the parent's histogram objects, the smaller one is
       * marked unsplittable where the parent is) */
      if (larger >= 0) {
        const char* pf = splittable + (size_t)left_leaf * F; /* parent flags live under the left (= parent) leaf id */
        char keep[4096];
        memcpy(keep, pf, F);
        memcpy(splittable + (size_t)smaller * F, keep, F);
        memcpy(splittable + (size_t)larger * F, keep, F);
      }
      double* hs = (double*)malloc(sizeof(double) * (size_t)F * 512);
      orc_construct_hist(bins, n, F, num_bin, indices + leaf_begin[smaller], leaf_cnt[smaller], grad, hess_const, hs);
      hist_of[smaller] = hs;
      if (larger >= 0) { /* parent - smaller, in place in the parent's buffer */
        for (size_t t = 0; t < (size_t)F * 512; ++t) parent_hist[t] -= hs[t];
        hist_of[larger] = parent_hist;
      }
      for (int pass = 0; pass < 2; ++pass) {
        const int leaf = pass == 0 ? smaller : larger;
        if (leaf < 0) continue;
        orc_split bl; bl.gain = -INFINITY; bl.feature = -1;
        for (int f = 0; f < F; ++f) {
          if (!splittable[(size_t)leaf * F + f]) continue;
          orc_split s;
          int sp = 0;
          orc_find_best_threshold(hist_of[leaf] + (size_t)f * 512, num_bin[f], leaf_sum_g[leaf], leaf_sum_h[leaf], leaf_cnt[leaf], cfg, &s, &sp);
          splittable[(size_t)leaf * F + f] = (char)sp;
          s.feature = f;
          if (orc_split_better(&s, &bl)) bl = s;
        }
        best[leaf] = bl;
      }
    }
    /* ArgMax over leaves (first maximum) */
    int best_leaf = 0;
    for (int l = 1; l < num_leaves; ++l) if (orc_split_better(&best[l], &best[best_leaf])) best_leaf = l;
    /* ArrayArgs::ArgMax uses operator> on SplitInfo: equal gain -> smaller feature; then lower index stays */
    const orc_split bs = best[best_leaf];
    if (!(bs.gain > 0.0)) break;
    /* partition (stable): left keeps the leaf id */
    const int b = leaf_begin[best_leaf], c = leaf_cnt[best_leaf];
    const uint8_t* col = bins + (size_t)bs.feature * n;
    int nleft = 0, nright = 0;
    for (int j = 0; j < c; ++j) { const int r = indices[b + j]; if (col[r] <= bs.threshold) indices[b + nleft++] = r; else tmp[nright++] = r; }
    memcpy(indices + b + nleft, tmp, sizeof(int32_t) * nright);
    const int new_leaf = num_leaves;
    leaf_cnt[best_leaf] = nleft; leaf_begin[new_leaf] = b + nleft; leaf_cnt[new_leaf] = nright;
    /* Tree::Split */
    const int node = num_leaves - 1;
    const int parent = leaf_parent[best_leaf];
    if (parent >= 0) { if (left_child[parent] == ~best_leaf) left_child[parent] = node; else right_child[parent] = node; }
    split_feature[node] = bs.feature; threshold_bin[node] = bs.threshold;
    split_gain[node] = (float)(bs.gain + cfg->min_gain_to_split);
    left_child[node] = ~best_leaf; right_child[node] = ~new_leaf;
    leaf_parent[best_leaf] = node; leaf_parent[new_leaf] = node;
    leaf_value[best_leaf] = isnan(bs.left_output) ? 0. : bs.left_output; leaf_count[best_leaf] = nleft;
    leaf_value[new_leaf] = isnan(bs.right_output) ? 0. : bs.right_output; leaf_count[new_leaf] = nright;
    leaf_depth[new_leaf] = leaf_depth[best_leaf] + 1; leaf_depth[best_leaf]++;
    leaf_sum_g[best_leaf] = bs.left_sum_gradient; leaf_sum_h[best_leaf] = bs.left_sum_hessian;
    leaf_sum_g[new_leaf] = bs.right_sum_gradient; leaf_sum_h[new_leaf] = bs.right_sum_hessian;
    best[best_leaf].gain = -INFINITY; best[best_leaf].feature = -1; /* recomputed next round */
    best[new_leaf].gain = -INFINITY; best[new_leaf].feature = -1;
    ++num_leaves;
    left_leaf = best_leaf; right_leaf = new_leaf;
  }
  for (int l = 0; l < num_leaves; ++l)
    for (int j = 0; j < leaf_cnt[l]; ++j) leaf_of_row[indices[leaf_begin[l] + j]] = l;
  for (int l = 0; l < L; ++l) free(hist_of[l]);
  free(indices); free(tmp); free(leaf_begin); free(leaf_cnt); free(leaf_depth); free(leaf_parent); free(leaf_sum_g); free(leaf_sum_h);
  free(best); free(hist_of); free(splittable);
  return num_leaves;
}
