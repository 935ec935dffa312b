// TEST INFRASTRUCTURE ONLY (oracle). The reference's "random" Vecchia ordering is
// std::shuffle(indices 0..n-1, std::mt19937(seed)) — re_model_template.h:159-161 (rng_ = RNG_t(seed)),
// Vecchia_utils.cpp:1129-1131. The permutation is defined by the C++ standard library's shuffle, so the
// restatement is the same library call (libstdc++ on this image).
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>
extern "C" void orc_vecchia_random_order(int n, int seed, int32_t* perm) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::mt19937 rng(seed);
  std::shuffle(idx.begin(), idx.end(), rng);
  for (int i = 0; i < n; ++i) perm[i] = idx[i];
}

// Indices sorted by value exactly as the reference's SortIndeces does (include/GPBoost/utils.h:230-238):
// std::sort on an iota'd std::vector<int> with the comparator v[i1] < v[i2]. Ties are left in whatever order
// libstdc++'s introsort produces, which the neighbour search's visiting order (and hence its tie-breaks on
// lattice data) inherits — so the restatement makes the same library call on the same input.
extern "C" void orc_sort_indices(const double* v, int n, int32_t* idx_out) {
  std::vector<double> vv(v, v + n);
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&vv](int i1, int i2) { return vv[i1] < vv[i2]; });
  for (int i = 0; i < n; ++i) idx_out[i] = idx[i];
}
