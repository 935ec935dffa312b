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

// Probe vectors of the stochastic Lanczos quadrature, exactly as GenRandVecNormalParallel draws them
// (src/GPBoost/CG_utils.cpp:978-994): column col_i comes from mt19937 seeded with
// seed_seq{base_seed, run_id lo, run_id hi, col_i} through std::normal_distribution<double>. Output column-major n x t.
extern "C" void orc_gen_rand_normal(int base_seed, unsigned long long run_id, int n, int t, double* out) {
  const uint32_t b32 = static_cast<uint32_t>(base_seed);
  for (int col = 0; col < t; ++col) {
    std::normal_distribution<double> ndist(0.0, 1.0);
    std::seed_seq seq{b32, static_cast<uint32_t>(run_id), static_cast<uint32_t>(run_id >> 32), static_cast<uint32_t>(col)};
    std::mt19937 gen(seq);
    for (int row = 0; row < n; ++row) out[(size_t)col * n + row] = ndist(gen);
  }
}
