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
