#include "dataset.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <set>
#include <sstream>
#include <stdexcept>

namespace gpb200 {

namespace {
constexpr double kZeroThreshold = 1e-35f;  // include/LightGBM/meta.h:56
[[noreturn]] void Fatal(const std::string& m) { throw std::runtime_error(m); }

// include/LightGBM/utils/random.h:41-109 — the linear congruential generator behind Dataset row sampling
class Random {
 public:
  explicit Random(int seed) : x(seed) {}
  int NextInt(int lo, int hi) { return RandInt32() % (hi - lo) + lo; }
  float NextFloat() { return static_cast<float>(RandInt16()) / (32768.0f); }
  std::vector<int> Sample(int N, int K) {
    std::vector<int> ret;
    ret.reserve(K);
    if (K > N || K <= 0) return ret;
    if (K == N) {
      for (int i = 0; i < N; ++i) ret.push_back(i);
    } else if (K > 1 && K > (N / std::log2(K))) {
      for (int i = 0; i < N; ++i) {
        double prob = (K - ret.size()) / static_cast<double>(N - i);
        if (NextFloat() < prob) ret.push_back(i);
      }
    } else {
      std::set<int> sample_set;
      for (int r = N - K; r < N; ++r) {
        int v = NextInt(0, r);
        if (!sample_set.insert(v).second) sample_set.insert(r);
      }
      for (int v : sample_set) ret.push_back(v);
    }
    return ret;
  }

 private:
  int RandInt16() { x = (214013 * x + 2531011); return static_cast<int>((x >> 16) & 0x7FFF); }
  int RandInt32() { x = (214013 * x + 2531011); return static_cast<int>(x & 0x7FFFFFFF); }
  unsigned int x;
};

inline bool CheckDoubleEqualOrdered(double a, double b) { return b <= std::nextafter(a, INFINITY); }  // common.h:861
inline double GetDoubleUpperBound(double a) { return std::nextafter(a, INFINITY); }                   // common.h:866

// bin.cpp:78-155
std::vector<double> GreedyFindBin(const double* distinct_values, const int* counts, int num_distinct_values, int max_bin,
                                  size_t total_cnt, int min_data_in_bin) {
  std::vector<double> bin_upper_bound;
  if (!(max_bin > 0)) Fatal("Check failed: max_bin > 0");
  if (num_distinct_values <= max_bin) {
    int cur_cnt_inbin = 0;
    for (int i = 0; i < num_distinct_values - 1; ++i) {
      cur_cnt_inbin += counts[i];
      if (cur_cnt_inbin >= min_data_in_bin) {
        auto val = GetDoubleUpperBound((distinct_values[i] + distinct_values[i + 1]) / 2.0);
        if (bin_upper_bound.empty() || !CheckDoubleEqualOrdered(bin_upper_bound.back(), val)) {
          bin_upper_bound.push_back(val);
          cur_cnt_inbin = 0;
        }
      }
    }
    bin_upper_bound.push_back(std::numeric_limits<double>::infinity());
  } else {
    if (min_data_in_bin > 0) {
      max_bin = std::min(max_bin, static_cast<int>(total_cnt / min_data_in_bin));
      max_bin = std::max(max_bin, 1);
    }
    double mean_bin_size = static_cast<double>(total_cnt) / max_bin;
    int rest_bin_cnt = max_bin;
    int rest_sample_cnt = static_cast<int>(total_cnt);
    std::vector<bool> is_big_count_value(num_distinct_values, false);
    for (int i = 0; i < num_distinct_values; ++i) {
      if (counts[i] >= mean_bin_size) {
        is_big_count_value[i] = true;
        --rest_bin_cnt;
        rest_sample_cnt -= counts[i];
      }
    }
    mean_bin_size = static_cast<double>(rest_sample_cnt) / rest_bin_cnt;
    std::vector<double> upper_bounds(max_bin, std::numeric_limits<double>::infinity());
    std::vector<double> lower_bounds(max_bin, std::numeric_limits<double>::infinity());
    int bin_cnt = 0;
    lower_bounds[bin_cnt] = distinct_values[0];
    int cur_cnt_inbin = 0;
    for (int i = 0; i < num_distinct_values - 1; ++i) {
      if (!is_big_count_value[i]) rest_sample_cnt -= counts[i];
      cur_cnt_inbin += counts[i];
      if (is_big_count_value[i] || cur_cnt_inbin >= mean_bin_size ||
          (is_big_count_value[i + 1] && cur_cnt_inbin >= std::max(1.0, mean_bin_size * 0.5f))) {
        upper_bounds[bin_cnt] = distinct_values[i];
        ++bin_cnt;
        lower_bounds[bin_cnt] = distinct_values[i + 1];
        if (bin_cnt >= max_bin - 1) break;
        cur_cnt_inbin = 0;
        if (!is_big_count_value[i]) {
          --rest_bin_cnt;
          mean_bin_size = rest_sample_cnt / static_cast<double>(rest_bin_cnt);
        }
      }
    }
    ++bin_cnt;
    for (int i = 0; i < bin_cnt - 1; ++i) {
      auto val = GetDoubleUpperBound((upper_bounds[i] + lower_bounds[i + 1]) / 2.0);
      if (bin_upper_bound.empty() || !CheckDoubleEqualOrdered(bin_upper_bound.back(), val)) bin_upper_bound.push_back(val);
    }
    bin_upper_bound.push_back(std::numeric_limits<double>::infinity());
  }
  return bin_upper_bound;
}

// bin.cpp:241-297
std::vector<double> FindBinWithZeroAsOneBin(const double* distinct_values, const int* counts, int num_distinct_values, int max_bin,
                                            size_t total_sample_cnt, int min_data_in_bin) {
  std::vector<double> bin_upper_bound;
  int left_cnt_data = 0, cnt_zero = 0, right_cnt_data = 0;
  for (int i = 0; i < num_distinct_values; ++i) {
    if (distinct_values[i] <= -kZeroThreshold) left_cnt_data += counts[i];
    else if (distinct_values[i] > kZeroThreshold) right_cnt_data += counts[i];
    else cnt_zero += counts[i];
  }
  int left_cnt = -1;
  for (int i = 0; i < num_distinct_values; ++i) {
    if (distinct_values[i] > -kZeroThreshold) { left_cnt = i; break; }
  }
  if (left_cnt < 0) left_cnt = num_distinct_values;
  if ((left_cnt > 0) && (max_bin > 1)) {
    int left_max_bin = static_cast<int>(static_cast<double>(left_cnt_data) / (total_sample_cnt - cnt_zero) * (max_bin - 1));
    left_max_bin = std::max(1, left_max_bin);
    bin_upper_bound = GreedyFindBin(distinct_values, counts, left_cnt, left_max_bin, left_cnt_data, min_data_in_bin);
    if (bin_upper_bound.size() > 0) bin_upper_bound.back() = -kZeroThreshold;
  }
  int right_start = -1;
  for (int i = left_cnt; i < num_distinct_values; ++i) {
    if (distinct_values[i] > kZeroThreshold) { right_start = i; break; }
  }
  int right_max_bin = max_bin - 1 - static_cast<int>(bin_upper_bound.size());
  if (right_start >= 0 && right_max_bin > 0) {
    auto right_bounds = GreedyFindBin(distinct_values + right_start, counts + right_start, num_distinct_values - right_start,
                                      right_max_bin, right_cnt_data, min_data_in_bin);
    bin_upper_bound.push_back(kZeroThreshold);
    bin_upper_bound.insert(bin_upper_bound.end(), right_bounds.begin(), right_bounds.end());
  } else {
    bin_upper_bound.push_back(std::numeric_limits<double>::infinity());
  }
  if (!(bin_upper_bound.size() <= static_cast<size_t>(max_bin))) Fatal("Check failed: bin_upper_bound.size() <= max_bin");
  return bin_upper_bound;
}

// bin.cpp:53-66 (numerical)
bool NeedFilter(const std::vector<int>& cnt_in_bin, int total_cnt, int filter_cnt) {
  int sum_left = 0;
  for (size_t i = 0; i + 1 < cnt_in_bin.size(); ++i) {
    sum_left += cnt_in_bin[i];
    if (sum_left >= filter_cnt && total_cnt - sum_left >= filter_cnt) return false;
  }
  return true;
}

// BinMapper::FindBin, numerical, use_missing but no NaN present -> MissingType::None (bin.cpp:325-520)
FeatureBins FindBin(std::vector<double>& values, size_t total_sample_cnt, int max_bin, int min_data_in_bin, int min_split_data,
                    bool pre_filter) {
  FeatureBins fb;
  int num_sample_values = (int)values.size();
  for (double v : values)
    if (std::isnan(v)) Fatal("Missing values (NaN) in the feature matrix are not supported by the B200 tree learner yet");
  const int zero_cnt = static_cast<int>(total_sample_cnt - num_sample_values);
  std::vector<double> distinct_values;
  std::vector<int> counts;
  std::stable_sort(values.begin(), values.end());
  if (num_sample_values == 0 || (values[0] > 0.0f && zero_cnt > 0)) { distinct_values.push_back(0.0f); counts.push_back(zero_cnt); }
  if (num_sample_values > 0) { distinct_values.push_back(values[0]); counts.push_back(1); }
  for (int i = 1; i < num_sample_values; ++i) {
    if (!CheckDoubleEqualOrdered(values[i - 1], values[i])) {
      if (values[i - 1] < 0.0f && values[i] > 0.0f) { distinct_values.push_back(0.0f); counts.push_back(zero_cnt); }
      distinct_values.push_back(values[i]);
      counts.push_back(1);
    } else {
      distinct_values.back() = values[i];
      ++counts.back();
    }
  }
  if (num_sample_values > 0 && values[num_sample_values - 1] < 0.0f && zero_cnt > 0) { distinct_values.push_back(0.0f); counts.push_back(zero_cnt); }
  fb.min_val = distinct_values.front();
  fb.max_val = distinct_values.back();
  const int num_distinct_values = (int)distinct_values.size();
  fb.upper_bounds = FindBinWithZeroAsOneBin(distinct_values.data(), counts.data(), num_distinct_values, max_bin, total_sample_cnt,
                                            min_data_in_bin);
  fb.num_bin = (int)fb.upper_bounds.size();
  std::vector<int> cnt_in_bin(fb.num_bin, 0);
  int i_bin = 0;
  for (int i = 0; i < num_distinct_values; ++i) {
    if (distinct_values[i] > fb.upper_bounds[i_bin]) ++i_bin;
    cnt_in_bin[i_bin] += counts[i];
  }
  if (!(fb.num_bin <= max_bin)) Fatal("Check failed: num_bin_ <= max_bin");
  fb.trivial = fb.num_bin <= 1;
  if (!fb.trivial && pre_filter && NeedFilter(cnt_in_bin, static_cast<int>(total_sample_cnt), min_split_data)) fb.trivial = true;
  return fb;
}

template <typename T>
inline double At(const void* data, int32_t nrow, int32_t ncol, int is_row_major, int64_t i, int j) {
  const T* p = static_cast<const T*>(data);
  return is_row_major ? (double)p[i * ncol + j] : (double)p[(int64_t)j * nrow + i];
}

}  // namespace

Params Params::Parse(const char* s) {
  Params p;
  if (!s) return p;
  std::istringstream is(s);
  std::string tok;
  while (is >> tok) {
    auto eq = tok.find('=');
    if (eq == std::string::npos) continue;
    p.kv[tok.substr(0, eq)] = tok.substr(eq + 1);
  }
  return p;
}
static const std::string* Find(const std::map<std::string, std::string>& kv, const std::string& k, std::initializer_list<const char*> al) {
  auto it = kv.find(k);
  if (it != kv.end()) return &it->second;
  for (const char* a : al) {
    it = kv.find(a);
    if (it != kv.end()) return &it->second;
  }
  return nullptr;
}
int Params::GetInt(const std::string& k, int d, std::initializer_list<const char*> al) const {
  const std::string* v = Find(kv, k, al);
  return v ? std::stoi(*v) : d;
}
double Params::GetDouble(const std::string& k, double d, std::initializer_list<const char*> al) const {
  const std::string* v = Find(kv, k, al);
  return v ? std::stod(*v) : d;
}
bool Params::GetBool(const std::string& k, bool d, std::initializer_list<const char*> al) const {
  const std::string* v = Find(kv, k, al);
  if (!v) return d;
  return *v == "true" || *v == "True" || *v == "1" || *v == "+";
}
std::string Params::GetString(const std::string& k, const std::string& d, std::initializer_list<const char*> al) const {
  const std::string* v = Find(kv, k, al);
  return v ? *v : d;
}

uint32_t FeatureBins::ValueToBin(double value) const {  // bin.h:465-488 (numerical, MissingType::None)
  if (std::isnan(value)) value = 0.0f;
  int l = 0, r = num_bin - 1;
  while (l < r) {
    int m = (r + l - 1) / 2;
    if (value <= upper_bounds[m]) r = m; else l = m + 1;
  }
  return (uint32_t)l;
}

Dataset::Dataset(const void* data, int data_type, int32_t nrow, int32_t ncol, int is_row_major, const Params& params)
    : num_data_(nrow), num_total_features_(ncol), params_(params) {
  if (data == nullptr || nrow <= 0 || ncol <= 0) Fatal("LGBM_DatasetCreateFromMat: empty data");
  if (data_type != 0 && data_type != 1) Fatal("Unknown data type in LGBM_DatasetCreateFromMat (float32 / float64 supported)");
  const int max_bin = params.GetInt("max_bin", 255);
  if (max_bin < 2 || max_bin > 255) Fatal("max_bin must be in [2, 255] for the B200 tree learner (uint8 bins)");
  const int min_data_in_bin = params.GetInt("min_data_in_bin", 3);
  const int sample_cnt_cfg = params.GetInt("bin_construct_sample_cnt", 200000, {"subsample_for_bin"});
  const int data_random_seed = params.GetInt("data_random_seed", 1, {"data_seed"});
  const bool pre_filter = params.GetBool("feature_pre_filter", true);
  const int min_data_in_leaf = params.GetInt("min_data_in_leaf", 20, {"min_data_per_leaf", "min_data", "min_child_samples"});
  auto at = [&](int64_t i, int j) {
    return data_type == 0 ? At<float>(data, nrow, ncol, is_row_major, i, j) : At<double>(data, nrow, ncol, is_row_major, i, j);
  };
  // ---- row sample (c_api.cpp:1182-1206)
  Random rand(data_random_seed);
  int sample_cnt = nrow < sample_cnt_cfg ? nrow : sample_cnt_cfg;
  std::vector<int> sample_indices = rand.Sample(nrow, sample_cnt);
  sample_cnt = (int)sample_indices.size();
  const int filter_cnt = static_cast<int>(static_cast<double>(min_data_in_leaf * (int64_t)sample_cnt) / nrow);  // dataset_loader.cpp:644
  bins_.resize(ncol);
#pragma omp parallel for schedule(dynamic)
  for (int j = 0; j < ncol; ++j) {
    std::vector<double> vals;
    vals.reserve(sample_cnt);
    for (int s = 0; s < sample_cnt; ++s) {
      const double v = at(sample_indices[s], j);
      if (std::fabs(v) > kZeroThreshold || std::isnan(v)) vals.push_back(v);
    }
    try {
      bins_[j] = FindBin(vals, (size_t)sample_cnt, max_bin, min_data_in_bin, filter_cnt, pre_filter);
    } catch (...) {
      bins_[j].num_bin = -1;  // re-raised below (no exceptions across the OpenMP region)
    }
  }
  for (int j = 0; j < ncol; ++j)
    if (bins_[j].num_bin < 0) Fatal("Missing values (NaN) in the feature matrix are not supported by the B200 tree learner yet");
  for (int j = 0; j < ncol; ++j)
    if (!bins_[j].trivial) used_features_.push_back(j);
  // ---- value -> bin for every row (Dataset::PushOneRow -> BinMapper::ValueToBin), feature-major like DenseBin
  bin_data_.resize((size_t)used_features_.size() * nrow);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < (int)used_features_.size(); ++k) {
    const int j = used_features_[k];
    uint8_t* col = bin_data_.data() + (size_t)k * nrow;
    for (int64_t i = 0; i < nrow; ++i) col[i] = (uint8_t)bins_[j].ValueToBin(at(i, j));
  }
}

void Dataset::SetLabel(const float* label, int n) {
  if (n != num_data_) Fatal("Length of label is not same with #data");
  label_.assign(label, label + n);
}

std::vector<std::string> Dataset::feature_infos() const {
  std::vector<std::string> out;
  for (int j = 0; j < num_total_features_; ++j) {
    const FeatureBins& fb = bins_[j];
    if (fb.trivial) { out.push_back("none"); continue; }
    char buf[96];
    std::snprintf(buf, sizeof(buf), "[%.17g:%.17g]", fb.min_val, fb.max_val);
    out.push_back(buf);
  }
  return out;
}

}  // namespace gpb200
