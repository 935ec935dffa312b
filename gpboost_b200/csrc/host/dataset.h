// Host-side Dataset of the B200 build: parameter parsing, bin finding and value->bin mapping for dense numerical
// matrices; produces the feature-major uint8 bins the device tree learner consumes.
//
// Restates, for numerical features without missing values:
//   LGBM_DatasetCreateFromMat / ...Mats           src/LightGBM/c_api.cpp:1112-1232 (row sampling with Random(data_random_seed))
//   DatasetLoader::ConstructFromSampleData        src/LightGBM/io/dataset_loader.cpp:600-700 (filter_cnt, FindBin per column)
//   BinMapper::FindBin                            src/LightGBM/io/bin.cpp:325-520
//   FindBinWithZeroAsOneBin / GreedyFindBin       bin.cpp:241-311 / :78-155
//   NeedFilter                                    bin.cpp:53-75
//   BinMapper::ValueToBin                         include/LightGBM/bin.h:465-503
//   Random::Sample / NextFloat / NextInt          include/LightGBM/utils/random.h:41-109
// Bin finding runs once per dataset on the host, as in the reference (device binning is a later row).
#ifndef GPB200_DATASET_H_
#define GPB200_DATASET_H_
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace gpb200 {

// key=value parameters ("parameters" strings of the LGBM_* API; Config::Str2Map, src/LightGBM/io/config.cpp)
struct Params {
  std::map<std::string, std::string> kv;
  static Params Parse(const char* s);
  int GetInt(const std::string& k, int dflt, std::initializer_list<const char*> aliases = {}) const;
  double GetDouble(const std::string& k, double dflt, std::initializer_list<const char*> aliases = {}) const;
  bool GetBool(const std::string& k, bool dflt, std::initializer_list<const char*> aliases = {}) const;
  std::string GetString(const std::string& k, const std::string& dflt, std::initializer_list<const char*> aliases = {}) const;
};

struct FeatureBins {
  std::vector<double> upper_bounds;  // bin_upper_bound_
  int num_bin = 0;
  bool trivial = false;
  double min_val = 0., max_val = 0.;
  uint32_t ValueToBin(double v) const;
};

class Dataset {
 public:
  // data: nrow x ncol, float32 or float64 (C_API_DTYPE_*), row- or column-major
  Dataset(const void* data, int data_type, int32_t nrow, int32_t ncol, int is_row_major, const Params& params);
  int32_t num_data() const { return num_data_; }
  int num_total_features() const { return num_total_features_; }
  int num_features() const { return (int)used_features_.size(); }   // non-trivial ("inner") features
  int real_feature_index(int inner) const { return used_features_[inner]; }
  const FeatureBins& feature(int inner) const { return bins_[used_features_[inner]]; }
  const std::vector<uint8_t>& bins_feature_major() const { return bin_data_; }  // num_features() x num_data
  void SetLabel(const float* label, int n);
  const std::vector<float>& label() const { return label_; }
  bool has_label() const { return !label_.empty(); }
  const Params& params() const { return params_; }
  // feature names (Dataset::set_feature_names / feature_names, include/LightGBM/dataset.h): default "Column_<i>"
  const std::vector<std::string>& feature_names() const { return feature_names_; }
  void set_feature_names(const std::vector<std::string>& names) { feature_names_ = names; }
  // "[min:max]" per feature, "none" for a trivial one (BinMapper::bin_info_string, include/LightGBM/bin.h:195-215)
  std::vector<std::string> feature_infos() const;

 private:
  int32_t num_data_ = 0;
  int num_total_features_ = 0;
  Params params_;
  std::vector<FeatureBins> bins_;      // per real feature
  std::vector<int> used_features_;     // inner -> real
  std::vector<uint8_t> bin_data_;
  std::vector<float> label_;
  std::vector<std::string> feature_names_;
};

}  // namespace gpb200
#endif
