"""Feature matrices for the binning parity test (shared by tests/golden/make_binning_golden.py and tests/test_binning_gpu.py)."""
import numpy as np

CASES = [
    dict(name="all_rows", n=3000, seed=11, params=dict(max_bin=255, min_data_in_bin=3, min_data_in_leaf=20)),
    dict(name="selection_sample", n=4000, seed=12, params=dict(max_bin=255, min_data_in_bin=3, min_data_in_leaf=20, bin_construct_sample_cnt=1500)),
    dict(name="floyd_sample", n=4000, seed=13, params=dict(max_bin=63, min_data_in_bin=5, min_data_in_leaf=10, bin_construct_sample_cnt=300)),
    dict(name="few_bins", n=2500, seed=14, params=dict(max_bin=15, min_data_in_bin=1, min_data_in_leaf=50, data_random_seed=7)),
    dict(name="pre_filter", n=2000, seed=15, params=dict(max_bin=255, min_data_in_bin=3, min_data_in_leaf=700, feature_pre_filter=True)),
]


def make_matrix(n, seed):
    r = np.random.default_rng(seed)
    cols = [
        r.random(n),                                                       # continuous, positive
        r.standard_normal(n),                                              # both signs
        r.integers(0, 10, n).astype(float),                                # low cardinality with zeros
        np.where(r.random(n) < 0.7, 0.0, r.exponential(2.0, n)),           # mostly zero
        np.where(r.random(n) < 0.3, 1.5, r.random(n) * 4),                 # one heavy value among many distinct ones
        -r.gamma(2.0, 1.0, n),                                             # negative only
        np.full(n, 3.25),                                                  # constant: filtered out
        1.0 + r.integers(0, 40, n) * np.finfo(float).eps,                  # neighbours one ulp apart
        (r.random(n) < 0.4).astype(float),                                 # binary
        np.round(r.standard_normal(n), 1),                                 # repeated values of both signs and exact zeros
        r.integers(-3, 4, n).astype(float) * 1e-36,                        # inside the zero band
    ]
    return np.ascontiguousarray(np.stack(cols, axis=1))
