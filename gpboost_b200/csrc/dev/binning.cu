// Device binning: value -> bin for every (row, used feature) of a dense host matrix, written straight into the
// row-major n x Fpad uint8 layout the tree learner reads. C ABI in include/gpboost_b200_dev.h.
//
// Replaces the n x F host pass of LGBM_DatasetCreateFromMat (src/LightGBM/c_api.cpp:1134-1232: PushOneRow ->
// BinMapper::ValueToBin, include/LightGBM/bin.h:465-503, numerical feature, MissingType::None) and the feature-major -> row-major
// transposition + H2D of the bins that followed it. The bin BOUNDARIES stay on the host (sample-based, csrc/host/dataset.cpp).
//
// Semantics: bin(v) = the smallest l with v <= upper_bound[l] (the reference's binary search finds the same l because the bounds
// increase strictly and the last one is +inf); NaN counts as 0. Here: bin = #{l : upper_bound[l] < v}, found by a branch-free
// power-of-two descent over the feature's bounds padded to 256 entries with +inf — integer output, bit-exact by construction.
//
// Layout / traffic: the host matrix is streamed in row chunks (pageable memory -> device staging buffer, cudaMemcpy(2D)); one
// thread produces one 4-feature word of one row, a warp writes 128 contiguous bytes of the bin matrix. Algorithmic bytes per row:
// ncol * sizeof(T) read + Fpad written; the bounds (2 KB per feature) stay in L1/L2.
#include "../../../include/gpboost_b200_dev.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace {

thread_local std::string g_bin_err;
int bfail(const std::string& m) { g_bin_err = m; return -1; }
#define BCUDA(expr)                                                                                          \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return bfail(std::string("CUDA error at " __FILE__ ":") + std::to_string(__LINE__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

constexpr int kBoundsPerFeature = 256;

__device__ __forceinline__ uint32_t bin_of(double v, const double* __restrict__ ub) {
  v = (v != v) ? 0. : v;  // NaN -> 0 (bin.h:466-468 for MissingType::None)
  int pos = 0;
#pragma unroll
  for (int step = kBoundsPerFeature / 2; step >= 1; step >>= 1)
    if (__ldg(ub + pos + step - 1) < v) pos += step;
  return (uint32_t)pos;  // <= 255: the padding (+inf) is never < v
}

// chunk: rows [0, rows) of the staged block. ROWMAJOR: stage[r * ncol + c]; else stage[c * ld + r] (ld = rows of the staged block).
// word w of row r = bins of used features 4w .. 4w+3 (0 for padding features).
template <typename T, bool ROWMAJOR>
__global__ void __launch_bounds__(256) bin_rows_kernel(const T* __restrict__ stage, int64_t rows, int ncol, int64_t ld, int F, int words,
                                                       const int32_t* __restrict__ real_feature, const double* __restrict__ bounds,
                                                       uint32_t* __restrict__ out) {
  const int64_t total = rows * words;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r;
    int w;
    if (ROWMAJOR) { r = e / words; w = (int)(e - r * words); }  // consecutive threads: consecutive words (coalesced write, near-contiguous read)
    else { w = (int)(e / rows); r = e - (int64_t)w * rows; }      // consecutive threads: consecutive rows (coalesced column reads)
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int f = 4 * w + k;
      if (f < F) {
        const int c = real_feature[f];
        const double v = (double)(ROWMAJOR ? stage[r * ncol + c] : stage[(int64_t)c * ld + r]);
        packed |= bin_of(v, bounds + (size_t)f * kBoundsPerFeature) << (8 * k);
      }
    }
    out[r * words + w] = packed;
  }
}

template <typename T>
int bin_typed(const T* host, int64_t nrow, int ncol, int is_row_major, int F, int Fpad, const int32_t* real_dev, const double* bounds_dev,
              uint8_t* bins_dev, cudaStream_t stream, int num_sms) {
  // staging block of at most ~256 MB; rows of a column-major matrix are gathered by a 2-D copy (one strip per column)
  const int64_t row_bytes = (int64_t)ncol * (int64_t)sizeof(T);
  int64_t chunk = std::max<int64_t>(1, (int64_t)(256ll << 20) / row_bytes);
  chunk = std::min(chunk, nrow);
  T* stage = nullptr;
  BCUDA(cudaMalloc(&stage, (size_t)chunk * row_bytes));
  const int words = Fpad / 4;
  int rc = 0;
  for (int64_t r0 = 0; r0 < nrow && rc == 0; r0 += chunk) {
    const int64_t rows = std::min(chunk, nrow - r0);
    cudaError_t e;
    if (is_row_major) e = cudaMemcpyAsync(stage, host + r0 * ncol, (size_t)rows * row_bytes, cudaMemcpyHostToDevice, stream);
    else e = cudaMemcpy2DAsync(stage, (size_t)rows * sizeof(T), host + r0, (size_t)nrow * sizeof(T), (size_t)rows * sizeof(T), (size_t)ncol,
                               cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { rc = bfail(std::string("gpbdev_bin_matrix: host -> device copy failed: ") + cudaGetErrorString(e)); break; }
    const int grid = (int)std::min<int64_t>((rows * words + 255) / 256, (int64_t)num_sms * 16);
    uint32_t* out = reinterpret_cast<uint32_t*>(bins_dev + (size_t)r0 * Fpad);
    if (is_row_major) bin_rows_kernel<T, true><<<grid, 256, 0, stream>>>(stage, rows, ncol, rows, F, words, real_dev, bounds_dev, out);
    else bin_rows_kernel<T, false><<<grid, 256, 0, stream>>>(stage, rows, ncol, rows, F, words, real_dev, bounds_dev, out);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // the staging block is reused by the next chunk
    if (e != cudaSuccess) rc = bfail(std::string("gpbdev_bin_matrix: binning kernel failed: ") + cudaGetErrorString(e));
  }
  cudaFree(stage);
  return rc;
}

}  // namespace

extern "C" {

const char* gpbdev_bin_last_error(void) { return g_bin_err.c_str(); }

int gpbdev_bin_matrix(int device, const void* data_host, int data_type, int64_t nrow, int ncol, int is_row_major, int F,
                      const int32_t* real_feature, const int32_t* num_bin, const double* upper_bounds, int upper_bounds_stride,
                      int Fpad, uint8_t** bins_dev_out) {
  if (!data_host || !real_feature || !num_bin || !upper_bounds || !bins_dev_out) return bfail("gpbdev_bin_matrix: null argument");
  if (nrow <= 0 || ncol <= 0 || F <= 0 || Fpad < F || (Fpad % 32) != 0) return bfail("gpbdev_bin_matrix: bad shape");
  if (data_type != 0 && data_type != 1) return bfail("gpbdev_bin_matrix: data_type must be 0 (float32) or 1 (float64)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    cudaGetLastError();
    return bfail("gpbdev_bin_matrix: no CUDA device " + std::to_string(device) + " — the B200 Dataset has no CPU fallback");
  }
  BCUDA(cudaSetDevice(device));
  // bounds padded to 256 per feature with +inf
  std::vector<double> padded((size_t)F * kBoundsPerFeature, INFINITY);
  for (int f = 0; f < F; ++f) {
    if (num_bin[f] < 1 || num_bin[f] > kBoundsPerFeature) return bfail("gpbdev_bin_matrix: num_bin must be in [1, 256]");
    if (real_feature[f] < 0 || real_feature[f] >= ncol) return bfail("gpbdev_bin_matrix: feature index out of range");
    for (int b = 0; b < num_bin[f]; ++b) padded[(size_t)f * kBoundsPerFeature + b] = upper_bounds[(size_t)f * upper_bounds_stride + b];
  }
  cudaDeviceProp prop;
  BCUDA(cudaGetDeviceProperties(&prop, device));
  cudaStream_t stream;
  BCUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  double* bounds_dev = nullptr;
  int32_t* real_dev = nullptr;
  uint8_t* bins = nullptr;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(bounds_dev); cudaFree(real_dev); cudaStreamDestroy(stream); };
  cudaError_t e = cudaMalloc(&bounds_dev, padded.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&real_dev, sizeof(int32_t) * F);
  if (e == cudaSuccess) e = cudaMalloc(&bins, (size_t)nrow * Fpad);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bounds_dev, padded.data(), padded.size() * sizeof(double), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(real_dev, real_feature, sizeof(int32_t) * F, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) {
    cleanup(); cudaFree(bins);
    return bfail(std::string("gpbdev_bin_matrix: device allocation failed: ") + cudaGetErrorString(e));
  }
  if (data_type == 0) rc = bin_typed(static_cast<const float*>(data_host), nrow, ncol, is_row_major, F, Fpad, real_dev, bounds_dev, bins, stream, prop.multiProcessorCount);
  else rc = bin_typed(static_cast<const double*>(data_host), nrow, ncol, is_row_major, F, Fpad, real_dev, bounds_dev, bins, stream, prop.multiProcessorCount);
  cleanup();
  if (rc != 0) { cudaFree(bins); return rc; }
  *bins_dev_out = bins;
  return 0;
}

int gpbdev_bin_free(int device, uint8_t* bins_dev) {
  if (!bins_dev) return 0;
  cudaSetDevice(device);
  cudaFree(bins_dev);
  return 0;
}

int gpbdev_bin_download(int device, const uint8_t* bins_dev, int64_t nrow, int Fpad, uint8_t* out_host) {
  if (!bins_dev || !out_host) return bfail("gpbdev_bin_download: null argument");
  BCUDA(cudaSetDevice(device));
  BCUDA(cudaMemcpy(out_host, bins_dev, (size_t)nrow * Fpad, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
