// Device engine for the exact (dense) Gaussian-process likelihood — SURVEY §8 row a6, BASELINE config 1.
//
// Replaces RECompGP::CalcSigma (include/GPBoost/re_comp.h:1273 -> CovFunction::CalculateCovMat cov_fcts.h:635-755),
// CalcZSigmaZt (re_model_template.h:9273: Psi = I + Sigma), CalcChol (:6492, Eigen::LLT), the two triangular solves of
// CalcYAux (:9894) / CalcYTPsiIInvY (:10002) and the log-determinant (:3127).
//
// Formulation: the (n+1) x (n+1) matrix [[Psi, y],[y^T, *]] (lower triangle, row-major, 64 x 64 tiles) is factorised by a
// right-looking blocked Cholesky; the response rides along as row n, so after the factorisation
//     log|Psi| = 2 sum_{i<n} log L_ii      and      y^T Psi^-1 y = sum_{c<n} L[n][c]^2
// without a separate forward solve. Per block column k: (1) POTRF of the diagonal tile in shared memory, (2) TRSM of
// the tiles below it, (3) trailing update A_ij -= L_ik L_jk^T on the FP64 tensor cores (mma.sync m8n8k4 f64 = DMMA; tcgen05
// has no FP64 kind), operands staged in shared memory. Psi^-1 y = L^-T z by a blocked back substitution.
// Bound: tensor/FP64 pipe for the trailing update (n^3/3 flop), HBM for the Gram build (8 n^2 bytes written once).
#include "../../../include/gpboost_b200_dev.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "vecchia_factor.cuh"  // cov_eval, rsqrt_fast

namespace {
thread_local std::string g_dense_err;
int dfail(const std::string& m) { g_dense_err = m; return -1; }
#define DCUDA(expr)                                                                                          \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return dfail(std::string("CUDA error at " __FILE__ ":") + std::to_string(__LINE__) + ": " + cudaGetErrorString(e__)); \
  } while (0)

constexpr int NB = 64;

// ---- Gram build: lower-triangular tiles of Psi = I + Sigma, plus the response row n
template <int COV>
__global__ void gram_kernel(const double* __restrict__ coords, int n, int d, const double* __restrict__ y, double var, double range,
                            double* __restrict__ A, int ld) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  const int N = n + 1;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = ti * NB + e / NB, c = tj * NB + e % NB;
    if (r >= ld || c >= ld) continue;
    double v = 0.;
    if (r < n && c < n) {
      if (r == c) v = var + 1.;
      else if (c < r) {
        double d2 = 0.;
        for (int k = 0; k < d; ++k) { const double df = coords[(size_t)r * d + k] - coords[(size_t)c * d + k]; d2 = fma(df, df, d2); }
        const double dist = d2 * gpb::rsqrt_fast(d2 + 1e-300);
        double g;
        v = gpb::cov_eval<COV, false>(dist, var, range, g);
      }
    } else if (r == n && c < n) {
      v = y[c];
    } else if (r == c) {
      v = 1.;  // padding rows/cols (and the response row's own diagonal): identity, never used as pivots that matter
    }
    A[(size_t)r * ld + c] = v;
    (void)N;
  }
}

// ---- POTRF of the diagonal tile k (one CTA, tile in shared memory). Rows >= npiv of the tile are eliminated but not pivoted.
__global__ void potrf_tile_kernel(double* __restrict__ A, int ld, int k, int npiv, int* __restrict__ info) {
  __shared__ double T[NB][NB + 1];
  double* base = A + (size_t)k * NB * ld + (size_t)k * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) T[e / NB][e % NB] = base[(size_t)(e / NB) * ld + e % NB];
  __syncthreads();
  for (int p = 0; p < npiv; ++p) {
    const double dpp = T[p][p];
    if (threadIdx.x == 0 && !(dpp > 0.)) atomicAdd(info, 1);
    const double rs = gpb::rsqrt_fast(dpp);
    __syncthreads();
    for (int r = p + threadIdx.x; r < NB; r += blockDim.x) T[r][p] *= rs;  // column p (incl. the diagonal: sqrt)
    __syncthreads();
    for (int e = threadIdx.x; e < (NB - p - 1) * (NB - p - 1); e += blockDim.x) {
      const int r = p + 1 + e / (NB - p - 1), c = p + 1 + e % (NB - p - 1);
      if (c <= r) T[r][c] -= T[r][p] * T[c][p];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    if (c <= r) base[(size_t)r * ld + c] = T[r][c];
  }
}

// ---- TRSM: tiles (i, k), i > k:  X = A_ik L_kk^-T  (each thread owns one row of the tile; L_kk in shared memory)
__global__ void trsm_tile_kernel(double* __restrict__ A, int ld, int k, int npiv) {
  __shared__ double L[NB][NB + 1];
  const double* lkk = A + (size_t)k * NB * ld + (size_t)k * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) L[e / NB][e % NB] = lkk[(size_t)(e / NB) * ld + e % NB];
  __syncthreads();
  const int i = k + 1 + blockIdx.x;
  const int r = threadIdx.x;
  if (r >= NB) return;
  double* row = A + ((size_t)i * NB + r) * ld + (size_t)k * NB;
  double x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = row[c];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    if (c < npiv) {
      double s = x[c];
#pragma unroll
      for (int t = 0; t < c; ++t) s -= x[t] * L[c][t];
      x[c] = s / L[c][c];
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = x[c];
}

// ---- trailing update on the FP64 tensor cores: A_ij -= L_ik L_jk^T for i >= j > k. One CTA (8 warps) per tile; warp w owns
// the 8-row band w of the tile and sweeps its 8 column blocks with mma.sync.m8n8k4.f64.
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) syrk_tile_kernel(double* __restrict__ A, int ld, int k, int nt) {
  // map the linear block index to a lower-triangular tile (i, j) of the trailing matrix, i >= j > k
  const int m = nt - k - 1;
  int t = blockIdx.x;
  int bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  const int bj = t - bi * (bi + 1) / 2;
  if (bi >= m) return;
  const int i = k + 1 + bi, j = k + 1 + bj;
  extern __shared__ __align__(16) double syrk_sm[];
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(syrk_sm);
  double (*Lj)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(syrk_sm + NB * (NB + 1));
  const double* pi = A + (size_t)i * NB * ld + (size_t)k * NB;
  const double* pj = A + (size_t)j * NB * ld + (size_t)k * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    Li[e / NB][e % NB] = pi[(size_t)(e / NB) * ld + e % NB];
    Lj[e / NB][e % NB] = pj[(size_t)(e / NB) * ld + e % NB];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ar = warp * 8 + (lane >> 2);   // A fragment: row lane/4 of the band, k-column lane%4
  const int kc = lane & 3;
  double acc[8][2];
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) { acc[cb][0] = 0.; acc[cb][1] = 0.; }
#pragma unroll 4
  for (int k0 = 0; k0 < NB; k0 += 4) {
    const double a = Li[ar][k0 + kc];
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
      const double b = Lj[cb * 8 + (lane >> 2)][k0 + kc];  // B[k][c] = L_j[c][k], column c = lane/4 of block cb
      dmma8x8x4(acc[cb][0], acc[cb][1], a, b);
    }
  }
  double* out = A + (size_t)i * NB * ld + (size_t)j * NB;
  const int orow = warp * 8 + (lane >> 2);
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    const int oc = cb * 8 + (lane & 3) * 2;
    double* o = out + (size_t)orow * ld + oc;
    o[0] -= acc[cb][0];
    o[1] -= acc[cb][1];
  }
}

// ---- log-det and quadratic form from the factor: out[0] = sum_{c<n} L[n][c]^2, out[1] = 2 sum_{i<n} log L_ii
__global__ void dense_sums_kernel(const double* __restrict__ A, int ld, int n, double* __restrict__ out) {
  __shared__ double s0[256], s1[256];
  double q = 0., l = 0.;
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per, e = min(b + per, n);
  for (int c = b; c < e; ++c) { const double z = A[(size_t)n * ld + c]; q += z * z; l += log(A[(size_t)c * ld + c]); }
  s0[threadIdx.x] = q; s1[threadIdx.x] = l;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = s0[0]; out[1] = 2. * s1[0]; }
}

// ---- Psi^-1 y = L^-T z, z = L[n][0..n): one CTA; x lives in shared memory; columns are swept from the last to the first
__global__ void dense_backsolve_kernel(const double* __restrict__ A, int ld, int n, double scale, double* x_out) {
  double* xs = x_out;  // global (L2-resident) work vector: one CTA, ordered by __syncthreads
  for (int c = threadIdx.x; c < n; c += blockDim.x) xs[c] = A[(size_t)n * ld + c];
  __syncthreads();
  for (int r = n - 1; r >= 0; --r) {
    if (threadIdx.x == 0) xs[r] /= A[(size_t)r * ld + r];
    __syncthreads();
    const double xr = xs[r];
    const double* lrow = A + (size_t)r * ld;  // row r of L holds L[r][c] for c < r: contiguous, coalesced
    for (int c = threadIdx.x; c < r; c += blockDim.x) xs[c] -= lrow[c] * xr;
    __syncthreads();
  }
  for (int c = threadIdx.x; c < n; c += blockDim.x) xs[c] *= scale;
}

using GramKernel = void (*)(const double*, int, int, const double*, double, double, double*, int);
GramKernel pick_gram(int cov) {
  switch (cov) {
    case 0: return gram_kernel<0>;
    case 1: return gram_kernel<1>;
    case 2: return gram_kernel<2>;
    default: return gram_kernel<3>;
  }
}
}  // namespace

struct gpbdev_dense {
  int device = 0;
  int n = 0, d = 0, ld = 0, nt = 0;
  cudaStream_t stream = nullptr;
  double *coords = nullptr, *y = nullptr, *A = nullptr, *out = nullptr, *x = nullptr;
  int* info = nullptr;
  double* out_host = nullptr;
  int64_t launches = 0;
  bool factored = false;
};

extern "C" {

const char* gpbdev_dense_last_error(void) { return g_dense_err.c_str(); }

int gpbdev_dense_create(gpbdev_dense_t* out, int device, int n, int d, const double* coords_rowmajor) {
  if (!out || !coords_rowmajor) return dfail("gpbdev_dense_create: null argument");
  if (n <= 0 || d <= 0) return dfail("gpbdev_dense_create: need n > 0 and dim > 0");
  if (n > 46000) return dfail("gpbdev_dense_create: the dense engine keeps an n x n fp64 matrix; n is limited to 46000 (use gp_approx = 'vecchia')");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    cudaGetLastError();
    return dfail("gpbdev_dense_create: no CUDA device " + std::to_string(device) + " — the B200 engine has no CPU fallback");
  }
  DCUDA(cudaSetDevice(device));
  gpbdev_dense* h = new gpbdev_dense();
  h->device = device; h->n = n; h->d = d;
  h->nt = (n + 1 + NB - 1) / NB;
  h->ld = h->nt * NB;
  DCUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  DCUDA(cudaMalloc(&h->coords, sizeof(double) * (size_t)n * d));
  DCUDA(cudaMalloc(&h->y, sizeof(double) * n));
  DCUDA(cudaMalloc(&h->x, sizeof(double) * n));
  DCUDA(cudaMalloc(&h->A, sizeof(double) * (size_t)h->ld * h->ld));
  DCUDA(cudaMalloc(&h->out, sizeof(double) * 4));
  DCUDA(cudaMalloc(&h->info, sizeof(int)));
  DCUDA(cudaMallocHost(&h->out_host, sizeof(double) * 4));
  DCUDA(cudaMemcpy(h->coords, coords_rowmajor, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice));
  DCUDA(cudaMemset(h->y, 0, sizeof(double) * n));
  DCUDA(cudaFuncSetAttribute(syrk_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * NB * (NB + 1))));
  *out = h;
  return 0;
}

int gpbdev_dense_free(gpbdev_dense_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->coords); cudaFree(h->y); cudaFree(h->x); cudaFree(h->A); cudaFree(h->out); cudaFree(h->info);
  cudaFreeHost(h->out_host);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int gpbdev_dense_set_y(gpbdev_dense_t h, const double* y_host) {
  if (!h || !y_host) return dfail("gpbdev_dense_set_y: null argument");
  DCUDA(cudaSetDevice(h->device));
  DCUDA(cudaMemcpyAsync(h->y, y_host, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
  DCUDA(cudaStreamSynchronize(h->stream));
  h->factored = false;
  return 0;
}

int gpbdev_dense_eval(gpbdev_dense_t h, int cov_type, double var, double range, double* out3) {
  if (!h || !out3) return dfail("gpbdev_dense_eval: null argument");
  if (cov_type < 0 || cov_type > 3) return dfail("gpbdev_dense_eval: unknown covariance id");
  if (!(var > 0.) || !(range > 0.)) return dfail("gpbdev_dense_eval: covariance parameters must be positive");
  DCUDA(cudaSetDevice(h->device));
  const int n = h->n, nt = h->nt, ld = h->ld;
  DCUDA(cudaMemsetAsync(h->info, 0, sizeof(int), h->stream));
  pick_gram(cov_type)<<<dim3(nt, nt), 256, 0, h->stream>>>(h->coords, n, h->d, h->y, var, range, h->A, ld);
  DCUDA(cudaGetLastError());
  h->launches += 1;
  for (int k = 0; k < nt; ++k) {
    // pivots of this tile that belong to Psi (the response row n and the padding are eliminated, never pivoted)
    const int npiv = std::max(0, std::min(NB, n - k * NB));
    if (npiv == 0) break;
    potrf_tile_kernel<<<1, 256, 0, h->stream>>>(h->A, ld, k, npiv, h->info);
    if (k + 1 < nt) {
      trsm_tile_kernel<<<nt - k - 1, NB, 0, h->stream>>>(h->A, ld, k, npiv);
      const int m = nt - k - 1;
      syrk_tile_kernel<<<m * (m + 1) / 2, 256, sizeof(double) * 2 * NB * (NB + 1), h->stream>>>(h->A, ld, k, nt);
      h->launches += 2;
    }
    h->launches += 1;
    DCUDA(cudaGetLastError());
  }
  dense_sums_kernel<<<1, 256, 0, h->stream>>>(h->A, ld, n, h->out);
  h->launches += 1;
  int info = 0;
  DCUDA(cudaMemcpyAsync(h->out_host, h->out, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
  DCUDA(cudaMemcpyAsync(&info, h->info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  DCUDA(cudaStreamSynchronize(h->stream));
  out3[0] = h->out_host[0];
  out3[1] = h->out_host[1];
  out3[2] = (double)info;  // non-positive pivots (matrix not positive definite)
  h->factored = true;
  return 0;
}

int gpbdev_dense_yaux(gpbdev_dense_t h, double scale, double* yaux_host) {
  if (!h || !yaux_host) return dfail("gpbdev_dense_yaux: null argument");
  if (!h->factored) return dfail("gpbdev_dense_yaux: call gpbdev_dense_eval first");
  DCUDA(cudaSetDevice(h->device));
  dense_backsolve_kernel<<<1, 1024, 0, h->stream>>>(h->A, h->ld, h->n, scale, h->x);
  DCUDA(cudaGetLastError());
  h->launches += 1;
  DCUDA(cudaMemcpyAsync(yaux_host, h->x, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
  DCUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t gpbdev_dense_launch_count(gpbdev_dense_t h) { return h ? h->launches : 0; }

}  // extern "C"
