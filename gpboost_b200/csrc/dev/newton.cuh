// Newton update of the leaf values in GPBoost, Vecchia GP with a Gaussian likelihood.
// Included at the end of dev_api.cu (shares the engine struct and the stored factor).
//
// Replaces REModelTemplate::NewtonUpdateLeafValues, Vecchia branch (include/GPBoost/re_model_template.h:4982-5063; called from
// GBDT::TrainOneIter, src/LightGBM/boosting/gbdt.cpp:470-478): with H the n x L leaf incidence matrix of the new tree,
//     leaf values = (H^T Psi^-1 H)^-1 H^T Psi^-1 (y - F),      Psi^-1 = B^T D^-1 B  (transformed scale).
// The reference forms the sparse product B H and then (BH)^T D^-1 (BH). Here row i of B H is built on the fly from the resident
// factor — +1 at leaf(i), -A[i,k] at leaf(nn[i,k]) — and its weighted outer product is accumulated into 64 x 64 tiles of the
// L x L Gram matrix in shared memory; chunk partials are summed in chunk order (deterministic). The right-hand side is the
// per-leaf sum of the gradient Psi^-1 (F - y) / sigma^2 the boosting loop already holds (HTYAux, :5008). The L x L solve (L <= 256)
// is done by the caller on the host.
// HBM traffic: one pass over A and nn ((8 + 4) m bytes per row) per tile pair; leaf ids (4 MB at n = 1e6) stay in L2.
namespace gpn {

constexpr int kTile = 64;
constexpr int kWarps = 8;
constexpr int kMaxLeaves = 256;

// partial[chunk][pair][64 x 64]: pair = (ta, tb), ta >= tb, of the lower block triangle
__global__ void __launch_bounds__(kWarps * 32) newton_gram_kernel(const double* __restrict__ A, const int32_t* __restrict__ nn,
                                                                  const double* __restrict__ Dinv, const int32_t* __restrict__ perm,
                                                                  const int32_t* __restrict__ leaf_of_row, int m, int64_t n, int L,
                                                                  int64_t rows_per_chunk, double* __restrict__ partial) {
  __shared__ double M[kTile * kTile];
  __shared__ double ra[kWarps][kTile], rb[kWarps][kTile];  // the row of B H restricted to the leaf ranges of the two tiles
  __shared__ double dw[kWarps];
  __shared__ int sleaf[kWarps][32];
  __shared__ double scoef[kWarps][32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int chunk = blockIdx.x;
  // tile pair of this CTA
  int ta = 0, tb = 0;
  { int p = blockIdx.y; while (p > ta) { p -= ta + 1; ++ta; } tb = p; }
  for (int e = tid; e < kTile * kTile; e += blockDim.x) M[e] = 0.;
  const int64_t r0 = (int64_t)chunk * rows_per_chunk, r1 = min(r0 + rows_per_chunk, n);
  for (int64_t base = r0; base < r1; base += kWarps) {
    const int64_t i = base + w;
    // ---- row i of B H restricted to the leaves of tile ta (ra) and tile tb (rb): zero, then the entries in neighbour order
    // (one lane: fixed order of additions)
    for (int l = lane; l < kTile; l += 32) { ra[w][l] = 0.; rb[w][l] = 0.; }
    if (i < r1) {
      const int32_t j = lane < m ? nn[i * m + lane] : -1;
      sleaf[w][lane] = j >= 0 ? leaf_of_row[perm[j]] : -1;
      scoef[w][lane] = j >= 0 ? -A[i * m + lane] : 0.;
    }
    __syncwarp();
    if (lane == 0) {
      if (i < r1) {
        auto add = [&](int l, double c) {
          if (l / kTile == ta) ra[w][l % kTile] += c;
          if (l / kTile == tb) rb[w][l % kTile] += c;
        };
        add(leaf_of_row[perm[i]], 1.);
        for (int k = 0; k < m; ++k) { const int l = sleaf[w][k]; if (l >= 0) add(l, scoef[w][k]); }
        dw[w] = Dinv[i];
      } else {
        dw[w] = 0.;
      }
    }
    __syncthreads();
    // ---- M[a][b] += sum_w D_w r_w[a] r_w[b] over this batch, warps in order
    for (int e = tid; e < kTile * kTile; e += blockDim.x) {
      const int a = ta * kTile + e / kTile, b = tb * kTile + e % kTile;
      if (a < L && b < L) {
        double acc = M[e];
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) acc += dw[ww] * ra[ww][e / kTile] * rb[ww][e % kTile];
        M[e] = acc;
      }
    }
    __syncthreads();
  }
  double* out = partial + ((size_t)chunk * gridDim.y + blockIdx.y) * kTile * kTile;
  for (int e = tid; e < kTile * kTile; e += blockDim.x) out[e] = M[e];
}

// M[a][b] (row-major L x L, both triangles) = sum over chunks, chunk order
__global__ void newton_reduce_kernel(const double* __restrict__ partial, int nchunks, int npairs, int L, double* __restrict__ Mout) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L * L) return;
  int a = e / L, b = e % L;
  if (b > a) { const int t = a; a = b; b = t; }  // lower triangle holds the value
  const int ta = a / kTile, tb = b / kTile;
  const int pair = ta * (ta + 1) / 2 + tb;
  int ia = a % kTile, ib = b % kTile;
  // inside a diagonal tile both orders were accumulated; off-diagonal tiles are (ta > tb): a along rows
  double s = 0.;
  for (int c = 0; c < nchunks; ++c) s += partial[((size_t)c * npairs + pair) * kTile * kTile + ia * kTile + ib];
  Mout[e] = s;
}

// rhs[l] = sum of g over the rows of leaf l (original row order), fixed order: thread = leaf, chunks then rows ascending
__global__ void newton_leafsum_kernel(const double* __restrict__ g, const int32_t* __restrict__ leaf_of_row, int64_t n, int L,
                                      int64_t rows_per_chunk, double* __restrict__ partial) {
  const int l = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(r0 + rows_per_chunk, n);
  double s = 0.;
  if (l < L)
    for (int64_t i = r0; i < r1; ++i) s += leaf_of_row[i] == l ? g[i] : 0.;
  if (l < L) partial[(size_t)blockIdx.x * kMaxLeaves + l] = s;
}
__global__ void newton_leafsum_reduce_kernel(const double* __restrict__ partial, int nchunks, int L, double* __restrict__ out) {
  const int l = threadIdx.x;
  if (l >= L) return;
  double s = 0.;
  for (int c = 0; c < nchunks; ++c) s += partial[(size_t)c * kMaxLeaves + l];
  out[l] = s;
}

}  // namespace gpn

extern "C" {

// After gpbdev_vecchia_eval(mode = STORE) at the current covariance parameters (what the gradient computation leaves behind):
// M_host (L x L, row-major) = H^T B^T D^-1 B H, rhs_host (L) = H^T g, with H from leaf_of_row_dev (n int32, ORIGINAL row order)
// and g = grad_dev (n doubles, original order). Replaces re_model_template.h:4996-5012 (HTPsiInvH, HTYAux before the sign/scale).
int gpbdev_vecchia_newton_system(gpbdev_vecchia_t h, const int32_t* leaf_of_row_dev, int num_leaves, const double* grad_dev,
                                 double* M_host, double* rhs_host) {
  if (!h || !leaf_of_row_dev || !grad_dev || !M_host || !rhs_host) return fail("gpbdev_vecchia_newton_system: null argument");
  if (num_leaves < 1 || num_leaves > gpn::kMaxLeaves) return fail("gpbdev_vecchia_newton_system: num_leaves must be in [1, 256]");
  if (!h->factor_stored) return fail("gpbdev_vecchia_newton_system: the factor is not resident (gpbdev_vecchia_eval with mode = STORE first)");
  if (h->row_begin != 0 || h->row_end != h->n) return fail("gpbdev_vecchia_newton_system: row-sharded engines are not supported yet");
  if (h->m > 32) return fail("gpbdev_vecchia_newton_system: num_neighbors must be <= 32");
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t n = h->n;
  const int L = num_leaves;
  const int nt = (L + gpn::kTile - 1) / gpn::kTile, npairs = nt * (nt + 1) / 2;
  const int nchunks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)std::max(1, h->num_sms * 2 / npairs));
  const int64_t rpc = (n + nchunks - 1) / nchunks;
  double *partial = nullptr, *Mdev = nullptr, *lpart = nullptr, *rdev = nullptr;
  auto release = [&]() { cudaFree(partial); cudaFree(Mdev); cudaFree(lpart); cudaFree(rdev); };
  cudaError_t e = cudaMalloc(&partial, sizeof(double) * (size_t)nchunks * npairs * gpn::kTile * gpn::kTile);
  if (e == cudaSuccess) e = cudaMalloc(&Mdev, sizeof(double) * L * L);
  const int lchunks = h->num_sms * 4;
  const int64_t lrpc = (n + lchunks - 1) / lchunks;
  if (e == cudaSuccess) e = cudaMalloc(&lpart, sizeof(double) * (size_t)lchunks * gpn::kMaxLeaves);
  if (e == cudaSuccess) e = cudaMalloc(&rdev, sizeof(double) * L);
  if (e != cudaSuccess) { release(); return fail(std::string("gpbdev_vecchia_newton_system: ") + cudaGetErrorString(e)); }
  gpn::newton_gram_kernel<<<dim3(nchunks, npairs), gpn::kWarps * 32, 0, h->stream>>>(h->A, h->nn, h->Dinv, h->perm, leaf_of_row_dev, h->m, n, L, rpc, partial);
  gpn::newton_reduce_kernel<<<(L * L + 255) / 256, 256, 0, h->stream>>>(partial, nchunks, npairs, L, Mdev);
  gpn::newton_leafsum_kernel<<<lchunks, gpn::kMaxLeaves, 0, h->stream>>>(grad_dev, leaf_of_row_dev, n, L, lrpc, lpart);
  gpn::newton_leafsum_reduce_kernel<<<1, gpn::kMaxLeaves, 0, h->stream>>>(lpart, lchunks, L, rdev);
  e = cudaGetLastError();
  h->launches += 4;
  if (e == cudaSuccess) e = cudaMemcpyAsync(M_host, Mdev, sizeof(double) * L * L, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rhs_host, rdev, sizeof(double) * L, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  release();
  if (e != cudaSuccess) return fail(std::string("gpbdev_vecchia_newton_system: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
