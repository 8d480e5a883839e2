// On-device builders for the synthetic benchmark inputs: kNN graph -> CSR with spectral normalisation, and the
// diffusion-process data of the k-step prediction task.  Host-side references: Utils/graphTools.py:516-634 (graph construction),
// :110-149 (normalisation by the largest eigenvalue, kStepPredGRNNs.py:768), Utils/dataTools.py:1259-1319 (x_{t+1} = x_t A + w_t).
// At N = 1e5 the reference's dense matrices / numpy eig are impossible; gated_gcrnns_b200/graphs.py did this with scipy on the host
// (cKDTree + 50 sparse power iterations: ~10 s), here it is a uniform-grid search and power iteration on the GPU.
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace gcrnn {
namespace build {

constexpr int KMAX = 32;

// index of cell (x, y) of a 2^bits x 2^bits grid along a Hilbert curve (same construction as graphs._hilbert_index)
__host__ __device__ inline unsigned hilbert(unsigned x, unsigned y, int bits) {
  unsigned d = 0;
  for (unsigned s = 1u << (bits - 1); s > 0; s >>= 1) {
    const unsigned rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
    d += s * s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
      const unsigned t = x; x = y; y = t;
    }
    x &= s - 1; y &= s - 1;
  }
  return d;
}
__device__ __forceinline__ unsigned cell_of(float v, int G) {
  int c = (int)(v * (float)G);
  return (unsigned)(c < 0 ? 0 : (c >= G ? G - 1 : c));
}

__global__ void cell_key_k(const float* __restrict__ xy, unsigned* __restrict__ key, int* __restrict__ count, int N, int G, int bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned k = hilbert(cell_of(xy[2 * i], G), cell_of(xy[2 * i + 1], G), bits);
  key[i] = k;
  atomicAdd(count + k, 1);
}
// exclusive scan of count[0..n) into start[0..n] by ONE block (n <= 2^20 cells: a few passes of 1024 threads)
__global__ void scan_k(const int* __restrict__ count, int* __restrict__ start, int n) {
  __shared__ int part[1024];
  const int per = (n + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(n, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += count[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (int i = lo; i < hi; ++i) { start[i] = run; run += count[i]; }
  if (threadIdx.x == 1023) start[n] = part[1023];
}
__global__ void scatter_k(const unsigned* __restrict__ key, const int* __restrict__ start, int* __restrict__ cursor, int* __restrict__ sorted, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  sorted[start[key[i]] + atomicAdd(cursor + key[i], 1)] = i;
}
// deterministic order inside a cell: ascending original index (cells hold a handful of points)
__global__ void cell_sort_k(const int* __restrict__ start, int* __restrict__ sorted, int ncell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  for (int i = start[c] + 1; i < start[c + 1]; ++i) {
    const int v = sorted[i];
    int j = i - 1;
    while (j >= start[c] && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; }
    sorted[j + 1] = v;
  }
}
__global__ void inverse_k(const int* __restrict__ sorted, int* __restrict__ rank_of, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) rank_of[sorted[i]] = i;
}

// one thread per point (visited in sorted order): k nearest neighbours by expanding square rings of grid cells.
// nbr / d2: [N][k] indexed by the point's position in the sorted order; neighbours are ORIGINAL indices, ascending distance.
__global__ void knn_k(const float* __restrict__ xy, const int* __restrict__ sorted, const int* __restrict__ start, int* __restrict__ nbr,
                      float* __restrict__ d2out, int N, int k, int G, int bits) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= N) return;
  const int i = sorted[pos];
  const float px = xy[2 * i], py = xy[2 * i + 1];
  const int cx = (int)cell_of(px, G), cy = (int)cell_of(py, G);
  const float cell = 1.f / (float)G;
  float bd[KMAX]; int bi[KMAX];
  int have = 0;
  for (int r = 0; r < G; ++r) {
    for (int dy = -r; dy <= r; ++dy) {
      const int y = cy + dy;
      if (y < 0 || y >= G) continue;
      const int step = (dy == -r || dy == r) ? 1 : 2 * r;          // full row on the ring's top / bottom, else its two ends
      for (int dx = -r; dx <= r; dx += (step > 0 ? step : 1)) {
        const int x = cx + dx;
        if (x < 0 || x >= G) continue;
        const unsigned c = hilbert((unsigned)x, (unsigned)y, bits);
        for (int q = start[c]; q < start[c + 1]; ++q) {
          const int j = sorted[q];
          if (j == i) continue;
          const float ex = xy[2 * j] - px, ey = xy[2 * j + 1] - py;
          const float dd = ex * ex + ey * ey;
          if (have < k || dd < bd[have - 1] || (dd == bd[have - 1] && j < bi[have - 1])) {
            int m = have < k ? have++ : k - 1;
            while (m > 0 && (bd[m - 1] > dd || (bd[m - 1] == dd && bi[m - 1] > j))) { bd[m] = bd[m - 1]; bi[m] = bi[m - 1]; --m; }
            bd[m] = dd; bi[m] = j;
          }
        }
      }
    }
    // every point outside the rings 0..r is at least r * cell away
    if (have == k && bd[k - 1] <= (float)r * cell * (float)r * cell) break;
  }
  for (int m = 0; m < k; ++m) { nbr[(size_t)pos * k + m] = m < have ? bi[m] : -1; d2out[(size_t)pos * k + m] = m < have ? bd[m] : 0.f; }
}
__global__ void kth_sum_k(const float* __restrict__ d2, double* acc, int N, int k) {
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) s += (double)d2[(size_t)i * k + k - 1];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}
// CSR rows in the OUTPUT numbering: row id = reorder ? position in the sorted order : original index; columns likewise, ascending
__global__ void assemble_k(const int* __restrict__ sorted, const int* __restrict__ rank_of, const int* __restrict__ nbr, const float* __restrict__ d2,
                           int* __restrict__ col, float* __restrict__ val, int N, int k, float inv_sigma2, int reorder) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= N) return;
  const int row = reorder ? pos : sorted[pos];
  int c[KMAX]; float w[KMAX];
  for (int m = 0; m < k; ++m) {
    const int j = nbr[(size_t)pos * k + m];
    c[m] = reorder ? rank_of[j] : j;
    w[m] = expf(-d2[(size_t)pos * k + m] * inv_sigma2);
  }
  for (int a = 1; a < k; ++a) {                                    // ascending column order within the row
    const int cv = c[a]; const float wv = w[a];
    int b = a - 1;
    while (b >= 0 && c[b] > cv) { c[b + 1] = c[b]; w[b + 1] = w[b]; --b; }
    c[b + 1] = cv; w[b + 1] = wv;
  }
  for (int m = 0; m < k; ++m) { col[(size_t)row * k + m] = c[m]; val[(size_t)row * k + m] = w[m]; }
}
// y = A v (rows of k entries), z += A^T y by atomics
__global__ void av_k(const int* __restrict__ col, const float* __restrict__ val, const float* __restrict__ v, float* __restrict__ y, int N, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float s = 0.f;
  for (int m = 0; m < k; ++m) s = fmaf(val[(size_t)i * k + m], v[col[(size_t)i * k + m]], s);
  y[i] = s;
}
__global__ void atv_k(const int* __restrict__ col, const float* __restrict__ val, const float* __restrict__ y, float* __restrict__ z, int N, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float yi = y[i];
  for (int m = 0; m < k; ++m) atomicAdd(z + col[(size_t)i * k + m], val[(size_t)i * k + m] * yi);
}
__global__ void sqnorm_k(const float* __restrict__ v, double* acc, int N) {
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) s += (double)v[i] * (double)v[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}
__global__ void scale_k(float* __restrict__ v, float s, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] *= s;
}
__global__ void fill_k(float* __restrict__ v, float s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = s;
}

// out[t+1][r][n] = sum_p val[p] out[t][r][idx[p]] (p over the gather list of n: (x S)[n]) + noise[t][r][n]
__global__ void diffuse_k(const int* __restrict__ ptr, const int* __restrict__ idx, const float* __restrict__ val, const float* __restrict__ in,
                          const float* __restrict__ noise, float* __restrict__ out, long long R, int N) {
  const long long total = R * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / N; const int n = (int)(e - r * N);
    const float* row = in + r * N;
    float s = noise ? noise[e] : 0.f;
    for (int p = ptr[n]; p < ptr[n + 1]; ++p) s = fmaf(val[p], row[idx[p]], s);
    out[e] = s;
  }
}

struct DevBuf {
  std::vector<void*> ptrs;
  template <class T> T* get(size_t n) { void* p = nullptr; CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); ptrs.push_back(p); return (T*)p; }
  ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
};

}  // namespace build
}  // namespace gcrnn

using namespace gcrnn;
using namespace gcrnn::build;

extern "C" int gcrnn_build_knn_csr(int32_t N, int32_t k, const float* xy_dev, float sigma2, int32_t power_iters, int32_t reorder,
                                   int32_t device, int64_t* rowptr, int32_t* colidx, float* vals, int32_t* perm, float* sigma2_out,
                                   float* lambda_out) {
  try {
    GCRNN_CHECK(N > 1 && k >= 1 && k <= KMAX && k < N && xy_dev && rowptr && colidx && vals, "bad arguments (1 <= k <= %d, k < N)", KMAX);
    DeviceScope dev(device);
    cudaStream_t st = nullptr;
    int bits = 1;
    while ((1 << bits) * (1 << bits) * 2 < N && bits < 10) ++bits;          // ~2 points per cell, at most 1024 x 1024 cells
    const int G = 1 << bits, ncell = G * G;
    const int TPB = 256, nb = (N + TPB - 1) / TPB;
    DevBuf d;
    unsigned* key = d.get<unsigned>(N);
    int* count = d.get<int>(ncell); int* start = d.get<int>(ncell + 1); int* cursor = d.get<int>(ncell);
    int* sorted = d.get<int>(N); int* rank_of = d.get<int>(N);
    int* nbr = d.get<int>((size_t)N * k); float* d2 = d.get<float>((size_t)N * k);
    int* col = d.get<int>((size_t)N * k); float* val = d.get<float>((size_t)N * k);
    float* v = d.get<float>(N); float* y = d.get<float>(N); float* z = d.get<float>(N);
    double* acc = d.get<double>(2);
    CUDA_OK(cudaMemsetAsync(count, 0, ncell * sizeof(int), st));
    CUDA_OK(cudaMemsetAsync(cursor, 0, ncell * sizeof(int), st));
    cell_key_k<<<nb, TPB, 0, st>>>(xy_dev, key, count, N, G, bits);
    scan_k<<<1, 1024, 0, st>>>(count, start, ncell);
    scatter_k<<<nb, TPB, 0, st>>>(key, start, cursor, sorted, N);
    cell_sort_k<<<(ncell + TPB - 1) / TPB, TPB, 0, st>>>(start, sorted, ncell);
    inverse_k<<<nb, TPB, 0, st>>>(sorted, rank_of, N);
    knn_k<<<(N + 127) / 128, 128, 0, st>>>(xy_dev, sorted, start, nbr, d2, N, k, G, bits);
    count_launch(6);
    if (!(sigma2 > 0.f)) {                                          // mean squared distance to the k-th neighbour (graphs.knn_csr)
      CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), st));
      kth_sum_k<<<148, 256, 0, st>>>(d2, acc, N, k);
      count_launch();
      double s = 0.0;
      CUDA_OK(cudaMemcpy(&s, acc, sizeof(double), cudaMemcpyDeviceToHost));
      sigma2 = (float)(s / N);
    }
    assemble_k<<<nb, TPB, 0, st>>>(sorted, rank_of, nbr, d2, col, val, N, k, 1.f / sigma2, reorder);
    count_launch();
    // spectral norm: power iteration on A^T A, lambda = sqrt(|A^T A v| / |v|) as in graphs.knn_csr
    float lam = 1.f;
    if (power_iters > 0) {
      fill_k<<<nb, TPB, 0, st>>>(v, 1.f / std::sqrt((float)N), N);
      for (int it = 0; it < power_iters; ++it) {
        av_k<<<nb, TPB, 0, st>>>(col, val, v, y, N, k);
        CUDA_OK(cudaMemsetAsync(z, 0, N * sizeof(float), st));
        atv_k<<<nb, TPB, 0, st>>>(col, val, y, z, N, k);
        CUDA_OK(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
        sqnorm_k<<<148, 256, 0, st>>>(z, acc, N);
        sqnorm_k<<<148, 256, 0, st>>>(v, acc + 1, N);
        double h[2];
        CUDA_OK(cudaMemcpy(h, acc, 2 * sizeof(double), cudaMemcpyDeviceToHost));
        const double nz = std::sqrt(h[0]), nv = std::sqrt(h[1]);
        lam = (float)std::sqrt(nz / std::max(nv, 1e-30));
        CUDA_OK(cudaMemcpyAsync(v, z, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
        scale_k<<<148, 256, 0, st>>>(v, (float)(1.0 / std::max(nz, 1e-30)), N);
        count_launch(5);
      }
      scale_k<<<148 * 4, 256, 0, st>>>(val, 1.f / lam, (long long)N * k);
      count_launch(2);
    }
    CUDA_OK(cudaGetLastError());
    std::vector<int> hcol((size_t)N * k);
    CUDA_OK(cudaMemcpy(hcol.data(), col, hcol.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(vals, val, (size_t)N * k * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < hcol.size(); ++i) { GCRNN_CHECK(hcol[i] >= 0 && hcol[i] < N, "kNN search left a row short of neighbours"); colidx[i] = hcol[i]; }
    for (int i = 0; i <= N; ++i) rowptr[i] = (int64_t)i * k;
    if (perm) {                                                     // perm[new] = old (identity when reorder == 0)
      if (reorder) CUDA_OK(cudaMemcpy(perm, sorted, N * sizeof(int), cudaMemcpyDeviceToHost));
      else for (int i = 0; i < N; ++i) perm[i] = i;
    }
    if (sigma2_out) *sigma2_out = sigma2;
    if (lambda_out) *lambda_out = lam;
  } catch (const gcrnn::Error& e) { set_last_error("%s", e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error("%s", e.what()); return -1; }
  return 0;
}

extern "C" int gcrnn_data_diffusion(const gcrnn_graph* g, const float* x0, const float* noise, float* out, int64_t R, int32_t T, void* stream) {
  try {
    GCRNN_CHECK(g && x0 && out && R > 0 && T >= 0, "bad arguments");
    GCRNN_CHECK(g->E == 1, "diffusion data needs one shift operator (E == 1)");
    DeviceScope dev(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long RN = R * g->N;
    CUDA_OK(cudaMemcpyAsync(out, x0, RN * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const Gather& fw = g->fwd[0];                                  // (x S)[n] = sum over the gather list of n
    for (int t = 0; t < T; ++t) {
      diffuse_k<<<(unsigned)std::min<long long>((RN + 255) / 256, 148 * 16), 256, 0, st>>>(fw.ptr, fw.idx, fw.val, out + (size_t)t * RN,
                                                                                          noise ? noise + (size_t)t * RN : nullptr,
                                                                                          out + (size_t)(t + 1) * RN, R, g->N);
      count_launch();
    }
    CUDA_OK(cudaGetLastError());
  } catch (const gcrnn::Error& e) { set_last_error("%s", e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error("%s", e.what()); return -1; }
  return 0;
}
