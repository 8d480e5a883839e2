// Persistent fused recurrence for SMALL graphs (fp32 exact): one CTA runs the WHOLE sequence of one sample with the shift
// operator's gather lists, every filter tap, the time-gate weights and the state h_t resident on chip across all T steps.
//
// This is the north star's "persistent fused forward kernel" (and its matching reverse-time kernel) at the sizes where it fits
// one SM: the reference's own configurations (kStepPredGRNNs.py: N = 80, F = 20, K = 5, T = 5, B = 100).  There a forward +
// backward through the per-op kernels is ~250 dependent launches of a few microseconds each — launch- and latency-bound even
// when replayed as a CUDA graph (76 k sequences/s); here it is TWO launches.
//
//   forward  (Utils/graphML.py:2336-2428): per step  z_k = z_{k-1} S (gathers from shared memory),  a = A(S)x_t + b,
//            r = B(S)h_{t-1} + b, time gates g = sigmoid(W_g . tanh(A_g(S)x_t + B_g(S)h0 + 2 b_g) + c_g) (block reduction),
//            h_t = tanh(g_i a + g_f r) -> H[b, t] and back into shared memory as the next step's chain input.
//   backward: reverse sweep per sample, recomputing the step's filter outputs from H (the output itself) and the saved gate
//            values: dpre = (dH_t + dh)(1 - h_t^2); weight gradients as reductions over the nodes accumulated in shared memory
//            across all steps (one atomicAdd per parameter and CTA at the end); dh_{t-1} by Horner's rule with S^T;
//            gate path: d logit -> dW_g, dc_g, dA_g, and the T-invariant term dc0 -> dB_g, db_g, dh0 once per sequence.
//
// Supported: E = 1, no spatial gating, time gating on or off, no dX (the reference never asks for it: train_rnn.py:256);
// sizes such that everything fits the 227 KB of one SM (persist_smem_floats).  Everything else takes the per-op kernels.
#pragma once
#include "common.cuh"

namespace gcrnn {
namespace persist {

constexpr int PT = 512;       // threads per CTA

struct Args {
  int N, F, G, Kin, Kst, tg, has_bias;
  long long B, T;
  const int *cptr, *cidx; const float* cval;     // gather form of z @ S   (CSC of S):  out[n] = sum_p cval[p] in[cidx[p]]
  const int *rptr, *ridx; const float* rval;     // gather form of g @ S^T (CSR of S)
  int nnz;                                       // entries of S
  int lists_smem;                                // 1: both gather lists are staged in shared memory (16-bit indices)
  const float *A, *Bw, *bias;                    // [F,Kin,G] [F,Kst,F] [F]
  const float *tA[2], *tB[2], *tb[2], *tW[2], *tc[2];
  const float *X, *h0;                           // [B,T,G,N] [B,F,N]
  float* H;                                      // [B,T,F,N]  (forward: out; backward: in)
  float* gt;                                     // [2][B][T] time-gate values (forward: out; backward: in)
  // backward
  const float* dH; long long dH_bstride, dH_tstride; int dh_last_only;   // dH[b,t] = dH + b*bstride + t*tstride ([F][N]); last_only: zero for t < T-1
  float *dA, *dBw, *dbias, *dtA[2], *dtB[2], *dtb[2], *dtW[2], *dtc[2];
  float* dh0;                                    // [B,F,N] or null
};

// shared-memory floats of the two kernels (same formula on host and device)
__host__ __device__ inline long long weights_floats(int F, int G, int Kin, int Kst, int N, int tg) {
  const long long cell = (long long)F * Kin * G + (long long)F * Kst * F + F;
  return cell + (tg ? 2 * cell + 2LL * F * N : 0);
}
__host__ __device__ inline long long fwd_floats(int F, int G, int Kin, int Kst, int N, int tg) {
  return weights_floats(F, G, Kin, Kst, N, tg) + (long long)Kin * G * N + (long long)Kst * F * N + (long long)F * N /*hn*/ +
         (tg ? 2LL * F * N : 0) /*c0*/ + 64;
}
__host__ __device__ inline long long bwd_floats(int F, int G, int Kin, int Kst, int N, int tg) {
  const long long FN = (long long)F * N;
  return weights_floats(F, G, Kin, Kst, N, tg) /*weights*/ + weights_floats(F, G, Kin, Kst, N, tg) /*gradient accumulators*/ +
         (long long)Kin * G * N + (long long)Kst * F * N + 5 * FN /*da dr dh b1 b2*/ + (tg ? 3 * FN : 0) /*c0 x2... see kernel*/ +
         (tg ? 2 * FN : 0) + 64;
}

// shared-memory bytes of ONE staged gather list: ptr[N+1] (int), val[nnz] (float), idx[nnz] (u16), 16-byte aligned pieces
__host__ __device__ inline long long list_bytes(int N, int nnz) {
  return (((long long)(N + 1) * 4 + 15) & ~15LL) + (((long long)nnz * 4 + 15) & ~15LL) + (((long long)nnz * 2 + 15) & ~15LL);
}
struct List {                                    // a gather list, in shared memory (idx16) or global memory (idx32)
  const int* ptr; const float* val; const unsigned short* idx16; const int* idx32;
  __device__ __forceinline__ int idx(int p) const { return idx16 ? (int)idx16[p] : __ldg(idx32 + p); }
};
__device__ __forceinline__ List stage_list(unsigned char*& sp, const int* gptr, const int* gidx, const float* gval, int N, int nnz, bool to_smem) {
  List l;
  if (!to_smem) { l.ptr = gptr; l.val = gval; l.idx16 = nullptr; l.idx32 = gidx; return l; }
  int* sptr = reinterpret_cast<int*>(sp); sp += ((N + 1) * 4 + 15) & ~15;
  float* sval = reinterpret_cast<float*>(sp); sp += (nnz * 4 + 15) & ~15;
  unsigned short* sidx = reinterpret_cast<unsigned short*>(sp); sp += (nnz * 2 + 15) & ~15;
  for (int i = threadIdx.x; i <= N; i += PT) sptr[i] = gptr[i];
  for (int i = threadIdx.x; i < nnz; i += PT) { sval[i] = gval[i]; sidx[i] = (unsigned short)gidx[i]; }
  l.ptr = sptr; l.val = sval; l.idx16 = sidx; l.idx32 = nullptr;
  return l;
}

__device__ __forceinline__ float block_sum(float v, float* red) {      // red: >= 33 floats; result broadcast to all threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = l < PT / 32 ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// out[r][n] = sum_p val[p] in[r][idx[p]], p in [ptr[n], ptr[n+1])   (rows r < R; in / out in shared memory, [R][N])
__device__ __forceinline__ void shift(const List& l, const float* in, float* out, int R, int N) {
  for (int e = threadIdx.x; e < R * N; e += PT) {
    const int r = e / N, n = e - r * N;
    const float* row = in + r * N;
    float s0 = 0.f, s1 = 0.f;                                     // two chains: the loop is latency-bound on the dependent FMA
    const int p1 = l.ptr[n + 1];
    int p = l.ptr[n];
    for (; p + 1 < p1; p += 2) { s0 = fmaf(l.val[p], row[l.idx(p)], s0); s1 = fmaf(l.val[p + 1], row[l.idx(p + 1)], s1); }
    if (p < p1) s0 = fmaf(l.val[p], row[l.idx(p)], s0);
    out[e] = s0 + s1;
  }
}
// z[k] = z[k-1] S for k = 1..K-1, z: [K][R][N]
__device__ __forceinline__ void chain(const List& fw, float* z, int K, int R, int N) {
  for (int k = 1; k < K; ++k) {
    shift(fw, z + (size_t)(k - 1) * R * N, z + (size_t)k * R * N, R, N);
    __syncthreads();
  }
}
// y[f][n] = sum_{k,g} W[f][k][g] z[k][g][n]   for one (f, n)
__device__ __forceinline__ float contract(const float* W, const float* z, int f, int n, int K, int C, int N) {
  const float* w = W + (size_t)f * K * C;
  const float* zc = z + n;
  const int KC = K * C;                                           // rows (k, g) of z are contiguous: z[(k*C + g)*N + n]
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;                   // four independent chains
  int i = 0;
  for (; i + 3 < KC; i += 4) {
    s0 = fmaf(w[i], zc[(size_t)i * N], s0);
    s1 = fmaf(w[i + 1], zc[(size_t)(i + 1) * N], s1);
    s2 = fmaf(w[i + 2], zc[(size_t)(i + 2) * N], s2);
    s3 = fmaf(w[i + 3], zc[(size_t)(i + 3) * N], s3);
  }
  for (; i < KC; ++i) s0 = fmaf(w[i], zc[(size_t)i * N], s0);
  return (s0 + s1) + (s2 + s3);
}

struct Weights { float *A, *Bw, *bias, *tA[2], *tB[2], *tb[2], *tW[2]; };
__device__ __forceinline__ float* carve(Weights& w, float* p, const Args& a) {
  const int nA = a.F * a.Kin * a.G, nB = a.F * a.Kst * a.F;
  w.A = p; p += nA; w.Bw = p; p += nB; w.bias = p; p += a.F;
  for (int g = 0; g < 2; ++g) {
    if (a.tg) { w.tA[g] = p; p += nA; w.tB[g] = p; p += nB; w.tb[g] = p; p += a.F; w.tW[g] = p; p += a.F * a.N; }
    else { w.tA[g] = w.tB[g] = w.tb[g] = w.tW[g] = nullptr; }
  }
  return p;
}
__device__ __forceinline__ void load_weights(const Weights& w, const Args& a) {
  const int nA = a.F * a.Kin * a.G, nB = a.F * a.Kst * a.F, FN = a.F * a.N;
  for (int i = threadIdx.x; i < nA; i += PT) w.A[i] = a.A[i];
  for (int i = threadIdx.x; i < nB; i += PT) w.Bw[i] = a.Bw[i];
  for (int i = threadIdx.x; i < a.F; i += PT) w.bias[i] = a.has_bias ? a.bias[i] : 0.f;
  if (a.tg)
    for (int g = 0; g < 2; ++g) {
      for (int i = threadIdx.x; i < nA; i += PT) w.tA[g][i] = a.tA[g][i];
      for (int i = threadIdx.x; i < nB; i += PT) w.tB[g][i] = a.tB[g][i];
      for (int i = threadIdx.x; i < a.F; i += PT) w.tb[g][i] = a.has_bias ? a.tb[g][i] : 0.f;
      for (int i = threadIdx.x; i < FN; i += PT) w.tW[g][i] = a.tW[g][i];
    }
}

__global__ void __launch_bounds__(PT, 1) persist_fwd_k(const Args a) {
  extern __shared__ __align__(16) float psm[];
  const int N = a.N, F = a.F, FN = F * N, GN = a.G * N;
  const long long b = blockIdx.x;
  Weights w;
  float* p = carve(w, psm, a);
  float* zx = p; p += (size_t)a.Kin * GN;
  float* zh = p; p += (size_t)a.Kst * FN;
  float* hn = p; p += FN;
  float* c0 = p; p += a.tg ? 2 * FN : 0;
  float* red = p; p += 64;
  unsigned char* sp = reinterpret_cast<unsigned char*>(p);
  const List fw = stage_list(sp, a.cptr, a.cidx, a.cval, N, a.nnz, a.lists_smem != 0);
  load_weights(w, a);
  for (int e = threadIdx.x; e < FN; e += PT) zh[e] = a.h0[b * FN + e];
  __syncthreads();
  if (a.tg) {                                                     // T-invariant gate term: B_g(S) h0 + 2 b_g   (graphML.py:2362, :2417-2423)
    chain(fw, zh, a.Kst, F, N);
    for (int g = 0; g < 2; ++g)
      for (int e = threadIdx.x; e < FN; e += PT) {
        const int f = e / N, n = e - f * N;
        c0[g * FN + e] = contract(w.tB[g], zh, f, n, a.Kst, F, N) + 2.f * w.tb[g][f];
      }
    __syncthreads();
  }
  for (long long t = 0; t < a.T; ++t) {
    const float* xt = a.X + (b * a.T + t) * GN;
    for (int e = threadIdx.x; e < GN; e += PT) zx[e] = xt[e];
    __syncthreads();
    chain(fw, zx, a.Kin, a.G, N);
    float gi = 1.f, gf = 1.f;
    if (a.tg) {
      for (int g = 0; g < 2; ++g) {
        float part = 0.f;
        for (int e = threadIdx.x; e < FN; e += PT) {
          const int f = e / N, n = e - f * N;
          const float u = tanhf(contract(w.tA[g], zx, f, n, a.Kin, a.G, N) + c0[g * FN + e]);
          part = fmaf(w.tW[g][e], u, part);
        }
        const float logit = block_sum(part, red) + (a.has_bias ? __ldg(a.tc[g]) : 0.f);
        const float gv = 1.f / (1.f + expf(-logit));
        if (g == 0) gi = gv; else gf = gv;
        if (threadIdx.x == 0) a.gt[((long long)g * a.B + b) * a.T + t] = gv;
      }
    }
    if (!(a.tg && t == 0)) chain(fw, zh, a.Kst, F, N);                // at t = 0 with gating the h0 chain is already there
    float* Ht = a.H + (b * a.T + t) * FN;
    for (int e = threadIdx.x; e < FN; e += PT) {
      const int f = e / N, n = e - f * N;
      const float av = contract(w.A, zx, f, n, a.Kin, a.G, N) + w.bias[f];
      const float rv = contract(w.Bw, zh, f, n, a.Kst, F, N) + w.bias[f];           // the same bias in both filters (:2405-2407)
      const float h = tanhf(fmaf(gi, av, gf * rv));
      Ht[e] = h;
      hn[e] = h;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < FN; e += PT) zh[e] = hn[e];
    __syncthreads();
  }
}

// acc[f,k,g] += sum_n d[f][n] z[k][g][n]   (every output owned by one thread; acc in shared memory)
__device__ __forceinline__ void wgrad_acc(float* acc, const float* d, const float* z, int F, int K, int C, int N) {
  for (int o = threadIdx.x; o < F * K * C; o += PT) {
    const int f = o / (K * C), kc = o - f * (K * C);
    const float* dr = d + (size_t)f * N;
    const float* zr = z + (size_t)kc * N;
    float s0 = 0.f, s1 = 0.f;
    int nn = threadIdx.x % N;                                      // skewed start: the lanes of a warp read different banks
    int n = 0;
    for (; n + 1 < N; n += 2) {
      s0 = fmaf(dr[nn], zr[nn], s0); if (++nn == N) nn = 0;
      s1 = fmaf(dr[nn], zr[nn], s1); if (++nn == N) nn = 0;
    }
    if (n < N) s0 = fmaf(dr[nn], zr[nn], s0);
    acc[o] += s0 + s1;
  }
}
// dh[g][n] (+)= Horner over k of ( sum_f W[f][k][g] d[f][n] ) with S^T:  out = u_0 + (u_1 + (... u_{K-1} S^T ...) S^T) S^T
// b1 / b2: [C][N] scratch; result ADDED to `out` if accumulate else written
__device__ __forceinline__ void adjoint_chain(const Args& a, const List& bw, const float* W, const float* d, float* b1, float* b2, float* out,
                                              int K, int C, bool accumulate) {
  const int N = a.N, F = a.F;
  float* cur = b1; float* nxt = b2;
  for (int k = K - 1; k >= 0; --k) {
    // nxt[g][n] = (cur S^T)[g][n] (if k < K-1) + sum_f W[f][k][g] d[f][n]
    for (int e = threadIdx.x; e < C * N; e += PT) {
      const int g = e / N, n = e - g * N;
      float s = 0.f;
      if (k < K - 1) {
        const float* row = cur + (size_t)g * N;
        const int p1 = bw.ptr[n + 1];
        for (int p = bw.ptr[n]; p < p1; ++p) s = fmaf(bw.val[p], row[bw.idx(p)], s);
      }
      float t0 = 0.f, t1 = 0.f;
      int f = 0;
      for (; f + 1 < F; f += 2) {
        t0 = fmaf(W[((size_t)f * K + k) * C + g], d[(size_t)f * N + n], t0);
        t1 = fmaf(W[((size_t)(f + 1) * K + k) * C + g], d[(size_t)(f + 1) * N + n], t1);
      }
      if (f < F) t0 = fmaf(W[((size_t)f * K + k) * C + g], d[(size_t)f * N + n], t0);
      s += t0 + t1;
      if (k == 0) { if (accumulate) out[e] += s; else out[e] = s; }
      else nxt[e] = s;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

__global__ void __launch_bounds__(PT, 1) persist_bwd_k(const Args a) {
  extern __shared__ __align__(16) float psm[];
  const int N = a.N, F = a.F, FN = F * N, GN = a.G * N;
  const int nA = F * a.Kin * a.G, nB = F * a.Kst * F;
  const long long b = blockIdx.x;
  Weights w, gacc;
  float* p = carve(w, psm, a);
  p = carve(gacc, p, a);                                          // gradient accumulators, same layout as the weights
  float* zx = p; p += (size_t)a.Kin * GN;
  float* zh = p; p += (size_t)a.Kst * FN;
  float* da = p; p += FN;
  float* dr = p; p += FN;
  float* dh = p; p += FN;
  float* b1 = p; p += FN;
  float* b2 = p; p += FN;
  float* c0 = p; p += a.tg ? 2 * FN : 0;
  float* dc0 = p; p += a.tg ? 2 * FN : 0;
  float* dpu = p; p += a.tg ? FN : 0;
  float* red = p; p += 64;
  unsigned char* sp = reinterpret_cast<unsigned char*>(p);
  const List fw = stage_list(sp, a.cptr, a.cidx, a.cval, N, a.nnz, a.lists_smem != 0);
  const List bw = stage_list(sp, a.rptr, a.ridx, a.rval, N, a.nnz, a.lists_smem != 0);
  load_weights(w, a);
  for (float* q = gacc.A; q < zx; q += PT) { if (q + threadIdx.x < zx) q[threadIdx.x] = 0.f; }      // zero every accumulator
  for (int e = threadIdx.x; e < FN; e += PT) { dh[e] = 0.f; if (a.tg) { dc0[e] = 0.f; dc0[FN + e] = 0.f; } }
  float dtc[2] = {0.f, 0.f};
  __syncthreads();
  if (a.tg) {                                                     // c0 of both gates (needed to recompute u at every step)
    for (int e = threadIdx.x; e < FN; e += PT) zh[e] = a.h0[b * FN + e];
    __syncthreads();
    chain(fw, zh, a.Kst, F, N);
    for (int g = 0; g < 2; ++g)
      for (int e = threadIdx.x; e < FN; e += PT) {
        const int f = e / N, n = e - f * N;
        c0[g * FN + e] = contract(w.tB[g], zh, f, n, a.Kst, F, N) + 2.f * w.tb[g][f];
      }
    __syncthreads();
  }
  for (long long t = a.T - 1; t >= 0; --t) {
    const float* hprev = t > 0 ? a.H + (b * a.T + t - 1) * FN : a.h0 + b * FN;
    const float* xt = a.X + (b * a.T + t) * GN;
    for (int e = threadIdx.x; e < FN; e += PT) zh[e] = hprev[e];
    for (int e = threadIdx.x; e < GN; e += PT) zx[e] = xt[e];
    __syncthreads();
    chain(fw, zh, a.Kst, F, N);
    chain(fw, zx, a.Kin, a.G, N);
    float gi = 1.f, gf = 1.f;
    if (a.tg) { gi = a.gt[((long long)0 * a.B + b) * a.T + t]; gf = a.gt[((long long)1 * a.B + b) * a.T + t]; }
    const float* Ht = a.H + (b * a.T + t) * FN;
    const bool has_dH = !(a.dh_last_only && t < a.T - 1);
    const float* dHt = a.dH + b * a.dH_bstride + (a.dh_last_only ? 0 : t * a.dH_tstride);
    float sgi = 0.f, sgf = 0.f;
    for (int e = threadIdx.x; e < FN; e += PT) {
      const int f = e / N, n = e - f * N;
      const float av = contract(w.A, zx, f, n, a.Kin, a.G, N) + w.bias[f];
      const float rv = contract(w.Bw, zh, f, n, a.Kst, F, N) + w.bias[f];
      const float h = Ht[e];
      const float dp = ((has_dH ? dHt[e] : 0.f) + dh[e]) * (1.f - h * h);
      sgi = fmaf(dp, av, sgi); sgf = fmaf(dp, rv, sgf);
      da[e] = gi * dp; dr[e] = gf * dp;
    }
    float dgi = 0.f, dgf = 0.f;
    if (a.tg) { dgi = block_sum(sgi, red); dgf = block_sum(sgf, red); }
    __syncthreads();
    wgrad_acc(gacc.A, da, zx, F, a.Kin, a.G, N);
    wgrad_acc(gacc.Bw, dr, zh, F, a.Kst, F, N);
    for (int f = threadIdx.x; f < F; f += PT) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += da[(size_t)f * N + n] + dr[(size_t)f * N + n];
      gacc.bias[f] += s;
    }
    adjoint_chain(a, bw, w.Bw, dr, b1, b2, dh, a.Kst, F, false);      // dh_{t-1} (recurrent part)
    if (a.tg) {
      for (int g = 0; g < 2; ++g) {
        const float gv = g == 0 ? gi : gf;
        const float dl = (g == 0 ? dgi : dgf) * gv * (1.f - gv);
        dtc[g] += dl;
        for (int e = threadIdx.x; e < FN; e += PT) {
          const int f = e / N, n = e - f * N;
          const float u = tanhf(contract(w.tA[g], zx, f, n, a.Kin, a.G, N) + c0[g * FN + e]);
          gacc.tW[g][e] += dl * u;
          const float d = dl * w.tW[g][e] * (1.f - u * u);
          dc0[g * FN + e] += d;
          dpu[e] = d;
        }
        __syncthreads();
        wgrad_acc(gacc.tA[g], dpu, zx, F, a.Kin, a.G, N);
        __syncthreads();
      }
    }
  }
  // ---- T-invariant gate term: after t = 0 the zh buffers hold h0's chain ------------------------------------------------------
  if (a.tg) {
    for (int g = 0; g < 2; ++g) {
      const float* v = dc0 + (size_t)g * FN;
      wgrad_acc(gacc.tB[g], v, zh, F, a.Kst, F, N);
      for (int f = threadIdx.x; f < F; f += PT) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += v[(size_t)f * N + n];
        gacc.tb[g][f] += 2.f * s;                                  // the sub-cell adds its bias twice (:2421-2422)
      }
      if (a.dh0) adjoint_chain(a, bw, w.tB[g], v, b1, b2, dh, a.Kst, F, true);
      __syncthreads();
    }
  }
  if (a.dh0) for (int e = threadIdx.x; e < FN; e += PT) a.dh0[b * FN + e] = dh[e];
  // ---- one atomicAdd per parameter and CTA ---------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < nA; i += PT) if (a.dA) atomicAdd(a.dA + i, gacc.A[i]);
  for (int i = threadIdx.x; i < nB; i += PT) if (a.dBw) atomicAdd(a.dBw + i, gacc.Bw[i]);
  for (int i = threadIdx.x; i < F; i += PT) if (a.dbias) atomicAdd(a.dbias + i, gacc.bias[i]);
  if (a.tg)
    for (int g = 0; g < 2; ++g) {
      for (int i = threadIdx.x; i < nA; i += PT) if (a.dtA[g]) atomicAdd(a.dtA[g] + i, gacc.tA[g][i]);
      for (int i = threadIdx.x; i < nB; i += PT) if (a.dtB[g]) atomicAdd(a.dtB[g] + i, gacc.tB[g][i]);
      for (int i = threadIdx.x; i < F; i += PT) if (a.dtb[g]) atomicAdd(a.dtb[g] + i, gacc.tb[g][i]);
      for (int i = threadIdx.x; i < FN; i += PT) if (a.dtW[g]) atomicAdd(a.dtW[g] + i, gacc.tW[g][i]);
      if (threadIdx.x == 0 && a.dtc[g]) atomicAdd(a.dtc[g], dtc[g]);
    }
}

}  // namespace persist
}  // namespace gcrnn
