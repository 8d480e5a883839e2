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
//            node gates (graphML.py:2379-2407): q = sigmoid(GraphFilter_{F->1}(tanh(A_n(S)x_t + B_n(S)h0 + 2 b_n))) per node, the
//            F -> 1 head contracted first and shifted as a scalar signal (Horner), applied as g_i q_i[n] a + g_f q_f[n] r.
//
//            edge gates (graphML.py:2409-2416, 521-627): both filter outputs pass through a one-head graph attention layer over
//            the pattern of S + I (scores leaky_relu(a2.Wy_i + a1.Wy_j), row softmax, S'-weighted aggregation, ReLU); the pattern
//            (row view + column view with edge ids) sits in shared memory next to the gather lists, attention weights are
//            recomputed in the reverse sweep.
//
// Supported: E = 1, time gating on or off, node OR edge gating or neither, no dX (the reference never asks for it:
// train_rnn.py:256); sizes such that everything fits the 227 KB of one SM.  Everything else takes the per-op kernels.
#pragma once
#include "common.cuh"

namespace gcrnn {
namespace persist {

constexpr int PT = 512;       // threads per CTA

struct Args {
  int N, F, G, Kin, Kst, tg, node, edge, has_bias;
  long long B, T;
  const int *cptr, *cidx; const float* cval;     // gather form of z @ S   (CSC of S):  out[n] = sum_p cval[p] in[cidx[p]]
  const int *rptr, *ridx; const float* rval;     // gather form of g @ S^T (CSR of S)
  int nnz;                                       // entries of S
  int lists_smem;                                // 1: both gather lists are staged in shared memory (16-bit indices)
  const float *A, *Bw, *bias;                    // [F,Kin,G] [F,Kst,F] [F]
  const float *tA[2], *tB[2], *tb[2], *tW[2], *tc[2];          // time-gate sub-cells + MLP
  const float *nA[2], *nB[2], *nb[2], *nhw[2], *nhb[2];         // node-gate sub-cells + F -> 1 head [Kst][F], [1]
  // edge gates: attention pattern of S + I (row view; column view carrying row and edge id), weight [F][F], mixer [2F] per gate
  const int *arptr, *acol; const float* aval; const int *acptr, *acrow, *aceid; int annz;
  const float *eW[2], *em[2];
  const float *X, *h0;                           // [B,T,G,N] [B,F,N]
  float* H;                                      // [B,T,F,N]  (forward: out; backward: in)
  float* gt;                                     // [2][B][T] time-gate values (forward: out; backward: in)
  float* qn;                                     // [2][B][T][N] node-gate values (forward: out; backward: in)
  // backward
  const float* dH; long long dH_bstride, dH_tstride; int dh_last_only;   // dH[b,t] = dH + b*bstride + t*tstride ([F][N]); last_only: zero for t < T-1
  float *dA, *dBw, *dbias, *dtA[2], *dtB[2], *dtb[2], *dtW[2], *dtc[2];
  float *dnA[2], *dnB[2], *dnb[2], *dnhw[2], *dnhb[2];
  float *deW[2], *dem[2];
  float* dh0;                                    // [B,F,N] or null
};

// shared-memory floats of the two kernels (same formula on host and device)
struct Shape { int F, G, Kin, Kst, N, tg, node, edge, annz; };
__host__ __device__ inline long long weights_floats(const Shape& s) {
  const long long cell = (long long)s.F * s.Kin * s.G + (long long)s.F * s.Kst * s.F + s.F;
  return cell + (s.tg ? 2 * cell + 2LL * s.F * s.N : 0) + (s.node ? 2 * cell + 2LL * s.Kst * s.F : 0) + (s.edge ? 2LL * (s.F * s.F + 2 * s.F) : 0);
}
__host__ __device__ inline long long annz4(const Shape& s) { return (s.annz + 3) & ~3; }
__host__ __device__ inline long long zx_floats(int Kin, int G, int N) { return ((long long)Kin * G * N + 3) & ~3LL; }
__host__ __device__ inline long long fwd_floats(const Shape& s) {
  const long long FN = (long long)s.F * s.N;
  return weights_floats(s) + zx_floats(s.Kin, s.G, s.N) + (long long)s.Kst * FN + FN /*hn*/ +
         (s.tg ? 2 * FN : 0) /*c0*/ + (s.node ? 3 * FN + (long long)s.Kst * s.N + 2LL * s.N : 0) /*c0n, s, pk, q*/ +
         (s.edge ? 5 * FN + 2LL * s.N + annz4(s) : 0) /*ya, yr, Wx, out_a, out_r, rr, cc, al*/ + 64;
}
__host__ __device__ inline long long bwd_floats(const Shape& s) {
  const long long FN = (long long)s.F * s.N;
  return 2 * weights_floats(s) /*weights + gradient accumulators*/ +
         zx_floats(s.Kin, s.G, s.N) + (long long)s.Kst * FN + 5 * FN /*da dr dh b1 b2*/ + (s.tg ? 4 * FN : 0) /*c0, dc0*/ +
         ((s.tg || s.node) ? FN : 0) /*dpu*/ + (s.node ? 5 * FN + (long long)s.Kst * s.N + 4LL * s.N : 0) /*c0n, dc0n, s, vch, q, dq*/ +
         (s.edge ? 6 * FN + 4LL * s.N + 2 * annz4(s) : 0) /*ya, yr, dp, Wx, out/dy, dWx, rr, cc, drr, dcc, al, tmp*/ + 64;
}
// shared-memory bytes of the staged attention pattern: rptr, cptr (int), val (float), col, crow, ceid, erow (u16)
__host__ __device__ inline long long att_bytes(int N, int annz) {
  const long long p = ((long long)(N + 1) * 4 + 15) & ~15LL, v = ((long long)annz * 4 + 15) & ~15LL, h = ((long long)annz * 2 + 15) & ~15LL;
  return 2 * p + v + 4 * h;
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

// ---- blocked inner loops ------------------------------------------------------------------------------------------------
// Every loop below is register-blocked so that one shared-memory load feeds several FMAs (the first version issued two loads
// per FMA and ran at 40 % issue utilisation with 16 warps: profiles/r02_ncu_persist_*.raw.csv): NB = 4 consecutive nodes per thread with
// 16-byte loads where the contraction runs over rows (needs N % 4 == 0), RB = 4 rows per thread where a gather list is shared.

// out[r][n] = sum_p val[p] in[r][idx[p]], p in [ptr[n], ptr[n+1])   (rows r < R; in / out in shared memory, [R][N])
template <int RB>
__device__ __forceinline__ void shift(const List& l, const float* in, float* out, int R, int N) {
  const int RG = R / RB;
  for (int e = threadIdx.x; e < RG * N; e += PT) {
    const int rg = e / N, n = e - rg * N;
    const float* row = in + (size_t)rg * RB * N;
    float s[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) s[j] = 0.f;
    const int p1 = l.ptr[n + 1];
    for (int p = l.ptr[n]; p < p1; ++p) {
      const float v = l.val[p];
      const int i = l.idx(p);
#pragma unroll
      for (int j = 0; j < RB; ++j) s[j] = fmaf(v, row[(size_t)j * N + i], s[j]);
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) out[((size_t)rg * RB + j) * N + n] = s[j];
  }
}
// z[k] = z[k-1] S for k = 1..K-1, z: [K][R][N]
__device__ __forceinline__ void chain(const List& fw, float* z, int K, int R, int N) {
  for (int k = 1; k < K; ++k) {
    if (R % 4 == 0) shift<4>(fw, z + (size_t)(k - 1) * R * N, z + (size_t)k * R * N, R, N);
    else shift<1>(fw, z + (size_t)(k - 1) * R * N, z + (size_t)k * R * N, R, N);
    __syncthreads();
  }
}
// y[j] = sum_i W[f][i] z[i][n0 + j], j < NB   (i over the KC = K * C rows of z)
template <int NB>
__device__ __forceinline__ void contract(const float* W, const float* z, int f, int n0, int KC, int N, float* y) {
  const float* w = W + (size_t)f * KC;
  const float* zc = z + n0;
#pragma unroll
  for (int j = 0; j < NB; ++j) y[j] = 0.f;
  if (NB == 4) {
    float y2[4] = {0.f, 0.f, 0.f, 0.f};                           // second set of chains for the odd rows
    int i = 0;
    for (; i + 1 < KC; i += 2) {
      const float4 a = *reinterpret_cast<const float4*>(zc + (size_t)i * N);
      const float4 c = *reinterpret_cast<const float4*>(zc + (size_t)(i + 1) * N);
      const float w0 = w[i], w1 = w[i + 1];
      y[0] = fmaf(w0, a.x, y[0]); y[1] = fmaf(w0, a.y, y[1]); y[2] = fmaf(w0, a.z, y[2]); y[3] = fmaf(w0, a.w, y[3]);
      y2[0] = fmaf(w1, c.x, y2[0]); y2[1] = fmaf(w1, c.y, y2[1]); y2[2] = fmaf(w1, c.z, y2[2]); y2[3] = fmaf(w1, c.w, y2[3]);
    }
    if (i < KC) {
      const float4 a = *reinterpret_cast<const float4*>(zc + (size_t)i * N);
      const float w0 = w[i];
      y[0] = fmaf(w0, a.x, y[0]); y[1] = fmaf(w0, a.y, y[1]); y[2] = fmaf(w0, a.z, y[2]); y[3] = fmaf(w0, a.w, y[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] += y2[j];
  } else {
    float s0 = 0.f, s1 = 0.f;
    int i = 0;
    for (; i + 1 < KC; i += 2) { s0 = fmaf(w[i], zc[(size_t)i * N], s0); s1 = fmaf(w[i + 1], zc[(size_t)(i + 1) * N], s1); }
    if (i < KC) s0 = fmaf(w[i], zc[(size_t)i * N], s0);
    y[0] = s0 + s1;
  }
}

// ---- "quad" layout of the state-side slabs (QZ, needs F % 4 == 0) ---------------------------------------------------------------
// The shift of a 4-row group gathers in[r..r+3][idx] per edge: four scalar loads at stride N plus their address arithmetic were the
// hot spot of these kernels (profiles/r02_ncu_persist_*.source_top.txt).  With the four rows of a group interleaved per node,
//   element (c, n) of a slab at  ((c / 4) * N + n) * 4 + c % 4,
// an edge costs ONE 16-byte load, and the contractions read a weight quad and a signal quad per four FMAs.  Slabs k = 0..K-1 are
// contiguous, so row i = k*C + c of the stacked signal lives in quad i / 4.
__device__ __forceinline__ int quad_index(int c, int n, int N) { return (((c >> 2) * N + n) << 2) + (c & 3); }
template <bool QZ>
__device__ __forceinline__ void put_state(float* z0, const float* src, int F, int N) {      // slab 0 <- src [F][N] (global or shared)
  for (int e = threadIdx.x; e < F * N; e += PT) {
    if (QZ) { const int c = e / N, n = e - c * N; z0[quad_index(c, n, N)] = src[e]; }
    else z0[e] = src[e];
  }
}
__device__ __forceinline__ void shift_quad(const List& l, const float* in, float* out, int R, int N) {
  const float4* in4 = reinterpret_cast<const float4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int e = threadIdx.x; e < (R >> 2) * N; e += PT) {
    const int rg = e / N, n = e - rg * N;
    const float4* row = in4 + (size_t)rg * N;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p1 = l.ptr[n + 1];
    for (int p = l.ptr[n]; p < p1; ++p) {
      const float v = l.val[p];
      const float4 x = row[l.idx(p)];
      s.x = fmaf(v, x.x, s.x); s.y = fmaf(v, x.y, s.y); s.z = fmaf(v, x.z, s.z); s.w = fmaf(v, x.w, s.w);
    }
    out4[e] = s;
  }
}
template <bool QZ>
__device__ __forceinline__ void chain_z(const List& fw, float* z, int K, int R, int N) {
  if constexpr (!QZ) {
    chain(fw, z, K, R, N);
  } else {
    for (int k = 1; k < K; ++k) {
      shift_quad(fw, z + (size_t)(k - 1) * R * N, z + (size_t)k * R * N, R, N);
      __syncthreads();
    }
  }
}
// y[j] = sum_i W[f][i] z[i][n0 + j], j < NB, z in quad layout (KC % 4 == 0)
template <int NB>
__device__ __forceinline__ void contract_quad(const float* W, const float* z, int f, int n0, int KC, int N, float* y) {
  const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)f * KC);
  const float4* z4 = reinterpret_cast<const float4*>(z) + n0;
  float acc[NB][2];
#pragma unroll
  for (int j = 0; j < NB; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; }
  const int Q = KC >> 2;
  for (int q = 0; q < Q; ++q) {
    const float4 w = w4[q];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float4 x = z4[(size_t)q * N + j];
      acc[j][0] = fmaf(w.x, x.x, fmaf(w.y, x.y, acc[j][0]));
      acc[j][1] = fmaf(w.z, x.z, fmaf(w.w, x.w, acc[j][1]));
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) y[j] = acc[j][0] + acc[j][1];
}
template <int NB, bool QZ>
__device__ __forceinline__ void contract_z(const float* W, const float* z, int f, int n0, int KC, int N, float* y) {
  if (QZ) contract_quad<NB>(W, z, f, n0, KC, N, y); else contract<NB>(W, z, f, n0, KC, N, y);
}
// acc[f][4q..4q+3] += sum_n d[f][n] z4[q][n]: a thread owns one feature and one quad of stacked rows; the lanes of a warp walk the
// nodes from different starting points so that their 16-byte loads fall into different banks
__device__ __forceinline__ void wgrad_quad(float* acc, const float* d, const float* z, int F, int KC, int N) {
  const int Q = KC >> 2;
  const float4* z4 = reinterpret_cast<const float4*>(z);
  for (int o = threadIdx.x; o < F * Q; o += PT) {
    const int f = o / Q, q = o - f * Q;
    const float* dr = d + (size_t)f * N;
    const float4* zr = z4 + (size_t)q * N;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int n = threadIdx.x % N;
    for (int i = 0; i < N; ++i) {
      const float dv = dr[n];
      const float4 x = zr[n];
      s.x = fmaf(dv, x.x, s.x); s.y = fmaf(dv, x.y, s.y); s.z = fmaf(dv, x.z, s.z); s.w = fmaf(dv, x.w, s.w);
      if (++n == N) n = 0;
    }
    float4* ap = reinterpret_cast<float4*>(acc + (size_t)f * KC) + q;
    float4 t = *ap;
    t.x += s.x; t.y += s.y; t.z += s.z; t.w += s.w;
    *ap = t;
  }
}

struct Sub { float *A, *B, *b; };                 // an ungated sub-cell: input taps, state taps, bias
struct Weights {
  float *A, *Bw, *bias;
  Sub ts[2]; float* tW[2];                       // time gates: sub-cell + MLP weights [F*N]
  Sub ns[2]; float* nh[2];                       // node gates: sub-cell + head taps [Kst][F]
  float* eW[2]; float* em[2];                    // edge gates: attention weight [F][F], mixer [2F]
};
__device__ __forceinline__ float* carve(Weights& w, float* p, const Args& a) {
  const int nA = a.F * a.Kin * a.G, nB = a.F * a.Kst * a.F;
  w.A = p; p += nA; w.Bw = p; p += nB; w.bias = p; p += a.F;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    w.ts[g].A = w.ts[g].B = w.ts[g].b = w.tW[g] = p;
    if (a.tg) { w.ts[g].A = p; p += nA; w.ts[g].B = p; p += nB; w.ts[g].b = p; p += a.F; w.tW[g] = p; p += a.F * a.N; }
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    w.ns[g].A = w.ns[g].B = w.ns[g].b = w.nh[g] = p;
    if (a.node) { w.ns[g].A = p; p += nA; w.ns[g].B = p; p += nB; w.ns[g].b = p; p += a.F; w.nh[g] = p; p += a.Kst * a.F; }
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    w.eW[g] = w.em[g] = p;
    if (a.edge) { w.eW[g] = p; p += a.F * a.F; w.em[g] = p; p += 2 * a.F; }
  }
  return p;
}
__device__ __forceinline__ void load_sub(const Sub& s, const float* A, const float* B, const float* b, const Args& a) {
  const int nA = a.F * a.Kin * a.G, nB = a.F * a.Kst * a.F;
  for (int i = threadIdx.x; i < nA; i += PT) s.A[i] = A[i];
  for (int i = threadIdx.x; i < nB; i += PT) s.B[i] = B[i];
  for (int i = threadIdx.x; i < a.F; i += PT) s.b[i] = a.has_bias ? b[i] : 0.f;
}
__device__ __forceinline__ void load_weights(const Weights& w, const Args& a) {
  const int nA = a.F * a.Kin * a.G, nB = a.F * a.Kst * a.F, FN = a.F * a.N;
  for (int i = threadIdx.x; i < nA; i += PT) w.A[i] = a.A[i];
  for (int i = threadIdx.x; i < nB; i += PT) w.Bw[i] = a.Bw[i];
  for (int i = threadIdx.x; i < a.F; i += PT) w.bias[i] = a.has_bias ? a.bias[i] : 0.f;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (a.tg) {
      load_sub(w.ts[g], a.tA[g], a.tB[g], a.tb[g], a);
      for (int i = threadIdx.x; i < FN; i += PT) w.tW[g][i] = a.tW[g][i];
    }
    if (a.node) {
      load_sub(w.ns[g], a.nA[g], a.nB[g], a.nb[g], a);
      for (int i = threadIdx.x; i < a.Kst * a.F; i += PT) w.nh[g][i] = a.nhw[g][i];
    }
    if (a.edge) {
      for (int i = threadIdx.x; i < a.F * a.F; i += PT) w.eW[g][i] = a.eW[g][i];
      for (int i = threadIdx.x; i < 2 * a.F; i += PT) w.em[g][i] = a.em[g][i];
    }
  }
}
// c0[f][n] = sum_{k,g} B_s[f][k][g] zh[k][g][n] + 2 b_s[f]: the T-invariant term of a sub-cell run from the initial state
template <int NB, bool QZ>
__device__ __forceinline__ void subcell_c0(const Sub& sc, const float* zh, float* c0, int F, int KCb, int N) {
  const int NQ = N / NB;
  for (int e = threadIdx.x; e < F * NQ; e += PT) {
    const int f = e / NQ, n0 = (e - f * NQ) * NB;
    float y[NB];
    contract_z<NB, QZ>(sc.B, zh, f, n0, KCb, N, y);
#pragma unroll
    for (int j = 0; j < NB; ++j) c0[f * N + n0 + j] = y[j] + 2.f * sc.b[f];
  }
}
// node-gate head, forward: s = tanh(A_n(S)x_t + c0n) (kept in `sbuf`), p_k[n] = sum_f wh[k][f] s[f][n], Horner with S:
// lin = p_0 + (p_1 + (... p_{K-1} S ...) S) S, q = sigmoid(lin + c)   -> qs[n]
template <int NB>
__device__ __forceinline__ void node_gate_fwd(const Args& a, const List& fw, const Sub& sc, const float* wh, float hb, const float* zx,
                                              const float* c0n, float* sbuf, float* pk, float* qs) {
  const int N = a.N, F = a.F, NQ = N / NB, KCa = a.Kin * a.G;
  for (int e = threadIdx.x; e < F * NQ; e += PT) {
    const int f = e / NQ, n0 = (e - f * NQ) * NB;
    float y[NB];
    contract<NB>(sc.A, zx, f, n0, KCa, N, y);
#pragma unroll
    for (int j = 0; j < NB; ++j) sbuf[f * N + n0 + j] = tanhf(y[j] + c0n[f * N + n0 + j]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < a.Kst * N; e += PT) {
    const int k = e / N, n = e - k * N;
    float s0 = 0.f, s1 = 0.f;
    int f = 0;
    for (; f + 1 < F; f += 2) { s0 = fmaf(wh[k * F + f], sbuf[f * N + n], s0); s1 = fmaf(wh[k * F + f + 1], sbuf[(f + 1) * N + n], s1); }
    if (f < F) s0 = fmaf(wh[k * F + f], sbuf[f * N + n], s0);
    pk[e] = s0 + s1;
  }
  __syncthreads();
  for (int k = a.Kst - 2; k >= 0; --k) {                            // pk[k] += pk[k+1] S
    const float* cur = pk + (size_t)(k + 1) * N;
    for (int n = threadIdx.x; n < N; n += PT) {
      float s = 0.f;
      const int p1 = fw.ptr[n + 1];
      for (int p = fw.ptr[n]; p < p1; ++p) s = fmaf(fw.val[p], cur[fw.idx(p)], s);
      pk[(size_t)k * N + n] += s;
    }
    __syncthreads();
  }
  for (int n = threadIdx.x; n < N; n += PT) qs[n] = 1.f / (1.f + expf(-(pk[n] + hb)));
  __syncthreads();
}

// ---- edge gates: one-head graph attention over the pattern of S' = S + I, all in shared memory -------------------------------
struct Att {                                     // row i: edges p in [rptr[i], rptr[i+1]) -> column col[p], value val[p];
  const int* rptr; const int* cptr;              // column j: entries q in [cptr[j], cptr[j+1]) -> row crow[q], edge id ceid[q]
  const float* val; const unsigned short *col, *crow, *ceid, *erow;
  int nnz;
};
__device__ __forceinline__ Att stage_att(unsigned char*& sp, const Args& a) {
  Att t; t.nnz = a.annz;
  const int N = a.N, nnz = a.annz;
  int* rp = reinterpret_cast<int*>(sp); sp += ((N + 1) * 4 + 15) & ~15;
  int* cp = reinterpret_cast<int*>(sp); sp += ((N + 1) * 4 + 15) & ~15;
  float* v = reinterpret_cast<float*>(sp); sp += (nnz * 4 + 15) & ~15;
  unsigned short* h[4];
  for (int i = 0; i < 4; ++i) { h[i] = reinterpret_cast<unsigned short*>(sp); sp += (nnz * 2 + 15) & ~15; }
  for (int i = threadIdx.x; i <= N; i += PT) { rp[i] = a.arptr[i]; cp[i] = a.acptr[i]; }
  for (int i = threadIdx.x; i < nnz; i += PT) {
    v[i] = a.aval[i]; h[0][i] = (unsigned short)a.acol[i]; h[1][i] = (unsigned short)a.acrow[i]; h[2][i] = (unsigned short)a.aceid[i];
  }
  for (int i = threadIdx.x; i < N; i += PT)
    for (int p = a.arptr[i]; p < a.arptr[i + 1]; ++p) h[3][p] = (unsigned short)i;
  t.rptr = rp; t.cptr = cp; t.val = v; t.col = h[0]; t.crow = h[1]; t.ceid = h[2]; t.erow = h[3];
  return t;
}
__device__ __forceinline__ float leaky02(float x) { return x > 0.f ? x : 0.2f * x; }     // graphML.py:521 negative_slope
// out[f][j] = sum_i Wy[f][i] S'_ij al_ij,  Wy = W y,  al = row softmax of leaky(a2.Wy_i + a1.Wy_j)   (graphML.py:586-625; no ReLU here)
template <int NB>
__device__ __forceinline__ void gat_fwd(int N, int F, const Att& t, const float* W, const float* mix, const float* y,
                                        float* Wy, float* rr, float* cc, float* al, float* out) {
  const int NQ = N / NB;
  for (int e = threadIdx.x; e < F * NQ; e += PT) {
    const int f = e / NQ, n0 = (e - f * NQ) * NB;
    float v[NB];
    contract<NB>(W, y, f, n0, F, N, v);
#pragma unroll
    for (int j = 0; j < NB; ++j) Wy[f * N + n0 + j] = v[j];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * N; e += PT) {
    const int h = e / N, n = e - h * N;
    float s0 = 0.f;
    for (int f = 0; f < F; ++f) s0 = fmaf(mix[h * F + f], Wy[f * N + n], s0);
    (h ? cc : rr)[n] = s0;                                         // mixer[:F] pairs with the column node j, mixer[F:] with the row node i
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += PT) {
    const int p0 = t.rptr[i], p1 = t.rptr[i + 1];
    float m = -INFINITY;
    for (int p = p0; p < p1; ++p) m = fmaxf(m, leaky02(cc[i] + rr[t.col[p]]));
    float den = 0.f;
    for (int p = p0; p < p1; ++p) { const float e = expf(leaky02(cc[i] + rr[t.col[p]]) - m); al[p] = e; den += e; }
    const float inv = 1.f / den;
    for (int p = p0; p < p1; ++p) al[p] *= inv;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F * N; o += PT) {
    const int f = o / N, j = o - f * N;
    float s0 = 0.f;
    for (int q = t.cptr[j]; q < t.cptr[j + 1]; ++q) { const int e = t.ceid[q]; s0 = fmaf(Wy[f * N + t.crow[q]], t.val[e] * al[e], s0); }
    out[o] = s0;
  }
  __syncthreads();
}

template <int NB, int SG, bool QZ>      // SG: spatial gating 0 none, 1 node, 2 edge; QZ: quad layout of the state-side slabs (F % 4 == 0)
__global__ void __launch_bounds__(PT, 1) persist_fwd_k(const Args a) {
  constexpr bool NODE = SG == 1, EDGE = SG == 2;
  extern __shared__ __align__(16) float psm[];
  const int N = a.N, F = a.F, FN = F * N, GN = a.G * N, NQ = N / NB;
  const int KCa = a.Kin * a.G, KCb = a.Kst * F;
  const long long b = blockIdx.x;
  Weights w;
  float* p = carve(w, psm, a);
  float* zx = p; p += zx_floats(a.Kin, a.G, N);
  float* zh = p; p += (size_t)a.Kst * FN;
  float* hn = p; p += FN;
  float* c0 = p; p += a.tg ? 2 * FN : 0;
  float* c0n = p; p += NODE ? 2 * FN : 0;
  float* sbuf = p; p += NODE ? FN : 0;
  float* pk = p; p += NODE ? a.Kst * N : 0;
  float* qs = p; p += NODE ? 2 * N : 0;
  float* ya = p; p += EDGE ? FN : 0;
  float* yr = p; p += EDGE ? FN : 0;
  float* Wy = p; p += EDGE ? FN : 0;
  float* oa = p; p += EDGE ? FN : 0;
  float* orr = p; p += EDGE ? FN : 0;
  float* rr = p; p += EDGE ? N : 0;
  float* cc = p; p += EDGE ? N : 0;
  float* al = p; p += EDGE ? ((a.annz + 3) & ~3) : 0;
  float* red = p; p += 64;
  unsigned char* sp = reinterpret_cast<unsigned char*>(p);
  const List fw = stage_list(sp, a.cptr, a.cidx, a.cval, N, a.nnz, a.lists_smem != 0);
  Att att{};
  if (EDGE) att = stage_att(sp, a);
  load_weights(w, a);
  put_state<QZ>(zh, a.h0 + b * FN, F, N);
  __syncthreads();
  const bool gated = a.tg || NODE;
  if (gated) {                                                    // T-invariant gate terms: B_s(S) h0 + 2 b_s   (graphML.py:2362, :2383, :2417-2423)
    chain_z<QZ>(fw, zh, a.Kst, F, N);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (a.tg) subcell_c0<NB, QZ>(w.ts[g], zh, c0 + g * FN, F, KCb, N);
      if (NODE) subcell_c0<NB, QZ>(w.ns[g], zh, c0n + g * FN, F, KCb, N);
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
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float part = 0.f;
        for (int e = threadIdx.x; e < F * NQ; e += PT) {
          const int f = e / NQ, n0 = (e - f * NQ) * NB;
          float y[NB];
          contract<NB>(w.ts[g].A, zx, f, n0, KCa, N, y);
#pragma unroll
          for (int j = 0; j < NB; ++j) part = fmaf(w.tW[g][f * N + n0 + j], tanhf(y[j] + c0[g * FN + f * N + n0 + j]), part);
        }
        const float logit = block_sum(part, red) + (a.has_bias ? __ldg(a.tc[g]) : 0.f);
        const float gv = 1.f / (1.f + expf(-logit));
        if (g == 0) gi = gv; else gf = gv;
        if (threadIdx.x == 0) a.gt[((long long)g * a.B + b) * a.T + t] = gv;
      }
    }
    if (NODE) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        node_gate_fwd<NB>(a, fw, w.ns[g], w.nh[g], a.has_bias ? __ldg(a.nhb[g]) : 0.f, zx, c0n + g * FN, sbuf, pk, qs + g * N);
        float* qo = a.qn + (((long long)g * a.B + b) * a.T + t) * N;
        for (int n = threadIdx.x; n < N; n += PT) qo[n] = qs[g * N + n];
      }
    }
    if (!(gated && t == 0)) chain_z<QZ>(fw, zh, a.Kst, F, N);           // at t = 0 with gating the h0 chain is already there
    float* Ht = a.H + (b * a.T + t) * FN;
    for (int e = threadIdx.x; e < F * NQ; e += PT) {
      const int f = e / NQ, n0 = (e - f * NQ) * NB;
      float av[NB], rv[NB];
      contract<NB>(w.A, zx, f, n0, KCa, N, av);
      contract_z<NB, QZ>(w.Bw, zh, f, n0, KCb, N, rv);
      const float bb = w.bias[f];                                  // the same bias in both filters (:2405-2407)
      if constexpr (EDGE) {
#pragma unroll
        for (int j = 0; j < NB; ++j) { ya[f * N + n0 + j] = av[j] + bb; yr[f * N + n0 + j] = rv[j] + bb; }
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float wi = NODE ? gi * qs[n0 + j] : gi, wf = NODE ? gf * qs[N + n0 + j] : gf;
          const float h = tanhf(fmaf(wi, av[j] + bb, wf * (rv[j] + bb)));
          Ht[f * N + n0 + j] = h;
          hn[f * N + n0 + j] = h;
        }
      }
    }
    __syncthreads();
    if (EDGE) {                                                  // :2409-2416: both filter outputs through their attention layer
      gat_fwd<NB>(N, F, att, w.eW[0], w.em[0], ya, Wy, rr, cc, al, oa);
      gat_fwd<NB>(N, F, att, w.eW[1], w.em[1], yr, Wy, rr, cc, al, orr);
      for (int o = threadIdx.x; o < FN; o += PT) {
        const float h = tanhf(fmaf(gi, fmaxf(oa[o], 0.f), gf * fmaxf(orr[o], 0.f)));
        Ht[o] = h;
        hn[o] = h;
      }
      __syncthreads();
    }
    put_state<QZ>(zh, hn, F, N);
    __syncthreads();
  }
}

// acc[f][kc] += sum_n d[f][n] z[kc][n]   (every output owned by one thread; acc in shared memory)
template <int NB>
__device__ __forceinline__ void wgrad_acc(float* acc, const float* d, const float* z, int F, int KC, int N) {
  for (int o = threadIdx.x; o < F * KC; o += PT) {
    const int f = o / KC, kc = o - f * KC;
    const float* dr = d + (size_t)f * N;
    const float* zr = z + (size_t)kc * N;
    float s0 = 0.f, s1 = 0.f;
    if (NB == 4) {
      const int NQ = N / 4;
      int q = threadIdx.x % NQ;                                    // skewed start: the lanes of a warp read different banks
      for (int i = 0; i < NQ; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(dr + 4 * q);
        const float4 c = *reinterpret_cast<const float4*>(zr + 4 * q);
        s0 = fmaf(a.x, c.x, s0); s1 = fmaf(a.y, c.y, s1); s0 = fmaf(a.z, c.z, s0); s1 = fmaf(a.w, c.w, s1);
        if (++q == NQ) q = 0;
      }
    } else {
      int nn = threadIdx.x % N;
      for (int n = 0; n < N; ++n) { if (n & 1) s1 = fmaf(dr[nn], zr[nn], s1); else s0 = fmaf(dr[nn], zr[nn], s0); if (++nn == N) nn = 0; }
    }
    acc[o] += s0 + s1;
  }
}
// out[g][n] (+)= Horner over k of ( sum_f W[f][k][g] d[f][n] ) with S^T:  out = u_0 + (u_1 + (... u_{K-1} S^T ...) S^T) S^T
// b1 / b2: [C][N] scratch.  GB features g per thread share the S^T gather list of node n and the loads of d[f][n].
template <int GB>
__device__ __forceinline__ void adjoint_chain(const Args& a, const List& bw, const float* W, const float* d, float* b1, float* b2, float* out,
                                              int K, int C, bool accumulate) {
  const int N = a.N, F = a.F, CG = C / GB;
  float* cur = b1; float* nxt = b2;
  for (int k = K - 1; k >= 0; --k) {
    for (int e = threadIdx.x; e < CG * N; e += PT) {
      const int cg = e / N, n = e - cg * N, g0 = cg * GB;
      float s[GB];
#pragma unroll
      for (int j = 0; j < GB; ++j) s[j] = 0.f;
      if (k < K - 1) {
        const float* row = cur + (size_t)g0 * N;
        const int p1 = bw.ptr[n + 1];
        for (int p = bw.ptr[n]; p < p1; ++p) {
          const float v = bw.val[p];
          const int i = bw.idx(p);
#pragma unroll
          for (int j = 0; j < GB; ++j) s[j] = fmaf(v, row[(size_t)j * N + i], s[j]);
        }
      }
      for (int f = 0; f < F; ++f) {
        const float dv = d[(size_t)f * N + n];
        const float* wp = W + ((size_t)f * K + k) * C + g0;
        if (GB == 4) {
          const float4 wv = *reinterpret_cast<const float4*>(wp);
          s[0] = fmaf(wv.x, dv, s[0]); s[1 % GB] = fmaf(wv.y, dv, s[1 % GB]); s[2 % GB] = fmaf(wv.z, dv, s[2 % GB]); s[3 % GB] = fmaf(wv.w, dv, s[3 % GB]);
        } else {
#pragma unroll
          for (int j = 0; j < GB; ++j) s[j] = fmaf(wp[j], dv, s[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < GB; ++j) {
        const size_t o = (size_t)(g0 + j) * N + n;
        if (k == 0) { if (accumulate) out[o] += s[j]; else out[o] = s[j]; }
        else nxt[o] = s[j];
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

// same with the scratch signals in quad layout (C % 4 == 0): the S^T gather of four features is one 16-byte load per edge
__device__ __forceinline__ void adjoint_chain_quad(const Args& a, const List& bw, const float* W, const float* d, float* b1, float* b2, float* out,
                                                   int K, int C, bool accumulate) {
  const int N = a.N, F = a.F, CG = C >> 2;
  float4* cur = reinterpret_cast<float4*>(b1); float4* nxt = reinterpret_cast<float4*>(b2);
  for (int k = K - 1; k >= 0; --k) {
    for (int e = threadIdx.x; e < CG * N; e += PT) {
      const int cg = e / N, n = e - cg * N, g0 = cg << 2;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K - 1) {
        const float4* row = cur + (size_t)cg * N;
        const int p1 = bw.ptr[n + 1];
        for (int p = bw.ptr[n]; p < p1; ++p) {
          const float v = bw.val[p];
          const float4 x = row[bw.idx(p)];
          s.x = fmaf(v, x.x, s.x); s.y = fmaf(v, x.y, s.y); s.z = fmaf(v, x.z, s.z); s.w = fmaf(v, x.w, s.w);
        }
      }
      for (int f = 0; f < F; ++f) {
        const float dv = d[(size_t)f * N + n];
        const float4 wv = *reinterpret_cast<const float4*>(W + ((size_t)f * K + k) * C + g0);
        s.x = fmaf(wv.x, dv, s.x); s.y = fmaf(wv.y, dv, s.y); s.z = fmaf(wv.z, dv, s.z); s.w = fmaf(wv.w, dv, s.w);
      }
      if (k == 0) {
        float* o = out + (size_t)g0 * N + n;
        if (accumulate) { o[0] += s.x; o[N] += s.y; o[2 * N] += s.z; o[3 * N] += s.w; }
        else { o[0] = s.x; o[N] = s.y; o[2 * N] = s.z; o[3 * N] = s.w; }
      } else {
        nxt[e] = s;
      }
    }
    __syncthreads();
    float4* t = cur; cur = nxt; nxt = t;
  }
}
template <int NB, bool QZ>
__device__ __forceinline__ void adjoint_z(const Args& a, const List& bw, const float* W, const float* d, float* b1, float* b2, float* out,
                                          int K, int C, bool accumulate) {
  if (QZ) adjoint_chain_quad(a, bw, W, d, b1, b2, out, K, C, accumulate);
  else adjoint_chain<NB>(a, bw, W, d, b1, b2, out, K, C, accumulate);
}
template <int NB, bool QZ>
__device__ __forceinline__ void wgrad_z(float* acc, const float* d, const float* z, int F, int KC, int N) {
  if (QZ) wgrad_quad(acc, d, z, F, KC, N); else wgrad_acc<NB>(acc, d, z, F, KC, N);
}

// T-invariant term of a sub-cell, backward: v = sum_t d pre_s.  dB_s += v zh0^T, db_s += 2 sum_n v, dh0 += adjoint chain
template <int NB, bool QZ>
__device__ __forceinline__ void subcell_c0_bwd(const Args& a, const List& bw, const Sub& sc, const Sub& gsc, const float* v, const float* zh,
                                               float* b1, float* b2, float* dh) {
  const int N = a.N, F = a.F;
  wgrad_z<NB, QZ>(gsc.B, v, zh, F, a.Kst * F, N);
  for (int f = threadIdx.x; f < F; f += PT) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += v[(size_t)f * N + n];
    gsc.b[f] += 2.f * s;                                           // the sub-cell adds its bias twice (:2421-2422)
  }
  if (a.dh0) adjoint_z<NB, QZ>(a, bw, sc.B, v, b1, b2, dh, a.Kst, F, true);
  __syncthreads();
}
__device__ __forceinline__ void flush_sub(const Sub& g, float* dA, float* dB, float* db, int nA, int nB, int F) {
  for (int i = threadIdx.x; i < nA; i += PT) if (dA) atomicAdd(dA + i, g.A[i]);
  for (int i = threadIdx.x; i < nB; i += PT) if (dB) atomicAdd(dB + i, g.B[i]);
  for (int i = threadIdx.x; i < F; i += PT) if (db) atomicAdd(db + i, g.b[i]);
}

// reverse of gat_fwd for one gate.  In: y (the filter output), Wy / rr / cc / al as gat_fwd left them, dy = d loss / d out (ReLU mask
// and gate scalar applied).  Out: dyin = d loss / d y; gW [F][F] and gmix [2F] accumulate.  tmp: [nnz], drr / dcc: [N], dWy: [F][N].
template <int NB>
__device__ __forceinline__ void gat_bwd(int N, int F, const Att& t, const float* W, const float* mix, float* gW, float* gmix,
                                        const float* y, const float* Wy, const float* dy, float* dWy, const float* rr, const float* cc,
                                        float* drr, float* dcc, const float* al, float* tmp, float* dyin) {
  for (int e = threadIdx.x; e < t.nnz; e += PT) {                  // d al_ij = S'_ij <Wy_i, dy_j>
    const int i = t.erow[e], j = t.col[e];
    float s0 = 0.f;
    for (int f = 0; f < F; ++f) s0 = fmaf(Wy[f * N + i], dy[f * N + j], s0);
    tmp[e] = t.val[e] * s0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += PT) {                      // softmax and leaky_relu backward, row sums
    const int p0 = t.rptr[i], p1 = t.rptr[i + 1];
    float sd = 0.f;
    for (int p = p0; p < p1; ++p) sd = fmaf(al[p], tmp[p], sd);
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
      float ds = al[p] * (tmp[p] - sd);
      ds *= (cc[i] + rr[t.col[p]] > 0.f) ? 1.f : 0.2f;
      tmp[p] = ds; acc += ds;
    }
    dcc[i] = acc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += PT) {                      // column sums
    float acc = 0.f;
    for (int q = t.cptr[j]; q < t.cptr[j + 1]; ++q) acc += tmp[t.ceid[q]];
    drr[j] = acc;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F * N; o += PT) {                  // d Wy
    const int f = o / N, i = o - f * N;
    float s0 = fmaf(mix[f], drr[i], mix[F + f] * dcc[i]);
    for (int p = t.rptr[i]; p < t.rptr[i + 1]; ++p) s0 = fmaf(t.val[p] * al[p], dy[f * N + t.col[p]], s0);
    dWy[o] = s0;
  }
  for (int h = threadIdx.x; h < 2 * F; h += PT) {                  // d mixer
    const float* dv = h < F ? drr : dcc;
    const int f = h < F ? h : h - F;
    float s0 = 0.f;
    for (int n = 0; n < N; ++n) s0 = fmaf(dv[n], Wy[f * N + n], s0);
    gmix[h] += s0;
  }
  __syncthreads();
  wgrad_acc<NB>(gW, dWy, y, F, F, N);                              // d W[f][g] += sum_n dWy[f][n] y[g][n]
  for (int o = threadIdx.x; o < F * N; o += PT) {                  // d y[g][n] = sum_f W[f][g] dWy[f][n]
    const int g = o / N, n = o - g * N;
    float s0 = 0.f, s1 = 0.f;
    int f = 0;
    for (; f + 1 < F; f += 2) { s0 = fmaf(W[f * F + g], dWy[f * N + n], s0); s1 = fmaf(W[(f + 1) * F + g], dWy[(f + 1) * N + n], s1); }
    if (f < F) s0 = fmaf(W[f * F + g], dWy[f * N + n], s0);
    dyin[o] = s0 + s1;
  }
  __syncthreads();
}

template <int NB, int SG, bool QZ>
__global__ void __launch_bounds__(PT, 1) persist_bwd_k(const Args a) {
  constexpr bool NODE = SG == 1, EDGE = SG == 2;
  extern __shared__ __align__(16) float psm[];
  const int N = a.N, F = a.F, FN = F * N, GN = a.G * N, NQ = N / NB;
  const int nA = F * a.Kin * a.G, nB = F * a.Kst * F;
  const int KCa = a.Kin * a.G, KCb = a.Kst * F;
  const long long b = blockIdx.x;
  Weights w, gacc;
  float* p = carve(w, psm, a);
  p = carve(gacc, p, a);                                          // gradient accumulators, same layout as the weights
  float* zx = p; p += zx_floats(a.Kin, a.G, N);
  float* zh = p; p += (size_t)a.Kst * FN;
  float* da = p; p += FN;
  float* dr = p; p += FN;
  float* dh = p; p += FN;
  float* b1 = p; p += FN;
  float* b2 = p; p += FN;
  float* c0 = p; p += a.tg ? 2 * FN : 0;
  float* dc0 = p; p += a.tg ? 2 * FN : 0;
  float* dpu = p; p += (a.tg || NODE) ? FN : 0;
  float* c0n = p; p += NODE ? 2 * FN : 0;
  float* dc0n = p; p += NODE ? 2 * FN : 0;
  float* sbuf = p; p += NODE ? FN : 0;
  float* vch = p; p += NODE ? a.Kst * N : 0;
  float* qs = p; p += NODE ? 2 * N : 0;
  float* dq = p; p += NODE ? 2 * N : 0;
  float* ya = p; p += EDGE ? FN : 0;
  float* yr = p; p += EDGE ? FN : 0;
  float* dpb = p; p += EDGE ? FN : 0;
  float* Wy = p; p += EDGE ? FN : 0;
  float* og = p; p += EDGE ? FN : 0;                             // a gate's attention output, then its gradient dy in place
  float* dWy = p; p += EDGE ? FN : 0;
  float* rr = p; p += EDGE ? N : 0;
  float* cc = p; p += EDGE ? N : 0;
  float* drr = p; p += EDGE ? N : 0;
  float* dcc = p; p += EDGE ? N : 0;
  float* al = p; p += EDGE ? ((a.annz + 3) & ~3) : 0;
  float* etmp = p; p += EDGE ? ((a.annz + 3) & ~3) : 0;
  float* red = p; p += 64;
  unsigned char* sp = reinterpret_cast<unsigned char*>(p);
  const List fw = stage_list(sp, a.cptr, a.cidx, a.cval, N, a.nnz, a.lists_smem != 0);
  const List bw = stage_list(sp, a.rptr, a.ridx, a.rval, N, a.nnz, a.lists_smem != 0);
  Att att{};
  if (EDGE) att = stage_att(sp, a);
  load_weights(w, a);
  for (float* q = gacc.A; q < zx; q += PT) { if (q + threadIdx.x < zx) q[threadIdx.x] = 0.f; }      // zero every accumulator
  for (int e = threadIdx.x; e < FN; e += PT) {
    dh[e] = 0.f;
    if (a.tg) { dc0[e] = 0.f; dc0[FN + e] = 0.f; }
    if (NODE) { dc0n[e] = 0.f; dc0n[FN + e] = 0.f; }
  }
  float dtc[2] = {0.f, 0.f}, dnc[2] = {0.f, 0.f};
  __syncthreads();
  const bool gated = a.tg || NODE;
  if (gated) {                                                    // c0 of every gate sub-cell (needed to recompute their states)
    put_state<QZ>(zh, a.h0 + b * FN, F, N);
    __syncthreads();
    chain_z<QZ>(fw, zh, a.Kst, F, N);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (a.tg) subcell_c0<NB, QZ>(w.ts[g], zh, c0 + g * FN, F, KCb, N);
      if (NODE) subcell_c0<NB, QZ>(w.ns[g], zh, c0n + g * FN, F, KCb, N);
    }
    __syncthreads();
  }
  for (long long t = a.T - 1; t >= 0; --t) {
    const float* hprev = t > 0 ? a.H + (b * a.T + t - 1) * FN : a.h0 + b * FN;
    const float* xt = a.X + (b * a.T + t) * GN;
    put_state<QZ>(zh, hprev, F, N);
    for (int e = threadIdx.x; e < GN; e += PT) zx[e] = xt[e];
    if (NODE)
      for (int e = threadIdx.x; e < 2 * N; e += PT) {
        const int g = e / N, n = e - g * N;
        qs[e] = a.qn[(((long long)g * a.B + b) * a.T + t) * N + n];
        dq[e] = 0.f;
      }
    __syncthreads();
    chain_z<QZ>(fw, zh, a.Kst, F, N);
    chain(fw, zx, a.Kin, a.G, N);
    float gi = 1.f, gf = 1.f;
    if (a.tg) { gi = a.gt[((long long)0 * a.B + b) * a.T + t]; gf = a.gt[((long long)1 * a.B + b) * a.T + t]; }
    const float* Ht = a.H + (b * a.T + t) * FN;
    const bool has_dH = !(a.dh_last_only && t < a.T - 1);
    const float* dHt = a.dH + b * a.dH_bstride + (a.dh_last_only ? 0 : t * a.dH_tstride);
    float sgi = 0.f, sgf = 0.f;
    for (int e = threadIdx.x; e < F * NQ; e += PT) {
      const int f = e / NQ, n0 = (e - f * NQ) * NB;
      float av[NB], rv[NB];
      contract<NB>(w.A, zx, f, n0, KCa, N, av);
      contract_z<NB, QZ>(w.Bw, zh, f, n0, KCb, N, rv);
      const float bb = w.bias[f];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int n = n0 + j, o = f * N + n;
        const float h = Ht[o];
        const float dp = ((has_dH ? dHt[o] : 0.f) + dh[o]) * (1.f - h * h);
        if (EDGE) { ya[o] = av[j] + bb; yr[o] = rv[j] + bb; dpb[o] = dp; continue; }
        const float qi = NODE ? qs[n] : 1.f, qf = NODE ? qs[N + n] : 1.f;
        const float ta = dp * (av[j] + bb), tr = dp * (rv[j] + bb);      // d pre / d (g_i q_i), d pre / d (g_f q_f)
        sgi = fmaf(ta, qi, sgi); sgf = fmaf(tr, qf, sgf);
        if (NODE) { atomicAdd(dq + n, gi * ta); atomicAdd(dq + N + n, gf * tr); }
        da[o] = gi * qi * dp; dr[o] = gf * qf * dp;
      }
    }
    if (EDGE) {                                                  // recompute each attention layer, then its reverse
      __syncthreads();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float* y = g == 0 ? ya : yr;
        gat_fwd<NB>(N, F, att, w.eW[g], w.em[g], y, Wy, rr, cc, al, og);
        const float gs = g == 0 ? gi : gf;
        float sg = 0.f;
        for (int o = threadIdx.x; o < FN; o += PT) {
          const float ov = og[o], dp = dpb[o];
          sg = fmaf(dp, fmaxf(ov, 0.f), sg);                       // d pre / d (time gate scalar)
          og[o] = ov > 0.f ? gs * dp : 0.f;
        }
        if (g == 0) sgi = sg; else sgf = sg;
        __syncthreads();
        gat_bwd<NB>(N, F, att, w.eW[g], w.em[g], gacc.eW[g], gacc.em[g], y, Wy, og, dWy, rr, cc, drr, dcc, al, etmp, g == 0 ? da : dr);
      }
    }
    float dgi = 0.f, dgf = 0.f;
    if (a.tg) { dgi = block_sum(sgi, red); dgf = block_sum(sgf, red); }
    __syncthreads();
    wgrad_acc<NB>(gacc.A, da, zx, F, KCa, N);
    wgrad_z<NB, QZ>(gacc.Bw, dr, zh, F, KCb, N);
    for (int f = threadIdx.x; f < F; f += PT) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s += da[(size_t)f * N + n] + dr[(size_t)f * N + n];
      gacc.bias[f] += s;
    }
    adjoint_z<NB, QZ>(a, bw, w.Bw, dr, b1, b2, dh, a.Kst, F, false);      // dh_{t-1} (recurrent part)
    if (a.tg) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float gv = g == 0 ? gi : gf;
        const float dl = (g == 0 ? dgi : dgf) * gv * (1.f - gv);
        dtc[g] += dl;
        for (int e = threadIdx.x; e < F * NQ; e += PT) {
          const int f = e / NQ, n0 = (e - f * NQ) * NB;
          float y[NB];
          contract<NB>(w.ts[g].A, zx, f, n0, KCa, N, y);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const int o = f * N + n0 + j;
            const float u = tanhf(y[j] + c0[g * FN + o]);
            gacc.tW[g][o] += dl * u;
            const float dd = dl * w.tW[g][o] * (1.f - u * u);
            dc0[g * FN + o] += dd;
            dpu[o] = dd;
          }
        }
        __syncthreads();
        wgrad_acc<NB>(gacc.ts[g].A, dpu, zx, F, KCa, N);
        __syncthreads();
      }
    }
    if (NODE) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        // d lin = dq q (1 - q);  v_k = v_{k-1} S^T (k < Kst) are the gradients of the head's tap outputs p_k
        float part = 0.f;
        for (int n = threadIdx.x; n < N; n += PT) {
          const float q = qs[g * N + n];
          const float dl = dq[g * N + n] * q * (1.f - q);
          vch[n] = dl;
          part += dl;
        }
        dnc[g] += block_sum(part, red);
        for (int k = 1; k < a.Kst; ++k) {
          const float* cur = vch + (size_t)(k - 1) * N;
          for (int n = threadIdx.x; n < N; n += PT) {
            float s = 0.f;
            const int p1 = bw.ptr[n + 1];
            for (int pp = bw.ptr[n]; pp < p1; ++pp) s = fmaf(bw.val[pp], cur[bw.idx(pp)], s);
            vch[(size_t)k * N + n] = s;
          }
          __syncthreads();
        }
        // recompute the sub-cell state s, then d wh[k][f] += sum_n v_k[n] s[f][n],  d pre_s = (sum_k wh[k][f] v_k[n]) (1 - s^2)
        for (int e = threadIdx.x; e < F * NQ; e += PT) {
          const int f = e / NQ, n0 = (e - f * NQ) * NB;
          float y[NB];
          contract<NB>(w.ns[g].A, zx, f, n0, KCa, N, y);
#pragma unroll
          for (int j = 0; j < NB; ++j) sbuf[f * N + n0 + j] = tanhf(y[j] + c0n[g * FN + f * N + n0 + j]);
        }
        __syncthreads();
        wgrad_acc<NB>(gacc.nh[g], vch, sbuf, a.Kst, F, N);
        for (int o = threadIdx.x; o < FN; o += PT) {
          const int f = o / N, n = o - f * N;
          float ds = 0.f;
          for (int k = 0; k < a.Kst; ++k) ds = fmaf(w.nh[g][k * F + f], vch[(size_t)k * N + n], ds);
          const float sv = sbuf[o];
          const float dd = ds * (1.f - sv * sv);
          dc0n[g * FN + o] += dd;
          dpu[o] = dd;
        }
        __syncthreads();
        wgrad_acc<NB>(gacc.ns[g].A, dpu, zx, F, KCa, N);
        __syncthreads();
      }
    }
  }
  // ---- T-invariant gate terms: after t = 0 the zh buffers hold h0's chain ---------------------------------------------------
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (a.tg) subcell_c0_bwd<NB, QZ>(a, bw, w.ts[g], gacc.ts[g], dc0 + (size_t)g * FN, zh, b1, b2, dh);
    if (NODE) subcell_c0_bwd<NB, QZ>(a, bw, w.ns[g], gacc.ns[g], dc0n + (size_t)g * FN, zh, b1, b2, dh);
  }
  if (a.dh0) for (int e = threadIdx.x; e < FN; e += PT) a.dh0[b * FN + e] = dh[e];
  // ---- one atomicAdd per parameter and CTA ---------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < nA; i += PT) if (a.dA) atomicAdd(a.dA + i, gacc.A[i]);
  for (int i = threadIdx.x; i < nB; i += PT) if (a.dBw) atomicAdd(a.dBw + i, gacc.Bw[i]);
  for (int i = threadIdx.x; i < F; i += PT) if (a.dbias) atomicAdd(a.dbias + i, gacc.bias[i]);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (a.tg) {
      flush_sub(gacc.ts[g], a.dtA[g], a.dtB[g], a.dtb[g], nA, nB, F);
      for (int i = threadIdx.x; i < FN; i += PT) if (a.dtW[g]) atomicAdd(a.dtW[g] + i, gacc.tW[g][i]);
      if (threadIdx.x == 0 && a.dtc[g]) atomicAdd(a.dtc[g], dtc[g]);
    }
    if (EDGE) {
      for (int i = threadIdx.x; i < F * F; i += PT) if (a.deW[g]) atomicAdd(a.deW[g] + i, gacc.eW[g][i]);
      for (int i = threadIdx.x; i < 2 * F; i += PT) if (a.dem[g]) atomicAdd(a.dem[g] + i, gacc.em[g][i]);
    }
    if (NODE) {
      flush_sub(gacc.ns[g], a.dnA[g], a.dnB[g], a.dnb[g], nA, nB, F);
      for (int i = threadIdx.x; i < a.Kst * F; i += PT) if (a.dnhw[g]) atomicAdd(a.dnhw[g] + i, gacc.nh[g][i]);
      if (threadIdx.x == 0 && a.dnhb[g]) atomicAdd(a.dnhb[g], dnc[g]);
    }
  }
}

}  // namespace persist
}  // namespace gcrnn
