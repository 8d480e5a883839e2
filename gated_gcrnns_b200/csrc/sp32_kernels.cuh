// Fused fp32 kernels of the sparse EDGE-GATED recurrence for F == 32 state features (the cfg5 shape of SURVEY.md §8d:
// CSR kNN graph, N = 1e5).  One warp owns one (sample, node) pair and lane == feature, so a node's signal is one
// coalesced 128-byte row and every neighbour gather is one full cache line.  Blocks walk CONTIGUOUS node ranges so
// that, with a locality-preserving node ordering, neighbour rows are re-used out of L1/L2 instead of HBM.
//
// Reference op sites (Utils/graphML.py): LSIGF shift :123 + contraction :134-139; graphAttention :586-625 with
// S' = S + I :577, leaky_relu(0.2) :603, masked softmax over j :611-622, aggregation over i :625; relu :2101;
// cell update h = tanh(Q_i(a) + Q_f(r)) :2402-2423.
//
// Algebra used to cut passes over HBM (exact in real arithmetic, fp32 rounding differs in the last bits only):
//   * Wu = W (sum_k B_k z_k + b) = sum_k (W B_k) z_k + W b : the filter output itself is never materialised;
//   * the weight gradients of BOTH the filter taps and the attention weight follow from one accumulated
//     outer product  M_k = sum_n dWu[n] (x) z_k[n]:   dB_k = W^T M_k,  dW = sum_k M_k B_k^T + (sum_n dWu) b^T;
//   * dh_{t-1} = sum_k B_k^T (d S^T^k) with d = W^T dWu : the adjoint chain shifts ONE signal, then contracts.
#pragma once
#include "common.cuh"

namespace gcrnn {
namespace e32 {

constexpr int F = 32;
constexpr int MAXKG = 16;           // Kin * G input taps handled by the fused kernels
constexpr int MAXK = 5;             // state taps
struct Chain { const float* p[MAXK]; };
struct ChainMut { float* p[MAXK]; };

// small per-call scratch of accumulated reductions (floats)
struct AccLayout {
  static constexpr int M = 0;                          // [Kst][32 g][32 f]
  static constexpr int MA = MAXK * 32 * 32;            // [MAXKG][32 f]
  static constexpr int SUMA = MA + MAXKG * 32;
  static constexpr int SUMR = SUMA + 32;
  static constexpr int M1A = SUMR + 32;                // sum_n dr_n Wu_a[n,f]
  static constexpr int M2A = M1A + 32;                 // sum_n dc_n Wu_a[n,f]
  static constexpr int M1R = M2A + 32;
  static constexpr int M2R = M1R + 32;
  static constexpr int TOTAL = M2R + 32;
};
// folded forward weights (floats): Cr[k][g][f] = (W_r B_k)[f,g], cr0 = W_r b, Ca[kg][f] = (W_a A)[f,kg], ca0 = W_a b
struct PrepLayout {
  static constexpr int CR = 0;
  static constexpr int CR0 = MAXK * 32 * 32;
  static constexpr int CA = CR0 + 32;
  static constexpr int CA0 = CA + MAXKG * 32;
  static constexpr int TOTAL = CA0 + 32;
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.2f * x; }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float hsum(float v) {     // sum within each 16-lane half
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum four per-lane values over the warp with 6 shuffles: totals land on lanes 0 (v0), 16 (v1), 8 (v2), 24 (v3).
__device__ __forceinline__ float wsum4(float v0, float v1, float v2, float v3, int lane) {
  const bool b4 = lane & 16, b3 = lane & 8;
  float t = (b4 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, b4 ? v0 : v1, 16);
  float u = (b4 ? v3 : v2) + __shfl_xor_sync(0xffffffffu, b4 ? v2 : v3, 16);
  float k = (b3 ? u : t) + __shfl_xor_sync(0xffffffffu, b3 ? t : u, 8);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
  return k;
}
// prod[s] summed over lanes, result of slot L on lane L (31 shuffles for 32 slots)
template <int M>
__device__ __forceinline__ void tree_step(float (&prod)[32], int lane) {
  const bool up = lane & M;
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const float lo = prod[k], hi = prod[k + M];
    prod[k] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, M);
  }
}
__device__ __forceinline__ float tree32(float (&prod)[32], int lane) {
  tree_step<16>(prod, lane); tree_step<8>(prod, lane); tree_step<4>(prod, lane); tree_step<2>(prod, lane); tree_step<1>(prod, lane);
  return prod[0];
}

// contiguous task range of this block; (r, n) of a task advanced incrementally (no 64-bit division per task)
struct Walk {
  long long task, hi, r;
  int n, N, step;
  __device__ __forceinline__ Walk(long long total, int N_, int warp, int nwarps) : N(N_), step(nwarps) {
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per;
    hi = lo + per < total ? lo + per : total;
    task = lo + warp;
    r = task / N; n = (int)(task - r * N);
  }
  __device__ __forceinline__ bool ok() const { return task < hi; }
  __device__ __forceinline__ void next() {
    task += step; n += step;
    while (n >= N) { n -= N; ++r; }
  }
};

// acc = sum_p val[p] * rows[idx[p]][lane] over the edges p0..p1 of one destination node (sequential order in p)
__device__ __forceinline__ float gather_row(const int* __restrict__ idx, const float* __restrict__ val, int p0, int p1,
                                            const float* __restrict__ rows /* + sample offset + lane */, int lane) {
  float acc = 0.f;
  for (int pb = p0; pb < p1; pb += 32) {
    const int cnt = min(32, p1 - pb);
    int mi = 0; float mv = 0.f;
    if (lane < cnt) { mi = __ldg(idx + pb + lane); mv = __ldg(val + pb + lane); }
    int q = 0;
    for (; q + 8 <= cnt; q += 8) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = __ldg(rows + (size_t)__shfl_sync(0xffffffffu, mi, q + u) * 32);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fmaf(__shfl_sync(0xffffffffu, mv, q + u), x[u], acc);
    }
    for (; q < cnt; ++q)
      acc = fmaf(__shfl_sync(0xffffffffu, mv, q), __ldg(rows + (size_t)__shfl_sync(0xffffffffu, mi, q) * 32), acc);
  }
  return acc;
}

// ---- sparse shift of a 32-channel node-major signal: out[r,d,:] = sum_p val[p] in[r, idx[p], :] ----------------
__global__ void __launch_bounds__(256) spmm32_k(const int* __restrict__ ptr, const int* __restrict__ idx,
                                                const float* __restrict__ val, const float* __restrict__ in,
                                                float* __restrict__ out, int N, long long RN) {
  const int lane = threadIdx.x & 31;
  for (Walk w(RN, N, threadIdx.x >> 5, blockDim.x >> 5); w.ok(); w.next())
    out[w.task * 32 + lane] = gather_row(idx, val, __ldg(ptr + w.n), __ldg(ptr + w.n + 1), in + w.r * N * 32 + lane, lane);
}

// ---- folded forward weights (one block) -------------------------------------------------------------------------
__global__ void prep_k(const float* __restrict__ A, const float* __restrict__ Bw, const float* __restrict__ bias,
                       const float* __restrict__ Wa, const float* __restrict__ Wr, float* __restrict__ prep, int KG, int Kst) {
  const int KF = Kst * 32;
  for (int o = threadIdx.x; o < KF * 32; o += blockDim.x) {          // o = (k*32+g)*32 + f
    const int f = o & 31, kg = o >> 5;
    float a = 0.f;
    for (int m = 0; m < 32; ++m) a = fmaf(Wr[f * 32 + m], Bw[m * KF + kg], a);
    prep[PrepLayout::CR + o] = a;
  }
  for (int o = threadIdx.x; o < KG * 32; o += blockDim.x) {
    const int f = o & 31, kg = o >> 5;
    float a = 0.f;
    for (int m = 0; m < 32; ++m) a = fmaf(Wa[f * 32 + m], A[m * KG + kg], a);
    prep[PrepLayout::CA + o] = a;
  }
  if (threadIdx.x < 32) {
    const int f = threadIdx.x;
    float a = 0.f, b = 0.f;
    if (bias) for (int m = 0; m < 32; ++m) { a = fmaf(Wr[f * 32 + m], bias[m], a); b = fmaf(Wa[f * 32 + m], bias[m], b); }
    prep[PrepLayout::CR0 + f] = a; prep[PrepLayout::CA0 + f] = b;
  }
}

// ---- forward, stage 1: both filters folded into the attention mixing; scores ------------------------------------
// Wu_r = sum_k Cr_k z_k + cr0 (z_{KST-1} gathered on the fly), Wu_a = Ca x-taps + ca0, rc = (a1.Wu_a, a2.Wu_a, a1.Wu_r, a2.Wu_r)
template <int KST>
__global__ void __launch_bounds__(128) filter_fwd_k(const int* __restrict__ ptr, const int* __restrict__ idx, const float* __restrict__ val,
                                                    Chain zc /* z_0 .. z_{KST-2}, [B,N,32] */, Chain xs /* x taps of this step, [B,N,G] */,
                                                    int Kin, int G, const float* __restrict__ prep,
                                                    const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                    float* __restrict__ wu_a, float* __restrict__ wu_r, float4* __restrict__ rc,
                                                    int N, long long RN) {
  __shared__ __align__(16) float zs[4][KST][32];
  __shared__ float cas[MAXKG * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int KG = Kin * G;
  for (int i = threadIdx.x; i < KG * 32; i += blockDim.x) cas[i] = prep[PrepLayout::CA + i];
  float cr[KST * 32];
#pragma unroll
  for (int i = 0; i < KST * 32; ++i) cr[i] = prep[PrepLayout::CR + i * 32 + lane];
  const float cr0 = prep[PrepLayout::CR0 + lane], ca0 = prep[PrepLayout::CA0 + lane];
  const float a1a = mix_a[lane], a2a = mix_a[32 + lane], a1r = mix_r[lane], a2r = mix_r[32 + lane];
  __syncthreads();
  for (Walk w(RN, N, warp, blockDim.x >> 5); w.ok(); w.next()) {
    const long long o = w.task * 32 + lane;
#pragma unroll
    for (int k = 0; k < KST - 1; ++k) zs[warp][k][lane] = zc.p[k][o];
    if (KST > 1)
      zs[warp][KST - 1][lane] = gather_row(idx, val, __ldg(ptr + w.n), __ldg(ptr + w.n + 1), zc.p[KST > 1 ? KST - 2 : 0] + w.r * N * 32 + lane, lane);
    __syncwarp();
    float ar = cr0;
#pragma unroll
    for (int k = 0; k < KST; ++k)
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(&zs[warp][k][g4 * 4]);
        ar = fmaf(cr[k * 32 + g4 * 4 + 0], v.x, ar); ar = fmaf(cr[k * 32 + g4 * 4 + 1], v.y, ar);
        ar = fmaf(cr[k * 32 + g4 * 4 + 2], v.z, ar); ar = fmaf(cr[k * 32 + g4 * 4 + 3], v.w, ar);
      }
    __syncwarp();
    float aa = ca0;
    for (int k = 0; k < Kin; ++k) {
      const float* xp = xs.p[k] + w.task * G;
      for (int g = 0; g < G; ++g) aa = fmaf(cas[(k * G + g) * 32 + lane], __ldg(xp + g), aa);
    }
    wu_a[o] = aa; wu_r[o] = ar;
    const float s = wsum4(a1a * aa, a2a * aa, a1r * ar, a2r * ar, lane);   // lanes 0, 16, 8, 24
    if ((lane & 7) == 0) reinterpret_cast<float*>(rc + w.task)[(lane >> 4) | ((lane >> 2) & 2)] = s;
  }
}

// ---- forward, stage 2: per-row softmax statistics of both gates --------------------------------------------------
// info[2*(r*N+i) + gate] = (r_i, c_i, max_j e_ij, 1 / sum_j exp(e_ij - max)),  e_ij = leaky(c_i + r_j) over the pattern of S + I
__global__ void __launch_bounds__(256) rowstats_k(const int* __restrict__ rptr, const int* __restrict__ col, const float4* __restrict__ rc,
                                                  float4* __restrict__ info, int N, long long RN) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < RN; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / N; const int i = (int)(t - r * N);
    const float4* rcr = rc + r * N;
    const float4 me = rcr[i];
    const int p0 = __ldg(rptr + i), p1 = __ldg(rptr + i + 1);
    float ma = -INFINITY, mr = -INFINITY;
    for (int p = p0; p < p1; ++p) {
      const float4 o = rcr[__ldg(col + p)];
      ma = fmaxf(ma, leaky(me.y + o.x)); mr = fmaxf(mr, leaky(me.w + o.z));
    }
    float da = 0.f, dr = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float4 o = rcr[__ldg(col + p)];
      da += expf(leaky(me.y + o.x) - ma); dr += expf(leaky(me.w + o.z) - mr);
    }
    info[2 * t] = make_float4(me.x, me.y, ma, 1.f / da);
    info[2 * t + 1] = make_float4(me.z, me.w, mr, 1.f / dr);
  }
}

// ---- forward, stage 3: attention aggregation of both gates + relu + tanh update -----------------------------------
// y_g[j,:] = relu( sum_{i -> j} S'_ij alpha^g_ij Wu_g[i,:] ),  h = tanh(y_a + y_r);  masks = sign bits of the two relus
__global__ void __launch_bounds__(256) aggregate_k(const int* __restrict__ cptr, const int* __restrict__ crow, const float* __restrict__ cval,
                                                   const float4* __restrict__ info, const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                   float* __restrict__ hn, uint2* __restrict__ masks, int N, long long RN) {
  const int lane = threadIdx.x & 31;
  for (Walk w(RN, N, threadIdx.x >> 5, blockDim.x >> 5); w.ok(); w.next()) {
    const long long rb = w.r * N;
    const float rja = info[2 * w.task].x, rjr = info[2 * w.task + 1].x;
    const float* pa = wu_a + rb * 32 + lane;
    const float* pr = wu_r + rb * 32 + lane;
    const int q0 = __ldg(cptr + w.n), q1 = __ldg(cptr + w.n + 1);
    float acc_a = 0.f, acc_r = 0.f;
    for (int qb = q0; qb < q1; qb += 32) {
      const int cnt = min(32, q1 - qb);
      int mi = 0; float ca = 0.f, cr = 0.f;
      if (lane < cnt) {
        mi = __ldg(crow + qb + lane);
        const float v = __ldg(cval + qb + lane);
        const float4 sa = info[2 * (rb + mi)], sr = info[2 * (rb + mi) + 1];
        ca = v * (expf(leaky(sa.y + rja) - sa.z) * sa.w);
        cr = v * (expf(leaky(sr.y + rjr) - sr.z) * sr.w);
      }
      int q = 0;
      for (; q + 4 <= cnt; q += 4) {
        float xa[4], xr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t off = (size_t)__shfl_sync(0xffffffffu, mi, q + u) * 32;
          xa[u] = __ldg(pa + off); xr[u] = __ldg(pr + off);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc_a = fmaf(__shfl_sync(0xffffffffu, ca, q + u), xa[u], acc_a);
          acc_r = fmaf(__shfl_sync(0xffffffffu, cr, q + u), xr[u], acc_r);
        }
      }
      for (; q < cnt; ++q) {
        const size_t off = (size_t)__shfl_sync(0xffffffffu, mi, q) * 32;
        acc_a = fmaf(__shfl_sync(0xffffffffu, ca, q), __ldg(pa + off), acc_a);
        acc_r = fmaf(__shfl_sync(0xffffffffu, cr, q), __ldg(pr + off), acc_r);
      }
    }
    const unsigned ba = __ballot_sync(0xffffffffu, acc_a > 0.f), br = __ballot_sync(0xffffffffu, acc_r > 0.f);
    hn[w.task * 32 + lane] = tanhf(fmaxf(acc_a, 0.f) + fmaxf(acc_r, 0.f));
    if (lane == 0) masks[w.task] = make_uint2(ba, br);
  }
}

// ---- backward, stage 0: dpre[r,n,f] = (dH[b,t,f,n] + dhrec[r,n,f]) (1 - h^2)   (tile transpose of the reference layout) ----
__global__ void __launch_bounds__(256) dpre_k(const float* __restrict__ dHt /* dH + t*F*N */, long long sample_stride,
                                              const float* __restrict__ dhrec, const float* __restrict__ hn,
                                              float* __restrict__ dpre, int N, long long R) {
  __shared__ float tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_n = (N + 31) / 32;
  const long long tiles = R * tiles_n;
  for (long long tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const long long r = tl / tiles_n; const int n0 = (int)(tl - r * tiles_n) * 32;
    const float* src = dHt + r * sample_stride;
    for (int f = warp; f < 32; f += 8) tile[f][lane] = (n0 + lane < N) ? src[(size_t)f * N + n0 + lane] : 0.f;
    __syncthreads();
    for (int nn = warp; nn < 32; nn += 8)
      if (n0 + nn < N) {
        const long long o = (r * N + n0 + nn) * 32 + lane;
        const float h = hn[o];
        float g = tile[lane][nn];
        if (dhrec) g += dhrec[o];
        dpre[o] = g * (1.f - h * h);
      }
    __syncthreads();
  }
}

// ---- backward, stage 1 (per source row i, both gates): softmax / leaky backward, partial dWu ---------------------------
// lanes 0-15 hold the edges of the row for the input gate, lanes 16-31 the same edges for the forget gate.
// Requires row degree of S + I <= 32 (two chunks of 16 edges).
__global__ void __launch_bounds__(256) bwd_rows_k(const int* __restrict__ rptr, const int* __restrict__ col, const float* __restrict__ val,
                                                  const float4* __restrict__ info, const uint2* __restrict__ masks,
                                                  const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                  const float* __restrict__ dpre, const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                  float* __restrict__ pa, float* __restrict__ pr, float* __restrict__ dr /* [R*N][2], zeroed */,
                                                  float* __restrict__ acc, int N, long long RN) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, e = lane & 15;
  const float a2a = mix_a[32 + lane], a2r = mix_r[32 + lane];
  float m2a = 0.f, m2r = 0.f;
  for (Walk w(RN, N, warp, blockDim.x >> 5); w.ok(); w.next()) {
    const long long rb = w.r * N;
    const long long o = w.task * 32 + lane;
    const float wua = wu_a[o], wur = wu_r[o];
    const float4 mine = info[2 * w.task + half];
    const int p0 = __ldg(rptr + w.n), deg = __ldg(rptr + w.n + 1) - p0;
    const float* dp = dpre + rb * 32 + lane;
    float al[2], dal[2], slope[2]; int jj[2]; bool valid[2];
    float parta = 0.f, partr = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int cnt = min(16, max(0, deg - c * 16));          // warp-uniform
      valid[c] = e < cnt;
      int j = 0; float v = 0.f; unsigned mk = 0u; float coef = 0.f;
      al[c] = 0.f; slope[c] = 0.f; dal[c] = 0.f;
      if (valid[c]) {
        j = __ldg(col + p0 + c * 16 + e); v = __ldg(val + p0 + c * 16 + e);
        const float rj = info[2 * (rb + j) + half].x;
        const uint2 m2 = masks[rb + j];
        mk = half ? m2.y : m2.x;
        const float s = mine.y + rj;
        al[c] = expf(leaky(s) - mine.z) * mine.w;
        slope[c] = s > 0.f ? 1.f : 0.2f;
        coef = v * al[c];
      }
      jj[c] = j;
      if (cnt == 0) continue;
      float dot;
      if (cnt > 2) {
        float prod[32];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (q < cnt) {
            const int jq = __shfl_sync(0xffffffffu, j, q);
            const unsigned mka = __shfl_sync(0xffffffffu, mk, q), mkr = __shfl_sync(0xffffffffu, mk, 16 + q);
            const float ca = __shfl_sync(0xffffffffu, coef, q), cr = __shfl_sync(0xffffffffu, coef, 16 + q);
            const float d = __ldg(dp + (size_t)jq * 32);
            const float dya = ((mka >> lane) & 1u) ? d : 0.f, dyr = ((mkr >> lane) & 1u) ? d : 0.f;
            prod[q] = dya * wua; prod[16 + q] = dyr * wur;
            parta = fmaf(ca, dya, parta); partr = fmaf(cr, dyr, partr);
          } else { prod[q] = 0.f; prod[16 + q] = 0.f; }
        }
        dot = tree32(prod, lane);                          // lane (half, e): <dy_g[j_e], Wu_g[i]>
      } else {
        dot = 0.f;
        for (int q = 0; q < cnt; ++q) {
          const int jq = __shfl_sync(0xffffffffu, j, q);
          const unsigned mka = __shfl_sync(0xffffffffu, mk, q), mkr = __shfl_sync(0xffffffffu, mk, 16 + q);
          const float ca = __shfl_sync(0xffffffffu, coef, q), cr = __shfl_sync(0xffffffffu, coef, 16 + q);
          const float d = __ldg(dp + (size_t)jq * 32);
          const float dya = ((mka >> lane) & 1u) ? d : 0.f, dyr = ((mkr >> lane) & 1u) ? d : 0.f;
          parta = fmaf(ca, dya, parta); partr = fmaf(cr, dyr, partr);
          const float da = wsum(dya * wua), drr = wsum(dyr * wur);
          if (e == q) dot = half ? drr : da;
        }
      }
      dal[c] = v * dot;                                      // d alpha_ij = S'_ij <dy_j, Wu_i>
    }
    const float S = hsum(fmaf(al[0], dal[0], al[1] * dal[1]));
    const float ds0 = al[0] * (dal[0] - S) * slope[0], ds1 = al[1] * (dal[1] - S) * slope[1];
    const float dc = hsum(ds0 + ds1);                        // d c_i of this half's gate
    if (valid[0]) atomicAdd(dr + 2 * (rb + jj[0]) + half, ds0);
    if (valid[1]) atomicAdd(dr + 2 * (rb + jj[1]) + half, ds1);
    const float dca = __shfl_sync(0xffffffffu, dc, 0), dcr = __shfl_sync(0xffffffffu, dc, 16);
    pa[o] = fmaf(a2a, dca, parta); pr[o] = fmaf(a2r, dcr, partr);
    m2a = fmaf(dca, wua, m2a); m2r = fmaf(dcr, wur, m2r);
  }
  red[warp][lane] = m2a; red[warp][32 + lane] = m2r;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k][threadIdx.x];
    atomicAdd(acc + (threadIdx.x < 32 ? AccLayout::M2A + threadIdx.x : AccLayout::M2R + threadIdx.x - 32), s);
  }
}

// ---- backward, stage 2 (per node): finish dWu, accumulate the outer products, d = W_r^T dWu_r ----------------------------
template <int KST>
__global__ void __launch_bounds__(128) bwd_node_k(const int* __restrict__ ptr, const int* __restrict__ idx, const float* __restrict__ val,
                                                  Chain zc, Chain xs, int Kin, int G,
                                                  const float* __restrict__ pa, const float* __restrict__ pr, const float2* __restrict__ dr,
                                                  const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                  const float* __restrict__ mix_a, const float* __restrict__ mix_r, const float* __restrict__ Wr,
                                                  float* __restrict__ dout, float* __restrict__ acc, int N, long long RN) {
  __shared__ __align__(16) float zs[4][KST][32];
  __shared__ __align__(16) float dws[4][32];
  __shared__ float wrs[32 * 32];
  __shared__ float red[KST * 32 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int KG = Kin * G;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) wrs[i] = Wr[i];
  for (int i = threadIdx.x; i < KST * 1024; i += blockDim.x) red[i] = 0.f;
  const float a1a = mix_a[lane], a1r = mix_r[lane];
  float M[KST * 32];
#pragma unroll
  for (int i = 0; i < KST * 32; ++i) M[i] = 0.f;
  float Ma[MAXKG];
#pragma unroll
  for (int i = 0; i < MAXKG; ++i) Ma[i] = 0.f;
  float suma = 0.f, sumr = 0.f, m1a = 0.f, m1r = 0.f;
  __syncthreads();
  for (Walk w(RN, N, warp, blockDim.x >> 5); w.ok(); w.next()) {
    const long long o = w.task * 32 + lane;
    const float2 drn = dr[w.task];
    const float dwa = fmaf(a1a, drn.x, pa[o]), dwr = fmaf(a1r, drn.y, pr[o]);
    m1a = fmaf(drn.x, wu_a[o], m1a); m1r = fmaf(drn.y, wu_r[o], m1r);
    suma += dwa; sumr += dwr;
#pragma unroll
    for (int k = 0; k < KST - 1; ++k) zs[warp][k][lane] = zc.p[k][o];
    if (KST > 1)
      zs[warp][KST - 1][lane] = gather_row(idx, val, __ldg(ptr + w.n), __ldg(ptr + w.n + 1), zc.p[KST > 1 ? KST - 2 : 0] + w.r * N * 32 + lane, lane);
    dws[warp][lane] = dwr;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KST; ++k)
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(&zs[warp][k][g4 * 4]);
        M[k * 32 + g4 * 4 + 0] = fmaf(dwr, v.x, M[k * 32 + g4 * 4 + 0]); M[k * 32 + g4 * 4 + 1] = fmaf(dwr, v.y, M[k * 32 + g4 * 4 + 1]);
        M[k * 32 + g4 * 4 + 2] = fmaf(dwr, v.z, M[k * 32 + g4 * 4 + 2]); M[k * 32 + g4 * 4 + 3] = fmaf(dwr, v.w, M[k * 32 + g4 * 4 + 3]);
      }
    float d = 0.f;
#pragma unroll
    for (int f4 = 0; f4 < 8; ++f4) {
      const float4 v = *reinterpret_cast<const float4*>(&dws[warp][f4 * 4]);
      d = fmaf(wrs[(f4 * 4 + 0) * 32 + lane], v.x, d); d = fmaf(wrs[(f4 * 4 + 1) * 32 + lane], v.y, d);
      d = fmaf(wrs[(f4 * 4 + 2) * 32 + lane], v.z, d); d = fmaf(wrs[(f4 * 4 + 3) * 32 + lane], v.w, d);
    }
    dout[o] = d;
    __syncwarp();
#pragma unroll
    for (int kg = 0; kg < MAXKG; ++kg)
      if (kg < KG) Ma[kg] = fmaf(dwa, __ldg(xs.p[kg / G] + w.task * G + (kg % G)), Ma[kg]);
  }
  // block reduction (shared atomics), then one global atomic per output per block
#pragma unroll
  for (int i = 0; i < KST * 32; ++i) atomicAdd(&red[i * 32 + lane], M[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < KST * 1024; i += blockDim.x) atomicAdd(acc + AccLayout::M + i, red[i]);
#pragma unroll
  for (int kg = 0; kg < MAXKG; ++kg)
    if (kg < KG) atomicAdd(acc + AccLayout::MA + kg * 32 + lane, Ma[kg]);
  atomicAdd(acc + AccLayout::SUMA + lane, suma); atomicAdd(acc + AccLayout::SUMR + lane, sumr);
  atomicAdd(acc + AccLayout::M1A + lane, m1a); atomicAdd(acc + AccLayout::M1R + lane, m1r);
}

// ---- backward, stage 3: dh_{t-1}[n,g] = sum_k sum_f B[f,k,g] w_k[n,f],  w_0 = d, w_k = w_{k-1} S^T (last one gathered here) ----
template <int KST>
__global__ void __launch_bounds__(128) dh_k(const int* __restrict__ ptr, const int* __restrict__ idx, const float* __restrict__ val,
                                            Chain wc /* w_0 .. w_{KST-2} */, const float* __restrict__ Bw, float* __restrict__ dh,
                                            int N, long long RN) {
  __shared__ __align__(16) float zs[4][KST][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float bt[KST * 32];
#pragma unroll
  for (int k = 0; k < KST; ++k)
#pragma unroll
    for (int f = 0; f < 32; ++f) bt[k * 32 + f] = Bw[f * (KST * 32) + k * 32 + lane];
  for (Walk w(RN, N, warp, blockDim.x >> 5); w.ok(); w.next()) {
    const long long o = w.task * 32 + lane;
#pragma unroll
    for (int k = 0; k < KST - 1; ++k) zs[warp][k][lane] = wc.p[k][o];
    if (KST > 1)
      zs[warp][KST - 1][lane] = gather_row(idx, val, __ldg(ptr + w.n), __ldg(ptr + w.n + 1), wc.p[KST > 1 ? KST - 2 : 0] + w.r * N * 32 + lane, lane);
    __syncwarp();
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < KST; ++k)
#pragma unroll
      for (int f4 = 0; f4 < 8; ++f4) {
        const float4 v = *reinterpret_cast<const float4*>(&zs[warp][k][f4 * 4]);
        a = fmaf(bt[k * 32 + f4 * 4 + 0], v.x, a); a = fmaf(bt[k * 32 + f4 * 4 + 1], v.y, a);
        a = fmaf(bt[k * 32 + f4 * 4 + 2], v.z, a); a = fmaf(bt[k * 32 + f4 * 4 + 3], v.w, a);
      }
    dh[o] = a;
    __syncwarp();
  }
}

// ---- backward, last: turn the accumulated reductions into parameter gradients (one block; grads are += ) ---------------
__global__ void finalize_k(const float* __restrict__ acc, const float* __restrict__ A, const float* __restrict__ Bw,
                           const float* __restrict__ bias, const float* __restrict__ Wa, const float* __restrict__ Wr,
                           float* gA, float* gB, float* gbias, float* gmix_a, float* gW_a, float* gmix_r, float* gW_r, int KG, int Kst) {
  const int KF = Kst * 32;
  const float* M = acc + AccLayout::M;      // [(k*32+g)][f]
  const float* Ma = acc + AccLayout::MA;    // [kg][f]
  const float* suma = acc + AccLayout::SUMA;
  const float* sumr = acc + AccLayout::SUMR;
  if (gB)
    for (int o = threadIdx.x; o < 32 * KF; o += blockDim.x) {      // o = m*KF + kg
      const int m = o / KF, kg = o % KF;
      float a = 0.f;
      for (int f = 0; f < 32; ++f) a = fmaf(Wr[f * 32 + m], M[kg * 32 + f], a);
      gB[o] += a;
    }
  if (gW_r)
    for (int o = threadIdx.x; o < 1024; o += blockDim.x) {         // o = f*32 + m
      const int f = o >> 5, m = o & 31;
      float a = bias ? sumr[f] * bias[m] : 0.f;
      for (int kg = 0; kg < KF; ++kg) a = fmaf(M[kg * 32 + f], Bw[m * KF + kg], a);
      gW_r[o] += a;
    }
  if (gA)
    for (int o = threadIdx.x; o < 32 * KG; o += blockDim.x) {
      const int m = o / KG, kg = o % KG;
      float a = 0.f;
      for (int f = 0; f < 32; ++f) a = fmaf(Wa[f * 32 + m], Ma[kg * 32 + f], a);
      gA[o] += a;
    }
  if (gW_a)
    for (int o = threadIdx.x; o < 1024; o += blockDim.x) {
      const int f = o >> 5, m = o & 31;
      float a = bias ? suma[f] * bias[m] : 0.f;
      for (int kg = 0; kg < KG; ++kg) a = fmaf(Ma[kg * 32 + f], A[m * KG + kg], a);
      gW_a[o] += a;
    }
  if (gbias && threadIdx.x < 32) {
    const int m = threadIdx.x;
    float a = 0.f;
    for (int f = 0; f < 32; ++f) a = fmaf(Wa[f * 32 + m], suma[f], fmaf(Wr[f * 32 + m], sumr[f], a));
    gbias[m] += a;
  }
  if (threadIdx.x < 32) {
    const int f = threadIdx.x;
    if (gmix_a) { gmix_a[f] += acc[AccLayout::M1A + f]; gmix_a[32 + f] += acc[AccLayout::M2A + f]; }
    if (gmix_r) { gmix_r[f] += acc[AccLayout::M1R + f]; gmix_r[32 + f] += acc[AccLayout::M2R + f]; }
  }
}

}  // namespace e32
}  // namespace gcrnn
