// Fused fp32 kernels of the sparse EDGE-GATED recurrence for F == 32 state features (the cfg5 shape of SURVEY.md §8d:
// CSR kNN graph, N = 1e5).  One warp owns one (sample, node) pair, a node's signal is one coalesced 128-byte row and
// every neighbour gather is one full cache line.  Blocks walk CONTIGUOUS node ranges so that, with a locality-
// preserving node ordering, neighbour rows are re-used out of L1/L2 instead of HBM.
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
//
// Two lane layouts are used.  "F-layout": lane == feature (contractions, whose weights sit in registers).
// "Q-layout": lane = 8 g + c reads the 16-byte chunk c of row (4 i + g) — one LDS.128 / LDG.128 covers FOUR neighbour
// rows — and a 3-shuffle reduce-scatter over the four groups leaves feature FEAT(lane) = 4 c + g on each lane.
#pragma once
#include "common.cuh"

namespace gcrnn {
namespace e32 {

constexpr int F = 32;
constexpr int MAXKG = 16;           // Kin * G input taps handled by the fused kernels
constexpr int MAXK = 5;             // state taps
struct Chain { const float* p[MAXK]; };
struct Gather3 { const int* ptr; const int* idx; const float* val; };

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

__device__ __forceinline__ int feat_of_lane(int lane) { return 4 * (lane & 7) + (lane >> 3); }
__device__ __forceinline__ int lane_of_feat(int f) { return 8 * (f & 3) + (f >> 2); }
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.2f * x; }
__device__ __forceinline__ float hsum(float v) {     // sum within each 16-lane half
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float osum(float v) {     // sum within each 8-lane group
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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
// prod[s] summed over lanes, result of slot L on lane L (31 shuffles for 32 slots): the "transposed" warp reduction
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
// Q-layout reduce-scatter: every lane holds a float4 partial of chunk c = lane & 7; returns the total of feature
// FEAT(lane) = 4 c + g summed over the four groups g = lane >> 3  (3 shuffles)
__device__ __forceinline__ float rs4(const float4& a, int lane) {
  const bool hi = lane & 16, odd = lane & 8;
  float k0 = (hi ? a.z : a.x) + __shfl_xor_sync(0xffffffffu, hi ? a.x : a.z, 16);
  float k1 = (hi ? a.w : a.y) + __shfl_xor_sync(0xffffffffu, hi ? a.y : a.w, 16);
  return (odd ? k1 : k0) + __shfl_xor_sync(0xffffffffu, odd ? k0 : k1, 8);
}
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& x) {
  acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y); acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
// acc2 += w2 * (v0, v1) on the packed fp32 pipe (one issue slot for two FMAs)
__device__ __forceinline__ void fma2(float2& acc, const float2& w, float v0, float v1) { acc = __ffma2_rn(w, make_float2(v0, v1), acc); }
// tanh for x >= 0 (sum of two relus): 1 - 2 / (e^{2x} + 1) with MUFU exp / rcp; abs. error ~1e-7
__device__ __forceinline__ float tanh_pos(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND) : "memory"); }

// ---- software pipeline ---------------------------------------------------------------------------------------------
// A warp walking "row pointers -> edge list -> neighbour rows -> math" one task at a time exposes three dependent
// global-load latencies per task.  Instead each warp runs a 4-stage pipeline over ITS task sequence: iteration `it`
// loads the row pointers of task it+3 (S0), the edge list of task it+2 (S1), issues cp.async copies of every row task
// it+1 needs into a warp-private double-buffered shared-memory stage (S2), and computes task it (S3).
constexpr int CAP = 24;             // neighbour rows per task that go through the cp.async stage (a longer edge list: synchronous tail)
struct Tk { int r, n; };
// this warp's task sequence inside the block's CONTIGUOUS task range: first + j * nw, j = 0 .. niter-1
struct WarpTasks {
  Tk t0, t1, t2, t3;           // coordinates of the tasks of stages S0 .. S3 in the current iteration
  int niter, nw, N;
  __device__ __forceinline__ WarpTasks(long long total, int N_, int warp, int nwarps) : nw(nwarps), N(N_) {
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per;
    const long long hi = lo + per < total ? lo + per : total;
    const long long first = lo + warp;
    niter = hi > first ? (int)((hi - first + nwarps - 1) / nwarps) : 0;
    t0.r = (int)(first / N); t0.n = (int)(first - (long long)t0.r * N);
    t1 = t2 = t3 = t0;
  }
  __device__ __forceinline__ void advance() {          // needs nw <= N
    t3 = t2; t2 = t1; t1 = t0;
    t0.n += nw;
    if (t0.n >= N) { t0.n -= N; ++t0.r; }
  }
  __device__ __forceinline__ size_t row(const Tk& t) const { return (size_t)t.r * N + t.n; }      // task index
  __device__ __forceinline__ size_t sample(const Tk& t) const { return (size_t)t.r * N; }
};

// Pipeline for "NS streamed rows + one gathered row" tasks:  gathered[f] = sum_p val[p] * gsrc[r, idx[p], f].
// The first CAP edges of a destination go through the cp.async stage, a longer tail is fetched synchronously.
template <int NS>
struct GatherPipe {
  static constexpr int ROWS = NS + CAP;
  float* wbuf;                                  // this warp's stage: [2][ROWS][32]
  Gather3 op;
  const float* gsrc;                            // gathered array [R, N, 32]
  const float* srow[NS];                        // streamed arrays [R, N, 32]
  const float* aux; int aux_stride;             // optional per-lane scalar of the task: aux[task * aux_stride]
  int lane;
  int a_p0, a_deg, b_p0, b_deg, b_mi, c_p0, c_deg, d_p0, d_deg;
  float b_mv, c_mv, d_mv, c_aux, d_aux;

  __device__ __forceinline__ void init(float* buf, const Gather3& g, const float* gathered_from, int lane_) {
    wbuf = buf; op = g; gsrc = gathered_from; aux = nullptr; aux_stride = 0; lane = lane_;
    a_p0 = a_deg = b_p0 = b_deg = b_mi = c_p0 = c_deg = d_p0 = d_deg = 0;
    b_mv = c_mv = d_mv = c_aux = d_aux = 0.f;
    for (int i = lane; i < 2 * ROWS * 32; i += 32) buf[i] = 0.f;     // stale rows are multiplied by 0: keep them finite
    __syncwarp();
  }
  // runs S2, S1, S0 of iteration `it`; afterwards d_* describe task `it` (coordinates wt.t3)
  __device__ __forceinline__ void advance(int it, const WarpTasks& wt) {
    d_p0 = c_p0; d_deg = c_deg; d_mv = c_mv; d_aux = c_aux;
    {                                           // S2: task it + 1
      const int j = it + 1;
      c_p0 = b_p0; c_deg = b_deg; c_mv = b_mv; c_aux = 0.f;
      if (j >= 0 && j < wt.niter) {
        const size_t task = wt.row(wt.t2);
        float* dst = wbuf + (j & 1) * (ROWS * 32) + lane * 4;           // + row * 32
        const int g = lane >> 3, ch = (lane & 7) * 4;
#pragma unroll
        for (int k0 = 0; k0 < NS; k0 += 4) {
          const int k = k0 + g;
          const float* sp = srow[0];
#pragma unroll
          for (int m = 1; m < NS; ++m) if (k == m) sp = srow[m];
          if (k < NS) cp_async16(dst + k0 * 32, sp + task * 32 + ch);
        }
        const int cnt = min(b_deg, CAP);
        const float* gb = gsrc + wt.sample(wt.t2) * 32 + ch;
        float* gd = dst + NS * 32;
        for (int q0 = 0; q0 < cnt; q0 += 4) {
          const int src = __shfl_sync(0xffffffffu, b_mi, (q0 + g) & 31);
          if (q0 + g < cnt) cp_async16(gd + q0 * 32, gb + (size_t)src * 32);
        }
        if (aux) c_aux = __ldg(aux + task * aux_stride);
      }
      cp_commit();
    }
    {                                           // S1: task it + 2
      const int j = it + 2;
      b_p0 = a_p0; b_deg = a_deg; b_mi = 0; b_mv = 0.f;
      if (j >= 0 && j < wt.niter && lane < min(a_deg, 32)) { b_mi = __ldg(op.idx + a_p0 + lane); b_mv = __ldg(op.val + a_p0 + lane); }
    }
    {                                           // S0: task it + 3
      a_p0 = 0; a_deg = 0;
      if (it + 3 < wt.niter) { a_p0 = __ldg(op.ptr + wt.t0.n); a_deg = __ldg(op.ptr + wt.t0.n + 1) - a_p0; }
    }
  }
  // S3 helpers (call after advance(it), it >= 0)
  __device__ __forceinline__ const float* rows(int it) const { return wbuf + (it & 1) * (ROWS * 32); }
  // value of feature FEAT(lane) of the gathered row (waits for the stage of task `it`)
  __device__ __forceinline__ float gathered(int it, const WarpTasks& wt) {
    cp_wait<1>();
    __syncwarp();
    const int g = lane >> 3;
    const float4* gr = reinterpret_cast<const float4*>(rows(it) + NS * 32) + lane;   // row q0 + g, chunk c : + q0 * 8
    const int cnt = min(d_deg, CAP);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q0 = 0; q0 < cnt; q0 += 4)
      fma4(acc, __shfl_sync(0xffffffffu, d_mv, (q0 + g) & 31), gr[q0 * 8]);     // val = 0 beyond the edge list
    if (d_deg > CAP) {                           // rare long tail: synchronous
      const float4* base = reinterpret_cast<const float4*>(gsrc + wt.sample(wt.t3) * 32) + (lane & 7);
      for (int pb = d_p0 + CAP; pb < d_p0 + d_deg; pb += 32) {
        const int c2 = min(32, d_p0 + d_deg - pb);
        int mi = 0; float mv = 0.f;
        if (lane < c2) { mi = __ldg(op.idx + pb + lane); mv = __ldg(op.val + pb + lane); }
        for (int q0 = 0; q0 < c2; q0 += 4) {
          const int src = __shfl_sync(0xffffffffu, mi, (q0 + g) & 31);
          fma4(acc, __shfl_sync(0xffffffffu, mv, (q0 + g) & 31), __ldg(base + (size_t)src * 8));
        }
      }
    }
    return rs4(acc, lane);
  }
};

// ---- sparse shift of a 32-channel node-major signal: out[r,d,:] = sum_p val[p] in[r, idx[p], :] ----------------
// register pipeline (row pointers two tasks ahead, edge list one task ahead), Q-layout LDG.128 gathers
__global__ void __launch_bounds__(256) spmm32_k(Gather3 op, const float* __restrict__ in, float* __restrict__ out, int N, long long RN) {
  const int lane = threadIdx.x & 31, g = lane >> 3;
  const int fo = feat_of_lane(lane);
  WarpTasks wt(RN, N, threadIdx.x >> 5, blockDim.x >> 5);
  int a_p0 = 0, a_deg = 0, b_p0 = 0, b_deg = 0, b_mi = 0; float b_mv = 0.f;
  wt.advance();                                  // this kernel has 3 stages: t1 (row pointers), t2 (edge list), t3 (compute)
  for (int it = -2; it < wt.niter; ++it) {
    const int d_p0 = b_p0, d_deg = b_deg, d_mi = b_mi; const float d_mv = b_mv;
    b_p0 = a_p0; b_deg = a_deg; b_mi = 0; b_mv = 0.f;
    if (it + 1 >= 0 && it + 1 < wt.niter && lane < min(a_deg, 32)) { b_mi = __ldg(op.idx + a_p0 + lane); b_mv = __ldg(op.val + a_p0 + lane); }
    a_p0 = 0; a_deg = 0;
    if (it + 2 < wt.niter) { a_p0 = __ldg(op.ptr + wt.t1.n); a_deg = __ldg(op.ptr + wt.t1.n + 1) - a_p0; }
    if (it >= 0) {
      const float4* base = reinterpret_cast<const float4*>(in + wt.sample(wt.t3) * 32) + (lane & 7);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int pb = 0; pb < d_deg; pb += 32) {
        int mi = d_mi; float mv = d_mv;
        const int cnt = min(32, d_deg - pb);
        if (pb > 0) { mi = 0; mv = 0.f; if (lane < cnt) { mi = __ldg(op.idx + d_p0 + pb + lane); mv = __ldg(op.val + d_p0 + pb + lane); } }
        int q0 = 0;
        for (; q0 + 8 <= cnt; q0 += 8) {          // two independent 4-row gathers in flight
          const float4 x0 = __ldg(base + (size_t)__shfl_sync(0xffffffffu, mi, q0 + g) * 8);
          const float4 x1 = __ldg(base + (size_t)__shfl_sync(0xffffffffu, mi, q0 + 4 + g) * 8);
          fma4(acc, __shfl_sync(0xffffffffu, mv, q0 + g), x0);
          fma4(acc, __shfl_sync(0xffffffffu, mv, q0 + 4 + g), x1);
        }
        for (; q0 < cnt; q0 += 4)
          fma4(acc, __shfl_sync(0xffffffffu, mv, (q0 + g) & 31), __ldg(base + (size_t)__shfl_sync(0xffffffffu, mi, (q0 + g) & 31) * 8));
      }
      out[wt.row(wt.t3) * 32 + fo] = rs4(acc, lane);
    }
    wt.advance();
  }
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
  for (int o = threadIdx.x; o < MAXKG * 32; o += blockDim.x) {       // zero padded to MAXKG taps
    const int f = o & 31, kg = o >> 5;
    float a = 0.f;
    if (kg < KG) for (int m = 0; m < 32; ++m) a = fmaf(Wa[f * 32 + m], A[m * KG + kg], a);
    prep[PrepLayout::CA + o] = a;
  }
  if (threadIdx.x < 32) {
    const int f = threadIdx.x;
    float a = 0.f, b = 0.f;
    if (bias) for (int m = 0; m < 32; ++m) { a = fmaf(Wr[f * 32 + m], bias[m], a); b = fmaf(Wa[f * 32 + m], bias[m], b); }
    prep[PrepLayout::CR0 + f] = a; prep[PrepLayout::CA0 + f] = b;
  }
}

// lane kg < KG carries input tap kg = k * G + g of the task: pointer to it for task 0, stride G per task
__device__ __forceinline__ const float* tap_pointer(const Chain& xs, int lane, int KG, int G) {
  if (lane >= KG) return nullptr;
  const int k = lane / G;
  const float* xp = xs.p[0];
#pragma unroll
  for (int m = 1; m < MAXK; ++m) if (k == m) xp = xs.p[m];
  return xp + (lane - k * G);
}

// ---- forward, stage 1: both filters folded into the attention mixing; scores ------------------------------------
// Wu_r = sum_k Cr_k z_k + cr0 (z_{KST-1} gathered on the fly), Wu_a = Ca x-taps + ca0, rc = (a1.Wu_a, a2.Wu_a, a1.Wu_r, a2.Wu_r)
template <int KST>
__global__ void __launch_bounds__(128) filter_fwd_k(Gather3 gop, Chain zc /* z_0 .. z_{KST-2}, [B,N,32] */,
                                                    Chain xs /* x taps of this step, [B,N,G] */, int Kin, int G,
                                                    const float* __restrict__ prep,
                                                    const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                    float* __restrict__ wu_a, float* __restrict__ wu_r, float4* __restrict__ rc,
                                                    int N, long long RN) {
  constexpr int NS = KST - 1;
  using Pipe = GatherPipe<NS>;
  __shared__ __align__(16) float bufs[4][2 * Pipe::ROWS * 32];
  __shared__ __align__(16) float zl[4][32];
  __shared__ float cas[MAXKG * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fo = feat_of_lane(lane);
  const int KG = Kin * G;
  for (int i = threadIdx.x; i < MAXKG * 32; i += blockDim.x) cas[i] = prep[PrepLayout::CA + i];
  float2 cr[KST * 16];                                // cr[k*16 + i] = (Cr[k][2i][lane], Cr[k][2i+1][lane])
#pragma unroll
  for (int i = 0; i < KST * 16; ++i) cr[i] = make_float2(prep[PrepLayout::CR + (2 * i) * 32 + lane], prep[PrepLayout::CR + (2 * i + 1) * 32 + lane]);
  const float cr0 = prep[PrepLayout::CR0 + lane], ca0 = prep[PrepLayout::CA0 + lane];
  const float a1a = mix_a[lane], a2a = mix_a[32 + lane], a1r = mix_r[lane], a2r = mix_r[32 + lane];
  WarpTasks wt(RN, N, warp, blockDim.x >> 5);
  Pipe pipe;
  pipe.init(bufs[warp], gop, zc.p[NS - 1], lane);
#pragma unroll
  for (int k = 0; k < NS; ++k) pipe.srow[k] = zc.p[k];
  pipe.aux = tap_pointer(xs, lane, KG, G); pipe.aux_stride = G;
  __syncthreads();
  for (int it = -3; it < wt.niter; ++it, wt.advance()) {
    pipe.advance(it, wt);
    if (it < 0) continue;
    zl[warp][fo] = pipe.gathered(it, wt);
    __syncwarp();
    const float* rows = pipe.rows(it);
    float2 ar[4] = {make_float2(cr0, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // 4 independent chains
#pragma unroll
    for (int k = 0; k < KST; ++k) {
      const float* zr = k < NS ? rows + k * 32 : zl[warp];
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(zr + g4 * 4);
        fma2(ar[(2 * g4) & 3], cr[k * 16 + g4 * 2], v.x, v.y);
        fma2(ar[(2 * g4 + 1) & 3], cr[k * 16 + g4 * 2 + 1], v.z, v.w);
      }
    }
    const float arr = ((ar[0].x + ar[0].y) + (ar[1].x + ar[1].y)) + ((ar[2].x + ar[2].y) + (ar[3].x + ar[3].y));
    float aa = ca0;
    for (int kg = 0; kg < KG; kg += 4) {              // taps padded with zero weights / zero values
#pragma unroll
      for (int u = 0; u < 4; ++u) aa = fmaf(cas[(kg + u) * 32 + lane], __shfl_sync(0xffffffffu, pipe.d_aux, kg + u), aa);
    }
    const size_t o = wt.row(wt.t3) * 32 + lane;
    wu_a[o] = aa; wu_r[o] = arr;
    const float s4 = wsum4(a1a * aa, a2a * aa, a1r * arr, a2r * arr, lane);   // lanes 0, 16, 8, 24
    if ((lane & 7) == 0) reinterpret_cast<float*>(rc + wt.row(wt.t3))[(lane >> 4) | ((lane >> 2) & 2)] = s4;
    __syncwarp();
  }
  cp_wait<0>();
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
      da += __expf(leaky(me.y + o.x) - ma); dr += __expf(leaky(me.w + o.z) - mr);
    }
    info[2 * t] = make_float4(me.x, me.y, ma, 1.f / da);
    info[2 * t + 1] = make_float4(me.z, me.w, mr, 1.f / dr);
  }
}

// ---- forward, stage 3: attention aggregation of both gates + relu + tanh update -----------------------------------
// y_g[j,:] = relu( sum_{i -> j} S'_ij alpha^g_ij Wu_g[i,:] ),  h = tanh(y_a + y_r);  masks = sign bits of the two relus
// (bit l of a mask word belongs to feature FEAT(l)).  4 warps / block, dynamic shared memory: per warp
// [2 stages][2 gates][CAP rows][32] floats (48 KB per block).
constexpr int AGG_STAGE = 2 * CAP * 32;
__global__ void __launch_bounds__(128) aggregate_k(const int* __restrict__ cptr, const int* __restrict__ crow, const float* __restrict__ cval,
                                                   const float4* __restrict__ info, const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                   float* __restrict__ hn, uint2* __restrict__ masks, int N, long long RN) {
  extern __shared__ __align__(16) float dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, ch = (lane & 7) * 4, fo = feat_of_lane(lane);
  float* wbuf = dyn + (size_t)warp * 2 * AGG_STAGE;
  for (int i = lane; i < 2 * AGG_STAGE; i += 32) wbuf[i] = 0.f;
  __syncwarp();
  WarpTasks wt(RN, N, warp, blockDim.x >> 5);
  int a_p0 = 0, a_deg = 0, b_p0 = 0, b_deg = 0, b_mi = 0, c_p0 = 0, c_deg = 0;
  float b_v = 0.f, c_v = 0.f, c_rja = 0.f, c_rjr = 0.f;
  float4 c_sa = make_float4(0.f, 0.f, 0.f, 1.f), c_sr = c_sa;
  for (int it = -3; it < wt.niter; ++it, wt.advance()) {
    const int d_p0 = c_p0, d_deg = c_deg;
    const float d_v = c_v, d_rja = c_rja, d_rjr = c_rjr;
    const float4 d_sa = c_sa, d_sr = c_sr;
    {                                             // S2: task it + 1
      const int j = it + 1;
      c_p0 = b_p0; c_deg = b_deg; c_v = b_v;
      if (j >= 0 && j < wt.niter) {
        const size_t rb = wt.sample(wt.t2);
        float* dst = wbuf + (j & 1) * AGG_STAGE + lane * 4;
        const int cnt = min(b_deg, CAP);
        const float* ga = wu_a + rb * 32 + ch;
        const float* gr = wu_r + rb * 32 + ch;
        for (int q0 = 0; q0 < cnt; q0 += 4) {
          const int src = __shfl_sync(0xffffffffu, b_mi, (q0 + g) & 31);
          if (q0 + g < cnt) {
            cp_async16(dst + q0 * 32, ga + (size_t)src * 32);
            cp_async16(dst + (CAP + q0) * 32, gr + (size_t)src * 32);
          }
        }
        if (lane < cnt) { c_sa = info[2 * (rb + b_mi)]; c_sr = info[2 * (rb + b_mi) + 1]; }
        const size_t task = wt.row(wt.t2);
        c_rja = info[2 * task].x; c_rjr = info[2 * task + 1].x;
      }
      cp_commit();
    }
    {                                             // S1: task it + 2
      const int j = it + 2;
      b_p0 = a_p0; b_deg = a_deg; b_mi = 0; b_v = 0.f;
      if (j >= 0 && j < wt.niter && lane < min(a_deg, 32)) { b_mi = __ldg(crow + a_p0 + lane); b_v = __ldg(cval + a_p0 + lane); }
    }
    {                                             // S0: task it + 3
      a_p0 = 0; a_deg = 0;
      if (it + 3 < wt.niter) { a_p0 = __ldg(cptr + wt.t0.n); a_deg = __ldg(cptr + wt.t0.n + 1) - a_p0; }
    }
    if (it < 0) continue;
    // S3
    const int cnt = min(d_deg, CAP);
    float ca = 0.f, cr = 0.f;                      // alpha * S' of edge `lane` (0 beyond the edge list)
    if (lane < cnt) {
      ca = d_v * (__expf(leaky(d_sa.y + d_rja) - d_sa.z) * d_sa.w);
      cr = d_v * (__expf(leaky(d_sr.y + d_rjr) - d_sr.z) * d_sr.w);
    }
    cp_wait<1>();
    __syncwarp();
    const float4* ra = reinterpret_cast<const float4*>(wbuf + (it & 1) * AGG_STAGE) + lane;     // row q0 + g: + q0 * 8
    const float4* rr = ra + CAP * 8;
    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_r = acc_a;
    for (int q0 = 0; q0 < cnt; q0 += 4) {
      fma4(acc_a, __shfl_sync(0xffffffffu, ca, (q0 + g) & 31), ra[q0 * 8]);
      fma4(acc_r, __shfl_sync(0xffffffffu, cr, (q0 + g) & 31), rr[q0 * 8]);
    }
    if (d_deg > CAP) {                             // rare long tail (hub columns): synchronous
      const size_t rb = wt.sample(wt.t3);
      const float4* pa = reinterpret_cast<const float4*>(wu_a + rb * 32) + (lane & 7);
      const float4* pr = reinterpret_cast<const float4*>(wu_r + rb * 32) + (lane & 7);
      for (int qb = d_p0 + CAP; qb < d_p0 + d_deg; qb += 32) {
        const int c2 = min(32, d_p0 + d_deg - qb);
        int mi = 0; float xa = 0.f, xr = 0.f;
        if (lane < c2) {
          mi = __ldg(crow + qb + lane);
          const float v = __ldg(cval + qb + lane);
          const float4 sa = info[2 * (rb + mi)], sr = info[2 * (rb + mi) + 1];
          xa = v * (__expf(leaky(sa.y + d_rja) - sa.z) * sa.w);
          xr = v * (__expf(leaky(sr.y + d_rjr) - sr.z) * sr.w);
        }
        for (int q0 = 0; q0 < c2; q0 += 4) {
          const size_t off = (size_t)__shfl_sync(0xffffffffu, mi, (q0 + g) & 31) * 8;
          fma4(acc_a, __shfl_sync(0xffffffffu, xa, (q0 + g) & 31), __ldg(pa + off));
          fma4(acc_r, __shfl_sync(0xffffffffu, xr, (q0 + g) & 31), __ldg(pr + off));
        }
      }
    }
    const float ya = rs4(acc_a, lane), yr = rs4(acc_r, lane);          // feature FEAT(lane)
    const unsigned ba = __ballot_sync(0xffffffffu, ya > 0.f), br = __ballot_sync(0xffffffffu, yr > 0.f);
    const size_t task = wt.row(wt.t3);
    hn[task * 32 + fo] = tanh_pos(fmaxf(ya, 0.f) + fmaxf(yr, 0.f));
    if (lane == 0) masks[task] = make_uint2(ba, br);
    __syncwarp();
  }
  cp_wait<0>();
}

// ---- backward, stage 0: g = (dH[b,t,f,n] + dhrec[r,n,f]) (1 - h^2);  dya = g [y_a > 0], dyr = g [y_r > 0] ----------------
// (tile transpose of the reference layout; the relu masks of the two gates are applied here, once, instead of per edge)
__global__ void __launch_bounds__(256) dpre_k(const float* __restrict__ dHt /* dH + t*F*N */, long long sample_stride,
                                              const float* __restrict__ dhrec, const float* __restrict__ hn,
                                              const uint2* __restrict__ masks, float* __restrict__ dya, float* __restrict__ dyr,
                                              int N, long long R, int node_major /* 1: dHt is already [R][N][32] */) {
  __shared__ float tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bit = lane_of_feat(lane);
  const int tiles_n = (N + 31) / 32;
  const long long tiles = R * tiles_n;
  for (long long tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const long long r = tl / tiles_n; const int n0 = (int)(tl - r * tiles_n) * 32;
    const float* src = dHt + r * sample_stride;
    if (!node_major)
      for (int f = warp; f < 32; f += 8) tile[f][lane] = (n0 + lane < N) ? src[(size_t)f * N + n0 + lane] : 0.f;
    __syncthreads();
    for (int nn = warp; nn < 32; nn += 8)
      if (n0 + nn < N) {
        const long long task = r * N + n0 + nn;
        const long long o = task * 32 + lane;
        const float h = hn[o];
        float gq = node_major ? src[(size_t)(n0 + nn) * 32 + lane] : tile[lane][nn];
        if (dhrec) gq += dhrec[o];
        gq *= 1.f - h * h;
        const uint2 mk = masks[task];
        dya[o] = ((mk.x >> bit) & 1u) ? gq : 0.f;
        dyr[o] = ((mk.y >> bit) & 1u) ? gq : 0.f;
      }
    __syncthreads();
  }
}

// ---- backward, stage 1 (per source row i, both gates): softmax / leaky backward, partial dWu ---------------------------
// Edge-distributed data: lanes 0-15 hold edges e (and 16+e) of the row for the input gate, lanes 16-31 the same edges
// for the forget gate.  Row data: Q-layout.  Requires row degree of S + I <= 32.  4 warps / block, dynamic shared
// memory per warp: [2 stages][dya rows R | dyr rows R | Wu_a row | Wu_r row][32] floats, R = max row degree rounded up to 4.
__host__ __device__ inline int rows_stage_floats(int R) { return (2 * R + 2) * 32; }
__global__ void __launch_bounds__(128) bwd_rows_k(const int* __restrict__ rptr, const int* __restrict__ col, const float* __restrict__ val,
                                                  const float4* __restrict__ info, const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                  const float* __restrict__ dya, const float* __restrict__ dyr,
                                                  const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                  float* __restrict__ pa, float* __restrict__ pr, float* __restrict__ dr /* [R*N][2], zeroed */,
                                                  float* __restrict__ acc, int R, int N, long long RN) {
  extern __shared__ __align__(16) float dyn[];
  __shared__ float red[4][64];
  __shared__ float sdot[4][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, e = lane & 15;
  const int g = lane >> 3, ch = (lane & 7) * 4, fo = feat_of_lane(lane);
  const int ROWS_STAGE = rows_stage_floats(R);
  float* wbuf = dyn + (size_t)warp * 2 * ROWS_STAGE;
  for (int i = lane; i < 2 * ROWS_STAGE; i += 32) wbuf[i] = 0.f;
  __syncwarp();
  const float a2a = mix_a[32 + fo], a2r = mix_r[32 + fo];
  float m2a = 0.f, m2r = 0.f;
  WarpTasks wt(RN, N, warp, blockDim.x >> 5);
  int a_p0 = 0, a_deg = 0, b_deg = 0, b_j0 = 0, b_j1 = 0, c_deg = 0, c_j0 = 0, c_j1 = 0;
  float b_v0 = 0.f, b_v1 = 0.f, c_v0 = 0.f, c_v1 = 0.f, c_rj0 = 0.f, c_rj1 = 0.f;
  float4 c_mine = make_float4(0.f, 0.f, 0.f, 1.f);
  for (int it = -3; it < wt.niter; ++it, wt.advance()) {
    const int d_deg = c_deg, d_j0 = c_j0, d_j1 = c_j1;
    const float d_v0 = c_v0, d_v1 = c_v1, d_rj0 = c_rj0, d_rj1 = c_rj1;
    const float4 mine = c_mine;
    {                                             // S2: task it + 1
      const int j = it + 1;
      c_deg = b_deg; c_j0 = b_j0; c_j1 = b_j1; c_v0 = b_v0; c_v1 = b_v1;
      c_rj0 = c_rj1 = 0.f;
      if (j >= 0 && j < wt.niter) {
        const size_t rb = wt.sample(wt.t2), task = wt.row(wt.t2);
        float* dst = wbuf + (j & 1) * ROWS_STAGE + lane * 4;
        const float* ga = dya + rb * 32 + ch;
        const float* gr = dyr + rb * 32 + ch;
        for (int q0 = 0; q0 < b_deg; q0 += 4) {     // rows q0 + g; q0 < 16: chunk 0, else chunk 1
          const int src = __shfl_sync(0xffffffffu, q0 < 16 ? b_j0 : b_j1, (q0 + g) & 15);
          if (q0 + g < b_deg) {
            cp_async16(dst + q0 * 32, ga + (size_t)src * 32);
            cp_async16(dst + (R + q0) * 32, gr + (size_t)src * 32);
          }
        }
        if (g < 2) cp_async16(dst + 2 * R * 32, (g == 0 ? wu_a : wu_r) + task * 32 + ch);  // rows 2R (Wu_a), 2R+1 (Wu_r)
        if (e < b_deg) c_rj0 = info[2 * (rb + b_j0) + half].x;
        if (16 + e < b_deg) c_rj1 = info[2 * (rb + b_j1) + half].x;
        c_mine = info[2 * task + half];
      }
      cp_commit();
    }
    {                                             // S1: task it + 2
      const int j = it + 2;
      b_deg = a_deg; b_j0 = b_j1 = 0; b_v0 = b_v1 = 0.f;
      if (j >= 0 && j < wt.niter) {
        if (e < a_deg) { b_j0 = __ldg(col + a_p0 + e); b_v0 = __ldg(val + a_p0 + e); }
        if (16 + e < a_deg) { b_j1 = __ldg(col + a_p0 + 16 + e); b_v1 = __ldg(val + a_p0 + 16 + e); }
      }
    }
    {                                             // S0: task it + 3
      a_p0 = 0; a_deg = 0;
      if (it + 3 < wt.niter) { a_p0 = __ldg(rptr + wt.t0.n); a_deg = __ldg(rptr + wt.t0.n + 1) - a_p0; }
    }
    if (it < 0) continue;
    // S3 -- edge coefficients in the (half, e) layout
    float al[2], slope[2], coef[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const bool valid = c * 16 + e < d_deg;
      const float sc = mine.y + (c ? d_rj1 : d_rj0);
      al[c] = valid ? __expf(leaky(sc) - mine.z) * mine.w : 0.f;
      slope[c] = valid ? (sc > 0.f ? 1.f : 0.2f) : 0.f;
      coef[c] = (c ? d_v1 : d_v0) * al[c];
    }
    cp_wait<1>();
    __syncwarp();
    const float4* st = reinterpret_cast<const float4*>(wbuf + (it & 1) * ROWS_STAGE);
    const float4 wua4 = st[2 * R * 8 + (lane & 7)], wur4 = st[(2 * R + 1) * 8 + (lane & 7)];
    const float4* ra = st + lane;                  // row q0 + g, chunk c: + q0 * 8
    const float4* rr = ra + R * 8;
    float4 parta = make_float4(0.f, 0.f, 0.f, 0.f), partr = parta;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int cnt = min(16, max(0, d_deg - c * 16));          // warp-uniform
      for (int q0 = 0; q0 < cnt; q0 += 8) {         // rows q0 + g and q0 + 4 + g (stale rows beyond cnt: coef = 0, dots unused)
        const float4 xa = ra[(c * 16 + q0) * 8], xr = rr[(c * 16 + q0) * 8];
        const bool two = q0 + 4 < cnt;               // warp-uniform: the second 4-row group exists
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ya = two ? ra[(c * 16 + q0 + 4) * 8] : zero4, yr = two ? rr[(c * 16 + q0 + 4) * 8] : zero4;
        fma4(parta, __shfl_sync(0xffffffffu, coef[c], (q0 + g) & 15), xa);            // coef = 0 beyond the edge list
        fma4(partr, __shfl_sync(0xffffffffu, coef[c], 16 + ((q0 + g) & 15)), xr);
        fma4(parta, __shfl_sync(0xffffffffu, coef[c], (q0 + 4 + g) & 15), ya);
        fma4(partr, __shfl_sync(0xffffffffu, coef[c], 16 + ((q0 + 4 + g) & 15)), yr);
        float d0 = dot4(xa, wua4), d1 = dot4(xr, wur4), d2 = dot4(ya, wua4), d3 = dot4(yr, wur4);   // <dy_g[j], Wu_g[i]>
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {             // four independent 8-lane reductions
          d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
          d2 += __shfl_xor_sync(0xffffffffu, d2, o); d3 += __shfl_xor_sync(0xffffffffu, d3, o);
        }
        if ((lane & 7) == 0) {
          sdot[warp][0][c * 16 + q0 + g] = d0; sdot[warp][1][c * 16 + q0 + g] = d1;
          if (q0 + 4 < 16) { sdot[warp][0][c * 16 + q0 + 4 + g] = d2; sdot[warp][1][c * 16 + q0 + 4 + g] = d3; }
        }
      }
    }
    __syncwarp();
    const float dal0 = d_v0 * sdot[warp][half][e], dal1 = d_v1 * sdot[warp][half][16 + e];   // d alpha_ij = S'_ij <dy_j, Wu_i>
    const float S = hsum((e < d_deg ? al[0] * dal0 : 0.f) + (16 + e < d_deg ? al[1] * dal1 : 0.f));
    const float ds0 = e < d_deg ? al[0] * (dal0 - S) * slope[0] : 0.f;
    const float ds1 = 16 + e < d_deg ? al[1] * (dal1 - S) * slope[1] : 0.f;
    const float dc = hsum(ds0 + ds1);                        // d c_i of this half's gate
    const size_t rb = wt.sample(wt.t3);
    if (e < d_deg) atomicAdd(dr + 2 * (rb + d_j0) + half, ds0);
    if (16 + e < d_deg) atomicAdd(dr + 2 * (rb + d_j1) + half, ds1);
    const float dca = __shfl_sync(0xffffffffu, dc, 0), dcr = __shfl_sync(0xffffffffu, dc, 16);
    const size_t o = wt.row(wt.t3) * 32 + fo;
    const float* wrow = reinterpret_cast<const float*>(st + 2 * R * 8);
    pa[o] = fmaf(a2a, dca, rs4(parta, lane)); pr[o] = fmaf(a2r, dcr, rs4(partr, lane));
    m2a = fmaf(dca, wrow[fo], m2a); m2r = fmaf(dcr, wrow[32 + fo], m2r);
    __syncwarp();
  }
  cp_wait<0>();
  red[warp][fo] = m2a; red[warp][32 + fo] = m2r;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k][threadIdx.x];
    atomicAdd(acc + (threadIdx.x < 32 ? AccLayout::M2A + threadIdx.x : AccLayout::M2R + threadIdx.x - 32), s);
  }
}

// ---- backward, stage 2 (per node): finish dWu, accumulate the outer products, d = W_r^T dWu_r ----------------------------
template <int KST>
__global__ void __launch_bounds__(128, 3) bwd_node_k(Gather3 gop, Chain zc, Chain xs, int Kin, int G,
                                                  const float* __restrict__ pa, const float* __restrict__ pr, const float2* __restrict__ dr,
                                                  const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                  const float* __restrict__ mix_a, const float* __restrict__ mix_r, const float* __restrict__ Wr,
                                                  float* __restrict__ dout, float* __restrict__ acc, int N, long long RN) {
  constexpr int NZ = KST - 1;                  // streamed taps z_0 .. z_{KST-2}
  constexpr int NS = NZ + 4;                   // + pa, pr, Wu_a, Wu_r rows of the node
  using Pipe = GatherPipe<NS>;
  static_assert(4 * 2 * Pipe::ROWS * 32 >= KST * 32 * 32, "stage buffers double as the block reduction scratch");
  __shared__ __align__(16) float bufs[4][2 * Pipe::ROWS * 32];
  __shared__ __align__(16) float zl[4][32];
  __shared__ __align__(16) float dws[4][32];
  __shared__ __align__(8) float2 wrs[16 * 32];     // wrs[i*32 + g] = (Wr[2i][g], Wr[2i+1][g])
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fo = feat_of_lane(lane);
  const int KG = Kin * G;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) wrs[i] = make_float2(Wr[(2 * (i >> 5)) * 32 + (i & 31)], Wr[(2 * (i >> 5) + 1) * 32 + (i & 31)]);
  const float a1a = mix_a[lane], a1r = mix_r[lane];
  float2 M[KST * 16];                          // M[k*16 + i] = (M_k[f=lane][2i], M_k[f=lane][2i+1])
#pragma unroll
  for (int i = 0; i < KST * 16; ++i) M[i] = make_float2(0.f, 0.f);
  float Ma[MAXKG];
#pragma unroll
  for (int i = 0; i < MAXKG; ++i) Ma[i] = 0.f;
  float suma = 0.f, sumr = 0.f, m1a = 0.f, m1r = 0.f;
  WarpTasks wt(RN, N, warp, blockDim.x >> 5);
  Pipe pipe;
  pipe.init(bufs[warp], gop, zc.p[NZ - 1], lane);
#pragma unroll
  for (int k = 0; k < NZ; ++k) pipe.srow[k] = zc.p[k];
  pipe.srow[NZ] = pa; pipe.srow[NZ + 1] = pr; pipe.srow[NZ + 2] = wu_a; pipe.srow[NZ + 3] = wu_r;
  pipe.aux = tap_pointer(xs, lane, KG, G); pipe.aux_stride = G;      // lanes 0..KG-1: x taps; lanes 16 / 17: (dr_a, dr_r)
  if (lane == 16 || lane == 17) { pipe.aux = reinterpret_cast<const float*>(dr) + (lane - 16); pipe.aux_stride = 2; }
  __syncthreads();
  for (int it = -3; it < wt.niter; ++it, wt.advance()) {
    pipe.advance(it, wt);
    if (it < 0) continue;
    zl[warp][fo] = pipe.gathered(it, wt);
    const float* rows = pipe.rows(it);
    const float drx = __shfl_sync(0xffffffffu, pipe.d_aux, 16), dry = __shfl_sync(0xffffffffu, pipe.d_aux, 17);
    const float dwa = fmaf(a1a, drx, rows[NZ * 32 + lane]), dwr = fmaf(a1r, dry, rows[(NZ + 1) * 32 + lane]);
    m1a = fmaf(drx, rows[(NZ + 2) * 32 + lane], m1a); m1r = fmaf(dry, rows[(NZ + 3) * 32 + lane], m1r);
    suma += dwa; sumr += dwr;
    dws[warp][lane] = dwr;
    __syncwarp();
    const float2 dwr2 = make_float2(dwr, dwr);
#pragma unroll
    for (int k = 0; k < KST; ++k) {
      const float* zr = k < NZ ? rows + k * 32 : zl[warp];
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(zr + g4 * 4);
        M[k * 16 + g4 * 2] = __ffma2_rn(dwr2, make_float2(v.x, v.y), M[k * 16 + g4 * 2]);
        M[k * 16 + g4 * 2 + 1] = __ffma2_rn(dwr2, make_float2(v.z, v.w), M[k * 16 + g4 * 2 + 1]);
      }
    }
    float2 d2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int f4 = 0; f4 < 8; ++f4) {
      const float4 v = *reinterpret_cast<const float4*>(&dws[warp][f4 * 4]);
      fma2(d2[(2 * f4) & 3], wrs[(f4 * 2) * 32 + lane], v.x, v.y);
      fma2(d2[(2 * f4 + 1) & 3], wrs[(f4 * 2 + 1) * 32 + lane], v.z, v.w);
    }
    dout[wt.row(wt.t3) * 32 + lane] = ((d2[0].x + d2[0].y) + (d2[1].x + d2[1].y)) + ((d2[2].x + d2[2].y) + (d2[3].x + d2[3].y));
    const float tapv = lane < KG ? pipe.d_aux : 0.f;
#pragma unroll
    for (int kg = 0; kg < MAXKG; kg += 4)
      if (kg < KG) {
#pragma unroll
        for (int u = 0; u < 4; ++u) Ma[kg + u] = fmaf(dwa, __shfl_sync(0xffffffffu, tapv, kg + u), Ma[kg + u]);
      }
    __syncwarp();
  }
  cp_wait<0>();
  // block reduction in the (now idle) stage buffers, then one global atomic per output per block
  __syncthreads();
  float* red = &bufs[0][0];
  for (int i = threadIdx.x; i < KST * 1024; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < KST * 16; ++i) {           // M[k*16+i] covers g = 2i, 2i+1 of tap k: acc index ((k*32+g)*32 + f)
    atomicAdd(&red[(2 * i) * 32 + lane], M[i].x);
    atomicAdd(&red[(2 * i + 1) * 32 + lane], M[i].y);
  }
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
__global__ void __launch_bounds__(128) dh_k(Gather3 gop, Chain wc /* w_0 .. w_{KST-2} */, const float* __restrict__ Bw,
                                            float* __restrict__ dh, int N, long long RN) {
  constexpr int NS = KST - 1;
  using Pipe = GatherPipe<NS>;
  __shared__ __align__(16) float bufs[4][2 * Pipe::ROWS * 32];
  __shared__ __align__(16) float zl[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fo = feat_of_lane(lane);
  float2 bt[KST * 16];                           // bt[k*16 + i] = (B[2i, k, lane], B[2i+1, k, lane])
#pragma unroll
  for (int k = 0; k < KST; ++k)
#pragma unroll
    for (int i = 0; i < 16; ++i)
      bt[k * 16 + i] = make_float2(Bw[(2 * i) * (KST * 32) + k * 32 + lane], Bw[(2 * i + 1) * (KST * 32) + k * 32 + lane]);
  WarpTasks wt(RN, N, warp, blockDim.x >> 5);
  Pipe pipe;
  pipe.init(bufs[warp], gop, wc.p[NS - 1], lane);
#pragma unroll
  for (int k = 0; k < NS; ++k) pipe.srow[k] = wc.p[k];
  for (int it = -3; it < wt.niter; ++it, wt.advance()) {
    pipe.advance(it, wt);
    if (it < 0) continue;
    zl[warp][fo] = pipe.gathered(it, wt);
    __syncwarp();
    const float* rows = pipe.rows(it);
    float2 a[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};     // 4 independent chains
#pragma unroll
    for (int k = 0; k < KST; ++k) {
      const float* zr = k < NS ? rows + k * 32 : zl[warp];
#pragma unroll
      for (int f4 = 0; f4 < 8; ++f4) {
        const float4 v = *reinterpret_cast<const float4*>(zr + f4 * 4);
        fma2(a[(2 * f4) & 3], bt[k * 16 + f4 * 2], v.x, v.y);
        fma2(a[(2 * f4 + 1) & 3], bt[k * 16 + f4 * 2 + 1], v.z, v.w);
      }
    }
    dh[wt.row(wt.t3) * 32 + lane] = ((a[0].x + a[0].y) + (a[1].x + a[1].y)) + ((a[2].x + a[2].y) + (a[3].x + a[3].y));
    __syncwarp();
  }
  cp_wait<0>();
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
