// Node gates of the dense tensor-core path (Utils/graphML.py:2379-2407), batched over every (b, t):
//   s_t = tanh(A_n(S) x_t + c0),  c0 = B_n(S) h0 + 2 b_n        (sub-cell state; c0 once per sequence, tap kernel TAP_PLAIN)
//   p_k[n] = sum_f wh[k][f] s_t[f][n]                             (the F -> 1 head contracted FIRST: node_gate_fwd_kernel)
//   lin = p_0 + (p_1 + (... p_{K-1} S ...) S) S,  q = sigmoid(lin + c)   (Horner on scalar node signals: K-1 shift GEMMs with
//                                                                         B*T rows - 1/64 of the state filter's rows per step)
// and the adjoint: v_0 = d lin, v_k = v_{k-1} S^T (shift GEMMs), then node_gate_bwd_kernel recomputes s_t and forms
//   d wh[k][f] += sum v_k s,  d pre_s = (sum_k wh[k][f] v_k)(1 - s^2),  d A_n += d pre_s (x) taps,  d c0 = sum_t d pre_s.
// The gate values enter the state update per node (tc_tap.cuh TAP_FWD epilogue) and leave the reverse sweep through dpre_kernel
// (tc_cell.cuh), which emits d lin directly.  Everything here is elementwise / MUFU-bound work next to the GEMMs.
#pragma once
#include "tc_gemm.cuh"

namespace gcrnn {
namespace tc {

constexpr int NG_KG = 8;        // Kin * G taps kept in registers
constexpr int NG_KMAX = 6;      // head taps
constexpr int NG_FC = 4;        // features per warp in the backward kernel

struct NodeGateArgs {
  const float* A;               // sub-cell input taps [F][KG]
  const float* wh;              // head taps [Kst][F]
  const float* X; const float* zx; long long zx_kstride;        // X [B,T,G,N]; x_t S^k (k >= 1) at zx + (k-1) * zx_kstride
  const float* c0;              // [B][F][N]
  int Kin, G, F, N, Kst, exact; long long B, T;
  float* p;                     // forward out: [Kst][B*T][N]
  const float* v;               // backward in:  [Kst][B*T][N] adjoint head signals
  float* dA; float* dwh;        // += [F][KG], [Kst][F]
  float* dc0;                   // = [B][F][N]
};

__device__ __forceinline__ float ng_tanh(float x, int exact) {
  if (exact) return tanh_acc(x);
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ng_load_taps(const NodeGateArgs& a, long long b, long long t, int n, float* z) {
  const int KG = a.Kin * a.G;
#pragma unroll
  for (int kg = 0; kg < NG_KG; ++kg) {
    if (kg < KG) {
      const int k = kg / a.G, g = kg - k * a.G;
      const size_t row = ((size_t)(b * a.T + t) * a.G + g) * a.N + n;
      z[kg] = k == 0 ? __ldg(a.X + row) : __ldg(a.zx + (size_t)(k - 1) * a.zx_kstride + row);
    } else {
      z[kg] = 0.f;
    }
  }
}

// per-(sample, node) base pointers of the taps at t = 0; step t is `tstride` floats further (address arithmetic out of the t loop:
// with it inside, index math was more than a third of the instructions of these kernels)
template <int KGT>
__device__ __forceinline__ void ng_tap_ptrs(const NodeGateArgs& a, long long b, int n, const float** zp) {
  const int KG = a.Kin * a.G;
#pragma unroll
  for (int kg = 0; kg < KGT; ++kg) {
    const int kq = kg < KG ? kg : 0;
    const int k = kq / a.G, g = kq - k * a.G;
    const size_t row = ((size_t)(b * a.T) * a.G + g) * a.N + n;
    zp[kg] = k == 0 ? a.X + row : a.zx + (size_t)(k - 1) * a.zx_kstride + row;
  }
}
template <int KGT>
__device__ __forceinline__ void ng_taps_at(const float* const* zp, size_t off, int KG, float* z) {
#pragma unroll
  for (int kg = 0; kg < KGT; ++kg) z[kg] = kg < KG ? __ldg(zp[kg] + off) : 0.f;
}

// shared-memory weight tables, 8 floats per feature each so that one feature costs two 16-byte broadcast loads per table
// (scalar loads made these kernels shared-memory-issue-bound: 13 LDS per feature against ~20 arithmetic instructions)
__device__ __forceinline__ void ng_stage_weights(const NodeGateArgs& a, float* sA, float* sW) {
  const int KG = a.Kin * a.G;
  for (int i = threadIdx.x; i < a.F * 8; i += blockDim.x) {
    const int f = i >> 3, j = i & 7;
    sA[i] = j < KG ? a.A[f * KG + j] : 0.f;
    sW[i] = j < a.Kst ? a.wh[j * a.F + f] : 0.f;                     // transposed: [f][k]
  }
}
__device__ __forceinline__ void ng_row8(const float* tab, int f, float* w) {
  const float4 lo = *reinterpret_cast<const float4*>(tab + f * 8), hi = *reinterpret_cast<const float4*>(tab + f * 8 + 4);
  w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
}

// thread <-> (b, n), loop over t and all features.  c0[b][f][n] is re-read per step (64 coalesced loads that stay in L1: 32 KB per
// block) instead of living in 64 registers: the register version ran 12 warps per SM and issued on 42 % of the cycles
// (profiles/r02_ncu_node_gate_fwd.raw.csv); the next step's taps are prefetched while this step computes.
template <int FMAX>
__global__ void __launch_bounds__(128) node_gate_fwd_kernel(const NodeGateArgs a) {
  extern __shared__ __align__(16) float ng_sm[];
  float* sA = ng_sm;                         // [F][8]
  float* sW = sA + a.F * 8;                  // [F][8]
  ng_stage_weights(a, sA, sW);
  __syncthreads();
  const int tiles_n = a.N / 128;
  const long long BT = a.B * a.T;
  for (long long item = blockIdx.x; item < a.B * tiles_n; item += gridDim.x) {
    const long long b = item / tiles_n;
    const int n = (int)(item - b * tiles_n) * 128 + threadIdx.x;
    const float* c0p = a.c0 + (size_t)b * a.F * a.N + n;
    const float* zp[NG_KG];
    ng_tap_ptrs<NG_KG>(a, b, n, zp);
    const size_t tstride = (size_t)a.G * a.N;
    const int KG = a.Kin * a.G;
    float z[NG_KG], zn[NG_KG];
    ng_taps_at<NG_KG>(zp, 0, KG, z);
    float* pout = a.p + ((size_t)b * a.T) * a.N + n;
    for (long long t = 0; t < a.T; ++t) {
      if (t + 1 < a.T) ng_taps_at<NG_KG>(zp, (size_t)(t + 1) * tstride, KG, zn);
      float pk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pk[k] = 0.f;
#pragma unroll 8
      for (int f = 0; f < a.F; ++f) {
        float wa[8], ww[8];
        ng_row8(sA, f, wa); ng_row8(sW, f, ww);
        float y = __ldg(c0p + (size_t)f * a.N);
#pragma unroll
        for (int kg = 0; kg < NG_KG; ++kg) y = fmaf(wa[kg], z[kg], y);
        const float s = ng_tanh(y, a.exact);
#pragma unroll
        for (int k = 0; k < NG_KMAX; ++k) pk[k] = fmaf(ww[k], s, pk[k]);
      }
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) if (k < a.Kst) pout[((size_t)k * BT + t) * a.N] = pk[k];
#pragma unroll
      for (int kg = 0; kg < NG_KG; ++kg) z[kg] = zn[kg];
    }
  }
}

// warp <-> chunk of NG_FC features, lane <-> node: a block of F / NG_FC warps covers every feature of 32 nodes of one sample, so the
// taps and head signals of a node are read from HBM once (the other warps hit L1).  Loop over t with the chunk's gradient
// accumulators in registers; they persist over all tiles of the block: one warp reduction and one atomicAdd per parameter at the end.
// KGT / KST: compile-time tap counts (exact for the common shapes, NG_KG / NG_KMAX padded otherwise: 37 vs 29 FMA-pipe
// instructions per (feature, step) at cfg3's Kin = Kst = 5).
template <int KGT, int KST>
__global__ void __launch_bounds__(512, 1) node_gate_bwd_kernel(const NodeGateArgs a) {
  extern __shared__ __align__(16) float ng_sm[];
  const int KG = a.Kin * a.G;
  float* sA = ng_sm;                         // [F][8]
  float* sW = sA + a.F * 8;                  // [F][8]
  ng_stage_weights(a, sA, sW);
  __syncthreads();
  const int tiles_n = a.N / 32;
  const long long BT = a.B * a.T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f0 = warp * NG_FC;
  float dA[NG_FC][KGT], dW[KST][NG_FC];
#pragma unroll
  for (int j = 0; j < NG_FC; ++j) {
#pragma unroll
    for (int kg = 0; kg < KGT; ++kg) dA[j][kg] = 0.f;
#pragma unroll
    for (int k = 0; k < KST; ++k) dW[k][j] = 0.f;
  }
  const size_t tstride = (size_t)a.G * a.N, kvs = (size_t)BT * a.N;
  for (long long item = blockIdx.x; item < a.B * tiles_n; item += gridDim.x) {
    const long long b = item / tiles_n;
    const int n = (int)(item - b * tiles_n) * 32 + lane;
    float c0[NG_FC], dc0[NG_FC];
#pragma unroll
    for (int j = 0; j < NG_FC; ++j) { c0[j] = __ldg(a.c0 + ((size_t)b * a.F + f0 + j) * a.N + n); dc0[j] = 0.f; }
    const float* zp[KGT];
    ng_tap_ptrs<KGT>(a, b, n, zp);
    const float* vp = a.v + ((size_t)b * a.T) * a.N + n;
    float z[KGT], vk[KST], zn[KGT], vn[KST];
    ng_taps_at<KGT>(zp, 0, KG, z);
#pragma unroll
    for (int k = 0; k < KST; ++k) vk[k] = k < a.Kst ? __ldg(vp + k * kvs) : 0.f;
    for (long long t = 0; t < a.T; ++t) {
      if (t + 1 < a.T) {                               // the next step's operands travel while this step computes
        ng_taps_at<KGT>(zp, (size_t)(t + 1) * tstride, KG, zn);
#pragma unroll
        for (int k = 0; k < KST; ++k) vn[k] = k < a.Kst ? __ldg(vp + k * kvs + (size_t)(t + 1) * a.N) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < NG_FC; ++j) {
        float wa[8], ww[8];
        ng_row8(sA, f0 + j, wa); ng_row8(sW, f0 + j, ww);
        float y0 = c0[j], y1 = 0.f;                      // two chains: the dependent FMA latency is what this kernel waits on
#pragma unroll
        for (int kg = 0; kg < KGT; ++kg) { if (kg & 1) y1 = fmaf(wa[kg], z[kg], y1); else y0 = fmaf(wa[kg], z[kg], y0); }
        const float s = ng_tanh(y0 + y1, a.exact);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int k = 0; k < KST; ++k) {
          if (k & 1) d1 = fmaf(ww[k], vk[k], d1); else d0 = fmaf(ww[k], vk[k], d0);
          dW[k][j] = fmaf(vk[k], s, dW[k][j]);
        }
        const float dps = (d0 + d1) * (1.f - s * s);
        dc0[j] += dps;
#pragma unroll
        for (int kg = 0; kg < KGT; ++kg) dA[j][kg] = fmaf(dps, z[kg], dA[j][kg]);
      }
#pragma unroll
      for (int kg = 0; kg < KGT; ++kg) z[kg] = zn[kg];
#pragma unroll
      for (int k = 0; k < KST; ++k) vk[k] = vn[k];
    }
#pragma unroll
    for (int j = 0; j < NG_FC; ++j) a.dc0[((size_t)b * a.F + f0 + j) * a.N + n] = dc0[j];
  }
#pragma unroll
  for (int j = 0; j < NG_FC; ++j) {
#pragma unroll
    for (int kg = 0; kg < KGT; ++kg) {
      float v = dA[j][kg];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && kg < KG) atomicAdd(a.dA + (size_t)(f0 + j) * KG + kg, v);
    }
#pragma unroll
    for (int k = 0; k < KST; ++k) {
      float v = dW[k][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && k < a.Kst) atomicAdd(a.dwh + (size_t)k * a.F + f0 + j, v);
    }
  }
}
inline size_t node_gate_smem_bytes(int F) { return (size_t)F * 16 * sizeof(float); }

// Input gradient through a gate sub-cell (only when the caller asks for dX): dxk[k][(b,t,g)][n] += sum_f A_s[f][k,g] d pre_s[f][n] with
//   time gate: d pre_s = dl[b,t] Wg[f][n] (1 - u^2)            node gate: d pre_s = (sum_k wh[k][f] v_k[n]) (1 - u^2)
// u recomputed once more (this path is rare: a cell that is not the first layer).  thread <-> (b, n), features in halves of 32 so that
// c0 / Wg stay in registers; a thread owns its (b, t, n) entries of dxk: plain read-modify-write, no atomics.
struct GateDxArgs {
  NodeGateArgs g;               // A, X, zx, c0, sizes; node mode: wh, v
  const float* Wg;              // time mode: [F][N]
  const float* dl;              // time mode: [B][T]
  float* dxk; long long dxk_kstride;     // [Kin][RX][N]
};
template <int MODE>             // 0: time gate, 1: node gate
__global__ void __launch_bounds__(128) gate_dx_kernel(const GateDxArgs q) {
  const NodeGateArgs& a = q.g;
  extern __shared__ __align__(16) float ng_sm[];
  float* sA = ng_sm;
  float* sW = sA + a.F * 8;
  if (MODE == 1) ng_stage_weights(a, sA, sW);
  else {
    const int KG = a.Kin * a.G;
    for (int i = threadIdx.x; i < a.F * 8; i += blockDim.x) { const int f = i >> 3, j = i & 7; sA[i] = j < KG ? a.A[f * KG + j] : 0.f; }
  }
  __syncthreads();
  constexpr int FH = 32;
  const int tiles_n = a.N / 128, KG = a.Kin * a.G;
  const long long BT = a.B * a.T;
  for (long long item = blockIdx.x; item < a.B * tiles_n; item += gridDim.x) {
    const long long b = item / tiles_n;
    const int n = (int)(item - b * tiles_n) * 128 + threadIdx.x;
    for (int fh = 0; fh < a.F; fh += FH) {
      float c0[FH], wg[FH];
#pragma unroll
      for (int j = 0; j < FH; ++j) {
        const bool ok = fh + j < a.F;
        c0[j] = ok ? __ldg(a.c0 + ((size_t)b * a.F + fh + j) * a.N + n) : 0.f;
        wg[j] = (MODE == 0 && ok) ? __ldg(q.Wg + (size_t)(fh + j) * a.N + n) : 0.f;
      }
      for (long long t = 0; t < a.T; ++t) {
        float z[NG_KG], vk[NG_KMAX], acc[NG_KG];
        ng_load_taps(a, b, t, n, z);
#pragma unroll
        for (int kg = 0; kg < NG_KG; ++kg) acc[kg] = 0.f;
        float dlv = 0.f;
        if (MODE == 0) dlv = __ldg(q.dl + b * a.T + t);
        else {
#pragma unroll
          for (int k = 0; k < NG_KMAX; ++k) vk[k] = k < a.Kst ? __ldg(a.v + ((size_t)k * BT + b * a.T + t) * a.N + n) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < FH; ++j) {
          if (fh + j < a.F) {
            float wa[8];
            ng_row8(sA, fh + j, wa);
            float y = c0[j];
#pragma unroll
            for (int kg = 0; kg < NG_KG; ++kg) y = fmaf(wa[kg], z[kg], y);
            const float u = ng_tanh(y, a.exact);
            float ds;
            if (MODE == 0) ds = dlv * wg[j];
            else {
              float ww[8];
              ng_row8(sW, fh + j, ww);
              ds = 0.f;
#pragma unroll
              for (int k = 0; k < NG_KMAX; ++k) ds = fmaf(ww[k], vk[k], ds);
            }
            const float dps = ds * (1.f - u * u);
#pragma unroll
            for (int kg = 0; kg < NG_KG; ++kg) acc[kg] = fmaf(wa[kg], dps, acc[kg]);
          }
        }
#pragma unroll
        for (int kg = 0; kg < NG_KG; ++kg) {
          if (kg < KG) {
            const int k = kg / a.G, g = kg - k * a.G;
            float* o = q.dxk + (size_t)k * q.dxk_kstride + ((size_t)(b * a.T + t) * a.G + g) * a.N + n;
            *o += acc[kg];
          }
        }
      }
    }
  }
}

// Horner step on scalar node signals: out = shifted + p  (fp32, n4 float4s), or the last one: q = sigmoid(shifted + p + c)
__global__ void node_head_add_kernel(const float4* __restrict__ shifted, const float4* __restrict__ p, float4* __restrict__ out, long long n4,
                                     const float* __restrict__ c, int sigmoid) {
  const float cv = (sigmoid && c) ? __ldg(c) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = p[i];
    if (shifted) { const float4 s = shifted[i]; v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
    if (sigmoid) {
      v.x = 1.f / (1.f + __expf(-(v.x + cv))); v.y = 1.f / (1.f + __expf(-(v.y + cv)));
      v.z = 1.f / (1.f + __expf(-(v.z + cv))); v.w = 1.f / (1.f + __expf(-(v.w + cv)));
    }
    out[i] = v;
  }
}
// out[0] += sum x
__global__ void sum_all_kernel(const float* __restrict__ x, float* out, long long n) {
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

}  // namespace tc
}  // namespace gcrnn
