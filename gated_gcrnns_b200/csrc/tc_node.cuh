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
constexpr int NG_FC = 8;        // features per thread in the backward kernel

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

// thread <-> (b, n): c0[b][:][n] in registers, loop over t.  Block = 128 nodes of one sample.
template <int FMAX>
__global__ void __launch_bounds__(128) node_gate_fwd_kernel(const NodeGateArgs a) {
  extern __shared__ float ng_sm[];
  const int KG = a.Kin * a.G;
  float* sA = ng_sm;                         // [F][NG_KG]
  float* sW = sA + a.F * NG_KG;              // [Kst][F]
  for (int i = threadIdx.x; i < a.F * NG_KG; i += blockDim.x) { const int f = i / NG_KG, kg = i - f * NG_KG; sA[i] = kg < KG ? a.A[f * KG + kg] : 0.f; }
  for (int i = threadIdx.x; i < a.Kst * a.F; i += blockDim.x) sW[i] = a.wh[i];
  __syncthreads();
  const int tiles_n = a.N / 128;
  const long long BT = a.B * a.T;
  for (long long item = blockIdx.x; item < a.B * tiles_n; item += gridDim.x) {
    const long long b = item / tiles_n;
    const int n = (int)(item - b * tiles_n) * 128 + threadIdx.x;
    float c0[FMAX];
#pragma unroll
    for (int f = 0; f < FMAX; ++f) c0[f] = f < a.F ? __ldg(a.c0 + ((size_t)b * a.F + f) * a.N + n) : 0.f;
    for (long long t = 0; t < a.T; ++t) {
      float z[NG_KG];
      ng_load_taps(a, b, t, n, z);
      float pk[NG_KMAX];
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) pk[k] = 0.f;
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        if (f < a.F) {
          float y = c0[f];
#pragma unroll
          for (int kg = 0; kg < NG_KG; ++kg) y = fmaf(sA[f * NG_KG + kg], z[kg], y);
          const float s = ng_tanh(y, a.exact);
#pragma unroll
          for (int k = 0; k < NG_KMAX; ++k) if (k < a.Kst) pk[k] = fmaf(sW[k * a.F + f], s, pk[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) if (k < a.Kst) a.p[((size_t)k * BT + b * a.T + t) * a.N + n] = pk[k];
    }
  }
}

// thread <-> (b, n, chunk of NG_FC features): loop over t with the chunk's gradient accumulators in registers; one block reduction and
// one atomicAdd per parameter and block at the end.  Block = 128 nodes of one (sample, feature chunk).
__global__ void __launch_bounds__(128) node_gate_bwd_kernel(const NodeGateArgs a) {
  extern __shared__ float ng_sm[];
  const int KG = a.Kin * a.G;
  float* sA = ng_sm;                         // [F][NG_KG]
  float* sW = sA + a.F * NG_KG;              // [Kst][F]
  float* sR = sW + a.Kst * a.F;              // reduction scratch [4 warps][NG_FC * (NG_KG + NG_KMAX)]
  for (int i = threadIdx.x; i < a.F * NG_KG; i += blockDim.x) { const int f = i / NG_KG, kg = i - f * NG_KG; sA[i] = kg < KG ? a.A[f * KG + kg] : 0.f; }
  for (int i = threadIdx.x; i < a.Kst * a.F; i += blockDim.x) sW[i] = a.wh[i];
  __syncthreads();
  const int tiles_n = a.N / 128, chunks = a.F / NG_FC;
  const long long BT = a.B * a.T;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long item = blockIdx.x; item < a.B * tiles_n * chunks; item += gridDim.x) {
    const int fc = (int)(item % chunks);
    const long long bn = item / chunks, b = bn / tiles_n;
    const int n = (int)(bn - b * tiles_n) * 128 + threadIdx.x;
    const int f0 = fc * NG_FC;
    float c0[NG_FC], dc0[NG_FC], dA[NG_FC][NG_KG], dW[NG_KMAX][NG_FC];
#pragma unroll
    for (int j = 0; j < NG_FC; ++j) {
      c0[j] = __ldg(a.c0 + ((size_t)b * a.F + f0 + j) * a.N + n); dc0[j] = 0.f;
#pragma unroll
      for (int kg = 0; kg < NG_KG; ++kg) dA[j][kg] = 0.f;
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) dW[k][j] = 0.f;
    }
    for (long long t = 0; t < a.T; ++t) {
      float z[NG_KG], vk[NG_KMAX];
      ng_load_taps(a, b, t, n, z);
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) vk[k] = k < a.Kst ? __ldg(a.v + ((size_t)k * BT + b * a.T + t) * a.N + n) : 0.f;
#pragma unroll
      for (int j = 0; j < NG_FC; ++j) {
        float y = c0[j];
#pragma unroll
        for (int kg = 0; kg < NG_KG; ++kg) y = fmaf(sA[(f0 + j) * NG_KG + kg], z[kg], y);
        const float s = ng_tanh(y, a.exact);
        float ds = 0.f;
#pragma unroll
        for (int k = 0; k < NG_KMAX; ++k) if (k < a.Kst) { ds = fmaf(sW[k * a.F + f0 + j], vk[k], ds); dW[k][j] = fmaf(vk[k], s, dW[k][j]); }
        const float dps = ds * (1.f - s * s);
        dc0[j] += dps;
#pragma unroll
        for (int kg = 0; kg < NG_KG; ++kg) dA[j][kg] = fmaf(dps, z[kg], dA[j][kg]);
      }
    }
#pragma unroll
    for (int j = 0; j < NG_FC; ++j) a.dc0[((size_t)b * a.F + f0 + j) * a.N + n] = dc0[j];
    // block reduction of the parameter gradients
    constexpr int PER = NG_FC * (NG_KG + NG_KMAX);
#pragma unroll
    for (int j = 0; j < NG_FC; ++j) {
#pragma unroll
      for (int kg = 0; kg < NG_KG; ++kg) {
        float v = dA[j][kg];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sR[warp * PER + j * NG_KG + kg] = v;
      }
#pragma unroll
      for (int k = 0; k < NG_KMAX; ++k) {
        float v = dW[k][j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sR[warp * PER + NG_FC * NG_KG + k * NG_FC + j] = v;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PER; i += blockDim.x) {
      const float v = sR[i] + sR[PER + i] + sR[2 * PER + i] + sR[3 * PER + i];
      if (i < NG_FC * NG_KG) {
        const int j = i / NG_KG, kg = i - j * NG_KG;
        if (kg < KG) atomicAdd(a.dA + (size_t)(f0 + j) * KG + kg, v);
      } else {
        const int r = i - NG_FC * NG_KG, k = r / NG_FC, j = r - k * NG_FC;
        if (k < a.Kst) atomicAdd(a.dwh + (size_t)k * a.F + f0 + j, v);
      }
    }
    __syncthreads();
  }
}
inline size_t node_gate_smem_bytes(int F, int Kst, bool bwd) {
  return ((size_t)F * NG_KG + (size_t)Kst * F + (bwd ? 4 * NG_FC * (NG_KG + NG_KMAX) : 0)) * sizeof(float);
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
