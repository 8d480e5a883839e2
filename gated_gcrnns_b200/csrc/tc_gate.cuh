// Time gates of the gated GCRNN (Utils/graphML.py:2357-2374) for ALL (b, t) at once — they depend on (x_t, h0) only:
//     u[b,t,f,n] = tanh( sum_{k,g} A_g[f,k,g] (x_t S^k)[b,g,n] + c0[b,f,n] ),   logit[b,t] = sum_{f,n} Wg[f,n] u
// with c0 = B_g(S) h0 + 2 bias_g computed once per sequence.  These are B*T*F*N tanh evaluations per gate (2.1e9 at
// cfg3 per 512 sequences) — CUDA-core + MUFU work, so the kernels are organised around the issue slots:
//   * work item = (64-node tile, sample, chunk of time steps); each CTA walks a contiguous range of items, the
//     x_t S^k rows of the NEXT item are staged into shared memory by cp.async while the current one is computed;
//   * a thread owns one node and F/4 features; the Kin*G filter taps of its features live in REGISTERS (the kernel
//     is templated on Kin*G), so one evaluation is Kin*G FMA + MUFU.TANH + 1 FMA with Kin*G shared loads per time
//     step shared by all of the thread's features;
//   * per-(b,t) sums over the 32 lanes use the transposed warp reduction (31 shuffles per 32 values).
#pragma once
#include "tc_cell.cuh"
#include "tc_tap.cuh"

namespace gcrnn {
namespace tc {

constexpr int TG_NT = 64;          // nodes per work item
constexpr int TG_FQ = 4;           // feature groups per CTA (256 threads = 64 nodes x 4 groups)
constexpr int TG_FMAX = 16;        // features per thread (F <= 64)
constexpr int TG_FB = 4;           // backward: features processed together
constexpr int TG_NWMAX = 16;       // warps per CTA of the widest variant (forward with 8 feature groups = 512 threads)

struct GateArgs {
  const float* A; int Kin, G, F, N; long long B, T;
  const float* X;                 // [B,T,G,N]
  const float* zx;                // [Kin-1][B*T*G][N]
  const float* c0;                // [B,F,N]  (includes both bias terms)
  const float* Wg;                // [F*N]
  float* logit;                   // fwd: [B,T] += partial
  const float* dl;                // bwd: dlogit [B,T]
  float* dWg;                     // bwd: [F*N] +=
  float* dc0;                     // bwd: [B,F,N] = sum_t dpre_u  (+= when nchunks > 1)
  float* dA;                      // bwd: [F,Kin,G] +=
  int TC, nchunks, nbuf;          // time steps per item, items per (tile, sample), staging buffers
  int bchunk;                     // generic kernel: samples per CTA
  int exact;                      // 1: tanh through ex2 + rcp (abs. error ~2e-7; split-bf16 mode), 0: tanh.approx (2^-11)
};
template <bool EX> __device__ __forceinline__ float gate_tanh(float x) { return EX ? tanh_acc(x) : tanh_fast(x); }

inline size_t gate_smem_bytes(int TC, int KG, int F, int nbuf) {
  return ((size_t)nbuf * TC * KG * TG_NT + TG_NWMAX * (size_t)TC + 2 * (size_t)TC + 2 * (size_t)F * KG) * sizeof(float);
}

struct GateItem { int tile, t_lo, tn; long long b; };
__device__ __forceinline__ GateItem gate_item(const GateArgs& a, long long item) {
  const long long per_tile = a.B * a.nchunks;
  GateItem it;
  it.tile = (int)(item / per_tile);
  const long long r = item % per_tile;
  it.b = r / a.nchunks;
  it.t_lo = (int)(r % a.nchunks) * a.TC;
  it.tn = min(a.TC, (int)a.T - it.t_lo);
  return it;
}

// x_t S^k rows of one item -> zs[t][kg][64 nodes]  (cp.async, 16 B per request)
template <int KG, int NTH = 256>
__device__ __forceinline__ void gate_stage(const GateArgs& a, const GateItem& it, float* zs, int tid) {
  const size_t kstride = (size_t)a.B * a.T * a.G * a.N;
  const int total = it.tn * KG * (TG_NT / 4);
  for (int i = tid; i < total; i += NTH) {
    const int c4 = i % (TG_NT / 4), kg = (i / (TG_NT / 4)) % KG, t = i / ((TG_NT / 4) * KG);
    const int k = kg / a.G, g = kg % a.G;
    const size_t row = ((size_t)it.b * a.T + it.t_lo + t) * a.G + g;
    const float* src = (k == 0) ? a.X + row * a.N : a.zx + (size_t)(k - 1) * kstride + row * a.N;
    cp_async16(smem_u32(zs + ((size_t)t * KG + kg) * TG_NT + c4 * 4), src + it.tile * TG_NT + c4 * 4);
  }
}

// FQ feature groups per CTA (64 * FQ threads, F / FQ features per thread).  FQ = 8 halves the per-thread tap registers so that
// 16 warps are resident instead of 8: with 8 warps the kernel issues on only half of the cycles (profiles/r01_ncu_gate.raw.csv:
// 2 warps per scheduler stalled on fixed-latency FMA -> MUFU dependencies).
template <int KG, int FQ, bool EX>
__global__ void __launch_bounds__(64 * FQ, 1) time_gate_fwd_kernel(const GateArgs a) {
  constexpr int NTH = 64 * FQ, NW = NTH / 32, FMAXV = 64 / FQ;
  extern __shared__ __align__(16) float gsm[];
  const int TC = a.TC;
  const size_t zbuf = (size_t)TC * KG * TG_NT;
  float* zs = gsm;                              // [nbuf][TC][KG][64]
  float* plog = zs + a.nbuf * zbuf;             // [NW warps][TC]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl = tid & 63, fq = tid >> 6;
  const int FG = a.F / FQ, f0 = fq * FG;
  const long long items = a.B * a.nchunks * (a.N / TG_NT);
  const long long per = (items + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = min(items, lo + per);
  if (lo >= hi) return;

  float ta[FMAXV][KG];
#pragma unroll
  for (int i = 0; i < FMAXV; ++i)
#pragma unroll
    for (int kg = 0; kg < KG; ++kg) ta[i][kg] = i < FG ? __ldg(a.A + (size_t)(f0 + i) * KG + kg) : 0.f;

  gate_stage<KG, NTH>(a, gate_item(a, lo), zs, tid);
  cp_async_commit();
  int cur_tile = -1, buf = 0;
  float wg[FMAXV];
  for (long long item = lo; item < hi; ++item) {
    const GateItem it = gate_item(a, item);
    const int n = it.tile * TG_NT + nl;
    if (it.tile != cur_tile) {
      cur_tile = it.tile;
#pragma unroll
      for (int i = 0; i < FMAXV; ++i) wg[i] = i < FG ? __ldg(a.Wg + (size_t)(f0 + i) * a.N + n) : 0.f;
    }
    float c0v[FMAXV];
#pragma unroll
    for (int i = 0; i < FMAXV; ++i) c0v[i] = i < FG ? __ldg(a.c0 + ((size_t)it.b * a.F + f0 + i) * a.N + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();                            // item's rows are visible; everyone is done with the other buffer
    if (a.nbuf == 2 && item + 1 < hi) gate_stage<KG, NTH>(a, gate_item(a, item + 1), zs + (buf ^ 1) * zbuf, tid);
    cp_async_commit();
    const float* zb = zs + buf * zbuf + nl;
    for (int t0 = 0; t0 < it.tn; t0 += 32) {
      float part[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float p = 0.f;
        if (t0 + j < it.tn) {
          const float* zt = zb + (size_t)(t0 + j) * KG * TG_NT;
          float z[KG];
#pragma unroll
          for (int kg = 0; kg < KG; ++kg) z[kg] = zt[kg * TG_NT];
#pragma unroll
          for (int ib = 0; ib < FMAXV; ib += 4) {
            if (ib < FG) {                      // F is a multiple of 16: a thread's features come in groups of 4
#pragma unroll
              for (int i = ib; i < ib + 4; ++i) {
                float pre = c0v[i];
#pragma unroll
                for (int kg = 0; kg < KG; ++kg) pre = fmaf(ta[i][kg], z[kg], pre);
                p = fmaf(wg[i], gate_tanh<EX>(pre), p);
              }
            }
          }
        }
        part[j] = p;
      }
      const float tot = warp_transpose_sum32(part, lane);
      if (t0 + lane < it.tn) plog[warp * TC + t0 + lane] = tot;
    }
    __syncthreads();
    for (int t = tid; t < it.tn; t += NTH) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += plog[w * TC + t];
      atomicAdd(a.logit + it.b * a.T + it.t_lo + t, s);
    }
    if (a.nbuf == 2) buf ^= 1;
    else if (item + 1 < hi) { gate_stage<KG, NTH>(a, gate_item(a, item + 1), zs, tid); cp_async_commit(); }
  }
  cp_async_wait<0>();
}

// backward: given dl[b,t] = d loss / d logit:  dpu = dl Wg (1 - u^2);  dWg += dl u;  dc0[b,f,n] = sum_t dpu;
// dA[f,kg] += sum_{b,t,n} dpu zx_kg.   u is recomputed (one MUFU) instead of being stored (8.6 GB per gate at cfg3).
template <int KG, int FQ, bool EX>
__global__ void __launch_bounds__(64 * FQ, 1) time_gate_bwd_kernel(const GateArgs a) {
  constexpr int NTH = 64 * FQ, FMAXV = 64 / FQ;
  extern __shared__ __align__(16) float gsm[];
  const int TC = a.TC;
  const size_t zbuf = (size_t)TC * KG * TG_NT;
  float* zs = gsm;                              // [nbuf][TC][KG][64]
  float* dls = zs + a.nbuf * zbuf + TG_NWMAX * TC;     // [2][TC]
  float* As = dls + 2 * TC;                     // [F][KG]
  float* dAs = As + a.F * KG;                   // [F][KG]
  const int tid = threadIdx.x, lane = tid & 31;
  const int nl = tid & 63, fq = tid >> 6;
  const int FG = a.F / FQ, f0 = fq * FG;
  const long long items = a.B * a.nchunks * (a.N / TG_NT);
  const long long per = (items + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = min(items, lo + per);
  if (lo >= hi) return;
  for (int i = tid; i < a.F * KG; i += NTH) { As[i] = a.A[i]; dAs[i] = 0.f; }

  {
    const GateItem it = gate_item(a, lo);
    gate_stage<KG, NTH>(a, it, zs, tid);
    for (int t = tid; t < it.tn; t += NTH) dls[t] = a.dl[it.b * a.T + it.t_lo + t];
  }
  cp_async_commit();
  int cur_tile = -1, buf = 0;
  float wg[FMAXV], dwg[FMAXV];
  auto flush_dwg = [&](int tile) {
#pragma unroll
    for (int i = 0; i < FMAXV; ++i) if (i < FG) atomicAdd(a.dWg + (size_t)(f0 + i) * a.N + tile * TG_NT + nl, dwg[i]);
  };
  for (long long item = lo; item < hi; ++item) {
    const GateItem it = gate_item(a, item);
    const int n = it.tile * TG_NT + nl;
    if (it.tile != cur_tile) {
      if (cur_tile >= 0) flush_dwg(cur_tile);
      cur_tile = it.tile;
#pragma unroll
      for (int i = 0; i < FMAXV; ++i) { wg[i] = i < FG ? __ldg(a.Wg + (size_t)(f0 + i) * a.N + n) : 0.f; dwg[i] = 0.f; }
    }
    float c0v[FMAXV];
#pragma unroll
    for (int i = 0; i < FMAXV; ++i) c0v[i] = i < FG ? __ldg(a.c0 + ((size_t)it.b * a.F + f0 + i) * a.N + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    if (a.nbuf == 2 && item + 1 < hi) {
      const GateItem nx = gate_item(a, item + 1);
      gate_stage<KG, NTH>(a, nx, zs + (buf ^ 1) * zbuf, tid);
      for (int t = tid; t < nx.tn; t += NTH) dls[(buf ^ 1) * TC + t] = a.dl[nx.b * a.T + nx.t_lo + t];
    }
    cp_async_commit();
    const float* zb = zs + buf * zbuf + nl;
    const float* dlb = dls + buf * TC;
#pragma unroll
    for (int fb = 0; fb < FMAXV; fb += TG_FB) {
      if (fb < FG) {
        float tb[TG_FB][KG], sA[TG_FB][KG], dw[TG_FB], dc[TG_FB];
#pragma unroll
        for (int i = 0; i < TG_FB; ++i) {
          dw[i] = 0.f; dc[i] = 0.f;
#pragma unroll
          for (int kg = 0; kg < KG; ++kg) { tb[i][kg] = As[(f0 + fb + i) * KG + kg]; sA[i][kg] = 0.f; }
        }
#pragma unroll 2
        for (int t = 0; t < it.tn; ++t) {
          const float* zt = zb + (size_t)t * KG * TG_NT;
          float z[KG];
#pragma unroll
          for (int kg = 0; kg < KG; ++kg) z[kg] = zt[kg * TG_NT];
          const float dlv = dlb[t];
#pragma unroll
          for (int i = 0; i < TG_FB; ++i) {
            float pre = c0v[fb + i];
#pragma unroll
            for (int kg = 0; kg < KG; ++kg) pre = fmaf(tb[i][kg], z[kg], pre);
            const float u = gate_tanh<EX>(pre);
            dw[i] = fmaf(dlv, u, dw[i]);
            const float dpu = (dlv * wg[fb + i]) * fmaf(-u, u, 1.f);
            dc[i] += dpu;
#pragma unroll
            for (int kg = 0; kg < KG; ++kg) sA[i][kg] = fmaf(dpu, z[kg], sA[i][kg]);
          }
        }
#pragma unroll
        for (int i = 0; i < TG_FB; ++i) {
          dwg[fb + i] += dw[i];
          float* o = a.dc0 + ((size_t)it.b * a.F + f0 + fb + i) * a.N + n;
          if (a.nchunks == 1) *o = dc[i]; else atomicAdd(o, dc[i]);
        }
        // sum the TG_FB*KG per-lane tap-gradient partials over the warp: lane l ends up with the total of entry l
        constexpr int NE = TG_FB * KG;
        static_assert(NE <= 32, "tap partials must fit one transposed reduction");
        float x[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = e < NE ? sA[e / KG][e % KG] : 0.f;
        const float tot = warp_transpose_sum32(x, lane);
        if (lane < NE) atomicAdd(dAs + (f0 + fb + lane / KG) * KG + lane % KG, tot);
      }
    }
    if (a.nbuf == 2) buf ^= 1;
    else {
      __syncthreads();
      if (item + 1 < hi) {
        const GateItem nx = gate_item(a, item + 1);
        gate_stage<KG, NTH>(a, nx, zs, tid);
        for (int t = tid; t < nx.tn; t += NTH) dls[t] = a.dl[nx.b * a.T + nx.t_lo + t];
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();
  if (cur_tile >= 0) flush_dwg(cur_tile);
  __syncthreads();
  for (int i = tid; i < a.F * KG; i += NTH) atomicAdd(a.dA + i, dAs[i]);
}

// =====================================================================================================
// generic variant for Kin*G > 8 (taps in shared memory, one sample at a time); same arithmetic as above
//   u = tanh(sum_{k,g} A_g[f,k,g] zx_k + c0[b,f,n]),   logit[b,t] = sum_{f,n} Wg[f,n] u
// CTA = (n-tile of 128, chunk of TG_FC features, b-split); thread = one n.
// =====================================================================================================
// CTA = (tile of 64 nodes, chunk of samples), 256 threads = 64 nodes x 4 feature groups (F/4 features each, <= 16).
// For every sample the x_t S^k rows of ALL T steps are staged once in shared memory ([T][Kin*G][64] fp32), so the
// L2 sees them once; taps A and dlogit are shared-memory broadcasts; the only per-thread state is c0/Wg (registers).
inline size_t gate_generic_smem_bytes(long long T, int KG, int F) {
  return ((size_t)T * KG * TG_NT + (size_t)F * KG + 8 * (size_t)T + (size_t)T + (size_t)F * KG) * sizeof(float);
}

template <bool BWD>
__global__ void __launch_bounds__(256) time_gate_generic_kernel(const GateArgs a) {
  extern __shared__ __align__(16) float gsm[];
  const int KG = a.Kin * a.G;
  const int T = (int)a.T;
  float* zs = gsm;                          // [T][KG][64]
  float* As = zs + (size_t)T * KG * TG_NT;  // [F][KG]
  float* plog = As + a.F * KG;              // [8 warps][T]   (fwd)
  float* dls = plog + 8 * T;                // [T]            (bwd)
  float* dAs = dls + T;                     // [F][KG]        (bwd)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl = tid & 63, fq = tid >> 6;
  const int n = blockIdx.x * TG_NT + nl;
  const int FG = a.F / TG_FQ;               // features per thread (<= 16)
  const int f0 = fq * FG;
  const long long b_lo = (long long)blockIdx.y * a.bchunk, b_hi = min(a.B, b_lo + a.bchunk);
  for (int i = tid; i < a.F * KG; i += 256) { As[i] = a.A[i]; if (BWD) dAs[i] = 0.f; }
  float wg[16], dwg[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { wg[i] = i < FG ? a.Wg[(size_t)(f0 + i) * a.N + n] : 0.f; dwg[i] = 0.f; }
  const size_t kstride = (size_t)a.B * a.T * a.G * a.N;
  for (long long b = b_lo; b < b_hi; ++b) {
    __syncthreads();                        // previous sample's readers are done with zs / plog / dls
    // stage x_t S^k for all t: zs[t][kg][0..63]
    for (int i = tid; i < T * KG * (TG_NT / 4); i += 256) {
      const int c4 = i % (TG_NT / 4), kg = (i / (TG_NT / 4)) % KG, t = i / ((TG_NT / 4) * KG);
      const int k = kg / a.G, g = kg % a.G;
      const size_t row = ((size_t)b * a.T + t) * a.G + g;
      const float* src = (k == 0) ? a.X + row * a.N : a.zx + (size_t)(k - 1) * kstride + row * a.N;
      cp_async16(smem_u32(zs + ((size_t)t * KG + kg) * TG_NT + c4 * 4), src + blockIdx.x * TG_NT + c4 * 4);
    }
    cp_async_commit();
    if (BWD) for (int t = tid; t < T; t += 256) dls[t] = a.dl[b * a.T + t];
    float c0v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c0v[i] = i < FG ? a.c0[((size_t)b * a.F + f0 + i) * a.N + n] : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    if (!BWD) {
      for (int t = 0; t < T; ++t) {
        const float* zt = zs + (size_t)t * KG * TG_NT + nl;
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < FG) {
            float pre = c0v[i];
            const float* ap = As + (f0 + i) * KG;
#pragma unroll 5
            for (int kg = 0; kg < KG; ++kg) pre = fmaf(ap[kg], zt[kg * TG_NT], pre);
            part = fmaf(wg[i], a.exact ? tanh_acc(pre) : tanh_fast(pre), part);
          }
        }
        part = warp_sum_f(part);
        if (lane == 0) plog[warp * T + t] = part;
      }
      __syncthreads();
      for (int t = tid; t < T; t += 256) {
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += plog[w * T + t];
        atomicAdd(a.logit + b * a.T + t, sacc);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < FG) {
          const float* ap = As + (f0 + i) * KG;
          float dc0 = 0.f, dw = 0.f;
          float sA[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int t = 0; t < T; ++t) {
            const float* zt = zs + (size_t)t * KG * TG_NT + nl;
            float pre = c0v[i];
#pragma unroll 5
            for (int kg = 0; kg < KG; ++kg) pre = fmaf(ap[kg], zt[kg * TG_NT], pre);
            const float u = a.exact ? tanh_acc(pre) : tanh_fast(pre);
            const float dlv = dls[t];
            dw = fmaf(dlv, u, dw);
            const float dpu = dlv * wg[i] * (1.f - u * u);
            dc0 += dpu;
            if (KG <= 8) {
#pragma unroll
              for (int kg = 0; kg < 8; ++kg) if (kg < KG) sA[kg] = fmaf(dpu, zt[kg * TG_NT], sA[kg]);
            } else {
              for (int kg = 0; kg < KG; ++kg) atomicAdd(dAs + (f0 + i) * KG + kg, dpu * zt[kg * TG_NT]);
            }
          }
          dwg[i] += dw;
          a.dc0[((size_t)b * a.F + f0 + i) * a.N + n] = dc0;
          if (KG <= 8) {
#pragma unroll
            for (int kg = 0; kg < 8; ++kg) {
              if (kg < KG) {
                const float v = warp_sum_f(sA[kg]);
                if (lane == 0) atomicAdd(dAs + (f0 + i) * KG + kg, v);
              }
            }
          }
        }
      }
    }
  }
  if (BWD) {
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i < FG) atomicAdd(a.dWg + (size_t)(f0 + i) * a.N + n, dwg[i]);
    __syncthreads();
    for (int i = tid; i < a.F * KG; i += 256) atomicAdd(a.dA + i, dAs[i]);
  }
}


}  // namespace tc
}  // namespace gcrnn
