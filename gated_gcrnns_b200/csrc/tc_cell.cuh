// Tensor-core cell path: everything around the tcgen05 shift GEMM for the dense, time-gated (or ungated) cell.
//
// Native reference layout throughout: a signal is [(b, f), n] row-major, so the shift GEMM's output rows are
// H[b, t, f, :] rows and no transposes are needed.  Per time step (Utils/graphML.py:2351-2427):
//     z_k = z_{k-1} @ S           (k = 1..Kst-1)   tcgen05 GEMM, bf16 in / bf16 out             [tc_gemm.cuh]
//     r   = sum_k B_k z_k ;  h_t = tanh(gi (A(S)x_t + b) + gf (r + b))     contract_mma<EPI_FWD>  [here]
// and in reverse (hand-derived adjoint, no recomputation of the forward chain):
//     dpre = (dH_t + dh_rec) (1 - h_t^2)                                    dpre_kernel           [here]
//     v_k = v_{k-1} @ S^T          (k = 1..Kst-1)   tcgen05 GEMM
//     q = sum_k B_k^T v_k ; dgf = <q, h_{t-1}> ; dh_rec = gf q              contract_mma<EPI_BWD>
//     dB_k += gf * v_k h_{t-1}^T                                            wgrad_mma
// The tap contractions are bf16 mma.sync (HMMA) kernels: they are ~8 % of the flops and HBM-bound.
#pragma once
#include "tc_gemm.cuh"

namespace gcrnn {
namespace tc {

// ---- small device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- fp32 -> bf16 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));   // MUFU.TANH, |err| <= 2^-10.99: below the bf16 operand noise
  return y;
}
// in: fp32 [rows][N]  ->  out: bf16 [rows][P*N], plane 0 = bf16(x), plane 1 = bf16(x - plane 0)   (N4 = N / 4)
__global__ void cvt_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4, int N4, int P) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    const long long row = i / N4, c = i % N4;
    uint2* o = reinterpret_cast<uint2*>(out) + row * P * N4 + c;
    uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
    o[0] = u;
    if (P > 1) {
      bf16x2_residual(u.x, v.x, v.y); bf16x2_residual(u.y, v.z, v.w);
      uint2 w; w.x = pack_bf16x2(v.x, v.y); w.y = pack_bf16x2(v.z, v.w);
      o[N4] = w;
    }
  }
}

// ---- weight preparation: W[f, k, g] fp32 -> bf16 [rows][ld] ------------------------------------------------------
// mode 0: out[f][k*G + g] = W[f,k,g]          (forward contraction, rows = output features)
// mode 1: out[g][k*F + f] = W[f,k,g]          (data-gradient contraction, rows = input features)
// P planes per row at column offsets q * pstride (row length ld = P * pstride): plane 1 = bf16 residual of plane 0
__global__ void prep_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int F, int K, int G, int ld, int mode,
                                   int P, int pstride) {
  const int total = F * K * G;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = i % G, k = (i / G) % K, f = i / (G * K);
    const float v = W[i];
    __nv_bfloat16* o = (mode == 0) ? out + (size_t)f * ld + k * G + g : out + (size_t)g * ld + k * F + f;
    store_planes(o, pstride, P, v);
  }
}

// =====================================================================================================
// argument block of the tap contraction  Y[b, m, n] = sum_{k, c} W[m, k*C + c] * Z_k[b, c, n]  (kernel: tc_tap.cuh)
// =====================================================================================================

struct ContractArgs {
  const __nv_bfloat16* W;    // [M][ldw] bf16, columns = slab-major
  const __nv_bfloat16* slab[8];   // K slabs (taps), each [B][C][P*N] bf16
  int K, C, M, N, ldw;
  int P;                          // operand planes per bf16 row (slabs, weights, bf16 outputs)
  long long B;
  // EPI_PLAIN: out_f32[b,m,n] = acc + bias_scale * bias[m]
  // EPI_FWD  : h = tanh(gi (a + bias) + gf (acc + bias)), a = sum_{k,g} A[m,k,g] zx_k[(b,t,g), n]
  // EPI_BWD  : dgf[b] += <acc, hprev[b]> ; out_f32[b,m,n] = gf[b] * acc (+ out_f32 if accumulate)
  float* out_f32; long long out_bstride;      // sample stride of out_f32 (H uses T*F*N)
  __nv_bfloat16* out_bf16;                    // [B][M][N] or null: bf16 copy of the new state (next step's GEMM operand)
  const float* bias; float bias_scale;
  const float* gi; const float* gf; long long gate_stride;     // gate value of sample b at gi[b*gate_stride]
  const float* A; int Kin, G;                                   // [M][Kin][G] fp32
  const float* x0; long long x0_bstride;                        // zx_0 = X[b, t]: [G][N] at x0 + b*x0_bstride
  const float* zx; long long zx_kstride, zx_bstride;            // zx_k (k>=1) at zx + (k-1)*kstride + b*bstride
  const float* hprev; long long hprev_bstride;                  // EPI_BWD: fp32 h_{t-1}[b] = hprev + b*bstride
  float* dgf; int accumulate; int scaled_chain;
  const float* dHn; long long dHn_bstride; const float* gfn; float* red;   // fused backward epilogue (tc_tap.cuh TAP_BWDF)
  const float* qi; const float* qf; long long q_bstride;                   // EPI_FWD: node gates of this step (null: 1), tc_node.cuh
};

// =====================================================================================================
// wgrad_mma: part[cta][k][f][g] += sum_{b in cta's tiles} scale[b] * sum_n V_k[b,f,n] * h[b,g,n]
// =====================================================================================================
constexpr int WG_NT = 64;
constexpr int WG_LD = WG_NT + 8;
constexpr int WG_MAXK = 5;

struct WgradArgs {
  const __nv_bfloat16* v0;   // slab 0         [B][F][N]
  const __nv_bfloat16* vc;   // slabs 1..K-1   [K-1][B][F][N]
  const float* h; long long h_bstride;          // fp32 h_{t-1}[b] = h + b*bstride, [F][N]
  const float* scale; long long scale_stride;   // per-sample scale (gf[b,t]) or null
  float* part;                                   // [grid][K][F][F] fp32, read-modify-write by its owner CTA
  int K, F, N; long long B;
  int P;                                         // planes per row of the v slabs (row stride P*N); plane 0 is used
};

__global__ void __launch_bounds__(256, 1) wgrad_mma_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) uint8_t wg_smem[];
  __nv_bfloat16* Vs = reinterpret_cast<__nv_bfloat16*>(wg_smem);          // [2][K][64][WG_LD]
  __nv_bfloat16* Hs = Vs + (size_t)2 * a.K * 64 * WG_LD;                  // [2][64][WG_LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;       // m-tile (16 rows of f), half of g (32 columns)
  const int tiles_n = a.N / WG_NT;
  const long long num_tiles = a.B * tiles_n;

  float acc[WG_MAXK][4][4];
#pragma unroll
  for (int k = 0; k < WG_MAXK; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[k][j][0] = acc[k][j][1] = acc[k][j][2] = acc[k][j][3] = 0.f; }

  auto load_tile = [&](long long tile, int buf) {
    const long long b = tile / tiles_n;
    const int n0 = (int)(tile % tiles_n) * WG_NT;
    __nv_bfloat16* vd = Vs + (size_t)buf * a.K * 64 * WG_LD;
    for (int i = tid; i < a.K * 64 * 8; i += 256) {
      const int ch = i & 7, row = (i >> 3) & 63, k = i >> 9;
      if (row < a.F) {
        const __nv_bfloat16* src = (k == 0 ? a.v0 : a.vc + (size_t)(k - 1) * a.B * a.F * a.P * a.N) + ((size_t)(b * a.F + row) * a.P * a.N + n0 + ch * 8);
        cp_async16(smem_u32(vd + ((size_t)k * 64 + row) * WG_LD + ch * 8), src);
      }
    }
    // h: fp32 -> (scale) -> bf16
    const float sc = a.scale ? a.scale[b * a.scale_stride] : 1.f;
    __nv_bfloat16* hd = Hs + (size_t)buf * 64 * WG_LD;
    for (int i = tid; i < 64 * 16; i += 256) {
      const int row = i >> 4, c4 = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.F) v = *reinterpret_cast<const float4*>(a.h + b * a.h_bstride + (size_t)row * a.N + n0 + c4 * 4);
      __nv_bfloat162 p = __floats2bfloat162_rn(v.x * sc, v.y * sc), q = __floats2bfloat162_rn(v.z * sc, v.w * sc);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p); u.y = *reinterpret_cast<uint32_t*>(&q);
      *reinterpret_cast<uint2*>(hd + row * WG_LD + c4 * 4) = u;
    }
  };
  // rows >= F of the V tiles are never written: zero them once
  for (int i = tid; i < 2 * a.K * 64 * WG_LD / 8; i += 256) reinterpret_cast<uint4*>(Vs)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  long long tile = blockIdx.x;
  if (tile < num_tiles) load_tile(tile, 0);
  cp_async_commit();
  int buf = 0;
  for (; tile < num_tiles; tile += gridDim.x, buf ^= 1) {
    const long long nxt = tile + gridDim.x;
    if (nxt < num_tiles) load_tile(nxt, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const __nv_bfloat16* vb = Vs + (size_t)buf * a.K * 64 * WG_LD;
    const __nv_bfloat16* hb = Hs + (size_t)buf * 64 * WG_LD;
    // A operand: V_k[f rows 16wm.., n]; B operand: h[g rows 32wn.., n] (both n-contiguous = K-contiguous)
    const uint32_t a_off = (uint32_t)((16 * wm + (lane & 7) + 8 * ((lane >> 3) & 1)) * WG_LD + 8 * (lane >> 4)) * 2;
    const uint32_t b_addr = smem_u32(hb + (32 * wn + (lane & 7) + 8 * (lane >> 4)) * WG_LD + 8 * ((lane >> 3) & 1));
#pragma unroll
    for (int ks = 0; ks < WG_NT / 16; ++ks) {
      uint32_t bf0[4], bf1[4];
      ldsm_x4(bf0, b_addr + ks * 32);
      ldsm_x4(bf1, b_addr + ks * 32 + 16 * WG_LD * 2);
#pragma unroll
      for (int k = 0; k < WG_MAXK; ++k) {
        if (k < a.K) {
          uint32_t af[4];
          ldsm_x4(af, smem_u32(vb + (size_t)k * 64 * WG_LD) + a_off + ks * 32);
          mma_bf16(acc[k][0], af, bf0[0], bf0[1]);
          mma_bf16(acc[k][1], af, bf0[2], bf0[3]);
          mma_bf16(acc[k][2], af, bf1[0], bf1[1]);
          mma_bf16(acc[k][3], af, bf1[2], bf1[3]);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  // accumulate into this CTA's private slice: part[cta][k][f][g]
  float* mine = a.part + (size_t)blockIdx.x * a.K * a.F * a.F;
  const int f0 = 16 * wm + (lane >> 2);
#pragma unroll
  for (int k = 0; k < WG_MAXK; ++k) {
    if (k < a.K) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = 32 * wn + 8 * j + 2 * (lane & 3);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int f = f0 + 8 * hh;
          if (f < a.F && g < a.F) {
            float* o = mine + ((size_t)k * a.F + f) * a.F + g;
            o[0] += acc[k][j][2 * hh];
            o[1] += acc[k][j][2 * hh + 1];
          }
        }
      }
    }
  }
}

// dW[f, k, g] += sum_cta part[cta][k][f][g]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* dW, int ncta, int K, int F) {
  const int total = K * F * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += part[(size_t)c * total + i];
    const int g = i % F, f = (i / F) % F, k = i / (F * F);
    dW[((size_t)f * K + k) * F + g] += s;
  }
}

// =====================================================================================================
// dpre kernel (one time step): dpre = (dH_t + dh_rec) * (1 - h_t^2), plus every reduction that needs dpre.
// One warp per (b, f) row, 8 rows (features) of one sample per CTA; float4 streams, warp-shuffle reductions.
// =====================================================================================================
struct DpreArgs {
  const float* dH; long long dH_bstride;        // dH[b, t]  : [F][N] at dH + b*bstride
  const float* Ht; long long H_bstride;         // h_t[b]
  const float* dhrec;                           // [B][F][N] or null (t = T-1)
  __nv_bfloat16* v0;                            // out: bf16 (g_f * dpre) [B][F][P*N]: input of the adjoint chain (scaled)
  int P;                                        // bf16 planes of v0
  const float* gi; const float* gf; long long gate_stride;
  const float* A; const float* bias; int Kin, G, F, N;
  const float* x0; long long x0_bstride; const float* zx; long long zx_kstride, zx_bstride;
  float* dgi; float* dgf;                       // [.. b*gate_stride]: dgi += <dpre, a + bias>; dgf += <dpre, bias> (bias part only)
  float* dA;                                    // [F][Kin][G]  += gi * sum_n dpre zx_k
  float* dbias;                                 // [F]          += (gi + gf) * sum_n dpre
  long long B;
  // node gates (tc_node.cuh): q_i, q_f of this step at q + b * q_bstride + n; the update is tanh(gi q_i (a + b) + gf q_f (r + b)).
  // Out: d lin_i[b][n] += (1 - q_i) gi q_i sum_f dpre (a + b),  d lin_f[b][n] += (1 - q_f) sum_f dpre (atanh(h) - gi q_i (a + b))
  // (gf q_f (r + b) recovered from the state itself: no recomputation of the state filter).  Needs Kin*G <= DP_KG and 2N floats
  // of dynamic shared memory.
  const float* qi; const float* qf; long long q_bstride;
  float* dlin_i; float* dlin_f;
  // input gradient: dxk[kg][b][n] += sum_f A[f][kg] gi q_i dpre[f][n] at dxk + kg * dxk_kstride + b * dxk_bstride + g_row * N + n
  // (the rows of one [B,T,G,N] layout, like zx); the caller shifts tap k by (S^T)^k afterwards.  Needs Kin*G <= DP_KG and
  // Kin*G*N more floats of dynamic shared memory (after the node gates' 2N, when present).
  float* dxk; long long dxk_kstride, dxk_bstride;
};
constexpr int DP_FC = 8;      // features (warps) per CTA
constexpr int DP_KG = 8;      // (k, g) pairs handled per pass

__global__ void __launch_bounds__(256) dpre_kernel(const DpreArgs a) {
  __shared__ float s_gi[DP_FC];
  extern __shared__ float s_dl[];               // node gates: [2][N] per-node sums over the CTA's DP_FC features
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KG = a.Kin * a.G;
  const int fgroups = a.F / DP_FC;
  const int N4 = a.N / 4;
  const bool node = a.qi != nullptr;
  const bool wantdx = a.dxk != nullptr;
  float* s_dx = s_dl + (node ? 2 * a.N : 0);      // [KG][N]
  if (node || wantdx) {
    for (int i = threadIdx.x; i < (node ? 2 * a.N : 0) + (wantdx ? KG * a.N : 0); i += blockDim.x) s_dl[i] = 0.f;
    __syncthreads();
  }
  for (long long item = blockIdx.x; item < a.B * fgroups; item += gridDim.x) {
    const long long b = item / fgroups;
    const int f = (int)(item % fgroups) * DP_FC + warp;
    const float vgi = a.gi ? a.gi[b * a.gate_stride] : 1.f;
    const float vgf = a.gf ? a.gf[b * a.gate_stride] : 1.f;
    const float bb = a.bias ? a.bias[f] : 0.f;
    const float4* pdH = reinterpret_cast<const float4*>(a.dH + b * a.dH_bstride + (size_t)f * a.N);
    const float4* pH = reinterpret_cast<const float4*>(a.Ht + b * a.H_bstride + (size_t)f * a.N);
    const float4* pR = a.dhrec ? reinterpret_cast<const float4*>(a.dhrec + ((size_t)b * a.F + f) * a.N) : nullptr;
    const float4* pQi = node ? reinterpret_cast<const float4*>(a.qi + b * a.q_bstride) : nullptr;
    const float4* pQf = node ? reinterpret_cast<const float4*>(a.qf + b * a.q_bstride) : nullptr;
    uint2* pV = reinterpret_cast<uint2*>(a.v0 + ((size_t)b * a.F + f) * a.P * a.N);
    float sdi = 0.f, sdf = 0.f, sa = 0.f;       // sum_n dpre q_i, sum_n dpre q_f, <dpre q_i, A(S)x>
    for (int kg0 = 0; kg0 < KG || kg0 == 0; kg0 += DP_KG) {
      float sz[DP_KG], Af[DP_KG];
#pragma unroll
      for (int j = 0; j < DP_KG; ++j) { sz[j] = 0.f; Af[j] = (kg0 + j < KG) ? a.A[(size_t)f * KG + kg0 + j] : 0.f; }
      for (int i = lane; i < N4; i += 32) {
        const float4 hv = pH[i];
        float4 d = pdH[i];
        if (pR) { const float4 r = pR[i]; d.x += r.x; d.y += r.y; d.z += r.z; d.w += r.w; }
        d.x *= 1.f - hv.x * hv.x; d.y *= 1.f - hv.y * hv.y; d.z *= 1.f - hv.z * hv.z; d.w *= 1.f - hv.w * hv.w;
        float4 qi4 = make_float4(1.f, 1.f, 1.f, 1.f), qf4 = qi4;
        if (node) { qi4 = pQi[i]; qf4 = pQf[i]; }
        if (kg0 == 0) {
          float s0 = vgf * qf4.x * d.x, s1 = vgf * qf4.y * d.y, s2 = vgf * qf4.z * d.z, s3 = vgf * qf4.w * d.w;
          uint2 u; u.x = pack_bf16x2(s0, s1); u.y = pack_bf16x2(s2, s3);
          pV[i] = u;
          if (a.P > 1) {
            bf16x2_residual(u.x, s0, s1); bf16x2_residual(u.y, s2, s3);
            uint2 w; w.x = pack_bf16x2(s0, s1); w.y = pack_bf16x2(s2, s3);
            pV[N4 + i] = w;
          }
          sdi += fmaf(d.x, qi4.x, d.y * qi4.y) + fmaf(d.z, qi4.z, d.w * qi4.w);
          sdf += fmaf(d.x, qf4.x, d.y * qf4.y) + fmaf(d.z, qf4.z, d.w * qf4.w);
        }
        const float4 dq = make_float4(d.x * qi4.x, d.y * qi4.y, d.z * qi4.z, d.w * qi4.w);
        float4 ax = make_float4(bb, bb, bb, bb);                        // A(S)x + b of this feature (node gates only)
#pragma unroll
        for (int j = 0; j < DP_KG; ++j) {
          const int kg = kg0 + j;
          if (kg < KG) {
            const int k = kg / a.G, g = kg % a.G;
            const float* zp = (k == 0) ? a.x0 + b * a.x0_bstride + (size_t)g * a.N
                                       : a.zx + (size_t)(k - 1) * a.zx_kstride + b * a.zx_bstride + (size_t)g * a.N;
            const float4 z = reinterpret_cast<const float4*>(zp)[i];
            sz[j] = fmaf(dq.x, z.x, fmaf(dq.y, z.y, fmaf(dq.z, z.z, fmaf(dq.w, z.w, sz[j]))));
            ax.x = fmaf(Af[j], z.x, ax.x); ax.y = fmaf(Af[j], z.y, ax.y); ax.z = fmaf(Af[j], z.z, ax.z); ax.w = fmaf(Af[j], z.w, ax.w);
          }
        }
        if (wantdx) {
#pragma unroll
          for (int j = 0; j < DP_KG; ++j)
            if (kg0 + j < KG) {
              float* sx = s_dx + (size_t)(kg0 + j) * a.N + 4 * i;
              const float w = vgi * Af[j];
              atomicAdd(sx, w * dq.x); atomicAdd(sx + 1, w * dq.y); atomicAdd(sx + 2, w * dq.z); atomicAdd(sx + 3, w * dq.w);
            }
        }
        if (node) {
          const float hx[4] = {hv.x, hv.y, hv.z, hv.w}, dx[4] = {d.x, d.y, d.z, d.w}, axx[4] = {ax.x, ax.y, ax.z, ax.w};
          const float qix[4] = {qi4.x, qi4.y, qi4.z, qi4.w}, qfx[4] = {qf4.x, qf4.y, qf4.z, qf4.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float wi = vgi * qix[c] * axx[c];                     // gi q_i (a + b)
            const float pre = atanhf(fminf(fmaxf(hx[c], -0.99999994f), 0.99999994f));
            atomicAdd(s_dl + 4 * i + c, (1.f - qix[c]) * dx[c] * wi);
            atomicAdd(s_dl + a.N + 4 * i + c, (1.f - qfx[c]) * dx[c] * (pre - wi));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < DP_KG; ++j) {
        const int kg = kg0 + j;
        if (kg < KG) {
          const float v = warp_sum_f(sz[j]);
          sa = fmaf(Af[j], v, sa);                                     // <dpre q_i, A(S)x> = sum_kg A[f,kg] <dpre q_i, zx_kg>
          if (lane == 0 && a.dA) atomicAdd(a.dA + (size_t)f * KG + kg, vgi * v);
        }
      }
    }
    sdi = warp_sum_f(sdi); sdf = warp_sum_f(sdf);
    if (lane == 0) {
      if (a.dbias) atomicAdd(a.dbias + f, vgi * sdi + vgf * sdf);
      s_gi[warp] = sa + bb * sdi;
      if (a.dgf) atomicAdd(a.dgf + b * a.gate_stride, bb * sdf);
    }
    __syncthreads();
    if (threadIdx.x == 0 && a.dgi) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < DP_FC; ++w) t += s_gi[w];
      atomicAdd(a.dgi + b * a.gate_stride, t);
    }
    if (node) {
      for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        atomicAdd(a.dlin_i + b * a.q_bstride + i, s_dl[i]);
        atomicAdd(a.dlin_f + b * a.q_bstride + i, s_dl[a.N + i]);
        s_dl[i] = 0.f; s_dl[a.N + i] = 0.f;
      }
    }
    if (wantdx) {
      for (int i = threadIdx.x; i < KG * a.N; i += blockDim.x) {
        const int kg = i / a.N, n = i - kg * a.N, k = kg / a.G, g = kg - k * a.G;
        atomicAdd(a.dxk + (size_t)k * a.dxk_kstride + b * a.dxk_bstride + (size_t)g * a.N + n, s_dx[i]);
        s_dx[i] = 0.f;
      }
    }
    __syncthreads();
  }
}

// g = sigmoid(logit + c)      /     dl = dg g (1-g), dc += sum dl
__global__ void gate_sigmoid_kernel(const float* __restrict__ logit, const float* __restrict__ c, float* __restrict__ g, long long n) {
  const float cv = c ? c[0] : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] = 1.f / (1.f + expf(-(logit[i] + cv)));
}
__global__ void gate_dlogit_kernel(const float* __restrict__ dg, const float* __restrict__ g, float* __restrict__ dl, float* dc, long long n) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float v = g[i]; const float d = dg[i] * v * (1.f - v); dl[i] = d; s += d; }
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0 && dc) { float t = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; atomicAdd(dc, t); }
}
// out[f] += scale * sum_{b,n} in[b,f,n]
__global__ void rowsum_bfn_kernel(const float* __restrict__ in, float* out, long long B, int F, int N, float scale) {
  __shared__ float red[8];
  for (long long item = blockIdx.x; item < B * F; item += gridDim.x) {
    const int f = (int)(item % F);
    const float* p = in + item * N;
    float s = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) s += p[n];
    s = warp_sum_f(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w]; atomicAdd(out + f, scale * t); }
    __syncthreads();
  }
}

}  // namespace tc
}  // namespace gcrnn
