// Second generation of the fused fp32 kernels of the sparse EDGE-GATED recurrence (F == 32, cfg5 of SURVEY.md §8d).
// Same algebra, same HBM arrays and the same accumulator layout as sp32_kernels.cuh (which stays selectable per stage through
// gcrnn_debug_set_option("sparse_v2", mask) for A/B tests); what changes is the work decomposition, chosen from the ncu
// exports of the first generation (profiles/r01_ncu_sparse_fwd/bwd*): those kernels were issue- and latency-bound at 330-670
// warp instructions per (sample, node) with 12 resident warps, not HBM-bound.  Measured at cfg5: 130 vs 60 sequences/s.
//
//   * Gathers: EIGHT lanes own one destination node (lane = 16-byte chunk of the 128-byte row), four nodes per warp, one
//     LDG.128 per neighbour row and no final reduce-scatter.  The edge list is loaded once per batch of eight edges (lane =
//     edge, coalesced, next batch prefetched) and broadcast with width-8 shuffles; 2-4 nodes of a thread run in lockstep
//     (8-16 row loads in flight per thread); ONE trip count per warp (the four node groups never serialise their different
//     in-degrees); pinned base pointers (SHFL, IMAD.WIDE, LDG.128, SHFL + 4 FFMA per gathered row).
//   * Contractions: a 256-thread block stages a tile of 128 consecutive nodes in shared memory (streamed taps by cp.async
//     while the block gathers the last one) and each warp contracts its 16 nodes against weights that sit in shared memory
//     once per block - on tensor cores with error compensation (3xTF32 mma.sync.m16n8k8, fp32 accumulate; conflict-free
//     fragment loads, row strides == 8 mod 32).  A register-tiled packed-FFMA2 version (thread = 4 nodes x 4 outputs) is kept
//     behind "sparse_v2_tc" = 0: it measured shared-memory-bound (every LDS.128 is four wavefronts).
//   * Weight-gradient outer products (M_k = sum_n dWu[n] (x) z_k[n]) are the transposed mma product of the same tile
//     (K dimension = the tile's nodes), kept in registers across all tiles of a block.
//   * Saved softmax statistics are (c, logsumexp) per gate and row: one 16-byte load per (edge, source row).
//   * The dh kernel's epilogue finishes the NEXT reverse step's dpre, so dh itself is never stored.
// What bounds them now (profiles/r01_ncu_sparse_v2_summary.txt): the L1 data pipe (l1tex data-pipe wavefronts 55-83 % of peak):
// every gathered neighbour row is one 128-byte wavefront, and shuffles / fragment loads share that pipe.
//
// Reference op sites (Utils/graphML.py): LSIGF shift :123 + contraction :134-139; graphAttention :586-625 with
// S' = S + I :577, leaky_relu(0.2) :603, masked softmax over j :611-622, aggregation over i :625; relu :2101;
// cell update h = tanh(Q_i(a) + Q_f(r)) :2402-2423.
#pragma once
#include "sp32_kernels.cuh"

namespace gcrnn {
namespace e32 {

// tile kernels: NT threads stage and contract a tile of TM = NT / 2 consecutive nodes (NT = 128 or 256)
constexpr int XS_LD = 16;                               // x taps per node in shared memory (zero padded, == MAXKG)
// row strides == 8 (mod 32): conflict-free for the LDS.128 rows of the FFMA2 contraction AND for both mma fragment patterns
// (A: bank 8 g + t, B of the transposed product: bank 8 t + g)
constexpr int LDD = 40;                                 // row stride of the [TM][32] tiles
__host__ __device__ constexpr int lda_of(int kst) { return kst * 32 + 8; }

// contiguous range of `unit`-node groups for this block (neighbouring groups share neighbour rows: L1/L2 reuse);
// walked as (sample r, group u within the sample) so that the loop carries no 64-bit division
struct Walk {
  long long left; long long r; int u, per_sample;
  __device__ __forceinline__ Walk(long long R, int N, int unit) {
    per_sample = (N + unit - 1) / unit;
    const long long total = R * per_sample, per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per;
    const long long hi = lo + per < total ? lo + per : total;
    left = hi > lo ? hi - lo : 0;
    r = lo / per_sample; u = (int)(lo - r * per_sample);
  }
  __device__ __forceinline__ bool more() const { return left > 0; }
  __device__ __forceinline__ void next() { --left; if (++u == per_sample) { u = 0; ++r; } }
};
// keep a derived base pointer materialised in one register pair (else the compiler re-adds its parts for every gathered row:
// 3 address instructions per load instead of one IMAD.WIDE)
template <class T> __device__ __forceinline__ const T* pin_ptr(const T* p) { asm volatile("" : "+l"(p)); return p; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float gsum8(float v, unsigned gmask) {       // sum over the 8 lanes of a node group
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}
// eight per-lane partial sums -> total of slot c on lane c of the 8-lane group (7 shuffles)
template <int M>
__device__ __forceinline__ void tree8_step(float (&v)[8], int c, unsigned gmask) {
  const bool up = c & M;
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const float lo = v[k], hi = v[k + M];
    v[k] = (up ? hi : lo) + __shfl_xor_sync(gmask, up ? lo : hi, M);
  }
}
__device__ __forceinline__ float tree8(float (&v)[8], int c, unsigned gmask) {
  tree8_step<4>(v, c, gmask); tree8_step<2>(v, c, gmask); tree8_step<1>(v, c, gmask);
  return v[0];
}

// Gather of NQ destination nodes per 8-lane group in lockstep:  acc[q] = sum_p val[p] * src[idx[p]] (this lane's 16-byte chunk).
// Lane c loads (idx, val) of edge (batch + c) ONCE per batch of eight edges (coalesced; the next batch is prefetched) and the
// group walks the batch with width-8 shuffles: per-edge index loads would cost as many L1 wavefronts as the rows themselves.
// NQ * NU row loads are in flight per thread.  Must be called by all 32 lanes of a warp.  Idle slots read row 0 of the sample with coefficient 0 (predicating them costs
// more instructions than it saves wavefronts).  n[q] < 0: no node.  `base` = sample base (float4 units) + chunk.
template <int NQ, int NU>
__device__ __forceinline__ void gather_nodes(const Gather3& op, const float4* base, const int (&n)[NQ], float4 (&acc)[NQ],
                                             unsigned gmask, int c) {
  base = pin_ptr(base);
  int p0[NQ], deg[NQ], md = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    p0[q] = 0; deg[q] = 0; acc[q] = f4zero();
    if (n[q] >= 0) { p0[q] = __ldg(op.ptr + n[q]); deg[q] = __ldg(op.ptr + n[q] + 1) - p0[q]; }
    md = max(md, deg[q]);
  }
  // one trip count per WARP: its four node groups would otherwise serialise their different-length edge loops (in-degrees vary)
  md = __reduce_max_sync(0xffffffffu, md);
  int mi[NQ]; float mv[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    mi[q] = 0; mv[q] = 0.f;
    if (c < deg[q]) { mi[q] = __ldg(op.idx + p0[q] + c); mv[q] = __ldg(op.val + p0[q] + c); }
  }
  for (int pb = 0; pb < md; pb += 8) {               // md is uniform within the warp
    int ni[NQ]; float nv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      ni[q] = 0; nv[q] = 0.f;
      if (pb + 8 + c < deg[q]) { ni[q] = __ldg(op.idx + p0[q] + pb + 8 + c); nv[q] = __ldg(op.val + p0[q] + pb + 8 + c); }
    }
#pragma unroll
    for (int e0 = 0; e0 < 8; e0 += NU) {
      if (pb + e0 < md) {
        float4 x[NQ][NU]; float w[NQ][NU];
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int ii = __shfl_sync(0xffffffffu, mi[q], e0 + u, 8);
            w[q][u] = __shfl_sync(0xffffffffu, mv[q], e0 + u, 8);
            x[q][u] = __ldg(base + (size_t)ii * 8);          // idle slots: row 0 of the sample (one shared line) with coefficient 0
          }
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
          for (int q = 0; q < NQ; ++q) fma4(acc[q], w[q][u], x[q][u]);
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) { mi[q] = ni[q]; mv[q] = nv[q]; }
  }
}

// ---- sparse shift of a 32-channel node-major signal: out[r,d,:] = sum_p val[p] in[r, idx[p], :] ----------------
// 256 threads = 32 groups of 8 lanes; a group owns two nodes (64 nodes per block pass)
__global__ void __launch_bounds__(256, 3) spmm32_v2_k(Gather3 op, const float* __restrict__ in, float* __restrict__ out, int N, long long R) {
  const int lane = threadIdx.x & 31, c = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const unsigned gmask = 0xFFu << (lane & 24);
  for (Walk w(R, N, 64); w.more(); w.next()) {
    const size_t rb = (size_t)w.r * N;
    int n[2]; float4 a[2];
    n[0] = w.u * 64 + slot; n[1] = n[0] + 32;
    if (n[0] >= N) n[0] = -1;
    if (n[1] >= N) n[1] = -1;
    gather_nodes<2, 4>(op, reinterpret_cast<const float4*>(in + rb * 32) + c, n, a, gmask, c);
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (n[q] >= 0) reinterpret_cast<float4*>(out + (rb + n[q]) * 32)[c] = a[q];
  }
}

// ---- tile staging helpers (NT-thread blocks, TM = NT / 2 nodes) ----------------------------------------------------
// rows n0 .. n0+TM-1 of a [R,N,32] array -> tile[node][col0 .. col0+31] by cp.async (zero rows beyond N); the caller
// commits / waits, so the copies fly while the block gathers
template <int NT>
__device__ __forceinline__ void stage_rows_async(float* tile, int ld, int col0, const float* __restrict__ src, size_t rb, int n0, int N) {
  const float* s = src + (rb + n0) * 32;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = threadIdx.x + NT * q, node = idx >> 3, ch = (idx & 7) * 4;
    float* dst = tile + node * ld + col0 + ch;
    if (n0 + node < N) cp_async16(dst, s + node * 32 + ch);
    else *reinterpret_cast<float4*>(dst) = f4zero();
  }
}
// gathered rows (src S)[n0 .. n0+TM-1] into registers: a thread's four nodes (slot + NT/8 * q) run in lockstep
template <int NT, int NU>
__device__ __forceinline__ void gather_tile(float4 (&a)[4], const Gather3& op, const float* __restrict__ src, size_t rb, int n0, int N) {
  const int c = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  int n[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { n[q] = n0 + slot + (NT / 8) * q; if (n[q] >= N) n[q] = -1; }
  gather_nodes<4, NU>(op, reinterpret_cast<const float4*>(src + rb * 32) + c, n, a, gmask, c);
}
template <int NT>
__device__ __forceinline__ void store_tile(float* tile, int ld, int col0, const float4 (&a)[4]) {
  const int c = threadIdx.x & 7, slot = threadIdx.x >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(tile + (slot + (NT / 8) * q) * ld + col0 + c * 4) = a[q];
}
// x taps of the tile's nodes -> Xs[node][kg], kg = k * G + g, zero padded to XS_LD.  A thread always serves the same kg
// (NT is a multiple of XS_LD): `xp` = its tap array + g, or nullptr beyond KG (tap_slot(), computed once per kernel).
__device__ __forceinline__ const float* tap_slot(const Chain& xs, int KG, int G) {
  const int kg = threadIdx.x & (XS_LD - 1);
  if (kg >= KG) return nullptr;
  const int k = kg / G;
  const float* xp = xs.p[0];
#pragma unroll
  for (int m = 1; m < MAXK; ++m) if (k == m) xp = xs.p[m];
  return xp + (kg - k * G);
}
template <int NT>
__device__ __forceinline__ void stage_taps(float* Xs, const float* __restrict__ xp, int G, size_t rb, int n0, int N) {
#pragma unroll
  for (int i = 0; i < (NT / 2) * XS_LD / NT; ++i) {
    const int idx = threadIdx.x + NT * i, node = idx >> 4;
    Xs[idx] = (xp != nullptr && n0 + node < N) ? __ldg(xp + (rb + n0 + node) * G) : 0.f;
  }
}
// weights w[kk][f] (kk < KK inputs, 32 outputs) -> shared "pair" layout: float4 ((kp*2 + half)*8 + fg) =
// (w[2kp][f0], w[2kp+1][f0], w[2kp][f0+1], w[2kp+1][f0+1]),  f0 = 4 fg + 2 half
template <class W>
__device__ __forceinline__ void stage_weights(float* Ws, int KK, W w) {
  for (int o = threadIdx.x; o < KK * 32; o += blockDim.x) {
    const int q = o & 3, fg = (o >> 2) & 7, half = (o >> 5) & 1, kp = o >> 6;
    Ws[o] = w(2 * kp + (q & 1), 4 * fg + 2 * half + (q >> 1));
  }
}
// o[i][j] = sum_kk tile[warp*16 + ngl + 4i][kk] * w[kk][4 fg + j]   (thread = 4 nodes x 4 outputs, FFMA2 over input pairs)
template <int KK>
__device__ __forceinline__ void tile_contract(const float* __restrict__ tile, int ld, const float* __restrict__ Ws, float (&o)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fg = lane & 7, ngl = lane >> 3;
  const float* arow = tile + (warp * 16 + ngl) * ld;
  const float4* wp = reinterpret_cast<const float4*>(Ws) + fg;
  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int kq = 0; kq < KK / 4; ++kq) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(arow + 4 * i * ld + 4 * kq);
    const float4 w00 = wp[(4 * kq + 0) * 8], w01 = wp[(4 * kq + 1) * 8], w10 = wp[(4 * kq + 2) * 8], w11 = wp[(4 * kq + 3) * 8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 lo = make_float2(a[i].x, a[i].y), hi = make_float2(a[i].z, a[i].w);
      acc[i][0] = __ffma2_rn(make_float2(w00.x, w00.y), lo, acc[i][0]);
      acc[i][1] = __ffma2_rn(make_float2(w00.z, w00.w), lo, acc[i][1]);
      acc[i][2] = __ffma2_rn(make_float2(w01.x, w01.y), lo, acc[i][2]);
      acc[i][3] = __ffma2_rn(make_float2(w01.z, w01.w), lo, acc[i][3]);
      acc[i][0] = __ffma2_rn(make_float2(w10.x, w10.y), hi, acc[i][0]);
      acc[i][1] = __ffma2_rn(make_float2(w10.z, w10.w), hi, acc[i][1]);
      acc[i][2] = __ffma2_rn(make_float2(w11.x, w11.y), hi, acc[i][2]);
      acc[i][3] = __ffma2_rn(make_float2(w11.z, w11.w), hi, acc[i][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = acc[i][j].x + acc[i][j].y;
}


// ---- 3xTF32 tensor-core contraction (mma.sync.m16n8k8, fp32 accumulate) ---------------------------------------------------
// x = hi + lo with hi = tf32(x), lo = x - hi (truncated to tf32 by the hardware): a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi keeps ~2^-21 relative error per
// product, i.e. fp32-class accuracy (the stated 1e-5 / 1e-4 bounds of the fp32 path hold, see tests), while the operands come
// out of shared memory as conflict-free fragments: 12 wavefronts per 16 x 32 x 8 block instead of 32 for the FFMA2 tile.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // hi = x rounded to 10 mantissa bits (round half away, as cvt.rna.tf32 does - which sm_100 emulates with 4+ instructions
  // because of its Inf/NaN path; the signals here are finite); lo = x - hi is exact in fp32 and the tensor core ignores its
  // low 13 mantissa bits itself, so it needs no rounding of its own.  3 instructions per element.
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al, bh0, bh1); mma_tf32(d, ah, bl0, bl1); mma_tf32(d, ah, bh0, bh1);      // small terms first
}
// weights w[kk][f] -> shared fragment layout: float4 ((ks*2 + half)*32 + lane) = (w[kk][g], w[kk][8+g], w[kk][16+g], w[kk][24+g]),
// kk = 8 ks + t + 4 half, g = lane >> 2, t = lane & 3  (the b0 / b1 registers of the four 8-wide output tiles)
template <class W>
__device__ __forceinline__ void stage_weights_tc(float* Ws, int KK, W w) {
  for (int o = threadIdx.x; o < KK * 32; o += blockDim.x) {
    const int nt = o & 3, lane = (o >> 2) & 31, half = (o >> 7) & 1, ks = o >> 8;
    Ws[o] = w(8 * ks + (lane & 3) + 4 * half, 8 * nt + (lane >> 2));
  }
}
// d[nt][.] = rows (warp*16 + g, warp*16 + g + 8) x columns (8 nt + 2 t, + 1) of  tile[16 rows of the warp][KK] * w[KK][32]
template <int KK>
__device__ __forceinline__ void tile_contract_tc(const float* __restrict__ tile, int ld, const float* __restrict__ Ws, float (&d)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const float* ap = tile + (warp * 16 + g) * ld + t;
  const float4* wp = reinterpret_cast<const float4*>(Ws) + lane;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f; }
#pragma unroll 2
  for (int ks = 0; ks < KK / 8; ++ks) {
    uint32_t ah[4], al[4];
    split_tf32(ap[8 * ks], ah[0], al[0]); split_tf32(ap[8 * ld + 8 * ks], ah[1], al[1]);
    split_tf32(ap[8 * ks + 4], ah[2], al[2]); split_tf32(ap[8 * ld + 8 * ks + 4], ah[3], al[3]);
    const float4 w0 = wp[(2 * ks) * 32], w1 = wp[(2 * ks + 1) * 32];
    const float b0[4] = {w0.x, w0.y, w0.z, w0.w}, b1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(b0[nt], bh0, bl0); split_tf32(b1[nt], bh1, bl1);
      mma3(d[nt], ah, al, bh0, bh1, bl0, bl1);
    }
  }
}

// ---- gather + contraction over a tile -----------------------------------------------------------------------------
// GC_FILTER (forward stage 1): Wu_r = sum_k Cr_k z_k + cr0 (z_{KST-1} = z_{KST-2} S gathered into the tile),
//   Wu_a = Ca x-taps + ca0, rc = (a1.Wu_a, a2.Wu_a, a1.Wu_r, a2.Wu_r);  weights from the folded `prep` block.
// GC_DH (backward stage 3): dh_{t-1}[n,g] = sum_k sum_f B[f,k,g] w_k[n,f], w_{KST-1} = w_{KST-2} S^T gathered into the tile.
enum { GC_FILTER = 0, GC_DH = 1 };
// GC_DH can finish the NEXT reverse step's dpre in its epilogue (tensor-core variant): with dh = dh_{t-1} in registers,
//   g = (dH[b,t-1,f,n] + dh[n,f]) (1 - h_{t-1}[n,f]^2),  dya = g [y_a > 0],  dyr = g [y_r > 0]   (what dpre_k computes from a stored dh)
// dHt == nullptr: plain mode (dh is written to `out`).
struct DpreFuse {
  const float* dHt; long long sample_stride;      // dH + (t-1) F N in the reference layout [B,T,F,N]
  const float* hn; const uint2* masks;            // h_{t-1} [B,N,32] and its relu masks
  float* dya; float* dyr;
  int node_major;                                 // 1: dHt is node-major [B,N,32] (the renumbered path converts it first), 0: reference layout
};
template <int KST, int NT>
__host__ __device__ constexpr int gc_smem_floats(int mode) {
  return KST * 32 * 32 + (NT / 2) * lda_of(KST) + (mode == GC_FILTER ? (NT / 2) * XS_LD + MAXKG * 32 + 6 * 32 : 0);
}
template <int KST, int MODE, int NT, bool TC>
__global__ void __launch_bounds__(NT, 512 / NT) gather_contract_k(Gather3 gop, Chain zc, Chain xs, int Kin, int G,
                                                            const float* __restrict__ wsrc /* FILTER: prep, DH: weight_B */,
                                                            const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                            float* __restrict__ out_a /* FILTER: Wu_a */, float* __restrict__ out /* FILTER: Wu_r, DH: dh */,
                                                            float4* __restrict__ rc, int N, long long R, DpreFuse fz) {
  constexpr int KK = KST * 32, LDA = lda_of(KST), NS = KST - 1, TM = NT / 2;
  extern __shared__ __align__(16) float dyn[];
  float* Ws = dyn;                       // [KK/2][2][8] float4
  float* As = Ws + KK * 32;              // [TM][LDA]
  float* Xs = As + TM * LDA;             // FILTER: [TM][XS_LD]
  float* Cas = Xs + TM * XS_LD;          // FILTER: [MAXKG][32]
  float* Cst = Cas + MAXKG * 32;         // FILTER: cr0, ca0, a1_a, a2_a, a1_r, a2_r  [6][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fg = lane & 7, ngl = lane >> 3;
  const int KG = Kin * G;
  const float* xp = nullptr;
  auto wf = [&](int kk, int f) { return MODE == GC_FILTER ? wsrc[PrepLayout::CR + kk * 32 + f] : wsrc[(kk & 31) * KK + (kk >> 5) * 32 + f]; };   // DH: B[f_in, k, g_out]
  if (TC) stage_weights_tc(Ws, KK, wf); else stage_weights(Ws, KK, wf);
  if (MODE == GC_FILTER) {
    for (int i = threadIdx.x; i < MAXKG * 32; i += NT) Cas[i] = wsrc[PrepLayout::CA + i];
    if (threadIdx.x < 32) {
      Cst[threadIdx.x] = wsrc[PrepLayout::CR0 + threadIdx.x]; Cst[32 + threadIdx.x] = wsrc[PrepLayout::CA0 + threadIdx.x];
      Cst[64 + threadIdx.x] = mix_a[threadIdx.x]; Cst[96 + threadIdx.x] = mix_a[32 + threadIdx.x];
      Cst[128 + threadIdx.x] = mix_r[threadIdx.x]; Cst[160 + threadIdx.x] = mix_r[32 + threadIdx.x];
    }
    xp = tap_slot(xs, KG, G);
  }
  for (Walk w(R, N, TM); w.more(); w.next()) {
    const int n0 = w.u * TM;
    const size_t rb = (size_t)w.r * N;
#pragma unroll
    for (int s = 0; s < NS; ++s) stage_rows_async<NT>(As, LDA, s * 32, zc.p[s], rb, n0, N);
    cp_commit();
    if (MODE == GC_FILTER) stage_taps<NT>(Xs, xp, G, rb, n0, N);
    {
      float4 gz[4];
      gather_tile<NT, 4>(gz, gop, zc.p[NS - 1], rb, n0, N);
      store_tile<NT>(As, LDA, NS * 32, gz);
    }
    cp_wait<0>();
    __syncthreads();
    if (TC) {
      float dd[4][4];
      tile_contract_tc<KK>(As, LDA, Ws, dd);
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {                 // rows g and g + 8 of the warp's 16 nodes
        const int node = warp * 16 + g + 8 * h, n = n0 + node;
        if (MODE == GC_FILTER) {
          float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
          float2 wr2[4], wa2[4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int f = 8 * nt + 2 * t;
            const float2 cr0 = *reinterpret_cast<const float2*>(Cst + f), ca0 = *reinterpret_cast<const float2*>(Cst + 32 + f);
            wr2[nt] = make_float2(dd[nt][2 * h] + cr0.x, dd[nt][2 * h + 1] + cr0.y);
            float2 wa = ca0;
            for (int kg = 0; kg < KG; ++kg) {
              const float xv = Xs[node * XS_LD + kg];
              const float2 cw = *reinterpret_cast<const float2*>(Cas + kg * 32 + f);
              wa.x = fmaf(xv, cw.x, wa.x); wa.y = fmaf(xv, cw.y, wa.y);
            }
            wa2[nt] = wa;
            const float2 m1a = *reinterpret_cast<const float2*>(Cst + 64 + f), m2a = *reinterpret_cast<const float2*>(Cst + 96 + f);
            const float2 m1r = *reinterpret_cast<const float2*>(Cst + 128 + f), m2r = *reinterpret_cast<const float2*>(Cst + 160 + f);
            s1 = fmaf(m1a.x, wa.x, fmaf(m1a.y, wa.y, s1)); s2 = fmaf(m2a.x, wa.x, fmaf(m2a.y, wa.y, s2));
            s3 = fmaf(m1r.x, wr2[nt].x, fmaf(m1r.y, wr2[nt].y, s3)); s4 = fmaf(m2r.x, wr2[nt].x, fmaf(m2r.y, wr2[nt].y, s4));
          }
#pragma unroll
          for (int of = 2; of > 0; of >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, of); s2 += __shfl_xor_sync(0xffffffffu, s2, of);
            s3 += __shfl_xor_sync(0xffffffffu, s3, of); s4 += __shfl_xor_sync(0xffffffffu, s4, of);
          }
          if (n < N) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              *reinterpret_cast<float2*>(out_a + (rb + n) * 32 + 8 * nt + 2 * t) = wa2[nt];
              *reinterpret_cast<float2*>(out + (rb + n) * 32 + 8 * nt + 2 * t) = wr2[nt];
            }
            if (t == 0) rc[rb + n] = make_float4(s1, s2, s3, s4);
          }
        } else if (n < N) {
          if (fz.dHt != nullptr) {                  // fused dpre of the next reverse step
            const uint2 mk = fz.masks[rb + n];
            const float* dHp = fz.dHt + w.r * fz.sample_stride + (fz.node_major ? (size_t)n * 32 : (size_t)n);
            const size_t fs = fz.node_major ? 1 : (size_t)N;      // stride between features
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const int f = 8 * nt + 2 * t;          // features f, f + 1: mask bits 8 (f & 3) + (f >> 2)
              const size_t o = (rb + n) * 32 + f;
              const float2 hv = *reinterpret_cast<const float2*>(fz.hn + o);
              const float g0 = (__ldg(dHp + (size_t)f * fs) + dd[nt][2 * h]) * (1.f - hv.x * hv.x);
              const float g1 = (__ldg(dHp + (size_t)(f + 1) * fs) + dd[nt][2 * h + 1]) * (1.f - hv.y * hv.y);
              const int b0 = 8 * (f & 3) + (f >> 2), b1 = 8 * ((f + 1) & 3) + ((f + 1) >> 2);
              *reinterpret_cast<float2*>(fz.dya + o) = make_float2(((mk.x >> b0) & 1u) ? g0 : 0.f, ((mk.x >> b1) & 1u) ? g1 : 0.f);
              *reinterpret_cast<float2*>(fz.dyr + o) = make_float2(((mk.y >> b0) & 1u) ? g0 : 0.f, ((mk.y >> b1) & 1u) ? g1 : 0.f);
            }
          } else {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              *reinterpret_cast<float2*>(out + (rb + n) * 32 + 8 * nt + 2 * t) = make_float2(dd[nt][2 * h], dd[nt][2 * h + 1]);
          }
        }
      }
    } else {
      float o[4][4];
      tile_contract<KK>(As, LDA, Ws, o);
      float4 c0r, c0a, a1a, a2a, a1r, a2r;
      if (MODE == GC_FILTER) {
        const float4* c4 = reinterpret_cast<const float4*>(Cst) + fg;
        c0r = c4[0]; c0a = c4[8]; a1a = c4[16]; a2a = c4[24]; a1r = c4[32]; a2r = c4[40];
      }
  #pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int node = warp * 16 + ngl + 4 * i, n = n0 + node;
        if (MODE == GC_FILTER) {
          const float4 wr = make_float4(o[i][0] + c0r.x, o[i][1] + c0r.y, o[i][2] + c0r.z, o[i][3] + c0r.w);
          float4 wa = c0a;
          for (int kg = 0; kg < KG; ++kg) fma4(wa, Xs[node * XS_LD + kg], *reinterpret_cast<const float4*>(Cas + kg * 32 + 4 * fg));
          float s1 = dot4(a1a, wa), s2 = dot4(a2a, wa), s3 = dot4(a1r, wr), s4 = dot4(a2r, wr);
  #pragma unroll
          for (int of = 4; of > 0; of >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, of); s2 += __shfl_xor_sync(0xffffffffu, s2, of);
            s3 += __shfl_xor_sync(0xffffffffu, s3, of); s4 += __shfl_xor_sync(0xffffffffu, s4, of);
          }
          if (n < N) {
            reinterpret_cast<float4*>(out_a + (rb + n) * 32)[fg] = wa;
            reinterpret_cast<float4*>(out + (rb + n) * 32)[fg] = wr;
            if (fg == 0) rc[rb + n] = make_float4(s1, s2, s3, s4);
          }
        } else if (n < N) {
          reinterpret_cast<float4*>(out + (rb + n) * 32)[fg] = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
        }
      }
    }
    __syncthreads();
  }
}

// ---- forward, stage 2: per-row softmax statistics of both gates, compact layout ---------------------------------------
// cl[r*N+i] = (c_a, lse_a, c_r, lse_r) with lse = log sum_j exp(e_ij) over the pattern of S + I, e_ij = leaky(c_i + r_j), so that
// alpha_ij = exp(e_ij - lse_i) needs ONE 16-byte load per (edge, source row);  rr[r*N+j] = (r_a, r_r).
__global__ void __launch_bounds__(256) rowstats_v2_k(const int* __restrict__ rptr, const int* __restrict__ col, const float4* __restrict__ rc,
                                                     float4* __restrict__ cl, float2* __restrict__ rr, int N, long long RN) {
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
    cl[t] = make_float4(me.y, ma + logf(da), me.w, mr + logf(dr));
    rr[t] = make_float2(me.x, me.z);
  }
}

// ---- forward, stage 3: attention aggregation of both gates + relu + tanh update -----------------------------------
// y_g[j,:] = relu( sum_{i -> j} S'_ij alpha^g_ij Wu_g[i,:] ),  h = tanh(y_a + y_r);  masks = sign bits of the two relus
// (bit 8 i + c of a mask word belongs to feature 4 c + i, the layout dpre_k reads).  Eight lanes per destination j: lane c
// computes alpha S' of edge (batch + c) once, the group then walks the batch with width-8 shuffles.
__global__ void __launch_bounds__(256, 4) aggregate_v2_k(const int* __restrict__ cptr, const int* __restrict__ crow, const float* __restrict__ cval,
                                                         const float4* __restrict__ cl, const float2* __restrict__ rr,
                                                         const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                         float* __restrict__ hn, uint2* __restrict__ masks, int N, long long R) {
  const int lane = threadIdx.x & 31, c = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const unsigned gmask = 0xFFu << (lane & 24);
  for (Walk w(R, N, 32); w.more(); w.next()) {
    const int j = w.u * 32 + slot;
    const bool live = j < N;                                    // uniform within the 8-lane group
    const size_t rb = (size_t)w.r * N;
    int p0 = 0, deg = 0;
    float2 rj = make_float2(0.f, 0.f);
    if (live) { p0 = __ldg(cptr + j); deg = __ldg(cptr + j + 1) - p0; rj = rr[rb + j]; }
    const int degw = __reduce_max_sync(0xffffffffu, deg);       // one trip count per warp (in-degrees vary: no serialised groups)
    const float rja = rj.x, rjr = rj.y;
    const float4* pa = pin_ptr(reinterpret_cast<const float4*>(wu_a + rb * 32) + c);
    const float4* pr = pin_ptr(reinterpret_cast<const float4*>(wu_r + rb * 32) + c);
    float4 acc_a = f4zero(), acc_r = f4zero();
    for (int pb = 0; pb < degw; pb += 8) {
      int mi = 0; float ca = 0.f, cr = 0.f;                     // idle lanes: row 0 of the sample with coefficient 0
      if (pb + c < deg) {
        mi = __ldg(crow + p0 + pb + c);
        const float v = __ldg(cval + p0 + pb + c);
        const float4 s4 = cl[rb + mi];                           // (c_a, lse_a, c_r, lse_r) of source row mi
        ca = v * __expf(leaky(s4.x + rja) - s4.y);
        cr = v * __expf(leaky(s4.z + rjr) - s4.w);
      }
#pragma unroll
      for (int e0 = 0; e0 < 8; e0 += 4) {
        if (pb + e0 < degw) {                                   // warp-uniform; four edges in flight
          int ii[4]; float wa[4], wr[4]; float4 xa[4], xr[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ii[u] = __shfl_sync(0xffffffffu, mi, e0 + u, 8);
            wa[u] = __shfl_sync(0xffffffffu, ca, e0 + u, 8); wr[u] = __shfl_sync(0xffffffffu, cr, e0 + u, 8);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) { xa[u] = __ldg(pa + (size_t)ii[u] * 8); xr[u] = __ldg(pr + (size_t)ii[u] * 8); }
#pragma unroll
          for (int u = 0; u < 4; ++u) { fma4(acc_a, wa[u], xa[u]); fma4(acc_r, wr[u], xr[u]); }
        }
      }
    }
    unsigned ma = (acc_a.x > 0.f ? 1u : 0u) | (acc_a.y > 0.f ? 0x100u : 0u) | (acc_a.z > 0.f ? 0x10000u : 0u) | (acc_a.w > 0.f ? 0x1000000u : 0u);
    unsigned mr = (acc_r.x > 0.f ? 1u : 0u) | (acc_r.y > 0.f ? 0x100u : 0u) | (acc_r.z > 0.f ? 0x10000u : 0u) | (acc_r.w > 0.f ? 0x1000000u : 0u);
    ma = __reduce_or_sync(gmask, ma << c); mr = __reduce_or_sync(gmask, mr << c);
    float4 h;
    h.x = tanh_pos(fmaxf(acc_a.x, 0.f) + fmaxf(acc_r.x, 0.f)); h.y = tanh_pos(fmaxf(acc_a.y, 0.f) + fmaxf(acc_r.y, 0.f));
    h.z = tanh_pos(fmaxf(acc_a.z, 0.f) + fmaxf(acc_r.z, 0.f)); h.w = tanh_pos(fmaxf(acc_a.w, 0.f) + fmaxf(acc_r.w, 0.f));
    if (live) {
      reinterpret_cast<float4*>(hn + (rb + j) * 32)[c] = h;
      if (c == 0) masks[rb + j] = make_uint2(ma, mr);
    }
  }
}

// ---- backward, stage 1 (per source row i, both gates): softmax / leaky backward, partial dWu ---------------------------
//   al_ij = softmax_j(leaky(c_i + r_j)),  dal_ij = S'_ij <dy_g[j], Wu_g[i]>,  ds_ij = al (dal - sum_j al dal) leaky',
//   dr_j += ds_ij (atomics), dc_i = sum_j ds_ij,  p_g[i,:] = sum_j S'_ij al_ij dy_g[j,:] + a2_g dc_i,  m2_g += dc_i Wu_g[i,:]
// Eight lanes per row; lane c keeps the edge data of edges c, 8 + c, 16 + c, 24 + c (row degree of S + I <= 32); the
// eight dot products of a batch are reduced "transposed" (7 shuffles) so that edge (batch + c) lands on lane c.
template <int BPS>
__global__ void __launch_bounds__(256, BPS) bwd_rows_v2_k(const int* __restrict__ rptr, const int* __restrict__ col, const float* __restrict__ val,
                                                        const float4* __restrict__ cl, const float2* __restrict__ rr,
                                                        const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                        const float* __restrict__ dya, const float* __restrict__ dyr,
                                                        const float* __restrict__ mix_a, const float* __restrict__ mix_r,
                                                        float* __restrict__ pa, float* __restrict__ pr, float2* __restrict__ dr /* [R*N], zeroed */,
                                                        float* __restrict__ acc, int N, long long R) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const float4 a2a = *reinterpret_cast<const float4*>(mix_a + 32 + 4 * c), a2r = *reinterpret_cast<const float4*>(mix_r + 32 + 4 * c);
  float4 m2a = f4zero(), m2r = f4zero();
  for (Walk w(R, N, 32); w.more(); w.next()) {
    const int i = w.u * 32 + slot;
    const bool live = i < N;                                    // uniform within the 8-lane group
    const size_t rb = (size_t)w.r * N;
    int p0 = 0, deg = 0;
    float4 mine = f4zero(), wua = f4zero(), wur = f4zero();      // mine = (c_a, lse_a, c_r, lse_r) of this row
    if (live) {
      p0 = __ldg(rptr + i); deg = __ldg(rptr + i + 1) - p0;
      mine = cl[rb + i];
      wua = reinterpret_cast<const float4*>(wu_a + (rb + i) * 32)[c]; wur = reinterpret_cast<const float4*>(wu_r + (rb + i) * 32)[c];
    }
    const int degw = __reduce_max_sync(0xffffffffu, deg);       // one trip count per warp: rows of different degree do not serialise
    const float4* ga = pin_ptr(reinterpret_cast<const float4*>(dya + rb * 32) + c);
    const float4* gr = pin_ptr(reinterpret_cast<const float4*>(dyr + rb * 32) + c);
    float4 parta = f4zero(), partr = f4zero();
    float tSa = 0.f, tSr = 0.f;
    int jj[4]; float ala[4], alr[4], dala[4], dalr[4], sla[4], slr[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      jj[b] = 0; ala[b] = alr[b] = dala[b] = dalr[b] = sla[b] = slr[b] = 0.f;
      if (8 * b < degw) {                                       // warp-uniform
        const int cnt = min(8, deg - 8 * b), cntw = min(8, degw - 8 * b);
        float v = 0.f, coa = 0.f, cor = 0.f;
        if (c < cnt) {
          jj[b] = __ldg(col + p0 + 8 * b + c);
          v = __ldg(val + p0 + 8 * b + c);
          const float2 rj = rr[rb + jj[b]];
          const float sca = mine.x + rj.x, scr = mine.z + rj.y;
          ala[b] = __expf(leaky(sca) - mine.y); sla[b] = sca > 0.f ? 1.f : 0.2f;
          alr[b] = __expf(leaky(scr) - mine.w); slr[b] = scr > 0.f ? 1.f : 0.2f;
          coa = v * ala[b]; cor = v * alr[b];
        }
        float pda[8], pdr[8];
#pragma unroll
        for (int e0 = 0; e0 < 8; e0 += 4) {
          if (e0 < cntw) {                                      // warp-uniform; idle edges: row 0 with coefficient 0
            int ii[4]; float wa[4], wr[4]; float4 xa[4], xr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              ii[u] = __shfl_sync(0xffffffffu, jj[b], e0 + u, 8);
              wa[u] = __shfl_sync(0xffffffffu, coa, e0 + u, 8); wr[u] = __shfl_sync(0xffffffffu, cor, e0 + u, 8);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { xa[u] = __ldg(ga + (size_t)ii[u] * 8); xr[u] = __ldg(gr + (size_t)ii[u] * 8); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              fma4(parta, wa[u], xa[u]); fma4(partr, wr[u], xr[u]);
              pda[e0 + u] = dot4(xa[u], wua); pdr[e0 + u] = dot4(xr[u], wur);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) { pda[e0 + u] = 0.f; pdr[e0 + u] = 0.f; }
          }
        }
        dala[b] = v * tree8(pda, c, 0xffffffffu); dalr[b] = v * tree8(pdr, c, 0xffffffffu);     // v == 0 on idle lanes
        tSa = fmaf(ala[b], dala[b], tSa); tSr = fmaf(alr[b], dalr[b], tSr);
      }
    }
    const float Sa = gsum8(tSa, 0xffffffffu), Sr = gsum8(tSr, 0xffffffffu);
    float dca = 0.f, dcr = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (8 * b + c < deg) {
        const float dsa = ala[b] * (dala[b] - Sa) * sla[b], dsr = alr[b] * (dalr[b] - Sr) * slr[b];
        atomicAdd(dr + rb + jj[b], make_float2(dsa, dsr));
        dca += dsa; dcr += dsr;
      }
    dca = gsum8(dca, 0xffffffffu); dcr = gsum8(dcr, 0xffffffffu);
    fma4(parta, dca, a2a); fma4(partr, dcr, a2r);
    if (live) {
      reinterpret_cast<float4*>(pa + (rb + i) * 32)[c] = parta;
      reinterpret_cast<float4*>(pr + (rb + i) * 32)[c] = partr;
    }
    fma4(m2a, dca, wua); fma4(m2r, dcr, wur);
  }
  // lanes c, c+8, c+16, c+24 of a warp hold the same features
  float m[8] = {m2a.x, m2a.y, m2a.z, m2a.w, m2r.x, m2r.y, m2r.z, m2r.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) { m[q] += __shfl_xor_sync(0xffffffffu, m[q], 8); m[q] += __shfl_xor_sync(0xffffffffu, m[q], 16); }
  if (lane < 8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { red[warp][4 * lane + q] = m[q]; red[warp][32 + 4 * lane + q] = m[4 + q]; }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    atomicAdd(acc + (threadIdx.x < 32 ? AccLayout::M2A + threadIdx.x : AccLayout::M2R + threadIdx.x - 32), s);
  }
}

// ---- backward, stage 2 (per tile): finish dWu, accumulate the outer products, d = W_r^T dWu_r ----------------------------
//   dWu_g[n,:] = p_g[n,:] + a1_g dr_g[n];  m1_g += dr_g[n] Wu_g[n,:];  sum_g += dWu_g[n,:]
//   M_k[f][g] += dWu_r[n,f] z_k[n,g] (z_{KST-1} gathered into the tile);  Ma[kg][f] += dWu_a[n,f] x[n,kg];  dout[n,:] = W_r^T dWu_r[n,:]
// dWu_a only feeds the small input-tap products: it is parked in the tile's gathered column block, consumed, and only then
// overwritten by the gathered tap (which waits in registers meanwhile) - no shared memory of its own.
template <int KST, int NT>
__host__ __device__ constexpr int bwd_node_smem_floats() { return 32 * 32 + (NT / 2) * lda_of(KST) + (NT / 2) * LDD + (NT / 2) * XS_LD + MAXKG * 32; }
template <int KST, int NT, bool TC>
__global__ void __launch_bounds__(NT, NT == 128 ? 3 : 2) bwd_node_v2_k(Gather3 gop, Chain zc, Chain xs, int Kin, int G,
                                                        const float* __restrict__ pa, const float* __restrict__ pr, const float2* __restrict__ dr,
                                                        const float* __restrict__ wu_a, const float* __restrict__ wu_r,
                                                        const float* __restrict__ mix_a, const float* __restrict__ mix_r, const float* __restrict__ Wr,
                                                        float* __restrict__ dout, float* __restrict__ acc, int N, long long R) {
  constexpr int KK = KST * 32, LDA = lda_of(KST), NS = KST - 1, TM = NT / 2;
  constexpr int KPT = KK / 16;                 // z columns per thread in the outer-product stage (pairs: KST float2)
  extern __shared__ __align__(16) float dyn[];
  float* Ws = dyn;                             // W_r in pair layout (32 inputs f, 32 outputs m)
  float* As = Ws + 32 * 32;                    // [TM][LDA]  z_0 .. z_{KST-1}
  float* Ds = As + TM * LDA;                   // [TM][LDD]  dWu_r
  float* Xs = Ds + TM * LDD;                   // [TM][XS_LD]
  float* Mas = Xs + TM * XS_LD;                // [MAXKG][32] block accumulators of the input-tap products
  float* Da = As + NS * 32;                    // dWu_a parked in the gathered column block (row stride LDA)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fg = lane & 7, ngl = lane >> 3;
  const int c = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int KG = Kin * G;
  auto wf = [&](int f, int m) { return Wr[f * 32 + m]; };
  if (TC) stage_weights_tc(Ws, 32, wf); else stage_weights(Ws, 32, wf);
  const float* xp = tap_slot(xs, KG, G);
  for (int i = threadIdx.x; i < MAXKG * 32; i += NT) Mas[i] = 0.f;
  // tensor-core outer products: warp = (16-feature half mh, quarter kq of the KK inputs), K dimension = the tile's nodes
  static_assert(!TC || NT == 256, "the mma outer-product mapping assumes 8 warps");
  const int mh = warp & 1, kq = warp >> 1;
  float Mt[KST][4];
#pragma unroll
  for (int u = 0; u < KST; ++u) { Mt[u][0] = Mt[u][1] = Mt[u][2] = Mt[u][3] = 0.f; }
  float4 m1a = f4zero(), m1r = f4zero();
  float2 M[4][KST];                            // M[fi][u]: f = 4 ft + fi, kk = KPT kt + 2u (+1), partial over this thread's 64-node half
#pragma unroll
  for (int fi = 0; fi < 4; ++fi)
#pragma unroll
    for (int u = 0; u < KST; ++u) M[fi][u] = make_float2(0.f, 0.f);
  float suma = 0.f, sumr = 0.f;
  const int ft = threadIdx.x & 7, kt = (threadIdx.x >> 3) & 15, nh = threadIdx.x >> 7;
  for (Walk w(R, N, TM); w.more(); w.next()) {
    const int n0 = w.u * TM;
    const size_t rb = (size_t)w.r * N;
#pragma unroll
    for (int s = 0; s < NS; ++s) stage_rows_async<NT>(As, LDA, s * 32, zc.p[s], rb, n0, N);
    cp_commit();
    stage_taps<NT>(Xs, xp, G, rb, n0, N);
    const float4 a1a = __ldg(reinterpret_cast<const float4*>(mix_a) + c), a1r = __ldg(reinterpret_cast<const float4*>(mix_r) + c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {                // finish dWu of both gates for (node, chunk c)
      const int node = slot + (NT / 8) * q, n = n0 + node;
      float4 dwa = f4zero(), dwr = f4zero();
      if (n < N) {
        const size_t o = (rb + n) * 8 + c;
        const float2 d2 = dr[rb + n];
        dwa = __ldg(reinterpret_cast<const float4*>(pa) + o); dwr = __ldg(reinterpret_cast<const float4*>(pr) + o);
        fma4(dwa, d2.x, a1a); fma4(dwr, d2.y, a1r);
        fma4(m1a, d2.x, __ldg(reinterpret_cast<const float4*>(wu_a) + o)); fma4(m1r, d2.y, __ldg(reinterpret_cast<const float4*>(wu_r) + o));
      }
      *reinterpret_cast<float4*>(Da + node * LDA + 4 * c) = dwa;
      *reinterpret_cast<float4*>(Ds + node * LDD + 4 * c) = dwr;
    }
    float4 gz[4];
    gather_tile<NT, 2>(gz, gop, zc.p[NS - 1], rb, n0, N);
    __syncthreads();
    {                                            // input-tap outer products and column sums: thread = feature `lane`, 16 nodes of the warp
      const float* dap = Da + warp * 16 * LDA + lane;
      const float* xsp = Xs + warp * 16 * XS_LD;
#pragma unroll
      for (int q = 0; q < 16; ++q) { suma += dap[q * LDA]; sumr += Ds[(warp * 16 + q) * LDD + lane]; }
      for (int kg = 0; kg < KG; ++kg) {
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) m = fmaf(dap[q * LDA], xsp[q * XS_LD + kg], m);
        atomicAdd(Mas + kg * 32 + lane, m);
      }
    }
    __syncthreads();
    store_tile<NT>(As, LDA, NS * 32, gz);
    cp_wait<0>();
    __syncthreads();
    if (TC) {
      {                                          // d = W_r^T dWu_r
        float dd[4][4];
        tile_contract_tc<32>(Ds, LDD, Ws, dd);
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = n0 + warp * 16 + g + 8 * h;
          if (n < N) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              *reinterpret_cast<float2*>(dout + (rb + n) * 32 + 8 * nt + 2 * t) = make_float2(dd[nt][2 * h], dd[nt][2 * h + 1]);
          }
        }
      }
      {                                          // M[f][kk] += sum_node dWu_r[node][f] z[node][kk]  as  (Ds^T) (As)
        const int g = lane >> 2, t = lane & 3;
        const float* ap = Ds + t * LDD + 16 * mh + g;
        const float* bp = As + t * LDA + kq * (KK / 4) + g;
#pragma unroll 2
        for (int ks = 0; ks < TM / 8; ++ks) {
          uint32_t ah[4], al[4];
          split_tf32(ap[(8 * ks) * LDD], ah[0], al[0]); split_tf32(ap[(8 * ks) * LDD + 8], ah[1], al[1]);
          split_tf32(ap[(8 * ks + 4) * LDD], ah[2], al[2]); split_tf32(ap[(8 * ks + 4) * LDD + 8], ah[3], al[3]);
#pragma unroll
          for (int u = 0; u < KST; ++u) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(bp[(8 * ks) * LDA + 8 * u], bh0, bl0); split_tf32(bp[(8 * ks + 4) * LDA + 8 * u], bh1, bl1);
            mma3(Mt[u], ah, al, bh0, bh1, bl0, bl1);
          }
        }
      }
    } else {
      {                                            // d = W_r^T dWu_r
        float o[4][4];
        tile_contract<32>(Ds, LDD, Ws, o);
  #pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int n = n0 + warp * 16 + ngl + 4 * i;
          if (n < N) reinterpret_cast<float4*>(dout + (rb + n) * 32)[fg] = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
        }
      }
      {                                            // outer products: thread = 4 outputs f x KPT inputs kk, reduction over 64 nodes of the tile
        const float* dp = Ds + nh * 64 * LDD + 4 * ft;
        const float* zp = As + nh * 64 * LDA + KPT * kt;
  #pragma unroll 4
        for (int node = 0; node < 64; ++node) {
          const float4 d4 = *reinterpret_cast<const float4*>(dp + node * LDD);
  #pragma unroll
          for (int u = 0; u < KST; ++u) {
            const float2 z2 = *reinterpret_cast<const float2*>(zp + node * LDA + 2 * u);
            M[0][u] = __ffma2_rn(make_float2(d4.x, d4.x), z2, M[0][u]);
            M[1][u] = __ffma2_rn(make_float2(d4.y, d4.y), z2, M[1][u]);
            M[2][u] = __ffma2_rn(make_float2(d4.z, d4.z), z2, M[2][u]);
            M[3][u] = __ffma2_rn(make_float2(d4.w, d4.w), z2, M[3][u]);
          }
        }
      }
    }
    __syncthreads();
  }
  // partial sums of distinct M entries (per 64-node half): one global atomic per entry per thread
  if (TC) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int u = 0; u < KST; ++u) {
      const int kk = kq * (KK / 4) + 8 * u + 2 * t, f = 16 * mh + g;
      atomicAdd(acc + AccLayout::M + kk * 32 + f, Mt[u][0]); atomicAdd(acc + AccLayout::M + (kk + 1) * 32 + f, Mt[u][1]);
      atomicAdd(acc + AccLayout::M + kk * 32 + f + 8, Mt[u][2]); atomicAdd(acc + AccLayout::M + (kk + 1) * 32 + f + 8, Mt[u][3]);
    }
  } else {
#pragma unroll
    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
      for (int u = 0; u < KST; ++u) {
        const int kk = KPT * kt + 2 * u, f = 4 * ft + fi;
        atomicAdd(acc + AccLayout::M + kk * 32 + f, M[fi][u].x);
        atomicAdd(acc + AccLayout::M + (kk + 1) * 32 + f, M[fi][u].y);
      }
  }
  for (int i = threadIdx.x; i < KG * 32; i += NT) atomicAdd(acc + AccLayout::MA + i, Mas[i]);   // last tile ended with a barrier
  atomicAdd(acc + AccLayout::SUMA + lane, suma); atomicAdd(acc + AccLayout::SUMR + lane, sumr);
  float m[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1r.x, m1r.y, m1r.z, m1r.w};
#pragma unroll
  for (int q = 0; q < 8; ++q) { m[q] += __shfl_xor_sync(0xffffffffu, m[q], 8); m[q] += __shfl_xor_sync(0xffffffffu, m[q], 16); }
  if (lane < 8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { atomicAdd(acc + AccLayout::M1A + 4 * lane + q, m[q]); atomicAdd(acc + AccLayout::M1R + 4 * lane + q, m[4 + q]); }
  }
}

}  // namespace e32
}  // namespace gcrnn
