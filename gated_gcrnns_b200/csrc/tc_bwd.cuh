// Fused reverse-time step of the dense tensor-core path (F = 64), one launch per time step t after the adjoint chain
// v_k = v_{k-1} S^T (tc_gemm2.cuh) has been formed from v_0 = bf16(gf_t dpre_t).  Adjoint of Utils/graphML.py:2402-2423.
//
// Per tile (sample b, 128 nodes) the K chain slabs are read ONCE from HBM and feed three tensor-core products:
//   MMA1  dh[n, g]     = sum_{k,f} v_k[f, n] B_k[f, g]           (= gf_t q_t = d loss / d h_{t-1}; A operand MN-major)
//   MMA2  dB_k[f, g]  += sum_n v_k[f, n] h_{t-1}[g, n]            (weight gradient; same smem tiles read K-major,
//                                                                 taps stacked in pairs to M = 128)
//   MMA3  dAx[f, j]   += sum_n v_0[f, n] Zs_t[j, n]               (input-filter tap + bias gradients: Zs_t holds the
//                                                                 rows (gi/gf) x_t S^k and the constant (gi/gf + 1))
// and the epilogue (16 warps, thread <-> node) turns dh into step t-1's
//   dpre = (dH_{t-1} + dh) (1 - h_{t-1}^2),  v_0' = bf16(gf_{t-1} dpre)  (next chain input),
//   dgf_t += <dh, h_{t-1}> / gf_t,  dgi_{t-1} += <dpre, A(S)x_{t-1} + b>,  dgf_{t-1} += <dpre, b>.
// MMA2 / MMA3 accumulate in TMEM over all tiles of the CTA; one read-modify-write of the CTA's private partial at
// the end (no atomics), reduced once per backward pass by wgrad_reduce_kernel / dax_reduce_kernel.
//
// Shared memory: ring of 4 "pair stages" (two taps x 128 nodes, 32 KB: [node half][tap][64 rows][128 B], SW128),
// ring of 2 aux stages ([node half]{h_{t-1} bf16 64 rows, Zs 16 rows}: 20 KB), resident weights.
//
// Instruction count matters here: an M = 128 tcgen05.mma with N = 16 or 64 occupies the tensor pipe about as long as one with
// N = 128 (profiles/r02_horner_notes.txt), and with separate instructions per operand plane and per product this kernel was
// bound by them, not by HBM (split-bf16: 124 per tile, 1059 us per launch for 4.8 GB).  Hence (i) MMA3 rides in MMA2: the B
// operand is the stacked [h; Zs] tile (N = 80), columns 64..79 of the first pair's accumulator are dAx; (ii) with split operands
// MMA1 multiplies signal plane 0 by the stacked weight planes [W0; W1] (N = 128) and the epilogue sums the two column halves.
#pragma once
#include "tc_tap.cuh"

namespace gcrnn {
namespace tc {

constexpr int BF_STAGES = 4;
constexpr int BF_STAGE_BYTES = 32768;
constexpr int BF_AUX = 2;
constexpr int BF_AUX_BYTES = 16384 + 4096;
constexpr int BF_THREADS = 64 + 16 * 32;
constexpr int BF_ZROWS = 16;                       // rows of a Zs tile: Kin*G taps, one constant row, zero padding
// dynamic shared memory: stage ring + aux ring + P weight planes of KB blocks + barriers / taps / bias / pointers
inline int bf_smem_bytes(int P, int KB, int stages) {
  return stages * BF_STAGE_BYTES + BF_AUX * BF_AUX_BYTES + P * KB * 8192 + 256 + (64 * 8 + 64) * 4 + 8 * 8 + 1024;
}

struct BwdFusedArgs {
  int K, N, KB;                 // taps, nodes, 64-row weight blocks per plane (= K)
  int P;                        // operand planes of the chain slabs, the weights and v0_out (1: bf16, 2: split hi + lo).
                                // dh (MMA1) uses v0 W0 + v0 W1 + v1 W0; the weight-gradient products (MMA2 / MMA3) sum both
                                // planes of v against plane 0 of h / Zs (their rounding noise averages over B*N summands)
  int stages;                   // pair-stage ring depth (<= BF_STAGES)
  long long B, R;               // samples, rows per slab (B*64)
  int last;                     // t == 0: no earlier step; write dh0 instead
  float* dh0;                   // [B][64][N] (last only, may be null)
  // step t
  const float* gf; long long gate_stride;          // gf[b, t] at gf[b * gate_stride] (null: 1)
  float* dgf;                                      // += <dh, h_{t-1}> / gf_t
  const float* hprev; long long hprev_bstride;     // fp32 h_{t-1}
  // step t-1 (unused when last)
  const float* dHn; long long dHn_bstride;
  const float* gfn;                                // gf[b, t-1]
  float* dgin; float* dgfn;                        // dgi[b, t-1], dgf[b, t-1] (null without time gating)
  const float* A; const float* bias;               // input taps [64][KG], bias [64] (null: 0)
  const float* x0; const float* zx; long long zx_kstride, z_bstride; int G;   // rows x_{t-1} S^k (fp32)
  __nv_bfloat16* v0_out;                           // [B][64][N] next chain input
  long long zs_row0, zs_rowb;                      // Zs tile of (b, t) starts at row zs_row0 + b * zs_rowb
  // node gates (tc_node.cuh), step t-1: q_i / q_f at q + b * q_bstride + n, gi[b, t-1] at gin[b * gate_stride] (null: 1).
  // v0' = bf16(gf q_f dpre); the per-node head gradients d lin_i / d lin_f (see dpre_kernel) are accumulated with one atomicAdd per
  // 16-feature group and node.  Zs then carries the per-node ratio (gi q_i) / (gf q_f) (zs_build_kernel).
  const float* qin; const float* qfn; long long q_bstride; const float* gin;
  float* dlin_i; float* dlin_f;
  // accumulators
  float* part;                  // [grid][K][64][64]
  float* partA;                 // [grid][64][16]
};

template <int KG, bool NODE>     // NODE: per-node gates (compile-time so that the ungated / time-gated kernel is untouched)
__global__ void __launch_bounds__(BF_THREADS, 1)
bwd_fused_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tmc,
                 const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmZ,
                 const __grid_constant__ CUtensorMap tmW, const BwdFusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sV = smem;                                              // [BF_STAGES][2 halves][2 taps][64 rows][128 B]
  uint8_t* sX = sV + a.stages * BF_STAGE_BYTES;                    // [BF_AUX][2 halves]{ h: [64][128 B], Zs: [16][128 B] }
  uint8_t* sW = sX + BF_AUX * BF_AUX_BYTES;                        // [KB][P][64 rows g][128 B]: the planes of a tap are stacked
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sW + a.P * a.KB * 8192);
  uint64_t* empty_bar = full_bar + BF_STAGES;
  uint64_t* aux_full = empty_bar + BF_STAGES;
  uint64_t* aux_empty = aux_full + BF_AUX;
  uint64_t* tmem_full = aux_empty + BF_AUX;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint64_t* done_bar = w_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  float* sAw = reinterpret_cast<float*>(tmem_slot + 4);            // [64][KG]
  float* sBias = sAw + 64 * 8;                                     // [64]
  const float** sZb = reinterpret_cast<const float**>(sBias + 64); // [8] fp32 row base pointers of x_{t-1} S^k

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = a.N / 128;
  const long long num_tiles = a.B * tiles_n;
  const long long per_cta = (num_tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_lo = blockIdx.x * per_cta;
  const long long tile_hi = tile_lo + per_cta < num_tiles ? tile_lo + per_cta : num_tiles;
  const int NP = (a.K + 1) / 2;
  constexpr uint32_t TMEM_D2 = 256, D2_STRIDE = 80, TMEM_COLS = 512;    // D1: 2 stages x 128 columns; D2: up to 3 pairs x 80

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm0); tma_prefetch_desc(&tmc); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmW);
    for (int s = 0; s < BF_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < BF_AUX; ++s) { mbar_init(aux_full + s, 1); mbar_init(aux_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 16); }   // one arrive per epilogue warp
    mbar_init(w_bar, 1); mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  // the unused second tap of an odd last pair is multiplied by zero weights: it must hold finite numbers
  for (int i = threadIdx.x; i < a.stages * BF_STAGE_BYTES / 16; i += BF_THREADS) reinterpret_cast<uint4*>(sV)[i] = make_uint4(0, 0, 0, 0);
  if (!a.last) {
    for (int i = threadIdx.x; i < 64 * KG; i += BF_THREADS) sAw[i] = a.A[i];
    for (int kg = threadIdx.x; kg < KG; kg += BF_THREADS) {
      const int k = kg / a.G, g = kg % a.G;
      sZb[kg] = (k == 0 ? a.x0 : a.zx + (size_t)(k - 1) * a.zx_kstride) + (size_t)g * a.N;
    }
  }
  for (int i = threadIdx.x; i < 64; i += BF_THREADS) sBias[i] = a.bias ? a.bias[i] : 0.f;
  fence_proxy_async();               // the zero fill is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = tile_lo < tile_hi;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && has_work) {
      mbar_expect_tx(w_bar, (uint32_t)(a.P * a.KB * 8192));
      for (int kb = 0; kb < a.KB; ++kb)
        for (int w = 0; w < a.P; ++w)                            // global: plane w = column blocks [w*KB, (w+1)*KB)
          tma_load_2d(sW + (kb * a.P + w) * 8192, &tmW, w_bar, (w * a.KB + kb) * 64, 0);
      int stage = 0; uint32_t phase = 0;
      int ax = 0; uint32_t aphase = 0;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const long long b = tile / tiles_n;
        const int n0 = (int)(tile % tiles_n) * 128;
        mbar_wait(aux_empty + ax, aphase ^ 1);
        uint8_t* xd = sX + ax * BF_AUX_BYTES;
        mbar_expect_tx(aux_full + ax, BF_AUX_BYTES);
        tma_load_2d(xd, &tmH, aux_full + ax, n0, (int)(b * 64));                                     // half 0: h rows 0..63
        tma_load_2d(xd + 8192, &tmZ, aux_full + ax, n0, (int)(a.zs_row0 + b * a.zs_rowb));          //         Zs rows 64..79
        tma_load_2d(xd + 10240, &tmH, aux_full + ax, n0 + 64, (int)(b * 64));                        // half 1
        tma_load_2d(xd + 10240 + 8192, &tmZ, aux_full + ax, n0 + 64, (int)(a.zs_row0 + b * a.zs_rowb));
        if (++ax == BF_AUX) { ax = 0; aphase ^= 1; }
        for (int pq = 0; pq < NP * a.P; ++pq) {                    // (tap pair p, signal plane q)
          const int p = pq / a.P, q = pq % a.P;
          const int ntap = (2 * p + 1 < a.K) ? 2 : 1;
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* dst = sV + stage * BF_STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, (uint32_t)(ntap * 16384));
          for (int j = 0; j < ntap; ++j) {
            const int k = 2 * p + j;
            const CUtensorMap* tm = (k == 0) ? &tm0 : &tmc;
            const int row = (k == 0) ? (int)(b * 64) : (int)((long long)(k - 1) * a.R + b * 64);
            tma_load_2d(dst + j * 8192, tm, full_bar + stage, q * a.N + n0, row);
            tma_load_2d(dst + 16384 + j * 8192, tm, full_bar + stage, q * a.N + n0 + 64, row);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && has_work) {
      constexpr uint32_t idesc1 = make_idesc_bf16_amn(128, 64);   // dh:  A MN-major (nodes), B K-major (weights), one plane
      constexpr uint32_t idesc1p = make_idesc_bf16_amn(128, 128); //      signal plane 0 against the stacked weight planes
      constexpr uint32_t idesc2 = make_idesc_bf16(128, 64 + BF_ZROWS);   // dB | dAx: both K-major (K = nodes), B = [h; Zs]
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int ax = 0; uint32_t aphase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      bool first = true;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        mbar_wait(aux_full + ax, aphase);
        tc_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)(acc * 128);
        const uint32_t sh = smem_u32(sX + ax * BF_AUX_BYTES);
        for (int pq = 0; pq < NP * a.P; ++pq) {
          const int p = pq / a.P, q = pq % a.P;
          const int ntap = (2 * p + 1 < a.K) ? 2 : 1;
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sv = smem_u32(sV + stage * BF_STAGE_BYTES);
          // MMA1: contraction rows (tap, f) of this pair, 16 at a time.  Split operands: plane 0 against the stacked [W0; W1]
          // (two column halves of D1, summed by the epilogue), plane 1 against W0 (first half)
          for (int ks = 0; ks < 4 * ntap; ++ks) {
            const uint64_t adesc = make_mnmajor_sw128_desc(sv + ks * 2048, 16384);
            const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sW + (2 * p + (ks >> 2)) * a.P * 8192)) + (uint64_t)(2 * (ks & 3));
            umma_f16(d1, adesc, bdesc, (a.P > 1 && q == 0) ? idesc1p : idesc1, (pq | ks) != 0);
          }
          // MMA2 + MMA3: K = 128 nodes = 2 halves x 4 steps, B = [h; Zs] (80 rows); both signal planes accumulate into the same D.
          // Columns 64..79 are the input-tap / bias gradients for pair 0 (rows of tap 0) and unused for the other pairs.
          const uint32_t d2 = tmem_base + TMEM_D2 + (uint32_t)(p * D2_STRIDE);
          const bool fresh = first && q == 0;
#pragma unroll
          for (int hs = 0; hs < 8; ++hs) {
            const int h = hs >> 2, ks = hs & 3;
            const uint64_t adesc = make_kmajor_sw128_desc(sv + h * 16384) + (uint64_t)(2 * ks);
            const uint64_t bdesc = make_kmajor_sw128_desc(sh + h * 10240) + (uint64_t)(2 * ks);
            umma_f16(d2, adesc, bdesc, idesc2, !(fresh && hs == 0));
          }
          umma_commit(empty_bar + stage);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        first = false;
        umma_commit(tmem_full + acc);
        umma_commit(aux_empty + ax);
        if (++ax == BF_AUX) { ax = 0; aphase ^= 1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    // ===== 16 epilogue warps: TMEM lane quarter q = warp % 4 (thread <-> node), feature group cg = 16 columns =====
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m0 = cg * 16;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long tile = tile_lo; tile < tile_hi; ++tile) {
      const long long b = tile / tiles_n;
      const int n = (int)(tile % tiles_n) * 128 + q * 32 + lane;
      float hp[16], dhn[16], z[KG];
      const float vgf = a.gf ? __ldg(a.gf + b * a.gate_stride) : 1.f;
      float vgfn = 1.f, vgin = 1.f, qi = 1.f, qf = 1.f;
      {
        const float* hb = a.hprev + b * a.hprev_bstride + (size_t)m0 * a.N + n;
#pragma unroll
        for (int i = 0; i < 16; ++i) hp[i] = __ldg(hb + (size_t)i * a.N);
      }
      if (!a.last) {
        const float* db = a.dHn + b * a.dHn_bstride + (size_t)m0 * a.N + n;
#pragma unroll
        for (int i = 0; i < 16; ++i) dhn[i] = __ldg(db + (size_t)i * a.N);
        if (a.gfn) vgfn = __ldg(a.gfn + b * a.gate_stride);
        if (NODE) {
          qi = __ldg(a.qin + b * a.q_bstride + n); qf = __ldg(a.qfn + b * a.q_bstride + n);
          if (a.gin) vgin = __ldg(a.gin + b * a.gate_stride);
        }
        const size_t zo = (size_t)b * a.z_bstride + n;
#pragma unroll
        for (int kg = 0; kg < KG; ++kg) z[kg] = __ldg(sZb[kg] + zo);
      }
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + m0), v);
      if (a.P > 1) {                                     // second column half: signal plane 0 x weight plane 1
        float v2[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128 + 64 + m0), v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += v2[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);      // accumulator is in registers: release the TMEM stage
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) part = fmaf(v[i], hp[i], part);
      if (a.last) {
        if (a.dh0) {
          float* of = a.dh0 + ((size_t)b * 64 + m0) * a.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) of[(size_t)i * a.N] = v[i];
        }
        if (a.dgf) {
          part = warp_sum_f(part);
          if (lane == 0) atomicAdd(a.dgf + b * a.gate_stride, vgf > 1e-30f ? part / vgf : 0.f);
        }
      } else {
        const long long ldo = (long long)a.P * a.N;
        __nv_bfloat16* ob = a.v0_out + ((size_t)b * 64 + m0) * ldo + n;
        const float* aw = sAw + m0 * KG;
        float sgi = 0.f, sgf = 0.f, li = 0.f, lf = 0.f;
        const float wfn = vgfn * qf, win = vgin * qi;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dp = (dhn[i] + v[i]) * fmaf(-hp[i], hp[i], 1.f);
          store_planes(ob + (size_t)i * ldo, a.N, a.P, wfn * dp);
          const float bb = sBias[m0 + i];
          float axb = bb;
#pragma unroll
          for (int kg = 0; kg < KG; ++kg) axb = fmaf(aw[i * KG + kg], z[kg], axb);
          sgi = fmaf(dp, axb, sgi);
          sgf = fmaf(dp, bb, sgf);
          if (NODE) {                                    // gf q_f (r + b) = atanh(h) - gi q_i (a + b): no state filter recomputation
            lf = fmaf(dp, atanhf(fminf(fmaxf(hp[i], -0.99999994f), 0.99999994f)) - win * axb, lf);
          }
        }
        if (NODE) {
          li = win * sgi;                                // sum_f dpre gi q_i (a + b) over this thread's 16 features
          atomicAdd(a.dlin_i + b * a.q_bstride + n, (1.f - qi) * li);
          atomicAdd(a.dlin_f + b * a.q_bstride + n, (1.f - qf) * lf);
          sgi *= qi; sgf *= qf;                          // the time gates see q_i (a + b) and q_f b
        }
        if (a.dgf) {                                     // time gating on: three per-sample scalars
          part = warp_sum_f(part); sgi = warp_sum_f(sgi); sgf = warp_sum_f(sgf);
          if (lane == 0) {
            atomicAdd(a.dgf + b * a.gate_stride, vgf > 1e-30f ? part / vgf : 0.f);
            atomicAdd(a.dgin + b * a.gate_stride, sgi);
            atomicAdd(a.dgfn + b * a.gate_stride, sgf);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    // ----- once per launch: fold the TMEM weight-gradient accumulators into this CTA's private partials -----
    if (has_work) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int row = q * 32 + lane;
      if (cg < NP) {
        const int k = 2 * cg + (row >> 6), f = row & 63;
        float* mine = a.part + (size_t)blockIdx.x * a.K * 64 * 64;
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          float w[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_D2 + (uint32_t)(cg * D2_STRIDE + c), w);
          if (k < a.K) {
            float4* o = reinterpret_cast<float4*>(mine + ((size_t)k * 64 + f) * 64 + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 t = o[i];
              t.x += w[4 * i]; t.y += w[4 * i + 1]; t.z += w[4 * i + 2]; t.w += w[4 * i + 3];
              o[i] = t;
            }
          }
        }
      } else if (cg == 3 && row < 64) {
        float w[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_D2 + 64, w);        // pair 0, rows of tap 0, columns 64..79
        float4* o = reinterpret_cast<float4*>(a.partA + ((size_t)blockIdx.x * 64 + row) * BF_ZROWS);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 t = o[i];
          t.x += w[4 * i]; t.y += w[4 * i + 1]; t.z += w[4 * i + 2]; t.w += w[4 * i + 3];
          o[i] = t;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// Zs[(b*T + t)*16 + j][n] (bf16): j < KG: (gi/gf)[b,t] * (x_t S^k)[b,g,n];  j == KG: (gi/gf)[b,t] + 1;  else 0.
// With v_0 = bf16(gf dpre):  sum_n v_0 Zs_j = gi <dpre, z_j>  (-> dA)  and  (gi + gf) sum_n dpre  (-> dbias).
// split == 1 (split-bf16 mode, KG <= 7): rows 8 + j hold the bf16 RESIDUAL of row j, so the tile carries both planes of Zs in
// its 16 rows and MMA3 returns the hi and lo partial products in columns j and 8 + j (summed by dax_reduce_kernel).
// Node gates (qi / qf != null, [BT][N]): the ratio is per node, (gi q_i[n]) / (gf q_f[n]).
__global__ void zs_build_kernel(const float* __restrict__ X, const float* __restrict__ zx, long long zx_kstride, int G, int KG,
                                const float* __restrict__ gi, const float* __restrict__ gf, __nv_bfloat16* __restrict__ Zs,
                                long long BT, int N, int split, const float* __restrict__ qi, const float* __restrict__ qf) {
  const int N8 = N / 8;
  const long long total = BT * BF_ZROWS * N8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % N8);
    const int jr = (int)((i / N8) % BF_ZROWS);
    const long long bt = i / ((long long)N8 * BF_ZROWS);
    const bool lo = split && jr >= 8;
    const int j = lo ? jr - 8 : jr;
    const float vgi = gi ? gi[bt] : 1.f, vgf = gf ? gf[bt] : 1.f;
    float ratio[8];
    if (qi) {
      const float4* a4 = reinterpret_cast<const float4*>(qi + (size_t)bt * N + c * 8);
      const float4* b4 = reinterpret_cast<const float4*>(qf + (size_t)bt * N + c * 8);
      const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
      const float qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, qb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) ratio[e] = (vgi * qa[e]) / fmaxf(vgf * qb[e], 1e-30f);
    } else {
      const float r = vgi / fmaxf(vgf, 1e-30f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ratio[e] = r;
    }
    float o[8];
    if (j < KG) {
      const int k = j / G, g = j % G;
      const float* src = (k == 0 ? X : zx + (size_t)(k - 1) * zx_kstride) + ((size_t)bt * G + g) * N + c * 8;
      const float4 p0 = reinterpret_cast<const float4*>(src)[0], p1 = reinterpret_cast<const float4*>(src)[1];
      o[0] = p0.x; o[1] = p0.y; o[2] = p0.z; o[3] = p0.w; o[4] = p1.x; o[5] = p1.y; o[6] = p1.z; o[7] = p1.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] *= ratio[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (j == KG) ? ratio[e] + 1.f : 0.f;
    }
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
    if (lo) {
      bf16x2_residual(u.x, o[0], o[1]); bf16x2_residual(u.y, o[2], o[3]); bf16x2_residual(u.z, o[4], o[5]); bf16x2_residual(u.w, o[6], o[7]);
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
    }
    reinterpret_cast<uint4*>(Zs)[i] = u;
  }
}

// dA[f, kg] += sum_cta partA[cta][f][kg] (kg < KG);  dbias[f] += sum_cta partA[cta][f][KG]   (+ columns 8 + j when split)
__global__ void dax_reduce_kernel(const float* __restrict__ partA, float* dA, float* dbias, int ncta, int KG, int split) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * BF_ZROWS) return;
  const int f = i / BF_ZROWS, j = i % BF_ZROWS;
  if (j > KG) return;
  float s = 0.f;
  for (int c = 0; c < ncta; ++c) {
    s += partA[(size_t)c * 64 * BF_ZROWS + i];
    if (split) s += partA[(size_t)c * 64 * BF_ZROWS + i + 8];
  }
  if (j < KG) { if (dA) dA[(size_t)f * KG + j] += s; }
  else if (dbias) dbias[f] += s;
}

}  // namespace tc
}  // namespace gcrnn
