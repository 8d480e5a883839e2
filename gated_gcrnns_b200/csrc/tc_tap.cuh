// tcgen05 tap contraction for sm_100a (per sample b):   Y^T[n, m] = sum_{k,c} Z_k[b, c, n] * W[m, k*C + c]
//
// This is the reference's `y = z.reshape(B,N,EKG) @ h.reshape(F,EKG)^T` (Utils/graphML.py:134-135) with the K
// shifted signals kept as K separate bf16 slabs [B][C][N] (no `cat`).  The contraction index (k, c) runs over slab
// ROWS, and n is contiguous in memory, so the signal tile is an MN-major ("transposed") UMMA operand:
//   A (M = 128 nodes, K = 64 (k,c) rows)  : TMA boxes {64 n, C rows} -> smem [64 rows][128 B] SW128, MN-major descriptor
//   B (N = M_out features, K = 64)        : prepared weights W[m][k*C+c], K-major SW128, resident in smem for the CTA
//   D (TMEM)                              : lane = node n, column = output feature m  (fp32)
// so an epilogue thread owns ONE node and all output features: stores to [b, m, n] are coalesced across the warp.
// Same warp-specialised persistent structure as tc_gemm.cuh (TMA producer warp, single-thread MMA issuer,
// 4 epilogue warps, double-buffered TMEM accumulators).  It is HBM-bound (reads K slabs once); the MMAs are ~5 % of
// the tile time.  Fused epilogues: EPI_FWD (input filter + bias + time gates + tanh, writes H[b,t] and the bf16 state),
// EPI_BWD (dg_f reduction, dh_rec = g_f q), EPI_PLAIN (gate sub-cell term).
#pragma once
#include "tc_gemm.cuh"

namespace gcrnn {
namespace tc {

constexpr int TAP_BM = 128;        // nodes per tile
constexpr int TAP_STAGES = 8;      // ring of 64-row stages (16 KB each); 6 with two operand planes (shared-memory budget)
constexpr int TAP_STAGE_BYTES = 2 * 64 * 128;
constexpr int TAP_MAX_KB = 6;      // K*C <= 384 contraction rows
constexpr int TAP_THREADS = 64 + 16 * 32;   // TMA warp + MMA warp + 16 epilogue warps

enum { TAP_PLAIN = 0, TAP_FWD = 1, TAP_BWD = 2, TAP_BWDF = 3, TAP_MIX = 4 };   // BWDF: BWD fused with the next step's dpre;
                                                                             // MIX: bf16 planes of the plain product (Horner seed w_{K-1} = B_{K-1} h)

struct TapArgs {
  int K, C, M, N, KB;              // slabs, channels per slab, output features, nodes, ceil(K*C/64)
  int P;                           // operand planes (1: bf16, 2: split bf16 hi + lo) of slabs, weights and bf16 outputs
  int stages;                      // stage ring depth (<= TAP_STAGES)
  int exact;                       // 1: tanh through ex2 + rcp (abs. error ~2e-7) instead of tanh.approx (2^-11)
  long long B, R;                  // samples, rows per slab (B*C)
  // epilogue
  float* out_f32; long long out_bstride;
  __nv_bfloat16* out_bf16;
  const float* bias; float bias_scale;
  const float* gi; const float* gf; long long gate_stride;
  const float* A; int Kin, G;
  const float* x0; long long x0_bstride;
  const float* zx; long long zx_kstride, zx_bstride;
  const float* hprev; long long hprev_bstride;
  float* dgf; int accumulate;
  int scaled_chain;              // TAP_BWD: the chain input was g_f * dpre, so acc = g_f q already
  // TAP_BWDF (step t): additionally forms step t-1's  dpre = (dH_{t-1} + acc) (1 - h_{t-1}^2), writes
  // out_bf16 = bf16(gf_{t-1} dpre) (the next chain input) and accumulates red[b][m][0] = sum_n dpre,
  // red[b][m][1+kg] = sum_n dpre * zx_kg (x0 / zx then point at step t-1's rows)
  const float* dHn; long long dHn_bstride;
  const float* gfn;              // gf[b, t-1] at gfn[b * gate_stride] (null: 1)
  float* red;                    // [B][M][8]
  // TAP_FWD with node gates (tc_node.cuh): h = tanh(gi q_i[n] (ax + b) + gf q_f[n] (v + b)), q at q + b * q_bstride + n
  const float* qi; const float* qf; long long q_bstride;
};

// sum over the 32 lanes of 32 per-lane quantities with 31 shuffles: afterwards lane l holds the total of x[l] in x[0]
__device__ __forceinline__ float warp_transpose_sum32(float* x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? x[i] : x[i + s];
      const float keep = up ? x[i + s] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

__device__ __forceinline__ float tap_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MN-major, 128B-swizzled operand: 64-element (128 B) blocks along M at stride LBO, 8-row groups along K at stride SBO
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_amn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// KGM: number of (k, g) input-filter taps the epilogue is unrolled for.  EXACT: Kin*G == KGM (no predicates in the
// unrolled tap loops); otherwise KGM is an upper bound.
template <int EPI, int KGM, bool EXACT>
__global__ void __launch_bounds__(TAP_THREADS, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tmc,
                const __grid_constant__ CUtensorMap tmW, const TapArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment by OFFSET (not by integer round-trip) so the compiler keeps the shared address space: LDS, not LD
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                              // [KB][P][M rows][128 B]: the planes of a block are STACKED
  uint8_t* sA = smem + a.P * a.KB * 8192;                          // [stages][2 halves][64 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + a.stages * TAP_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + TAP_STAGES;
  uint64_t* tmem_full = empty_bar + TAP_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* sAw = reinterpret_cast<float*>(tmem_slot + 4);            // [M][Kin*G] input-filter taps (<= 64*32)
  float* sBias = sAw + 64 * 32;                                    // [64]
  const float** sZb = reinterpret_cast<const float**>(sBias + 64); // [32] base pointer of input row (k, g): X or zx slab

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = a.N / TAP_BM;
  const long long num_tiles = a.B * tiles_n;
  // contiguous tile range per CTA: consecutive tiles are the node blocks of one sample (per-sample reductions
  // are then accumulated in registers and flushed once per sample)
  const long long per_cta = (num_tiles + gridDim.x - 1) / gridDim.x;
  const long long tile_lo = blockIdx.x * per_cta;
  const long long tile_hi = tile_lo + per_cta < num_tiles ? tile_lo + per_cta : num_tiles;
  const int KK = a.K * a.C;
  // Split operands: signal plane 0 meets BOTH weight planes in ONE instruction whose B operand is the stacked [W0; W1] block
  // (N = 2 M): its two column halves are summed by the epilogue.  An M = 128 tcgen05.mma with N = 64 occupies the tensor pipe
  // about as long as one with N = 128 (profiles/r02_horner_notes.txt), and with 3 separate plane products per k-step the
  // kernel was bound by the instruction count (60 per tile), not by HBM: 895 us per launch for 3.8 GB at cfg3.
  const int NW = a.P * a.M;                                        // accumulator columns written by a plane-0 instruction
  const int acc_stride = NW < 32 ? 32 : NW;                        // TMEM columns per accumulator stage
  const uint32_t tmem_cols = (2 * acc_stride <= 32) ? 32 : (2 * acc_stride <= 64) ? 64 : (2 * acc_stride <= 128) ? 128 : 256;
  const int wblk = a.M * 128;                                      // bytes of one plane of one 64-column weight block

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm0); tma_prefetch_desc(&tmc); tma_prefetch_desc(&tmW);
    for (int s = 0; s < TAP_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 128 * ((a.M + 15) / 16)); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  if (EPI == TAP_FWD) for (int i = threadIdx.x; i < a.M * a.Kin * a.G; i += TAP_THREADS) sAw[i] = a.A[i];
  if (EPI == TAP_FWD || EPI == TAP_BWDF) {
    for (int kg = threadIdx.x; kg < a.Kin * a.G; kg += TAP_THREADS) {
      const int k = kg / a.G, g = kg % a.G;
      sZb[kg] = (k == 0 ? a.x0 : a.zx + (size_t)(k - 1) * a.zx_kstride) + (size_t)g * a.N;
    }
  }
  for (int i = threadIdx.x; i < 64; i += TAP_THREADS) sBias[i] = (a.bias && i < a.M) ? a.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)(a.P * a.KB * a.M * 128));
      for (int kb = 0; kb < a.KB; ++kb)
        for (int w = 0; w < a.P; ++w)                              // global: plane w = column blocks [w*KB, (w+1)*KB)
          tma_load_2d(sW + (kb * a.P + w) * wblk, &tmW, w_bar, (w * a.KB + kb) * 64, 0);
      int stage = 0; uint32_t phase = 0;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const long long b = tile / tiles_n;
        const int n0 = (int)(tile % tiles_n) * TAP_BM;
        for (int sq = 0; sq < a.KB * a.P; ++sq) {                  // (64-row block s, signal plane q)
          const int s = sq / a.P, q = sq % a.P;
          const int rows = min(64, KK - 64 * s);
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* dst = sA + stage * TAP_STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, (uint32_t)(2 * rows * 128));
          for (int r0 = 0; r0 < rows; r0 += a.C) {                 // one box per slab touched by this stage
            const int k = (64 * s + r0) / a.C;
            const CUtensorMap* tm = (k == 0) ? &tm0 : &tmc;
            const int row = (k == 0) ? (int)(b * a.C) : (int)((long long)(k - 1) * a.R + b * a.C);
            tma_load_2d(dst + r0 * 128, tm, full_bar + stage, q * a.N + n0, row);
            tma_load_2d(dst + 8192 + r0 * 128, tm, full_bar + stage, q * a.N + n0 + 64, row);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16_amn(TAP_BM, a.M);    // one weight plane
      const uint32_t idescP = make_idesc_bf16_amn(TAP_BM, NW);     // signal plane 0 against the stacked weight planes
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
        for (int sq = 0; sq < a.KB * a.P; ++sq) {
          const int s = sq / a.P, q = sq % a.P;
          const int rows = min(64, KK - 64 * s);
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(sA + stage * TAP_STAGE_BYTES);
          // split operands: z W ~= z0 [W0; W1] (two column halves, summed in the epilogue) + z1 W0
          const uint32_t sb = smem_u32(sW + s * a.P * wblk);
          for (int j = 0; j < rows / 16; ++j) {
            const uint64_t adesc = make_mnmajor_sw128_desc(sa + j * 2048, 8192);   // 16 K-rows = 2048 B
            const uint64_t bdesc = make_kmajor_sw128_desc(sb) + (uint64_t)(2 * j); // 16 bf16 = 32 B along K
            // the very first instruction of a tile (plane 0) overwrites all NW columns
            umma_f16(d_tmem, adesc, bdesc, q == 0 ? idescP : idesc1, (sq | j) != 0);
          }
          umma_commit(empty_bar + stage);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== 16 epilogue warps: TMEM lane quarter q = warp % 4 (thread <-> node), feature group cg = 16 columns =====
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int m0 = cg * 16;
    if (m0 < a.M) {
      const int KG = EXACT ? KGM : a.Kin * a.G;
      int acc = 0; uint32_t acc_phase = 0;
      long long cur_b = -1;
      float racc[4] = {0.f, 0.f, 0.f, 0.f};
      for (long long tile = tile_lo; tile < tile_hi; ++tile) {
        const long long b = tile / tiles_n;
        const int n = (int)(tile % tiles_n) * TAP_BM + q * 32 + lane;
        // operands that do not depend on the accumulator are fetched before waiting for the MMAs
        float z[KGM], hp[16];
        float vgi = 1.f, vgf = 1.f;
        if (EPI != TAP_PLAIN) {
          if (a.gi) vgi = __ldg(a.gi + b * a.gate_stride);
          if (a.gf) vgf = __ldg(a.gf + b * a.gate_stride);
        }
        if (EPI == TAP_FWD) {
          const size_t zo = (size_t)b * a.x0_bstride + n;      // x0_bstride == zx_bstride (rows of one [B,T,G,N] layout)
#pragma unroll
          for (int kg = 0; kg < KGM; ++kg) z[kg] = (kg < KG) ? __ldg(sZb[kg] + zo) : 0.f;
        }
        float dhn[16];
        float vgfn = 1.f;
        if (EPI == TAP_BWD || EPI == TAP_BWDF) {
          const float* hb = a.hprev + b * a.hprev_bstride + (size_t)m0 * a.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) hp[i] = __ldg(hb + (size_t)i * a.N);
        }
        if (EPI == TAP_BWDF) {
          const float* db = a.dHn + b * a.dHn_bstride + (size_t)m0 * a.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) dhn[i] = __ldg(db + (size_t)i * a.N);
          if (a.gfn) vgfn = __ldg(a.gfn + b * a.gate_stride);
          {
            const size_t zo = (size_t)b * a.x0_bstride + n;
#pragma unroll
            for (int kg = 0; kg < KGM; ++kg) z[kg] = (kg < KG) ? __ldg(sZb[kg] + zo) : 0.f;
          }
          if (b != cur_b) {                       // new sample: flush the per-sample sums of the previous one
            if (cur_b >= 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) atomicAdd(a.red + ((size_t)cur_b * a.M + m0 + 4 * j + (lane >> 3)) * 8 + (lane & 7), racc[j]);
            }
            cur_b = b;
            racc[0] = racc[1] = racc[2] = racc[3] = 0.f;
          }
        }
        mbar_wait(tmem_full + acc, acc_phase);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride + m0), v);
        if (a.P > 1) {                                  // second column half: signal plane 0 x weight plane 1
          float v2[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride + a.M + m0), v2);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += v2[i];
        }
        tc_fence_before();
        mbar_arrive(tmem_empty + acc);                  // accumulator values are in registers: release the TMEM stage early
        float* of = a.out_f32 + b * a.out_bstride + (size_t)m0 * a.N + n;
        if (EPI == TAP_PLAIN) {
#pragma unroll
          for (int i = 0; i < 16; ++i) of[(size_t)i * a.N] = v[i] + a.bias_scale * sBias[m0 + i];
        } else if (EPI == TAP_MIX) {
          const long long ldo = (long long)a.P * a.N;
          __nv_bfloat16* ob = a.out_bf16 + ((size_t)b * a.M + m0) * ldo + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) store_planes(ob + (size_t)i * ldo, a.N, a.P, v[i]);
        } else if (EPI == TAP_FWD) {
          const long long ldo = (long long)a.P * a.N;          // bf16 row = P planes of N
          __nv_bfloat16* ob = a.out_bf16 + ((size_t)b * a.M + m0) * ldo + n;
          const float* aw = sAw + m0 * KG;
          if (a.qi) { vgi *= __ldg(a.qi + b * a.q_bstride + n); vgf *= __ldg(a.qf + b * a.q_bstride + n); }      // per-node gates
          const float gsum = vgi + vgf;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float ax = 0.f;
#pragma unroll
            for (int kg = 0; kg < KGM; ++kg) if (kg < KG) ax = fmaf(aw[i * KG + kg], z[kg], ax);
            // gi (ax + b) + gf (v + b)
            const float pre = fmaf(vgi, ax, fmaf(vgf, v[i], gsum * sBias[m0 + i]));
            const float h = a.exact ? tanh_acc(pre) : tap_tanh(pre);
            of[(size_t)i * a.N] = h;
            store_planes(ob + (size_t)i * ldo, a.N, a.P, h);
          }
        } else if (EPI == TAP_BWDF) {
          float part = 0.f;
          const long long ldo = (long long)a.P * a.N;
          __nv_bfloat16* ob = a.out_bf16 + ((size_t)b * a.M + m0) * ldo + n;
#pragma unroll
          for (int j = 0; j < 4; ++j) {                 // 4 features x 8 kinds = 32 per-lane quantities per group
            float x[32];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int i = 4 * j + ii;
              part = fmaf(v[i], hp[i], part);
              const float dp = (dhn[i] + v[i]) * (1.f - hp[i] * hp[i]);
              store_planes(ob + (size_t)i * ldo, a.N, a.P, vgfn * dp);
              x[8 * ii] = dp;
#pragma unroll
              for (int kg = 0; kg < 7; ++kg) x[8 * ii + 1 + kg] = (kg < KGM && kg < KG) ? dp * z[kg < KGM ? kg : 0] : 0.f;
            }
            racc[j] += warp_transpose_sum32(x, lane);
          }
          if (a.dgf) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            part = vgf > 1e-30f ? part / vgf : 0.f;
            if (lane == 0) atomicAdd(a.dgf + b * a.gate_stride, part);
          }
        } else {
          float part = 0.f;
          float old[16];
          if (a.accumulate) {
#pragma unroll
            for (int i = 0; i < 16; ++i) old[i] = of[(size_t)i * a.N];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            part = fmaf(v[i], hp[i], part);
            float r = a.scaled_chain ? v[i] : vgf * v[i];
            if (a.accumulate) r += old[i];
            of[(size_t)i * a.N] = r;
          }
          if (a.dgf) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (a.scaled_chain) part = vgf > 1e-30f ? part / vgf : 0.f;      // <q, h> = <g_f q, h> / g_f
            if (lane == 0) atomicAdd(a.dgf + b * a.gate_stride, part);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (EPI == TAP_BWDF && cur_b >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(a.red + ((size_t)cur_b * a.M + m0 + 4 * j + (lane >> 3)) * 8 + (lane & 7), racc[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// consumes (and clears) red[b][m][8] of one time step: dA += gi sz, dbias += (gi + gf) sdp,
// dgi[b] += sum_m (bias sdp + sum_kg A sz), dgf[b] += sum_m bias sdp        (block = 64 threads = features)
__global__ void dpre_finish_kernel(float* __restrict__ red, const float* __restrict__ gi, const float* __restrict__ gf,
                                   long long gate_stride, const float* __restrict__ A, const float* __restrict__ bias,
                                   float* dA, float* dbias, float* dgi, float* dgf, long long B, int M, int KG, int bchunk) {
  __shared__ float sh[2][2];
  const int m = threadIdx.x;
  const long long b0 = (long long)blockIdx.x * bchunk, b1 = b0 + bchunk < B ? b0 + bchunk : B;
  float aA[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, abias = 0.f;
  float wA[7];
#pragma unroll
  for (int kg = 0; kg < 7; ++kg) wA[kg] = (m < M && kg < KG) ? A[(size_t)m * KG + kg] : 0.f;
  const float bb = (bias && m < M) ? bias[m] : 0.f;
  for (long long b = b0; b < b1; ++b) {
    const float vgi = gi ? gi[b * gate_stride] : 1.f, vgf = gf ? gf[b * gate_stride] : 1.f;
    float si = 0.f, sf = 0.f;
    if (m < M) {
      float4* r = reinterpret_cast<float4*>(red + ((size_t)b * M + m) * 8);
      const float4 r0 = r[0], r1 = r[1];
      r[0] = make_float4(0.f, 0.f, 0.f, 0.f); r[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float sz[7] = {r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      const float sdp = r0.x;
      abias += (vgi + vgf) * sdp;
      float dot = 0.f;
#pragma unroll
      for (int kg = 0; kg < 7; ++kg) { aA[kg] = fmaf(vgi, sz[kg], aA[kg]); dot = fmaf(wA[kg], sz[kg], dot); }
      si = bb * sdp + dot; sf = bb * sdp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { si += __shfl_xor_sync(0xffffffffu, si, o); sf += __shfl_xor_sync(0xffffffffu, sf, o); }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = si; sh[threadIdx.x >> 5][1] = sf; }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (dgi) atomicAdd(dgi + b * gate_stride, sh[0][0] + sh[1][0]);
      if (dgf) atomicAdd(dgf + b * gate_stride, sh[0][1] + sh[1][1]);
    }
    __syncthreads();
  }
  if (m < M) {
    if (dbias) atomicAdd(dbias + m, abias);
    if (dA) for (int kg = 0; kg < KG; ++kg) atomicAdd(dA + (size_t)m * KG + kg, aA[kg]);
  }
}


// =====================================================================================================
// wgrad on tcgen05 (F = 64):  dB_k[f, g] += sum_{b, n} V_k[b, f, n] * h[b, g, n]   for all K taps at once
//   A (K-major): the K slabs' [64 f x 64 n] boxes stacked in smem -> pairs of taps form M = 128 operands
//   B (K-major): h_{t-1} bf16 [64 g x 64 n];  D_p (TMEM, 64 columns per tap pair) accumulates over the CTA's
//   whole list of (sample, 64-node block) tiles; one read-modify-write of the CTA's private partial at the end.
// =====================================================================================================
constexpr int WT_STAGES = 4;
struct WgradTcArgs {
  int K, N; long long B, R;
  int P;                       // planes of the v slabs (both are summed); h contributes its plane 0 only: rounding noise of h
                               // is independent across the B*N summands of a weight gradient and averages out
  float* part;                 // [grid][K][64][64]
};
__host__ __device__ constexpr int wt_stage_bytes(int K) { return (K + 1) * 8192; }

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tmc,
                const __grid_constant__ CUtensorMap tmH, const WgradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = wt_stage_bytes(a.K);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WT_STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + WT_STAGES;
  uint64_t* done_bar = empty_bar + WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = a.N / 64;
  const long long num_tiles = a.B * tiles_n;
  const int NP = (a.K + 1) / 2;
  const uint32_t tmem_cols = NP * 64 <= 64 ? 64 : NP * 64 <= 128 ? 128 : 256;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm0); tma_prefetch_desc(&tmc); tma_prefetch_desc(&tmH);
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = (long long)blockIdx.x < num_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long b = tile / tiles_n;
        const int n0 = (int)(tile % tiles_n) * 64;
        for (int q = 0; q < a.P; ++q) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* dst = smem + stage * stage_bytes;
          mbar_expect_tx(full_bar + stage, (uint32_t)stage_bytes);
          for (int k = 0; k < a.K; ++k) {
            const CUtensorMap* tm = (k == 0) ? &tm0 : &tmc;
            const int row = (k == 0) ? (int)(b * 64) : (int)((long long)(k - 1) * a.R + b * 64);
            tma_load_2d(dst + k * 8192, tm, full_bar + stage, q * a.N + n0, row);
          }
          tma_load_2d(dst + a.K * 8192, &tmH, full_bar + stage, n0, (int)(b * 64));
          if (++stage == WT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64);
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int q = 0; q < a.P; ++q) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + a.K * 8192);
          for (int p = 0; p < NP; ++p) {
            const uint64_t adesc = make_kmajor_sw128_desc(sa + p * 16384);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16(tmem_base + (uint32_t)(p * 64), adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, !(first && j == 0));
          }
          first = false;
          umma_commit(empty_bar + stage);
          if (++stage == WT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(done_bar);
    }
  } else if (has_work) {
    // epilogue: TMEM lane = row of the stacked pair (tap 2p rows 0..63, tap 2p+1 rows 64..127), column = g
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* mine = a.part + (size_t)blockIdx.x * a.K * 64 * 64;
    const int row = q * 32 + lane;
    for (int p = 0; p < NP; ++p) {
      const int k = 2 * p + (row >> 6);
      const int f = row & 63;
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * 64 + c), v);
        if (k < a.K) {
          float4* o = reinterpret_cast<float4*>(mine + ((size_t)k * 64 + f) * 64 + c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 t = o[i];
            t.x += v[4 * i]; t.y += v[4 * i + 1]; t.z += v[4 * i + 2]; t.w += v[4 * i + 3];
            o[i] = t;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dynamic shared memory of tap_gemm_kernel: P weight planes of KB blocks + the stage ring + barriers / taps / bias / pointers
inline int tap_smem_bytes(int P, int KB, int stages) {
  return P * KB * 8192 + stages * TAP_STAGE_BYTES + 256 + (64 * 32 + 64) * 4 + 32 * 8 + 1024;
}

}  // namespace tc
}  // namespace gcrnn
