// Horner-form forward step on tcgen05 (sm_100a, CTA pairs): the tap contraction of the state filter fused INTO the shift GEMMs.
//
// Reference: r = sum_k B_k (h S^k) + b  (Utils/graphML.py:117-139 inside GGCRNNCell.forward :2402-2423).  The feature mix B_k acts
// on rows, the shift S on columns, so they commute and Horner's rule applies:
//     w_{K-1} = B_{K-1} h ;   w_k = B_k h + w_{k+1} S   (k = K-2 .. 0) ;   r = w_0 + b.
// One launch of this kernel is one Horner stage  OUT = Z S + (I_B (x) W) h  for all samples:
//   * the accumulator is TRANSPOSED w.r.t. tc_gemm2.cuh: TMEM lane = NODE, TMEM column = signal row (b, f).  The main product
//     D[n, (b,f)] = sum_m S^T[n, m] Z[(b,f), m] therefore has the operator tile as the MMA "A" operand (M = 256 nodes per CTA
//     pair) and the signal tile as the "B" operand (N = 256 rows = 4 samples x 64 features), both plain K-major SW128 tiles;
//   * the per-sample mix  D[n, (b, f)] += sum_f' h[b, f', n] W[f, f']  is then an M = 256, N = 64, K = 64 product per sample
//     with the h tile as an MN-major A operand (node contiguous, exactly the operand form of tc_tap.cuh) and W K-major:
//     no block-diagonal padding, +6 % tensor work per operand-plane term instead of a separate HBM-bound kernel that re-reads
//     every chain slab;
//   * split-bf16 operands (tc_gemm.cuh "planes"): main products K-concatenated over (signal plane, operator plane), mix products
//     over (h plane q, weight plane w) with q + w < P; the operator is stored as S / scale, so the mix weights are W / scale
//     and the epilogue multiplies the whole accumulator by scale;
//   * epilogue (8 warps, thread <-> node).  A thread owns ONE node and 16 signal rows per chunk, but memory is node-contiguous:
//     storing element by element (16-bit stores, to global or to a staging tile for TMA) costs one LSU instruction per
//     element-row and made the kernel epilogue-bound (282 .. 327 us per launch against 166 us for the plain GEMM, whatever the
//     number of staging buffers).  Instead each warp transposes its [32 nodes][16 rows] bf16 block through a 1 KB shared-memory
//     buffer: two 16-byte stores per thread (node-major, nodes placed in a permuted slot order), then two ldmatrix.x4.trans, after
//     which a thread holds EIGHT consecutive nodes of one signal row in four registers -> one 16-byte global store (8 rows x 64 B
//     per warp instruction): 6 LSU instructions per chunk and plane instead of 16, no cross-warp synchronisation.
//     The LAST stage (k = 0) adds the input filter A(S)x_t, bias and time gates, applies tanh and writes H[b,t] (fp32, one
//     128-byte row segment per warp store) and the bf16 planes of the new state — the state update of tc_tap.cuh's TAP_FWD.
//
// MEASURED OUTCOME (B200, N = 1024, 131072 signal rows per launch; profiles/r02_horner_notes.txt): an intermediate stage takes
// 224 us (bf16) / 461 us (bf16x2) against 166 / 332 us for the plain pair GEMM, i.e. +35 % / +39 %, although the mix adds only
// +6 % / +9 % flops: an M = 256, N = 64 tcgen05.mma occupies the tensor pipe about as long as an N = 256 one (~130-190 cycles per
// instruction whether its A operand is MN- or K-major, dependent or interleaved accumulators), so the cost is the instruction
// COUNT: 16 (48) mix instructions next to 64 (128) main ones.  With the seed kernel (109 / 213 us) and the final stage's heavier
// epilogue (363 / 575 us with 16 epilogue warps) a forward step costs 1144 / 2171 us against 1016 / 2215 us for chain + tap
// kernel: no gain in bf16, 2 % in bf16x2.  The path is therefore OFF by default (option "fwd_fused" = 1 enables it); it is kept
// because it is correct (tests/test_gpu_tc.py::test_tc_horner_forward_matches_unfused) and documents what the hardware does.
//
// Structure per CTA (576 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM allocator, up to 16 epilogue warps;
// 6-stage ring of 32 KB stages shared by main and mix loads (a mix stage carries one h tile and this CTA's rows of W_k), 2 TMEM
// accumulator stages (512 columns).
#pragma once
#include "tc_gemm2.cuh"
#include "tc_tap.cuh"

namespace gcrnn {
namespace tc {

constexpr int HS_THREADS = 64 + 16 * 32;      // TMA warp + MMA warp + up to 16 epilogue warps
constexpr int HS_STAGES = 6;
constexpr int HS_OUT_BYTES = 16 * 1024;               // per epilogue warp: [32 node slots][16 rows] bf16 transpose buffer
constexpr int HS_SMEM = HS_STAGES * G2_STAGE_BYTES + HS_OUT_BYTES + 192 + (64 * 8 + 64) * 4 + 8 * 8 + 1024;

struct HShiftArgs {
  int M, N;                       // signal rows (B * 64), nodes
  ShiftSegs segs;                 // main products: a = signal plane, b = operator plane
  int P;                          // planes of the signal in / out, of h and of the weights
  float scale;                    // operator scale (the bf16 operator holds S / scale)
  int final_stage;                // 0: write the planes of w_k; 1: state update (k = 0)
  int wcol;                       // column of W_k's plane 0 in the prepared weight matrix (plane q at wcol + q * wpstride)
  int wpstride;
  int exact;
  int epi_warps;                  // 4, 8 or 16 epilogue warps take part (the others idle)
  __nv_bfloat16* out;             // [M][P * N] bf16 planes of the stage's result
  // final stage (time step t): operands of the TAP_FWD epilogue
  float* H; long long H_bstride;                   // fp32 h_t[b] = H + b * H_bstride, [64][N]
  const float* bias;
  const float* gi; const float* gf; long long gate_stride;
  const float* A; int KG, G;                       // input taps [64][KG]
  const float* x0; const float* zx; long long zx_kstride, z_bstride;
};

// TMA prefetch of one box into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HS_THREADS, 1)
hshift_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmH,
              const __grid_constant__ CUtensorMap tmW, const HShiftArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOut = smem + HS_STAGES * G2_STAGE_BYTES;                 // [8 epilogue warps][32 node slots][16 rows] bf16
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + HS_OUT_BYTES);
  uint64_t* empty_bar = full_bar + HS_STAGES;
  uint64_t* tmem_full = empty_bar + HS_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sAw = reinterpret_cast<float*>(tmem_slot + 4);              // [64][KG] input-filter taps (final stage)
  float* sBias = sAw + 64 * 8;                                       // [64]
  const float** sZb = reinterpret_cast<const float**>(sBias + 64);   // [8] base pointer of input row (k, g): X or zx slab

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_nodes = a.N / 256;
  const int tiles_rows = (a.M + 255) / 256;
  const int num_tiles = tiles_rows * tiles_nodes;
  const int num_k = a.N / BK;
  const int NS = 4;                                                  // samples (64-row blocks) per tile

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS); tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmW);
    for (int s = 0; s < HS_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 2 * a.epi_warps); }   // epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  if (a.final_stage) {
    for (int i = threadIdx.x; i < 64 * a.KG; i += HS_THREADS) sAw[i] = a.A[i];
    for (int kg = threadIdx.x; kg < a.KG; kg += HS_THREADS) {
      const int k = kg / a.G, g = kg % a.G;
      sZb[kg] = (k == 0 ? a.x0 : a.zx + (size_t)(k - 1) * a.zx_kstride) + (size_t)g * a.N;
    }
    for (int i = threadIdx.x; i < 64; i += HS_THREADS) sBias[i] = a.bias ? a.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 nodes of the operator, own half of the signal rows, own nodes of h =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int r0 = (tile / tiles_nodes) * 256;                               // first signal row of the tile
        const int n0 = (tile % tiles_nodes) * 256 + (int)rank * 128;             // this CTA's first node
        // The h tiles of this tile are its only loads that always miss L2 (every tile owns its node columns of h), and the ring's
        // look-ahead is ~6 stages x 135 ns < HBM latency: without this prefetch every mix stage cost ~1 us of tensor-pipe idle time
        // (ncu: tensor pipe 65 % active vs 83 % for the plain GEMM).  Pull them into L2 while the main loop runs.
        for (int s = 0; s < NS; ++s)
          for (int q = 0; q < a.P; ++q) {
            tma_prefetch_l2_2d(&tmH, q * a.N + n0, r0 + 64 * s);
            tma_prefetch_l2_2d(&tmH, q * a.N + n0 + 64, r0 + 64 * s);
          }
        for (int sg = 0; sg < a.segs.n; ++sg) {
          const int zcol = a.segs.a[sg] * a.N, srow = a.segs.b[sg] * a.N + n0;
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* sa = smem + stage * G2_STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * G2_STAGE_BYTES);
            tma_load_2d_pair(sa, &tmS, full_bar + stage, kb * BK, srow);
            tma_load_2d_pair(sa + G2_HALF_BYTES, &tmZ, full_bar + stage, zcol + kb * BK, r0 + (int)rank * 128);
            if (++stage == HS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        for (int s = 0; s < NS; ++s) {                        // h tiles of the tile's four samples
          for (int q = 0; q < a.P; ++q) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* sa = smem + stage * G2_STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * (G2_HALF_BYTES + a.P * 4096));
            tma_load_2d_pair(sa, &tmH, full_bar + stage, q * a.N + n0, r0 + 64 * s);            // rows >= M: zero filled
            tma_load_2d_pair(sa + 8192, &tmH, full_bar + stage, q * a.N + n0 + 64, r0 + 64 * s);
            for (int w = 0; w < a.P; ++w)                                        // this CTA's 32 rows of W_k, every plane (L2 hits)
              tma_load_2d_pair(sa + G2_HALF_BYTES + w * 4096, &tmW, full_bar + stage, a.wcol + w * a.wpstride, (int)rank * 32);
            if (++stage == HS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256);
      constexpr uint32_t idesc_mix = make_idesc_bf16_amn(256, 64);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        const int total_k = a.segs.n * num_k;
        for (int kb = 0; kb < total_k; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + G2_HALF_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit_pair(empty_bar + stage);
          if (++stage == HS_STAGES) { stage = 0; phase ^= 1; }
        }
        for (int s = 0; s < NS; ++s) {
          for (int q = 0; q < a.P; ++q) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
            for (int w = 0; w + q < a.P; ++w) {
              const uint64_t bdesc0 = make_kmajor_sw128_desc(sa + G2_HALF_BYTES + w * 4096);
#pragma unroll
              for (int j = 0; j < 4; ++j) {                                      // 64 contraction rows f' in steps of 16
                const uint64_t adesc = make_mnmajor_sw128_desc(sa + j * 2048, 8192);
                umma_f16_pair(d_tmem + (uint32_t)(64 * s), adesc, bdesc0 + (uint64_t)(2 * j), idesc_mix, 1u);
              }
            }
            umma_commit_pair(empty_bar + stage);
            if (++stage == HS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit_pair(tmem_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp - 2 < a.epi_warps) {
    // ===== epilogue warps: node quarter q = warp % 4 (TMEM lanes); column group g = (warp - 2) / 4 takes 1024 / epi_warps of the
    // 256 columns.  Intermediate stages run at the same speed with 4 or 8 warps; the final stage's epilogue (input filter, gates,
    // tanh, fp32 H, planes) is the heavy one: 720 us per launch with 4 warps, ~460 with 8, 363 with 16 (bf16).
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const int cols_per_warp = 1024 / a.epi_warps;                   // 256, 128 or 64 accumulator columns per warp
    const long long LD = (long long)a.P * a.N;
    // transpose buffer: lane l (node l of the warp's 32) owns slot 8j + 2c + e with l = 8c + 2j + e, so that after
    // ldmatrix.x4.trans thread t holds nodes 8 (t % 4) .. + 7 of signal row (t / 4) in its four registers
    const uint32_t tb = smem_u32(sOut + (warp - 2) * 1024);
    const uint32_t my_slot = (uint32_t)(8 * ((lane & 7) >> 1) + 2 * (lane >> 3) + (lane & 1));
    const uint32_t st_addr = tb + my_slot * 32;
    const uint32_t ld_addr = tb + (uint32_t)lane * 32;             // thread 8 i + m supplies row m (= slot 8 i + m) of matrix i
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int r0 = (tile / tiles_nodes) * 256;
      const int nw0 = (tile % tiles_nodes) * 256 + (int)rank * 128 + q * 32;     // first node of this warp
      const int n = nw0 + lane;                                     // this thread's node (TMEM lane)
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + g * cols_per_warp);
      const int nchunk = cols_per_warp / 16;
      float z[8];
      float vgi = 1.f, vgf = 1.f;
#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {                            // 16 accumulator columns = 16 signal rows (one sample, 16 features)
        const int row0 = r0 + g * cols_per_warp + c * 16;           // first signal row (b * 64 + f) of the chunk
        const long long b = row0 >> 6;
        const int f0 = row0 & 63;
        const bool live = row0 < a.M;
        if (a.final_stage && (c & 3) == 0 && live) {                // new sample: gates and the node's input-filter rows
          if (a.gi) vgi = __ldg(a.gi + b * a.gate_stride);
          if (a.gf) vgf = __ldg(a.gf + b * a.gate_stride);
          const size_t zo = (size_t)b * a.z_bstride + n;
#pragma unroll
          for (int kg = 0; kg < 8; ++kg) z[kg] = kg < a.KG ? __ldg(sZb[kg] + zo) : 0.f;
        }
        float v[16];
        tmem_ld16(t0 + (uint32_t)(c * 16), v);
        if (c == nchunk - 1) {                                      // accumulator fully read: hand the TMEM stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tmem_empty + acc);
        }
        if (!live) continue;                                        // rows beyond M (clipped last row tile); warp-uniform
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= a.scale;
        if (a.final_stage) {
          const float gsum = vgi + vgf;
          float* of = a.H + b * a.H_bstride + (size_t)f0 * a.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float* aw = sAw + (f0 + i) * a.KG;
            float ax = 0.f;
#pragma unroll
            for (int kg = 0; kg < 8; ++kg) if (kg < a.KG) ax = fmaf(aw[kg], z[kg], ax);
            const float pre = fmaf(vgi, ax, fmaf(vgf, v[i], gsum * sBias[f0 + i]));       // gi (ax + b) + gf (r + b)
            const float h = a.exact ? tanh_acc(pre) : tap_tanh(pre);
            of[(size_t)i * a.N] = h;
            v[i] = h;
          }
        }
        for (int pl = 0; pl < a.P; ++pl) {                          // plane 0 = bf16(v), plane 1 = bf16(v - plane 0)
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            pk[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
            if (pl + 1 < a.P) bf16x2_residual(pk[k], v[2 * k], v[2 * k + 1]);
          }
          __syncwarp();                                             // the previous ldmatrix reads of this buffer are done
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_addr + 16), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
          __syncwarp();
          __nv_bfloat16* ob = a.out + (size_t)(row0 + (lane >> 2)) * LD + (size_t)pl * a.N + nw0 + 8 * (lane & 3);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                          // signal rows row0 + 8 hh + lane / 4
            uint32_t r[4];
            ldsm_x4_trans(r, ld_addr + hh * 16);
            *reinterpret_cast<uint4*>(ob + (size_t)(8 * hh) * LD) = make_uint4(r[0], r[1], r[2], r[3]);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

// W[f, k, g] fp32 (the reference's weight_B[f, 0, k, g]) -> bf16 planes of ONE tap divided by `scale`:
// out[f][q * 64 + g], row length P * 64
__global__ void prep_tap_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int F, int K, int k, int P, float inv_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * F) return;
  const int g = i % F, f = i / F;
  store_planes(out + (size_t)f * P * F + g, F, P, W[((size_t)f * K + k) * F + g] * inv_scale);
}

}  // namespace tc
}  // namespace gcrnn
