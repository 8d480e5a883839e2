// tcgen05 shift GEMM, CTA-pair variant (cta_group::2) for sm_100a:  C[M, N] = A[M, K] * Bop[N, K]^T
//
// Same operation as tc_gemm.cuh (the dense graph shift `torch.matmul(x, S)`, Utils/graphML.py:123, batched over all
// (sample, feature) rows), but two CTAs of one cluster (the two SMs of a TPC) work on one 256 x 256 output tile:
//   * each CTA owns 128 of the 256 rows (its own A tile and its own TMEM accumulator [128 lanes x 256 columns]) and
//     loads only HALF of the B tile (128 of the 256 operator rows); the MMA unit of each SM reads the other half
//     from its peer's shared memory.  Per 128x256x64 block a CTA now pulls 32 KB from L2 instead of 48 KB: the
//     single-CTA kernel is capped by the L2->SM fabric (~14 TB/s aggregate at 1.19 PFLOP/s), this one is not;
//   * the leader CTA (cluster rank 0) issues `tcgen05.mma.cta_group::2` (M = 256) for the pair; TMA loads of both
//     CTAs complete on the LEADER's full barrier; `tcgen05.commit ... multicast::cluster` releases the smem stage
//     in both CTAs and publishes the accumulator to both epilogues.
// Structure per CTA (192 threads): warp 0 TMA producer, warp 1 MMA issuer (leader only) + TMEM allocator,
// warps 2..5 epilogue (TMEM -> registers -> swizzled smem staging -> TMA bulk store); 6-stage smem ring (32 KB per
// stage), 2 TMEM accumulator stages (512 columns).
//
// Split-bf16 operands (tc_gemm.cuh "planes"): the K loop runs over `segs.n` K-concatenated products
// sum_s A[:, plane a_s] * Bop[plane b_s]^T into the SAME TMEM accumulator (signal hi/lo x operator hi/lo), and the
// epilogue splits the fp32 accumulator into `out_planes` bf16 planes (hi, residual), one TMA store per plane.
#pragma once
#include "tc_gemm.cuh"

namespace gcrnn {
namespace tc {

constexpr int G2_BN = 256;
constexpr int G2_STAGES = 6;
constexpr int G2_HALF_BYTES = 128 * BK * 2;             // one 128-row K-major SW128 tile = 16 KB
constexpr int G2_STAGE_BYTES = 2 * G2_HALF_BYTES;       // A tile + B half
constexpr int G2_OUT_BYTES = 4 * 2 * 4096;               // 4 epilogue warps x 2 staging buffers x [32 rows][128 B]
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + G2_OUT_BYTES + 256 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;             // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (same smem offset, rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once every MMA issued so far has completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader CTA's barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

// bulk tensor store smem -> global (bf16 output tile through the TMA unit instead of 16 B-per-lane strided STGs:
// with direct stores the kernel was bound by the epilogue's store path, not by the tensor pipe — see profiles/)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
shift_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, Epi epi, int M, int N, const ShiftSegs segs, int out_planes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOut = smem + G2_STAGES * G2_STAGE_BYTES;                 // [4 warps][2][32 rows][128 B] SW128
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + G2_OUT_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_n = N / G2_BN;
  const int tiles_m = (M + 255) / 256;
  const int num_tiles = tiles_m * tiles_n;
  const int num_k = N / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows, own half of the operator rows =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m0 = (tile / tiles_n) * 256 + (int)rank * 128;
        const int n0 = (tile % tiles_n) * G2_BN + (int)rank * 128;
        for (int sg = 0; sg < segs.n; ++sg) {
          const int acol = segs.a[sg] * N, brow = segs.b[sg] * N + n0;
          for (int kb = 0; kb < num_k; ++kb) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* sa = smem + stage * G2_STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * G2_STAGE_BYTES);     // bytes of BOTH CTAs
            tma_load_2d_pair(sa, &tmA, full_bar + stage, acol + kb * BK, m0);
            tma_load_2d_pair(sa + G2_HALF_BYTES, &tmB, full_bar + stage, kb * BK, brow);
            if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both tensor cores =====
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, G2_BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G2_BN);
        const int total_k = segs.n * num_k;                 // K-concatenated products share the accumulator
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
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tmem_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..5 (both CTAs): own 128 rows, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const bool via_tma = epi.out_f32 == nullptr;      // bf16-only output (every GEMM of the recurrence): TMA store
    uint8_t* stage_out = sOut + q * 8192;
    int ob = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int m0 = (tile / tiles_n) * 256 + (int)rank * 128, n0 = (tile % tiles_n) * G2_BN;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G2_BN);
      if (via_tma) {
#pragma unroll 1
        for (int c = 0; c < G2_BN / 64; ++c) {
          float v[64];
          tmem_ld32(t0 + (uint32_t)(c * 64), v);
          tmem_ld32(t0 + (uint32_t)(c * 64 + 32), v + 32);
          if (c == G2_BN / 64 - 1) {                  // accumulator fully read: hand the TMEM stage back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tmem_empty + acc);
          }
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] *= epi.scale;
          for (int pl = 0; pl < out_planes; ++pl) {    // plane 0 = bf16(v), plane 1 = bf16(v - plane 0)
            if (lane == 0) tma_store_wait_read<1>();   // the store that last used this staging buffer has read it
            __syncwarp();
            uint8_t* dst = stage_out + ob * 4096 + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) = u;       // 128 B swizzle
              if (pl + 1 < out_planes) {
                bf16x2_residual(u.x, v[8 * j + 0], v[8 * j + 1]); bf16x2_residual(u.y, v[8 * j + 2], v[8 * j + 3]);
                bf16x2_residual(u.z, v[8 * j + 4], v[8 * j + 5]); bf16x2_residual(u.w, v[8 * j + 6], v[8 * j + 7]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stage_out + ob * 4096, pl * N + n0 + c * 64, m0 + q * 32);   // rows >= M are clipped by the TMA unit
              tma_store_commit();
            }
            ob ^= 1;
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < G2_BN / 32; ++c) {
          float v[32];
          tmem_ld32(t0 + (uint32_t)(c * 32), v);
          if (row < M) epi(row, n0 + c * 32, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tmem_empty + acc);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc2(tmem_base, 512); }
}

template <class Epi>
void launch_shift_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Epi& epi, int M, int N, int num_sms,
                        const ShiftSegs& segs, int out_planes, cudaStream_t st) {
  auto kern = shift_gemm2_kernel<Epi>;
  static DeviceOnce configured;
  if (configured.first()) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
  const int tiles = ((M + 255) / 256) * (N / G2_BN);
  int pairs = num_sms / 2;
  if (tiles < pairs) pairs = tiles;
  kern<<<2 * pairs, NUM_THREADS, G2_SMEM, st>>>(tmA, tmB, tmC, epi, M, N, segs, out_planes);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace tc
}  // namespace gcrnn
