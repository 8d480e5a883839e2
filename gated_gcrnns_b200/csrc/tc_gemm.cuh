// tcgen05 shift GEMM for sm_100a:  C[M, N] = A[M, K] * Bop[N, K]^T   (bf16 operands, fp32 accumulate in TMEM)
//
// This is the dense graph shift of the reference, `x = torch.matmul(x, S)` (Utils/graphML.py:123), batched over
// every (sample, feature) row: A = signals [(b,f), n] bf16 row-major (K-major), Bop = S^T (forward, z @ S) or S
// (backward, g @ S^T) stored row-major [n_out, n_in] bf16 (K-major), so both operands are plain K-major
// 128B-swizzled TMA tiles.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer   : 4-stage ring of {A 128x64, B BNx64} bf16 tiles (cp.async.bulk.tensor, SW128)
//   warp 1      MMA issuer     : tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 x4 per stage; 2 TMEM
//                                accumulator stages (2*BN <= 512 columns) so the epilogue of tile i overlaps
//                                the main loop of tile i+1
//   warps 2..5  epilogue       : tcgen05.ld 32x32b -> registers -> fused epilogue functor -> global stores
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace gcrnn {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 192;

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024 B)
//   [46,48) version = 1 (Blackwell) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = TILE_BYTES;
  static constexpr int TOTAL = TILE_BYTES + 256 + 1024;  // barriers + slack for the manual 1024 B alignment
};

// ---- split-bf16 operand planes ---------------------------------------------------------------------------------
// A bf16 "slab" holds P planes per row: row r = [plane 0: N values | plane 1: N values | ...], plane 0 = bf16(x),
// plane 1 = bf16(x - plane 0) (the rounding residual, exact in fp32), so plane 0 + plane 1 carries ~16 mantissa bits.
// P = 1 is the plain bf16 path.  Products of split operands are K-concatenated MMAs into one fp32 accumulator:
// x y ~= x0 y0 + x1 y0 + x0 y1 (the x1 y1 term is 2^-18 relative and dropped).
constexpr int MAX_PLANES = 2;
constexpr int MAX_SEGS = 3;
struct ShiftSegs {
  int n;                 // number of K-concatenated products (1..MAX_SEGS)
  int a[MAX_SEGS];       // signal plane of product s (column offset a * N in the slab)
  int b[MAX_SEGS];       // operator plane of product s (row offset b * N in the stacked operator)
};
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// residuals of a packed pair: (a - bf16(a), b - bf16(b)), exact in fp32
__device__ __forceinline__ void bf16x2_residual(uint32_t packed, float& a, float& b) {
  a -= __uint_as_float(packed << 16);
  b -= __uint_as_float(packed & 0xFFFF0000u);
}

// tanh(x) = 1 - 2 / (exp(2x) + 1): MUFU.EX2 + MUFU.RCP, absolute error ~2e-7 (saturates correctly at +-inf)
__device__ __forceinline__ float tanh_acc(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.f));
  return fmaf(-2.f, r, 1.f);
}
// bf16 planes of x at dst[0], dst[plane_stride], ...: plane 0 = bf16(x), plane 1 = bf16(x - plane 0)
__device__ __forceinline__ void store_planes(__nv_bfloat16* dst, long long plane_stride, int P, float x) {
  const __nv_bfloat16 hi = __float2bfloat16(x);
  dst[0] = hi;
  if (P > 1) dst[plane_stride] = __float2bfloat16(x - __bfloat162float(hi));
}

// ---- epilogue functors: called once per (row, 32-column chunk) with the fp32 accumulators ------------------
struct EpiStore {
  __nv_bfloat16* out_bf16;   // [M, planes * ld] or null
  float* out_f32;            // [M, ld] or null
  long long ld;
  float scale;               // the bf16 operator holds S / scale (exact for unweighted graphs); undone here in fp32
  int planes;                // bf16 planes written per row (1 or 2)
  __device__ __forceinline__ void operator()(int row, int col0, float* v) const {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= scale;
    if (out_f32) {
      float4* dst = reinterpret_cast<float4*>(out_f32 + (long long)row * ld + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    if (out_bf16) {
      for (int q = 0; q < planes; ++q) {
        uint4* dst = reinterpret_cast<uint4*>(out_bf16 + ((long long)row * planes + q) * ld + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          dst[i] = u;
          bf16x2_residual(u.x, v[8 * i + 0], v[8 * i + 1]); bf16x2_residual(u.y, v[8 * i + 2], v[8 * i + 3]);
          bf16x2_residual(u.z, v[8 * i + 4], v[8 * i + 5]); bf16x2_residual(u.w, v[8 * i + 6], v[8 * i + 7]);
        }
      }
    }
  }
};

template <int BN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
shift_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epi epi, int M, int N) {
  using L = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = N / BN;
  const int tiles_m = (M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_k = N / BK;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, L::STAGE_BYTES);
          tma_load_2d(sa, &tmA, full_bar + stage, kb * BK, m0);
          tma_load_2d(sa + L::A_BYTES, &tmB, full_bar + stage, kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(tmem_full + acc, acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(t0 + (uint32_t)(c * 32), v);
        if (row < M) epi(row, n0 + c * 32, v);
      }
      tc_fence_before();
      mbar_arrive(tmem_empty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();   // resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency)

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle
CUtensorMap make_tmap_bf16(const void* base, long long rows, long long cols, int box_rows);

template <int BN, class Epi>
void launch_shift_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const Epi& epi, int M, int N, int num_sms, cudaStream_t st) {
  using L = GemmSmem<BN>;
  auto kern = shift_gemm_kernel<BN, Epi>;
  static DeviceOnce configured;
  if (configured.first()) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(tmA, tmB, epi, M, N);
  count_launch();
  CUDA_OK(cudaGetLastError());
}

}  // namespace tc
}  // namespace gcrnn
