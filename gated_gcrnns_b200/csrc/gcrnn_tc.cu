// Dense tensor-core path (bf16 operands, fp32 accumulate) of the gated GCRNN recurrence for sm_100a.
// The graph shift z @ S (Utils/graphML.py:123) runs on tcgen05 (tc_gemm.cuh); see DESIGN.md §TC-path.
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "tc_cell.cuh"
#include "tc_tap.cuh"
#include "tc_gate.cuh"
#include "tc_node.cuh"
#include "tc_bwd.cuh"
#include "tc_hshift.cuh"
#include <cstdlib>
#include <cmath>
#include <algorithm>

namespace gcrnn {
namespace tc {

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    GCRNN_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

CUtensorMap make_tmap_bf16(const void* base, long long rows, long long cols, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCRNN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] at %p", (int)r, rows, cols, base);
  return tm;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, box_cols], NO swizzle (plain row-major shared-memory tile)
CUtensorMap make_tmap_bf16_plain(const void* base, long long rows, long long cols, int box_cols, int box_rows) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCRNN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (plain) failed (%d) for [%lld x %lld] at %p", (int)r, rows, cols, base);
  return tm;
}

int num_sms(int device) {
  static int cached[64] = {0};
  if (device < 64 && cached[device]) return cached[device];
  int n = 0;
  CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (device < 64) cached[device] = n;
  return n;
}

// out = A @ S (forward shift z @ S: Bop = S^T stored K-major) or A @ S^T (backward shift g @ S^T: Bop = S).
// A: bf16 [M][Pin * N] (Pin signal planes per row), out_bf16: [M][Pout * N], out_f32: [M][N].  With split operands the product
// is the K-concatenation  sum_a A_a S_0  (+ A_0 S_1 when the operator itself needs a residual plane: weighted graphs).
void shift_gemm(const gcrnn_graph* g, bool backward, const __nv_bfloat16* A, long long M, int Pin, __nv_bfloat16* out_bf16, int Pout,
                float* out_f32, cudaStream_t st) {
  const int N = g->N;
  GCRNN_CHECK(g->S_bf16 && g->St_bf16, "graph has no dense bf16 operator");
  GCRNN_CHECK(M > 0 && M < (1ll << 31), "row count out of range");
  GCRNN_CHECK(Pin >= 1 && Pin <= MAX_PLANES && Pout >= 1 && Pout <= MAX_PLANES, "shift_gemm: 1 or 2 operand planes");
  const __nv_bfloat16* Bop = backward ? g->S_bf16 : g->St_bf16;
  ShiftSegs segs{};
  for (int a = 0; a < Pin; ++a) { segs.a[segs.n] = a; segs.b[segs.n] = 0; ++segs.n; }
  if (Pin > 1 && g->s_planes > 1) { segs.a[segs.n] = 0; segs.b[segs.n] = 1; ++segs.n; }
  EpiStore epi{out_bf16, out_f32, (long long)N, g->dense_scale, Pout};
  const int sms = num_sms(g->device);
  const bool split = Pin > 1 || Pout > 1;
  GCRNN_CHECK(!split || N % 256 == 0, "split-bf16 operands need N %% 256 == 0 (N=%d)", N);
  if (N % 256 == 0 && (split || (opt().gemm_pair && M > 128))) {
    CUtensorMap tmA = make_tmap_bf16(A, M, (long long)Pin * N, 128), tmB = make_tmap_bf16(Bop, (long long)g->s_planes * N, N, 128);
    CUtensorMap tmC = out_bf16 ? make_tmap_bf16(out_bf16, M, (long long)Pout * N, 32) : tmA;      // bf16 output tiles leave through TMA stores
    launch_shift_gemm2<EpiStore>(tmA, tmB, tmC, epi, (int)M, N, sms, segs, Pout, st);
  } else if (N % 256 == 0) {
    CUtensorMap tmA = make_tmap_bf16(A, M, N, BM), tmB = make_tmap_bf16(Bop, (long long)g->s_planes * N, N, 256);
    launch_shift_gemm<256, EpiStore>(tmA, tmB, epi, (int)M, N, sms, st);
  } else {
    CUtensorMap tmA = make_tmap_bf16(A, M, N, BM), tmB = make_tmap_bf16(Bop, (long long)g->s_planes * N, N, 128);
    launch_shift_gemm<128, EpiStore>(tmA, tmB, epi, (int)M, N, sms, st);
  }
}


// ---- launch helpers -------------------------------------------------------------------------------------------
static void launched() { count_launch(); CUDA_OK(cudaGetLastError()); }
// fp32 [rows][N] -> bf16 [rows][P*N]
static void cvt_bf16(const float* in, __nv_bfloat16* out, long long rows, int N, int P, cudaStream_t st) {
  GCRNN_CHECK(N % 4 == 0, "cvt_bf16: row length must be a multiple of 4");
  const long long n4 = rows * (N / 4);
  cvt_bf16_kernel<<<(unsigned)std::min<long long>((n4 + 255) / 256, 148 * 16), 256, 0, st>>>(in, out, n4, N / 4, P);
  launched();
}
struct TcDims {
  int N, F, G, Kin, Kst, sms;
  int P;                 // operand planes: 1 = GCRNN_PREC_BF16_TC, 2 = GCRNN_PREC_BF16X2_TC (split hi + lo)
  long long LD;          // bf16 row length P * N
  long long B, T, R, RX, BT;
  bool tg, bias;
  bool node;             // node gates (tc_node.cuh)
};
static TcDims tc_dims(const gcrnn_cell* c, int64_t B, int64_t T) {
  TcDims d;
  d.N = c->g->N; d.F = c->d.F; d.G = c->d.G; d.Kin = c->d.Kin; d.Kst = c->d.Kst;
  d.B = B; d.T = T; d.R = B * d.F; d.RX = B * T * d.G; d.BT = B * T;
  d.tg = c->d.time_gating != 0; d.bias = c->d.bias != 0; d.node = c->d.spatial_gating == GCRNN_SPATIAL_NODE;
  d.P = c->d.precision == GCRNN_PREC_BF16X2_TC ? 2 : 1; d.LD = (long long)d.P * d.N;
  GCRNN_CHECK(d.P == 1 || d.N % 256 == 0, "split-bf16 tensor-core path: N %% 256 == 0 (N=%d)", d.N);
  GCRNN_CHECK(c->d.E == 1 && c->d.spatial_gating != GCRNN_SPATIAL_EDGE, "tensor-core path: E == 1, no edge gating");
  GCRNN_CHECK(!d.node || (d.Kin * d.G <= NG_KG && d.Kst <= NG_KMAX && d.F % NG_FC == 0 && d.N % 128 == 0),
              "tensor-core node gates: Kin*G <= %d, Kst <= %d", NG_KG, NG_KMAX);
  GCRNN_CHECK(d.F % 16 == 0 && d.F <= 64, "tensor-core path: F must be a multiple of 16 and <= 64 (F=%d)", d.F);
  GCRNN_CHECK(d.N % 128 == 0, "tensor-core path: N %% 128 == 0 (N=%d)", d.N);
  GCRNN_CHECK(d.Kin * d.G <= 32, "tensor-core path: Kin*G <= 32 (got %d); use the fp32 path", d.Kin * d.G);
  GCRNN_CHECK(d.Kst >= 1 && d.Kst <= WG_MAXK, "tensor-core path: Kst <= %d", WG_MAXK);
  GCRNN_CHECK(B > 0 && T > 0, "empty batch or sequence");
  d.sms = 0;
  return d;
}

struct TcSaved {
  float* zx;            // [Kin-1][RX][N]   x_t S^k, k >= 1
  float* gt;            // [2][B][T]        time-gate values
  __nv_bfloat16* Hb;    // [T][R][P*N]      bf16 planes of every state (GEMM / wgrad operand)
  float* qn;            // [2][B][T][N]     node-gate values
  void layout(Arena& a, const TcDims& d) {
    zx = a.get<float>((size_t)(d.Kin - 1) * d.RX * d.N);
    gt = d.tg ? a.get<float>(2 * d.BT) : nullptr;
    Hb = a.get<__nv_bfloat16>((size_t)d.T * d.R * d.LD);
    qn = d.node ? a.get<float>((size_t)2 * d.BT * d.N) : nullptr;
  }
};

// z_k = z_{k-1} @ S (forward) or @ S^T (backward), k = 1..K-1, bf16 slabs [K-1][rows][P*N]
static void chain(const gcrnn_graph* g, bool backward, const __nv_bfloat16* z0, __nv_bfloat16* zc, int K, long long rows, int P, cudaStream_t st) {
  const __nv_bfloat16* prev = z0;
  for (int k = 1; k < K; ++k) {
    __nv_bfloat16* out = zc + (size_t)(k - 1) * rows * P * g->N;
    shift_gemm(g, backward, prev, rows, P, out, P, nullptr, st);
    prev = out;
  }
}

static ContractArgs contract_base(const TcDims& d, const __nv_bfloat16* z0, const __nv_bfloat16* zc) {
  ContractArgs a{};
  a.K = d.Kst; a.C = d.F; a.M = d.F; a.N = d.N; a.B = d.B; a.P = d.P;
  a.slab[0] = z0;
  for (int k = 1; k < d.Kst; ++k) a.slab[k] = zc + (size_t)(k - 1) * d.R * d.LD;
  return a;
}

// tap contraction on tcgen05 (tc_tap.cuh).  Wp: prepared bf16 weights [M][KB*64] (zero padded), ca: the same
// argument block the mma.sync variant takes (slab pointers + epilogue operands).
template <int EPI>
static void launch_tap(const ContractArgs& ca, const __nv_bfloat16* Wp, int sms, cudaStream_t st) {
  TapArgs t{};
  t.K = ca.K; t.C = ca.C; t.M = ca.M; t.N = ca.N; t.KB = (ca.K * ca.C + 63) / 64; t.B = ca.B; t.R = ca.B * ca.C;
  t.P = ca.P; t.stages = ca.P > 1 ? 6 : TAP_STAGES; t.exact = ca.P > 1;
  GCRNN_CHECK(ca.P >= 1 && ca.P <= MAX_PLANES, "tap_gemm: 1 or 2 operand planes");
  GCRNN_CHECK(t.KB <= TAP_MAX_KB && (ca.C == 16 || ca.C == 32 || ca.C == 64) && ca.M % 16 == 0 && ca.M <= 64 && ca.N % TAP_BM == 0,
              "tap_gemm: unsupported sizes K=%d C=%d M=%d N=%d", ca.K, ca.C, ca.M, ca.N);
  t.out_f32 = ca.out_f32; t.out_bstride = ca.out_bstride; t.out_bf16 = ca.out_bf16; t.bias = ca.bias; t.bias_scale = ca.bias_scale;
  t.gi = ca.gi; t.gf = ca.gf; t.gate_stride = ca.gate_stride; t.A = ca.A; t.Kin = ca.Kin; t.G = ca.G;
  t.x0 = ca.x0; t.x0_bstride = ca.x0_bstride; t.zx = ca.zx; t.zx_kstride = ca.zx_kstride; t.zx_bstride = ca.zx_bstride;
  t.hprev = ca.hprev; t.hprev_bstride = ca.hprev_bstride; t.dgf = ca.dgf; t.accumulate = ca.accumulate; t.scaled_chain = ca.scaled_chain;
  t.dHn = ca.dHn; t.dHn_bstride = ca.dHn_bstride; t.gfn = ca.gfn; t.red = ca.red;
  t.qi = ca.qi; t.qf = ca.qf; t.q_bstride = ca.q_bstride;
  for (int k = 2; k < ca.K; ++k)
    GCRNN_CHECK(ca.slab[k] == ca.slab[1] + (size_t)(k - 1) * t.R * ca.P * ca.N, "tap_gemm: slabs 1..K-1 must be contiguous");
  const long long LD = (long long)ca.P * ca.N;
  const CUtensorMap tm0 = make_tmap_bf16(ca.slab[0], t.R, LD, ca.C);
  const CUtensorMap tmc = ca.K > 1 ? make_tmap_bf16(ca.slab[1], (long long)(ca.K - 1) * t.R, LD, ca.C) : tm0;
  const CUtensorMap tmW = make_tmap_bf16(Wp, ca.M, (long long)ca.P * t.KB * 64, ca.M);
  const int smem = tap_smem_bytes(t.P, t.KB, t.stages);
  GCRNN_CHECK(smem <= 227 * 1024, "tap_gemm: shared memory budget exceeded (%d B)", smem);
  const long long tiles = ca.B * (ca.N / TAP_BM);
  const int grid = (int)std::min<long long>(tiles, sms);
  GCRNN_CHECK(EPI != TAP_BWDF || ca.Kin * ca.G <= 7, "fused backward epilogue needs Kin*G <= 7");
  GCRNN_CHECK((EPI != TAP_FWD && EPI != TAP_BWDF) || ca.x0_bstride == ca.zx_bstride, "x0 / zx rows must share one sample stride");
  const int KG = ca.Kin * ca.G;
#define TAP_LAUNCH(KGM_, EXACT_)                                                                          \
  do {                                                                                                    \
    auto kern = tap_gemm_kernel<EPI, KGM_, EXACT_>;                                                       \
    static DeviceOnce once;                                                                              \
    if (once.first()) CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    kern<<<grid, TAP_THREADS, smem, st>>>(tm0, tmc, tmW, t);                                              \
  } while (0)
  if (EPI == TAP_FWD && KG >= 1 && KG <= 8) {
    switch (KG) {
      case 1: TAP_LAUNCH(1, true); break;
      case 2: TAP_LAUNCH(2, true); break;
      case 3: TAP_LAUNCH(3, true); break;
      case 4: TAP_LAUNCH(4, true); break;
      case 5: TAP_LAUNCH(5, true); break;
      case 6: TAP_LAUNCH(6, true); break;
      case 7: TAP_LAUNCH(7, true); break;
      default: TAP_LAUNCH(8, true); break;
    }
  } else if (EPI == TAP_FWD) {
    TAP_LAUNCH(32, false);
  } else {
    TAP_LAUNCH(8, false);
  }
#undef TAP_LAUNCH
  launched();
}

// prepared bf16 weights [64][P * KB*64] (zero padded; plane q at column q * KB*64): mode 0 rows = output features,
// mode 1 rows = input features
static size_t weight_elems(int F, int K, int P) { return (size_t)64 * P * (((K * F + 63) / 64) * 64); }
static int prep_contract_weight(const float* W, __nv_bfloat16* out, int F, int K, int mode, int P, cudaStream_t st) {
  const int pstride = ((K * F + 63) / 64) * 64, ld = P * pstride;
  CUDA_OK(cudaMemsetAsync(out, 0, (size_t)64 * ld * sizeof(__nv_bfloat16), st));
  prep_weight_kernel<<<(F * K * F + 255) / 256, 256, 0, st>>>(W, out, F, K, F, ld, mode, P, pstride);
  launched();
  return ld;
}

// dB_k += sum_{b,n} V_k h^T on tcgen05 (F = 64): hb = bf16 h_{t-1} [R][N]
static void launch_wgrad_tc(const TcDims& d, const __nv_bfloat16* v0, const __nv_bfloat16* vc, const __nv_bfloat16* hb, float* part,
                            cudaStream_t st) {
  WgradTcArgs w{};
  w.K = d.Kst; w.N = d.N; w.B = d.B; w.R = d.R; w.part = part; w.P = d.P;
  const CUtensorMap tm0 = make_tmap_bf16(v0, d.R, d.LD, 64);
  const CUtensorMap tmc = d.Kst > 1 ? make_tmap_bf16(vc, (long long)(d.Kst - 1) * d.R, d.LD, 64) : tm0;
  const CUtensorMap tmH = make_tmap_bf16(hb, d.R, d.LD, 64);
  const int sm = WT_STAGES * wt_stage_bytes(d.Kst) + 256 + 1024;
  static DeviceOnce once;
  if (once.first()) CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  wgrad_tc_kernel<<<d.sms, NUM_THREADS, sm, st>>>(tm0, tmc, tmH, w);
  launched();
}

// fused reverse-time step (tc_bwd.cuh), F = 64
template <int KG>
static void bwd_fused_launch_kg(const CUtensorMap& tm0, const CUtensorMap& tmc, const CUtensorMap& tmH, const CUtensorMap& tmZ,
                                const CUtensorMap& tmW, const BwdFusedArgs& a, int grid, cudaStream_t st) {
  static DeviceOnce once;
  if (once.first()) {
    CUDA_OK(cudaFuncSetAttribute(bwd_fused_kernel<KG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(bwd_fused_kernel<KG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int smem = bf_smem_bytes(a.P, a.KB, a.stages);
  GCRNN_CHECK(smem <= 227 * 1024, "fused backward step: shared memory budget exceeded (%d B)", smem);
  if (a.qin || (a.last && a.dlin_i)) bwd_fused_kernel<KG, true><<<grid, BF_THREADS, smem, st>>>(tm0, tmc, tmH, tmZ, tmW, a);
  else bwd_fused_kernel<KG, false><<<grid, BF_THREADS, smem, st>>>(tm0, tmc, tmH, tmZ, tmW, a);
}
// pair-stage ring depth that fits 227 KB next to P weight planes of K blocks
static int bf_stages(int P, int K) {
  for (int st = BF_STAGES; st >= 2; --st) if (bf_smem_bytes(P, K, st) <= 227 * 1024) return st;
  return 0;
}
static void launch_bwd_fused(const TcDims& d, BwdFusedArgs a, const __nv_bfloat16* v0, const __nv_bfloat16* vc, const __nv_bfloat16* hb,
                             const __nv_bfloat16* Zs, const __nv_bfloat16* Wp, cudaStream_t st) {
  a.K = d.Kst; a.N = d.N; a.KB = d.Kst; a.B = d.B; a.R = d.R; a.G = d.G; a.P = d.P; a.stages = bf_stages(d.P, d.Kst);
  const CUtensorMap tm0 = make_tmap_bf16(v0, d.R, d.LD, 64);
  const CUtensorMap tmc = d.Kst > 1 ? make_tmap_bf16(vc, (long long)(d.Kst - 1) * d.R, d.LD, 64) : tm0;
  const CUtensorMap tmH = make_tmap_bf16(hb, d.R, d.LD, 64);
  const CUtensorMap tmZ = make_tmap_bf16(Zs, d.BT * BF_ZROWS, d.N, BF_ZROWS);
  const CUtensorMap tmW = make_tmap_bf16(Wp, 64, (long long)d.P * d.Kst * 64, 64);
  const long long tiles = d.B * (d.N / 128);
  const int grid = (int)std::min<long long>(tiles, d.sms);
  switch (d.Kin * d.G) {
    case 1: bwd_fused_launch_kg<1>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 2: bwd_fused_launch_kg<2>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 3: bwd_fused_launch_kg<3>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 4: bwd_fused_launch_kg<4>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 5: bwd_fused_launch_kg<5>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 6: bwd_fused_launch_kg<6>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 7: bwd_fused_launch_kg<7>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    case 8: bwd_fused_launch_kg<8>(tm0, tmc, tmH, tmZ, tmW, a, grid, st); break;
    default: GCRNN_CHECK(false, "fused backward step: Kin*G must be <= 8");
  }
  launched();
}

// one Horner stage (tc_hshift.cuh): out = Zin S + (I (x) W_k) hprev  [+ the state update when `fin` is set]
static void launch_hshift(const gcrnn_graph* g, const TcDims& d, const __nv_bfloat16* Zin, const __nv_bfloat16* hprev, const __nv_bfloat16* Wk,
                          __nv_bfloat16* out, const HShiftArgs* fin, cudaStream_t st) {
  HShiftArgs a{};
  if (fin) a = *fin;
  a.M = (int)d.R; a.N = d.N; a.P = d.P; a.scale = g->dense_scale; a.final_stage = fin != nullptr; a.wcol = 0; a.wpstride = 64; a.exact = d.P > 1;
  a.epi_warps = fin ? 16 : 4;
  a.segs = ShiftSegs{};
  for (int q = 0; q < d.P; ++q) { a.segs.a[a.segs.n] = q; a.segs.b[a.segs.n] = 0; ++a.segs.n; }
  if (d.P > 1 && g->s_planes > 1) { a.segs.a[a.segs.n] = 0; a.segs.b[a.segs.n] = 1; ++a.segs.n; }
  const CUtensorMap tmS = make_tmap_bf16(g->St_bf16, (long long)g->s_planes * d.N, d.N, 128);
  const CUtensorMap tmZ = make_tmap_bf16(Zin, d.R, d.LD, 128);
  const CUtensorMap tmH = make_tmap_bf16(hprev, d.R, d.LD, 64);
  const CUtensorMap tmW = make_tmap_bf16(Wk, 64, (long long)d.P * 64, 32);
  a.out = out;
  static DeviceOnce once;
  if (once.first()) CUDA_OK(cudaFuncSetAttribute(hshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HS_SMEM));
  const int tiles = (int)((d.R + 255) / 256) * (d.N / 256);
  int pairs = d.sms / 2;
  if (tiles < pairs) pairs = tiles;
  hshift_kernel<<<2 * pairs, HS_THREADS, HS_SMEM, st>>>(tmS, tmZ, tmH, tmW, a);
  launched();
}

template <int KG, bool EX>
static void gate_launch_kg_ex(bool bwd, const GateArgs& ga, int grid, size_t sm, cudaStream_t st) {
  if (!bwd) {
    if (ga.F % 32 == 0 && opt().gate_fq8) {           // 8 feature groups: 512 threads, half the tap registers per thread
      CUDA_OK(cudaFuncSetAttribute(time_gate_fwd_kernel<KG, 8, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_fwd_kernel<KG, 8, EX><<<grid, 512, sm, st>>>(ga);
    } else {
      CUDA_OK(cudaFuncSetAttribute(time_gate_fwd_kernel<KG, 4, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_fwd_kernel<KG, 4, EX><<<grid, 256, sm, st>>>(ga);
    }
  } else {
    if (ga.F % 32 == 0 && opt().gate_fq8 >= 2) {
      CUDA_OK(cudaFuncSetAttribute(time_gate_bwd_kernel<KG, 8, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_bwd_kernel<KG, 8, EX><<<grid, 512, sm, st>>>(ga);
    } else {
      CUDA_OK(cudaFuncSetAttribute(time_gate_bwd_kernel<KG, 4, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_bwd_kernel<KG, 4, EX><<<grid, 256, sm, st>>>(ga);
    }
  }
}
template <int KG>
static void gate_launch_kg(bool bwd, const GateArgs& ga, int grid, size_t sm, cudaStream_t st) {
  if (ga.exact) gate_launch_kg_ex<KG, true>(bwd, ga, grid, sm, st);
  else gate_launch_kg_ex<KG, false>(bwd, ga, grid, sm, st);
}

static void gate_launch(bool bwd, GateArgs ga, const TcDims& d, cudaStream_t st) {
  GCRNN_CHECK(d.F % TG_FQ == 0 && d.F / TG_FQ <= TG_FMAX && d.N % TG_NT == 0, "time gate kernel: unsupported F=%d N=%d", d.F, d.N);
  const int KG = d.Kin * d.G;
  ga.exact = d.P > 1;
  if (KG > 8) {                      // generic kernel: taps in shared memory
    const size_t sm = gate_generic_smem_bytes(d.T, KG, d.F);
    GCRNN_CHECK(sm <= 200 * 1024, "time gate kernel: T*Kin*G too large for shared memory staging (%zu B)", sm);
    ga.bchunk = (int)std::max<long long>(1, std::min<long long>(16, d.B / 32));
    dim3 grid(d.N / TG_NT, (unsigned)((d.B + ga.bchunk - 1) / ga.bchunk));
    if (!bwd) {
      CUDA_OK(cudaFuncSetAttribute(time_gate_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_generic_kernel<false><<<grid, 256, sm, st>>>(ga);
    } else {
      CUDA_OK(cudaFuncSetAttribute(time_gate_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      time_gate_generic_kernel<true><<<grid, 256, sm, st>>>(ga);
    }
    launched();
    return;
  }
  // staging plan: whole sequences double-buffered if they fit, else single-buffered, else chunks of time steps
  const size_t budget = 200 * 1024;
  const int T = (int)d.T;
  if (gate_smem_bytes(T, KG, d.F, 2) <= budget) { ga.TC = T; ga.nbuf = 2; }
  else if (gate_smem_bytes(T, KG, d.F, 1) <= budget) { ga.TC = T; ga.nbuf = 1; }
  else {
    ga.nbuf = 2; ga.TC = T;
    while (ga.TC > 1 && gate_smem_bytes(ga.TC, KG, d.F, 2) > budget) ga.TC = (ga.TC + 1) / 2;
  }
  ga.nchunks = (T + ga.TC - 1) / ga.TC;
  const size_t sm = gate_smem_bytes(ga.TC, KG, d.F, ga.nbuf);
  GCRNN_CHECK(sm <= budget, "time gate kernel: Kin*G*F too large for shared memory staging (%zu B)", sm);
  if (bwd && ga.nchunks > 1) CUDA_OK(cudaMemsetAsync(ga.dc0, 0, (size_t)d.R * d.N * sizeof(float), st));
  const long long items = d.B * ga.nchunks * (d.N / TG_NT);
  const int grid = (int)std::min<long long>(items, d.sms);
  switch (KG) {
    case 1: gate_launch_kg<1>(bwd, ga, grid, sm, st); break;
    case 2: gate_launch_kg<2>(bwd, ga, grid, sm, st); break;
    case 3: gate_launch_kg<3>(bwd, ga, grid, sm, st); break;
    case 4: gate_launch_kg<4>(bwd, ga, grid, sm, st); break;
    case 5: gate_launch_kg<5>(bwd, ga, grid, sm, st); break;
    case 6: gate_launch_kg<6>(bwd, ga, grid, sm, st); break;
    case 7: gate_launch_kg<7>(bwd, ga, grid, sm, st); break;
    default: gate_launch_kg<8>(bwd, ga, grid, sm, st); break;
  }
  launched();
}

}  // namespace tc

using namespace tc;

size_t cell_forward_tc(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                       void* saved, size_t savedb, size_t* saved_used, void* ws, size_t wsb, int64_t B, int64_t T,
                       cudaStream_t st) {
  TcDims d = tc_dims(cell, B, T);
  const gcrnn_graph* g = cell->g;
  const int P = d.P;
  Arena a(ws, wsb);
  TcSaved s;
  { Arena sa(saved, savedb); s.layout(sa, d); if (saved_used) *saved_used = sa.off; }
  GCRNN_CHECK(a.dry() || saved, "forward needs the `saved` buffer");
  __nv_bfloat16* xb0 = a.get<__nv_bfloat16>((size_t)d.RX * d.LD);
  __nv_bfloat16* xb1 = a.get<__nv_bfloat16>((size_t)d.RX * d.LD);
  __nv_bfloat16* hb0 = a.get<__nv_bfloat16>((size_t)d.R * d.LD);
  __nv_bfloat16* zb = a.get<__nv_bfloat16>((size_t)(d.Kst - 1) * d.R * d.LD);
  __nv_bfloat16* Wb = a.get<__nv_bfloat16>(weight_elems(d.F, d.Kst, P));
  float* c0 = (d.tg || d.node) ? a.get<float>((size_t)d.R * d.N) : nullptr;
  float* logit = d.tg ? a.get<float>(2 * d.BT) : nullptr;
  // node gates: head signals p_k [Kst][B*T][N], Horner accumulator, shift output, bf16 planes of the shift input
  float* pbuf = d.node ? a.get<float>((size_t)d.Kst * d.BT * d.N) : nullptr;
  float* rcur = d.node ? a.get<float>((size_t)d.BT * d.N) : nullptr;
  float* rtmp = d.node ? a.get<float>((size_t)d.BT * d.N) : nullptr;
  __nv_bfloat16* rb = d.node ? a.get<__nv_bfloat16>((size_t)d.BT * d.LD) : nullptr;
  // Horner-form forward (tc_hshift.cuh): the state filter's tap contraction rides in the shift GEMMs
  const bool hfused = opt().fwd_fused && d.F == 64 && d.N % 256 == 0 && d.Kin * d.G <= 8 && d.Kst >= 2 && !d.node;
  __nv_bfloat16* wping = hfused ? a.get<__nv_bfloat16>((size_t)d.R * d.LD) : nullptr;
  __nv_bfloat16* wpong = hfused ? a.get<__nv_bfloat16>((size_t)d.R * d.LD) : nullptr;
  __nv_bfloat16* Wtaps = hfused ? a.get<__nv_bfloat16>((size_t)d.Kst * 64 * P * 64) : nullptr;     // [K][64][P*64]
  if (a.dry()) return a.off;
  d.sms = num_sms(g->device);
  const long long FN = (long long)d.F * d.N, GN = (long long)d.G * d.N;

  // ---- x_t S^k for every (b, t): one batched chain (rows = B*T*G) ------------------------------------------------
  if (d.Kin > 1) {
    cvt_bf16(X, xb0, d.RX, d.N, P, st);
    __nv_bfloat16* cur = xb0; __nv_bfloat16* nxt = xb1;
    for (int k = 1; k < d.Kin; ++k) {
      shift_gemm(g, false, cur, d.RX, P, (k < d.Kin - 1) ? nxt : nullptr, P, s.zx + (size_t)(k - 1) * d.RX * d.N, st);
      std::swap(cur, nxt);
    }
  }
  cvt_bf16(h0, hb0, d.R, d.N, P, st);
  // ---- time gates (graphML.py:2357-2374): depend on (x_t, h0) only -> all (b, t) at once ---------------------------
  if (d.tg || d.node) chain(g, false, hb0, zb, d.Kst, d.R, P, st);      // h0 S^k: the T-invariant term of every gate sub-cell
  if (d.tg) {
    CUDA_OK(cudaMemsetAsync(logit, 0, 2 * d.BT * sizeof(float), st));
    for (int gi = 0; gi < 2; ++gi) {
      prep_contract_weight(p->t_weight_B[gi], Wb, d.F, d.Kst, 0, P, st);
      ContractArgs ca = contract_base(d, hb0, zb);
      ca.out_f32 = c0; ca.out_bstride = FN; ca.bias = p->t_bias[gi]; ca.bias_scale = 2.f;   // bias enters twice (:2421-2422)
      launch_tap<TAP_PLAIN>(ca, Wb, d.sms, st);
      GateArgs ga{};
      ga.A = p->t_weight_A[gi]; ga.Kin = d.Kin; ga.G = d.G; ga.F = d.F; ga.N = d.N; ga.B = d.B; ga.T = d.T;
      ga.X = X; ga.zx = s.zx; ga.c0 = c0; ga.Wg = p->t_mlp_w[gi]; ga.logit = logit + gi * d.BT;
      gate_launch(false, ga, d, st);
      gate_sigmoid_kernel<<<(unsigned)((d.BT + 255) / 256), 256, 0, st>>>(logit + gi * d.BT, p->t_mlp_b[gi], s.gt + gi * d.BT, d.BT);
      launched();
    }
  }
  // ---- node gates (graphML.py:2379-2407): also functions of (x_t, h0) only -> all (b, t) at once (tc_node.cuh) -------------
  if (d.node) {
    const long long n4 = d.BT * (d.N / 4);
    const unsigned eg = (unsigned)std::min<long long>((n4 + 255) / 256, 148 * 16);
    for (int gi = 0; gi < 2; ++gi) {
      prep_contract_weight(p->n_weight_B[gi], Wb, d.F, d.Kst, 0, P, st);
      ContractArgs ca = contract_base(d, hb0, zb);
      ca.out_f32 = c0; ca.out_bstride = FN; ca.bias = p->n_bias[gi]; ca.bias_scale = 2.f;
      launch_tap<TAP_PLAIN>(ca, Wb, d.sms, st);
      NodeGateArgs na{};
      na.A = p->n_weight_A[gi]; na.wh = p->n_head_w[gi]; na.X = X; na.zx = s.zx; na.zx_kstride = d.RX * d.N; na.c0 = c0;
      na.Kin = d.Kin; na.G = d.G; na.F = d.F; na.N = d.N; na.Kst = d.Kst; na.exact = P > 1; na.B = d.B; na.T = d.T; na.p = pbuf;
      node_gate_fwd_kernel<64><<<(unsigned)std::min<long long>(d.B * (d.N / 128), 148 * 16), 128, node_gate_smem_bytes(d.F), st>>>(na);
      launched();
      float* q = s.qn + (size_t)gi * d.BT * d.N;
      const float* r = pbuf + (size_t)(d.Kst - 1) * d.BT * d.N;            // Horner: r = p_{K-1}; r = r S + p_k
      for (int k = d.Kst - 2; k >= 0; --k) {
        cvt_bf16(r, rb, d.BT, d.N, P, st);
        shift_gemm(g, false, rb, d.BT, P, nullptr, P, rtmp, st);
        node_head_add_kernel<<<eg, 256, 0, st>>>(reinterpret_cast<const float4*>(rtmp), reinterpret_cast<const float4*>(pbuf + (size_t)k * d.BT * d.N),
                                                reinterpret_cast<float4*>(k == 0 ? q : rcur), n4, p->n_head_b[gi], k == 0);
        launched();
        r = rcur;
      }
      if (d.Kst == 1) {
        node_head_add_kernel<<<eg, 256, 0, st>>>(nullptr, reinterpret_cast<const float4*>(pbuf), reinterpret_cast<float4*>(q), n4, p->n_head_b[gi], 1);
        launched();
      }
    }
  }
  // ---- the recurrence -------------------------------------------------------------------------------------------
  if (hfused) {
    // Horner: w_{K-1} = B_{K-1} h ; w_k = B_k h + w_{k+1} S ; h_t = tanh(gi (A(S)x_t + b) + gf (w_0 + b)).
    // Tap K-1 seeds the chain unscaled (tap kernel, MIX epilogue); taps 0..K-2 enter the shift GEMMs divided by the operator scale.
    for (int k = 0; k < d.Kst; ++k) {
      prep_tap_weight_kernel<<<(64 * 64 + 255) / 256, 256, 0, st>>>(p->weight_B, Wtaps + (size_t)k * 64 * P * 64, 64, d.Kst, k, P,
                                                                    k == d.Kst - 1 ? 1.f : 1.f / g->dense_scale);
      launched();
    }
    for (long long t = 0; t < d.T; ++t) {
      const __nv_bfloat16* hprev = t == 0 ? hb0 : s.Hb + (size_t)(t - 1) * d.R * d.LD;
      ContractArgs cm{};
      cm.K = 1; cm.C = d.F; cm.M = d.F; cm.N = d.N; cm.B = d.B; cm.P = P; cm.slab[0] = hprev; cm.out_bf16 = wping;
      launch_tap<TAP_MIX>(cm, Wtaps + (size_t)(d.Kst - 1) * 64 * P * 64, d.sms, st);
      __nv_bfloat16* cur = wping; __nv_bfloat16* nxt = wpong;
      for (int k = d.Kst - 2; k >= 1; --k) {
        launch_hshift(g, d, cur, hprev, Wtaps + (size_t)k * 64 * P * 64, nxt, nullptr, st);
        std::swap(cur, nxt);
      }
      HShiftArgs fa{};
      fa.H = H + t * FN; fa.H_bstride = d.T * FN; fa.bias = p->bias;
      fa.gi = d.tg ? s.gt + t : nullptr; fa.gf = d.tg ? s.gt + d.BT + t : nullptr; fa.gate_stride = d.T;
      fa.A = p->weight_A; fa.KG = d.Kin * d.G; fa.G = d.G;
      fa.x0 = X + t * GN; fa.zx = s.zx + t * GN; fa.zx_kstride = d.RX * d.N; fa.z_bstride = d.T * GN;
      launch_hshift(g, d, cur, hprev, Wtaps, s.Hb + (size_t)t * d.R * d.LD, &fa, st);
    }
    return a.off;
  }
  prep_contract_weight(p->weight_B, Wb, d.F, d.Kst, 0, P, st);
  for (long long t = 0; t < d.T; ++t) {
    const __nv_bfloat16* hprev = t == 0 ? hb0 : s.Hb + (size_t)(t - 1) * d.R * d.LD;
    if (!(t == 0 && (d.tg || d.node))) chain(g, false, hprev, zb, d.Kst, d.R, P, st);      // at t = 0 the gates' h0 chain is still in zb
    ContractArgs ca = contract_base(d, hprev, zb);
    ca.out_f32 = H + t * FN; ca.out_bstride = d.T * FN; ca.out_bf16 = s.Hb + (size_t)t * d.R * d.LD;
    ca.bias = p->bias;
    ca.gi = d.tg ? s.gt + t : nullptr; ca.gf = d.tg ? s.gt + d.BT + t : nullptr; ca.gate_stride = d.T;
    ca.A = p->weight_A; ca.Kin = d.Kin; ca.G = d.G;
    ca.x0 = X + t * GN; ca.x0_bstride = d.T * GN;
    ca.zx = s.zx + t * GN; ca.zx_kstride = d.RX * d.N; ca.zx_bstride = d.T * GN;
    if (d.node) { ca.qi = s.qn + t * d.N; ca.qf = s.qn + (size_t)d.BT * d.N + t * d.N; ca.q_bstride = d.T * d.N; }
    launch_tap<TAP_FWD>(ca, Wb, d.sms, st);
  }
  return a.off;
}

size_t cell_backward_tc(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0,
                        const float* H, const float* dH, const void* saved, size_t savedb,
                        const gcrnn_cell_params* gr, float* dX, float* dh0, void* ws, size_t wsb, int64_t B,
                        int64_t T, cudaStream_t st) {
  TcDims d = tc_dims(cell, B, T);
  const gcrnn_graph* g = cell->g;
  const int P = d.P;
  Arena a(ws, wsb);
  TcSaved s;
  { Arena sa(const_cast<void*>(saved), savedb); s.layout(sa, d); }
  GCRNN_CHECK(a.dry() || saved, "backward needs the buffer written by forward");
  const int max_sms = 256;
  __nv_bfloat16* vb0 = a.get<__nv_bfloat16>((size_t)d.R * d.LD);
  __nv_bfloat16* vb0b = a.get<__nv_bfloat16>((size_t)d.R * d.LD);
  float* red = a.get<float>((size_t)d.R * 8);
  __nv_bfloat16* vb = a.get<__nv_bfloat16>((size_t)(d.Kst - 1) * d.R * d.LD);
  float* dhrec = a.get<float>((size_t)d.R * d.N);
  const size_t wbuf = weight_elems(d.F, d.Kst, P);
  __nv_bfloat16* hb0 = a.get<__nv_bfloat16>((size_t)d.R * d.LD);
  __nv_bfloat16* WTb = a.get<__nv_bfloat16>(wbuf);
  float* part = a.get<float>((size_t)max_sms * d.Kst * d.F * d.F);
  // input gradients take the unfused reverse step (tap TAP_BWD + dpre_kernel + weight-gradient kernel)
  // (the workspace query passes a dX flag whenever ANY input gradient is wanted: buffers are sized for both variants)
  const bool fusable = opt().bwd_fused && d.F == 64 && d.Kin * d.G <= (P > 1 ? 7 : 8) && d.Kst <= 6 && d.N % 128 == 0 && bf_stages(P, d.Kst) >= 2;
  const bool fused = fusable && !dX;
  const int zs_split = P > 1;       // Zs tiles carry hi rows 0..7 and residual rows 8..15
  __nv_bfloat16* Zs = fusable ? a.get<__nv_bfloat16>((size_t)d.BT * BF_ZROWS * d.N) : nullptr;
  float* partA = fusable ? a.get<float>((size_t)max_sms * 64 * BF_ZROWS) : nullptr;
  float* zslab = cell->dh_last_only ? a.get<float>((size_t)d.F * d.N) : nullptr;
  float *dgt = nullptr, *c0 = nullptr, *dc0 = nullptr, *dl = nullptr;
  __nv_bfloat16* Wb = nullptr;
  if (d.tg || d.node) { c0 = a.get<float>((size_t)d.R * d.N); dc0 = a.get<float>((size_t)d.R * d.N); Wb = a.get<__nv_bfloat16>(wbuf); }
  if (d.tg) { dgt = a.get<float>(2 * d.BT); dl = a.get<float>(d.BT); }
  // node gates: d lin [2][B*T][N] (accumulated by dpre_kernel), adjoint head signals v_k [Kst][B*T][N], bf16 planes of a shift input
  float* dlin = d.node ? a.get<float>((size_t)2 * d.BT * d.N) : nullptr;
  float* vhead = d.node ? a.get<float>((size_t)d.Kst * d.BT * d.N) : nullptr;
  __nv_bfloat16* rb = d.node ? a.get<__nv_bfloat16>((size_t)d.BT * d.LD) : nullptr;
  // input gradient (dX != null; workspace sized for it whenever the query says so): per-tap contributions dxk [Kin][RX][N], then
  // dX = dxk_0 + (dxk_1 + (... dxk_{K-1} S^T ...) S^T) S^T by Horner with shift GEMMs on the RX = B*T*G rows
  const bool want_dx = dX != nullptr || (a.dry() && cell->need_dx);
  float* dxk = want_dx ? a.get<float>((size_t)d.Kin * d.RX * d.N) : nullptr;
  float* dxtmp = want_dx && d.Kin > 1 ? a.get<float>((size_t)d.RX * d.N) : nullptr;
  __nv_bfloat16* dxb = want_dx && d.Kin > 1 ? a.get<__nv_bfloat16>((size_t)d.RX * d.LD) : nullptr;
  if (a.dry()) return a.off;
  GCRNN_CHECK(dX == nullptr || (dxk != nullptr && d.Kin * d.G <= DP_KG), "tensor-core input gradients need Kin*G <= %d", DP_KG);
  if (dX) CUDA_OK(cudaMemsetAsync(dxk, 0, (size_t)d.Kin * d.RX * d.N * sizeof(float), st));
  d.sms = num_sms(g->device);
  GCRNN_CHECK(d.sms <= max_sms, "unexpected SM count %d", d.sms);
  const long long FN = (long long)d.F * d.N, GN = (long long)d.G * d.N;
  const size_t part_bytes = (size_t)d.sms * d.Kst * d.F * d.F * sizeof(float);
  if (zslab) CUDA_OK(cudaMemsetAsync(zslab, 0, (size_t)FN * sizeof(float), st));
  const DhView dv{dH, zslab, d.T, FN, cell->dh_last_only != 0};

  // dB_k += sum_{b,n} V_k h^T with the (already g_f-scaled) adjoint chain in vb0/vb
  auto wgrad_v = [&](const __nv_bfloat16* v0p, const float* h32, long long hstride, const __nv_bfloat16* h16) {
    if (d.F == 64) { launch_wgrad_tc(d, v0p, vb, h16, part, st); return; }
    WgradArgs w{};
    w.v0 = v0p; w.vc = vb; w.h = h32; w.h_bstride = hstride; w.scale = nullptr; w.scale_stride = 0;
    w.part = part; w.K = d.Kst; w.F = d.F; w.N = d.N; w.B = d.B; w.P = P;
    const size_t sm = ((size_t)2 * d.Kst * 64 * WG_LD + (size_t)2 * 64 * WG_LD) * sizeof(__nv_bfloat16);
    CUDA_OK(cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    wgrad_mma_kernel<<<d.sms, 256, sm, st>>>(w);
    launched();
  };
  auto wgrad_flush = [&](float* dW) {
    if (dW) { wgrad_reduce_kernel<<<(d.Kst * d.F * d.F + 255) / 256, 256, 0, st>>>(part, dW, d.sms, d.Kst, d.F); launched(); }
    CUDA_OK(cudaMemsetAsync(part, 0, part_bytes, st));
  };

  CUDA_OK(cudaMemsetAsync(part, 0, part_bytes, st));
  if (d.tg) CUDA_OK(cudaMemsetAsync(dgt, 0, 2 * d.BT * sizeof(float), st));
  if (d.node) CUDA_OK(cudaMemsetAsync(dlin, 0, (size_t)2 * d.BT * d.N * sizeof(float), st));
  prep_contract_weight(p->weight_B, WTb, d.F, d.Kst, 1, P, st);
  cvt_bf16(h0, hb0, d.R, d.N, P, st);

  // ---- reverse-time sweep -------------------------------------------------------------------------------------------
  CUDA_OK(cudaMemsetAsync(red, 0, (size_t)d.R * 8 * sizeof(float), st));
  const bool can_fuse = d.Kin * d.G <= 7 && !d.node && !dX;
  auto run_dpre = [&](long long t, const float* dhrec_in, __nv_bfloat16* v0_out) {
    DpreArgs da{};
    da.dH = dv.ptr(t); da.dH_bstride = dv.bstride(t); da.Ht = H + t * FN; da.H_bstride = d.T * FN;
    da.dhrec = dhrec_in; da.v0 = v0_out; da.P = P;
    da.gi = d.tg ? s.gt + t : nullptr; da.gf = d.tg ? s.gt + d.BT + t : nullptr; da.gate_stride = d.T;
    da.A = p->weight_A; da.bias = p->bias; da.Kin = d.Kin; da.G = d.G; da.F = d.F; da.N = d.N;
    da.x0 = X + t * GN; da.x0_bstride = d.T * GN; da.zx = s.zx + t * GN; da.zx_kstride = d.RX * d.N; da.zx_bstride = d.T * GN;
    da.dgi = d.tg ? dgt + t : nullptr; da.dgf = d.tg ? dgt + d.BT + t : nullptr;
    da.dA = fused ? nullptr : gr->weight_A; da.dbias = fused ? nullptr : gr->bias; da.B = d.B;   // fused: MMA3 covers every step
    if (d.node) {
      da.qi = s.qn + t * d.N; da.qf = s.qn + (size_t)d.BT * d.N + t * d.N; da.q_bstride = d.T * d.N;
      da.dlin_i = dlin + t * d.N; da.dlin_f = dlin + (size_t)d.BT * d.N + t * d.N;
    }
    if (dX) { da.dxk = dxk + t * GN; da.dxk_kstride = d.RX * d.N; da.dxk_bstride = d.T * GN; }
    const size_t dsm = ((d.node ? 2 * d.N : 0) + (dX ? (size_t)d.Kin * d.G * d.N : 0)) * sizeof(float);
    if (dsm > 48 * 1024) {
      static DeviceOnce once;
      if (once.first()) CUDA_OK(cudaFuncSetAttribute(dpre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      GCRNN_CHECK(dsm <= 100 * 1024, "dpre kernel: Kin*G*N too large for shared memory staging (%zu B)", dsm);
    }
    dpre_kernel<<<(unsigned)std::min<long long>(d.B * (d.F / DP_FC), 148 * 32), 256, dsm, st>>>(da);
    launched();
  };
  __nv_bfloat16* v0cur = vb0; __nv_bfloat16* v0nxt = vb0b;
  run_dpre(d.T - 1, nullptr, v0cur);
  if (fused) {
    CUDA_OK(cudaMemsetAsync(partA, 0, (size_t)d.sms * 64 * BF_ZROWS * sizeof(float), st));
    zs_build_kernel<<<148 * 8, 256, 0, st>>>(X, s.zx, d.RX * d.N, d.G, d.Kin * d.G, d.tg ? s.gt : nullptr, d.tg ? s.gt + d.BT : nullptr,
                                              Zs, d.BT, d.N, zs_split, d.node ? s.qn : nullptr, d.node ? s.qn + (size_t)d.BT * d.N : nullptr);
    launched();
  }
  for (long long t = d.T - 1; t >= 0; --t) {
    const float* hprev = t > 0 ? H + (t - 1) * FN : h0;
    const long long hstride = t > 0 ? d.T * FN : FN;
    const __nv_bfloat16* hprev16 = t > 0 ? s.Hb + (size_t)(t - 1) * d.R * d.LD : hb0;
    chain(g, true, v0cur, vb, d.Kst, d.R, P, st);
    if (fused) {
      BwdFusedArgs fa{};
      fa.last = t == 0; fa.dh0 = dh0 ? dhrec : nullptr;
      fa.gf = d.tg ? s.gt + d.BT + t : nullptr; fa.gate_stride = d.T; fa.dgf = d.tg ? dgt + d.BT + t : nullptr;
      fa.hprev = hprev; fa.hprev_bstride = hstride;
      if (t > 0) {
        fa.dHn = dv.ptr(t - 1); fa.dHn_bstride = dv.bstride(t - 1);
        fa.gfn = d.tg ? s.gt + d.BT + (t - 1) : nullptr;
        fa.dgin = d.tg ? dgt + (t - 1) : nullptr; fa.dgfn = d.tg ? dgt + d.BT + (t - 1) : nullptr;
        fa.A = p->weight_A; fa.x0 = X + (t - 1) * GN; fa.zx = s.zx + (t - 1) * GN; fa.zx_kstride = d.RX * d.N; fa.z_bstride = d.T * GN;
        fa.v0_out = v0nxt;
        if (d.node) {
          fa.qin = s.qn + (t - 1) * d.N; fa.qfn = s.qn + (size_t)d.BT * d.N + (t - 1) * d.N; fa.q_bstride = d.T * d.N;
          fa.gin = d.tg ? s.gt + (t - 1) : nullptr;
          fa.dlin_i = dlin + (t - 1) * d.N; fa.dlin_f = dlin + (size_t)d.BT * d.N + (t - 1) * d.N;
        }
      }
      fa.bias = p->bias;
      fa.zs_row0 = t * BF_ZROWS; fa.zs_rowb = d.T * BF_ZROWS;
      fa.part = part; fa.partA = partA;
      launch_bwd_fused(d, fa, v0cur, vb, hprev16, Zs, WTb, st);
      std::swap(v0cur, v0nxt);
      continue;
    }
    ContractArgs ca = contract_base(d, v0cur, vb);
    ca.gf = d.tg ? s.gt + d.BT + t : nullptr; ca.gate_stride = d.T;
    ca.hprev = hprev; ca.hprev_bstride = hstride; ca.dgf = d.tg ? dgt + d.BT + t : nullptr;
    ca.scaled_chain = 1;                       // the chain input is g_f * dpre: acc = g_f q = dh_{t-1} directly
    if (t > 0 && can_fuse) {
      // fused: the epilogue forms step t-1's dpre, its bf16 chain input and the per-(b, f) sums it needs
      ca.dHn = dv.ptr(t - 1); ca.dHn_bstride = dv.bstride(t - 1);
      ca.gfn = d.tg ? s.gt + d.BT + (t - 1) : nullptr;
      ca.Kin = d.Kin; ca.G = d.G;
      ca.x0 = X + (t - 1) * GN; ca.x0_bstride = d.T * GN;
      ca.zx = s.zx + (t - 1) * GN; ca.zx_kstride = d.RX * d.N; ca.zx_bstride = d.T * GN;
      ca.out_bf16 = v0nxt; ca.red = red;
      launch_tap<TAP_BWDF>(ca, WTb, d.sms, st);
      dpre_finish_kernel<<<(unsigned)((d.B + 7) / 8), 64, 0, st>>>(red, d.tg ? s.gt + (t - 1) : nullptr, d.tg ? s.gt + d.BT + (t - 1) : nullptr,
                                                                   d.T, p->weight_A, p->bias, gr->weight_A, gr->bias,
                                                                   d.tg ? dgt + (t - 1) : nullptr, d.tg ? dgt + d.BT + (t - 1) : nullptr,
                                                                   d.B, d.F, d.Kin * d.G, 8);
      launched();
    } else {
      ca.out_f32 = dhrec; ca.out_bstride = FN; ca.accumulate = 0;
      launch_tap<TAP_BWD>(ca, WTb, d.sms, st);
      if (t > 0) run_dpre(t - 1, dhrec, v0nxt);
    }
    if (gr->weight_B) wgrad_v(v0cur, hprev, hstride, hprev16);
    std::swap(v0cur, v0nxt);
  }
  wgrad_flush(gr->weight_B);
  if (fused) {
    dax_reduce_kernel<<<(64 * BF_ZROWS + 255) / 256, 256, 0, st>>>(partA, gr->weight_A, gr->bias, d.sms, d.Kin * d.G, zs_split);
    launched();
  }

  // T-invariant term of a gate sub-cell, forward again: c0 = B_s(S) h0 + 2 b_s   (the v slabs are free here: they hold h0's chain for a moment)
  auto subcell_c0 = [&](const float* wB, const float* bias) {
    chain(g, false, hb0, vb, d.Kst, d.R, P, st);
    prep_contract_weight(wB, Wb, d.F, d.Kst, 0, P, st);
    ContractArgs cc = contract_base(d, hb0, vb);
    cc.out_f32 = c0; cc.out_bstride = FN; cc.bias = bias; cc.bias_scale = 2.f;
    launch_tap<TAP_PLAIN>(cc, Wb, d.sms, st);
  };
  // ... and its adjoint from dc0 = sum_t d pre_s: db_s += 2 sum_n dc0; v_k = dc0 (S^T)^k; dB_s,k = v_k h0^T; dh0 += sum_k B_s,k^T v_k
  auto subcell_h0_path = [&](const float* wB, float* gwB, float* gbias) {
    if (gbias) {
      rowsum_bfn_kernel<<<(unsigned)std::min<long long>(d.R, 148 * 8), 256, 0, st>>>(dc0, gbias, d.B, d.F, d.N, 2.f);
      launched();
    }
    cvt_bf16(dc0, vb0, d.R, d.N, P, st);
    chain(g, true, vb0, vb, d.Kst, d.R, P, st);
    if (gwB) { wgrad_v(vb0, h0, FN, hb0); wgrad_flush(gwB); }
    if (dh0) {
      prep_contract_weight(wB, Wb, d.F, d.Kst, 1, P, st);
      ContractArgs cb = contract_base(d, vb0, vb);
      cb.out_f32 = dhrec; cb.out_bstride = FN; cb.hprev = h0; cb.hprev_bstride = FN; cb.accumulate = 1;
      launch_tap<TAP_BWD>(cb, Wb, d.sms, st);
    }
  };
  // ---- time gates, batched over (b, t) -----------------------------------------------------------------------------------
  if (d.tg) {
    for (int gi = 0; gi < 2; ++gi) {
      subcell_c0(p->t_weight_B[gi], p->t_bias[gi]);
      gate_dlogit_kernel<<<1, 1024, 0, st>>>(dgt + gi * d.BT, s.gt + gi * d.BT, dl, gr->t_mlp_b[gi], d.BT);
      launched();
      GateArgs ga{};
      ga.A = p->t_weight_A[gi]; ga.Kin = d.Kin; ga.G = d.G; ga.F = d.F; ga.N = d.N; ga.B = d.B; ga.T = d.T;
      ga.X = X; ga.zx = s.zx; ga.c0 = c0; ga.Wg = p->t_mlp_w[gi]; ga.dl = dl;
      ga.dWg = gr->t_mlp_w[gi]; ga.dc0 = dc0; ga.dA = gr->t_weight_A[gi];
      GCRNN_CHECK(ga.dWg && ga.dA, "time-gate gradient buffers missing");
      gate_launch(true, ga, d, st);
      if (dX) {                                                              // the gate sub-cell's own input-filter path into dX
        GateDxArgs gx{};
        gx.g.A = p->t_weight_A[gi]; gx.g.X = X; gx.g.zx = s.zx; gx.g.zx_kstride = d.RX * d.N; gx.g.c0 = c0;
        gx.g.Kin = d.Kin; gx.g.G = d.G; gx.g.F = d.F; gx.g.N = d.N; gx.g.Kst = d.Kst; gx.g.exact = P > 1; gx.g.B = d.B; gx.g.T = d.T;
        gx.Wg = p->t_mlp_w[gi]; gx.dl = dl; gx.dxk = dxk; gx.dxk_kstride = d.RX * d.N;
        gate_dx_kernel<0><<<(unsigned)std::min<long long>(d.B * (d.N / 128), 148 * 16), 128, node_gate_smem_bytes(d.F), st>>>(gx);
        launched();
      }
      subcell_h0_path(p->t_weight_B[gi], gr->t_weight_B[gi], gr->t_bias[gi]);
    }
  }
  // ---- node gates, batched over (b, t): adjoint head chain on scalar node signals, then the sub-cell (tc_node.cuh) -----------
  if (d.node) {
    const long long nelem = d.BT * d.N;
    for (int gi = 0; gi < 2; ++gi) {
      subcell_c0(p->n_weight_B[gi], p->n_bias[gi]);
      const float* dl_g = dlin + (size_t)gi * nelem;
      if (gr->n_head_b[gi]) { sum_all_kernel<<<148 * 4, 256, 0, st>>>(dl_g, gr->n_head_b[gi], nelem); launched(); }
      CUDA_OK(cudaMemcpyAsync(vhead, dl_g, (size_t)nelem * sizeof(float), cudaMemcpyDeviceToDevice, st));
      for (int k = 1; k < d.Kst; ++k) {                                       // v_k = v_{k-1} S^T
        cvt_bf16(vhead + (size_t)(k - 1) * nelem, rb, d.BT, d.N, P, st);
        shift_gemm(g, true, rb, d.BT, P, nullptr, P, vhead + (size_t)k * nelem, st);
      }
      NodeGateArgs na{};
      na.A = p->n_weight_A[gi]; na.wh = p->n_head_w[gi]; na.X = X; na.zx = s.zx; na.zx_kstride = d.RX * d.N; na.c0 = c0;
      na.Kin = d.Kin; na.G = d.G; na.F = d.F; na.N = d.N; na.Kst = d.Kst; na.exact = P > 1; na.B = d.B; na.T = d.T;
      na.v = vhead; na.dA = gr->n_weight_A[gi]; na.dwh = gr->n_head_w[gi]; na.dc0 = dc0;
      GCRNN_CHECK(na.dA && na.dwh, "node-gate gradient buffers missing");
      {
        const unsigned ng = (unsigned)std::min<long long>(d.B * (d.N / 32), d.sms * (d.F >= 64 ? 1 : 2)), nt = 32 * (d.F / NG_FC);
        const size_t nsm = node_gate_smem_bytes(d.F);
        const int kgx = d.Kin * d.G;
        if (kgx == 5 && d.Kst == 5) node_gate_bwd_kernel<5, 5><<<ng, nt, nsm, st>>>(na);
        else if (kgx <= 4 && d.Kst <= 4) node_gate_bwd_kernel<4, 4><<<ng, nt, nsm, st>>>(na);
        else node_gate_bwd_kernel<NG_KG, NG_KMAX><<<ng, nt, nsm, st>>>(na);
      }
      launched();
      if (dX) {
        GateDxArgs gx{};
        gx.g = na; gx.dxk = dxk; gx.dxk_kstride = d.RX * d.N;
        gate_dx_kernel<1><<<(unsigned)std::min<long long>(d.B * (d.N / 128), 148 * 16), 128, node_gate_smem_bytes(d.F), st>>>(gx);
        launched();
      }
      subcell_h0_path(p->n_weight_B[gi], gr->n_weight_B[gi], gr->n_bias[gi]);
    }
  }
  if (dh0) CUDA_OK(cudaMemcpyAsync(dh0, dhrec, (size_t)d.R * d.N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (dX) {                                                           // Horner over the input taps with S^T
    const long long n4 = d.RX * (d.N / 4);
    const unsigned eg = (unsigned)std::min<long long>((n4 + 255) / 256, 148 * 16);
    const float* r = dxk + (size_t)(d.Kin - 1) * d.RX * d.N;
    for (int k = d.Kin - 2; k >= 0; --k) {
      cvt_bf16(r, dxb, d.RX, d.N, P, st);
      shift_gemm(g, true, dxb, d.RX, P, nullptr, P, dxtmp, st);
      float* out = k == 0 ? dX : dxk + (size_t)(d.Kin - 1) * d.RX * d.N;           // the last tap's slab doubles as the accumulator
      node_head_add_kernel<<<eg, 256, 0, st>>>(reinterpret_cast<const float4*>(dxtmp), reinterpret_cast<const float4*>(dxk + (size_t)k * d.RX * d.N),
                                              reinterpret_cast<float4*>(out), n4, nullptr, 0);
      launched();
      r = out;
    }
    if (d.Kin == 1) CUDA_OK(cudaMemcpyAsync(dX, dxk, (size_t)d.RX * d.N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return a.off;
}

// bf16 copies of S and S^T, stacked planes [s_planes * N][N]: plane 0 = bf16(S / max|S|), plane 1 = bf16 of the residual.
// An unweighted graph (cfg3: S = W / lambda_max with 0/1 W) is exact in plane 0 and keeps one plane; the split-bf16 path adds
// the product with plane 1 for weighted graphs.
void tc_prepare_graph(gcrnn_graph* g, const float* S) {
  const int N = g->N;
  GCRNN_CHECK(N % 128 == 0, "the tensor-core path needs N %% 128 == 0 (N=%d)", N);
  const size_t NN = (size_t)N * N;
  std::vector<__nv_bfloat16> s(2 * NN), st(2 * NN);
  float mx = 0.f;
  for (size_t i = 0; i < NN; ++i) mx = std::max(mx, std::fabs(S[i]));
  g->dense_scale = mx > 0.f ? mx : 1.f;
  bool residual = false;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      const float x = S[(size_t)i * N + j] / g->dense_scale;
      const __nv_bfloat16 hi = __float2bfloat16(x);
      const __nv_bfloat16 lo = __float2bfloat16(x - __bfloat162float(hi));
      residual = residual || __bfloat162float(lo) != 0.f;
      s[(size_t)i * N + j] = hi; st[(size_t)j * N + i] = hi;
      s[NN + (size_t)i * N + j] = lo; st[NN + (size_t)j * N + i] = lo;
    }
  g->s_planes = residual ? 2 : 1;
  const size_t bytes = (size_t)g->s_planes * NN * sizeof(__nv_bfloat16);
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat16* d = nullptr;
    CUDA_OK(cudaMalloc(&d, bytes));
    g->owned.push_back(d);
    CUDA_OK(cudaMemcpy(d, which ? st.data() : s.data(), bytes, cudaMemcpyHostToDevice));
    (which ? g->St_bf16 : g->S_bf16) = d;
  }
  g->Npad = N;
}

}  // namespace gcrnn

extern "C" int gcrnn_debug_shift_gemm(const gcrnn_graph* g, int32_t backward, const void* A_bf16, int64_t M, int32_t planes_in,
                                      void* out_bf16, int32_t planes_out, float* out_f32, void* stream) {
  try {
    if (!g || !A_bf16) throw gcrnn::Error(-2, "null argument");
    gcrnn::DeviceScope dev(g->device);
    gcrnn::OptScope os(&g->opt);
    gcrnn::tc::shift_gemm(g, backward != 0, (const __nv_bfloat16*)A_bf16, M, planes_in, (__nv_bfloat16*)out_bf16, planes_out, out_f32,
                          (cudaStream_t)stream);
  } catch (const std::exception& e) {
    gcrnn::set_last_error("%s", e.what());
    return -1;
  }
  return 0;
}
