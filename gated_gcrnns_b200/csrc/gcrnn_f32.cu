// Host orchestration of the fp32 sparse exact path: LSIGF, graph attention and the gated GCRNN recurrence
// with its hand-derived reverse-time backward.  Follows Utils/graphML.py:47-140 (LSIGF), :521-627 + :2084-2116
// (attention), :2336-2428 (GGCRNNCell.forward); see DESIGN.md for the restructuring (gates hoisted out of the
// recurrence because they depend on (x_t, h0) only; recompute-from-H backward).
#include "kernels_f32.cuh"
#include "sp32_kernels.cuh"
#include "sp32_tile.cuh"
#include "persist_f32.cuh"
#include <climits>

namespace gcrnn {
using namespace k;

namespace {

struct Ctx {
  const gcrnn_graph* g;
  cudaStream_t st;
  bool dry;
};

constexpr int TPB = 256;

void zero(const Ctx& c, void* p, size_t bytes) {
  if (c.dry || bytes == 0) return;
  CUDA_OK(cudaMemsetAsync(p, 0, bytes, c.st));
}
void copy(const Ctx& c, void* dst, const void* src, size_t bytes) {
  if (c.dry || bytes == 0) return;
  CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c.st));
}
void check_launch() { count_launch(); CUDA_OK(cudaGetLastError()); }

// reference layout [.., C, N] -> node-major [.., N, C] (in_dir) or back, through the graph's node renumbering (graph.cu).
// C == 32: rows are full 128-byte lines, gathered (out) or scattered (in) whole; the reference-layout side stays coalesced.
struct NodeMap { const int* perm; const int* iperm; explicit operator bool() const { return perm != nullptr; } };
NodeMap node_map(const gcrnn_graph* g, const gcrnn_graph* view) { return view != g ? NodeMap{g->perm, g->iperm} : NodeMap{nullptr, nullptr}; }
void permute_nodes(const Ctx& c, bool in_dir, const float* in, float* out, NodeMap m, int C, int N, long long R1, long long R2,
                   long long is1, long long is2, long long os1, long long os2) {
  if (c.dry) return;
  const bool c32 = C == 32 && N >= 1024 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) && is1 % 4 == 0 && is2 % 4 == 0 &&
                   os1 % 4 == 0 && os2 % 4 == 0;
  const long long work = R1 * R2 * ((N + 127) / 128);
  const unsigned grid = (unsigned)std::min<long long>(work, 148LL * 64);
  if (c32 && in_dir) scatter_rows_c32_k<<<grid, 256, 0, c.st>>>(in, out, N, R1, R2, is1, is2, os1, os2, m.iperm);
  else if (c32) transpose_c32_k<<<grid, 256, 0, c.st>>>(in, out, N, R1, R2, is1, is2, os1, os2, m.iperm);
  else if (in_dir) permute_nodes_k<true><<<grid1d(R1 * R2 * N * C, TPB), TPB, 0, c.st>>>(in, out, m.perm, C, N, R1, R2, is1, is2, os1, os2);
  else permute_nodes_k<false><<<grid1d(R1 * R2 * N * C, TPB), TPB, 0, c.st>>>(in, out, m.perm, C, N, R1, R2, is1, is2, os1, os2);
  check_launch();
}

void transpose(const Ctx& c, const float* in, const float* add, float* out, int A, int Bd, long long R1, long long R2,
               long long is1, long long is2, long long os1, long long os2) {
  if (c.dry) return;
  if (A == 1 || Bd == 1) {                          // nothing to transpose inside a slab: strided batch of contiguous copies
    const long long L = (long long)A * Bd;
    dim3 grid((unsigned)std::min<long long>((L + 1023) / 1024, 64), (unsigned)std::min<long long>(R1 * R2, 32768));
    strided_copy_k<<<grid, 256, 0, c.st>>>(in, add, out, L, R1, R2, is1, is2, os1, os2);
    check_launch();
    return;
  }
  if (Bd == 32 && A >= 1024 && add == nullptr && ((uintptr_t)in % 16 == 0) && is1 % 4 == 0 && is2 % 4 == 0) {
    const long long work = R1 * R2 * ((A + 127) / 128);
    transpose_c32_k<<<(unsigned)std::min<long long>(work, 148LL * 64), 256, 0, c.st>>>(in, out, A, R1, R2, is1, is2, os1, os2, nullptr);
    check_launch();
    return;
  }
  dim3 grid((Bd + 31) / 32, ((A + 31) / 32 + TRANSPOSE_TY - 1) / TRANSPOSE_TY, (unsigned)std::min<long long>(R1 * R2, 32768));
  transpose_k<<<grid, dim3(32, 8), 0, c.st>>>(in, add, out, A, Bd, R1, R2, is1, is2, os1, os2);
  check_launch();
}

void spmm(const Ctx& c, const Gather& op, const float* in, const float* add, float* out, int C, long long R) {
  if (c.dry) return;
  const int N = c.g->N;
  const bool v4 = (C % 4 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) && (!add || (uintptr_t)add % 16 == 0);
  if (v4) spmm_k<4><<<grid1d(R * N * (C / 4), TPB), TPB, 0, c.st>>>(op.ptr, op.idx, op.val, in, add, out, N, C, R);
  else    spmm_k<1><<<grid1d(R * N * C, TPB), TPB, 0, c.st>>>(op.ptr, op.idx, op.val, in, add, out, N, C, R);
  check_launch();
}

size_t smem_opt_in(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  GCRNN_CHECK(bytes <= 200 * 1024, "filter taps do not fit in shared memory (%zu bytes)", bytes);
  return bytes;
}

void contract_fwd(const Ctx& c, const Slabs& z, const float* W, const float* bias, float bias_scale, const float* addb,
                  long long RB, float* y, long long R, int F, int S, int G) {
  if (c.dry) return;
  const int N = c.g->N;
  const size_t sm = (size_t)S * G * F * sizeof(float);
  if (addb) {
    smem_opt_in((const void*)contract_fwd_k<1>, sm);
    contract_fwd_k<1><<<grid1d(R * N * F, TPB), TPB, sm, c.st>>>(z, W, bias, bias_scale, addb, RB, y, R, N, F, S, G);
  } else {
    smem_opt_in((const void*)contract_fwd_k<0>, sm);
    contract_fwd_k<0><<<grid1d(R * N * F, TPB), TPB, sm, c.st>>>(z, W, bias, bias_scale, nullptr, 1, y, R, N, F, S, G);
  }
  check_launch();
}

void contract_bwd_data(const Ctx& c, const SlabsMut& dz, const float* W, const float* dy, long long RN, int F, int S, int G,
                       int accumulate) {
  if (c.dry) return;
  const size_t sm = smem_opt_in((const void*)contract_bwd_data_k, (size_t)S * G * F * sizeof(float));
  contract_bwd_data_k<<<grid1d(RN * S * G, TPB), TPB, sm, c.st>>>(dz, W, dy, RN, F, S, G, accumulate);
  check_launch();
}

void contract_wgrad(const Ctx& c, const Slabs& z, const float* dy, float* dW, long long RN, int F, int S, int G) {
  if (c.dry || !dW) return;
  const size_t sm = smem_opt_in((const void*)contract_wgrad_k, (size_t)WG_ITEMS * (F + S * G) * sizeof(float));
  const int outs = F * S * G;
  dim3 grid((unsigned)std::min<long long>((RN + WG_ITEMS - 1) / WG_ITEMS, 148 * 4), (outs + TPB * WG_OPT - 1) / (TPB * WG_OPT));
  contract_wgrad_k<<<grid, TPB, sm, c.st>>>(z, dy, dW, RN, F, S, G);
  check_launch();
}

void colsum(const Ctx& c, const float* in, float* out, long long rows, int C, float scale = 1.f) {
  if (c.dry || !out) return;
  colsum_k<<<(unsigned)std::min<long long>((rows + 63) / 64, 148 * 8), 64, 0, c.st>>>(in, out, rows, C, scale);
  check_launch();
}

void add_inplace(const Ctx& c, float* dst, const float* src, long long n) {
  if (c.dry) return;
  add_inplace_k<<<grid1d(n, TPB), TPB, 0, c.st>>>(dst, src, n);
  check_launch();
}

// Slab s = e*K + k of a K-tap chain: (e, 0) is the base signal itself, (e, k>=1) lives in `chain`.
Slabs chain_slabs(const float* base, const float* chain, int E, int K, long long slab_elems, long long base_off = 0) {
  GCRNN_CHECK(E * K <= MAX_SLABS, "E*K = %d exceeds the supported %d filter slabs", E * K, MAX_SLABS);
  Slabs s{};
  for (int e = 0; e < E; ++e)
    for (int kk = 0; kk < K; ++kk)
      s.p[e * K + kk] = (kk == 0) ? base + base_off
                                  : (chain ? chain + (long long)(e * (K - 1) + (kk - 1)) * slab_elems + base_off : nullptr);
  return s;
}
SlabsMut mut_slabs(float* buf, int S, long long slab_elems, long long off = 0) {
  GCRNN_CHECK(S <= MAX_SLABS, "too many slabs (%d)", S);
  SlabsMut s{};
  for (int i = 0; i < S; ++i) s.p[i] = buf ? buf + (long long)i * slab_elems + off : nullptr;
  return s;
}

// z_{e,k} = z_{e,k-1} @ S_e, k = 1..K-1, over R samples of C channels (node-major)
void shift_chain(const Ctx& c, const float* base, float* chain, int E, int K, int C, long long R) {
  const long long slab = R * c.g->N * C;
  for (int e = 0; e < E; ++e) {
    const float* prev = base;
    for (int kk = 1; kk < K; ++kk) {
      float* out = chain ? chain + (long long)(e * (K - 1) + (kk - 1)) * slab : nullptr;
      spmm(c, c.g->fwd[e], prev, nullptr, out, C, R);
      prev = out;
    }
  }
}

// Given dz_{e,k} (E*K distinct buffers): g_{e,k-1} = dz_{e,k-1} + g_{e,k} @ S_e^T (in place), dst (+)= sum_e g_{e,0}
void reverse_chain(const Ctx& c, float* dz, int E, int K, int C, long long R, float* dst, bool accumulate) {
  const long long slab = R * c.g->N * C;
  for (int e = 0; e < E; ++e)
    for (int kk = K - 1; kk >= 1; --kk) {
      float* hi = dz ? dz + (long long)(e * K + kk) * slab : nullptr;
      float* lo = dz ? dz + (long long)(e * K + kk - 1) * slab : nullptr;
      spmm(c, c.g->bwd[e], hi, lo, lo, C, R);
    }
  for (int e = 0; e < E; ++e) {
    float* g0 = dz ? dz + (long long)(e * K) * slab : nullptr;
    if (e == 0 && !accumulate) copy(c, dst, g0, slab * sizeof(float));
    else add_inplace(c, dst, g0, slab);
  }
}

// ---------------------------------------------------------------------------------------------------
// attention on node-major signals
// ---------------------------------------------------------------------------------------------------
struct GatBufs {
  float* wu = nullptr;     // [R][N][F]
  float2* rc = nullptr;    // [R][N]
  float* alpha = nullptr;  // [R][nnz_att]
  void alloc(Arena& a, long long R, int N, int F, long long nnz) {
    wu = a.get<float>(R * N * F); rc = a.get<float2>(R * N); alpha = a.get<float>(R * nnz);
  }
};

void gat_fwd_nm(const Ctx& c, const float* mixer, const float* weight, const float* un, float* yn, const GatBufs& b,
                int F, int G, long long R) {
  const gcrnn_graph* g = c.g;
  const int N = g->N;
  Slabs z{}; z.p[0] = un;
  contract_fwd(c, z, weight, nullptr, 0.f, nullptr, 1, b.wu, R, F, 1, G);              // Wu  (graphML.py:586-588)
  if (c.dry) return;
  gat_scores_k<<<grid1d(R * N, TPB), TPB, 0, c.st>>>(b.wu, mixer, b.rc, R * N, F);      // :591-594
  check_launch();
  gat_softmax_k<<<grid1d(R * N, 128), 128, 0, c.st>>>(g->att_rptr, g->att_col, b.rc, b.alpha, R, N, g->nnz_att);  // :597-622
  check_launch();
  gat_aggregate_k<<<grid1d(R * N * F, TPB), TPB, 0, c.st>>>(g->att_cptr, g->att_crow, g->att_ceid, g->att_val, b.alpha,
                                                            b.wu, yn, R, N, F, g->nnz_att);  // :625 + relu :2101
  check_launch();
}

struct GatBwdBufs {
  float *dyr = nullptr, *ds = nullptr, *dc = nullptr, *dwu = nullptr; float2* drc = nullptr;
  void alloc(Arena& a, long long R, int N, int F, long long nnz) {
    dyr = a.get<float>(R * N * F); ds = a.get<float>(R * nnz); dc = a.get<float>(R * N);
    dwu = a.get<float>(R * N * F); drc = a.get<float2>(R * N);
  }
};

// dyn: gradient w.r.t. the relu output yn.  Writes dun (may be null), accumulates dmixer / dweight (may be null).
void gat_bwd_nm(const Ctx& c, const float* mixer, const float* weight, const float* un, const float* yn, const float* dyn,
                const GatBufs& b, const GatBwdBufs& w, float* dun, float* dmixer, float* dweight, int F, int G, long long R) {
  const gcrnn_graph* g = c.g;
  const int N = g->N;
  if (!c.dry) {
    relu_mask_k<<<grid1d(R * N * F, TPB), TPB, 0, c.st>>>(dyn, yn, w.dyr, R * N * F);
    check_launch();
    gat_bwd_rows_k<<<grid1d(R * N, 128), 128, 0, c.st>>>(g->att_rptr, g->att_col, g->att_val, b.rc, b.alpha, b.wu, w.dyr,
                                                         w.ds, w.dc, R, N, F, g->nnz_att);
    check_launch();
    gat_bwd_dwu_k<<<grid1d(R * N * F, TPB), TPB, 0, c.st>>>(g->att_rptr, g->att_col, g->att_val, g->att_cptr, g->att_ceid,
                                                            b.alpha, w.ds, w.dc, w.dyr, mixer, w.dwu, w.drc, R, N, F, g->nnz_att);
    check_launch();
    if (dmixer) {
      gat_bwd_mixer_k<<<(unsigned)std::min<long long>((R * N + 63) / 64, 148 * 8), 64, 0, c.st>>>(w.drc, b.wu, dmixer, R * N, F);
      check_launch();
    }
  }
  Slabs z{}; z.p[0] = un;
  contract_wgrad(c, z, w.dwu, dweight, R * N, F, 1, G);
  if (dun) { SlabsMut dz{}; dz.p[0] = dun; contract_bwd_data(c, dz, weight, w.dwu, R * N, F, 1, G, 0); }
}

}  // namespace

// ===================================================================================================
// LSIGF, reference layout  x:[B,G,N] -> y:[B,F,N]
// ===================================================================================================
size_t lsigf_forward_f32(const gcrnn_graph* g, const float* h, const float* bias, const float* x, float* y,
                         int F, int K, int G, int64_t B, void* ws, size_t wsb, cudaStream_t st) {
  Arena a(ws, wsb);
  Ctx c{g, st, a.dry()};
  const int N = g->N, E = g->E;
  const long long slab = (long long)B * N * G;
  float* xn = (G == 1) ? const_cast<float*>(x) : a.get<float>(slab);
  float* chain = a.get<float>((size_t)E * (K - 1) * slab);
  float* yn = (F == 1) ? y : a.get<float>((size_t)B * N * F);
  if (G != 1) transpose(c, x, nullptr, xn, G, N, B, 1, (long long)G * N, 0, (long long)N * G, 0);
  shift_chain(c, xn, chain, E, K, G, B);
  contract_fwd(c, chain_slabs(xn, chain, E, K, slab), h, bias, 1.f, nullptr, 1, yn, B, F, E * K, G);
  if (F != 1) transpose(c, yn, nullptr, y, N, F, B, 1, (long long)N * F, 0, (long long)F * N, 0);
  return a.off;
}

size_t lsigf_backward_f32(const gcrnn_graph* g, const float* h, const float* x, const float* dy, float* dx, float* dh,
                          float* dbias, int F, int K, int G, int64_t B, void* ws, size_t wsb, cudaStream_t st) {
  Arena a(ws, wsb);
  Ctx c{g, st, a.dry()};
  const int N = g->N, E = g->E;
  const long long slab = (long long)B * N * G;
  float* xn = (G == 1) ? const_cast<float*>(x) : a.get<float>(slab);
  float* chain = a.get<float>((size_t)E * (K - 1) * slab);
  float* dyn = (F == 1) ? const_cast<float*>(dy) : a.get<float>((size_t)B * N * F);
  float* dz = a.get<float>((size_t)E * K * slab);
  float* dxn = (G == 1) ? dx : a.get<float>(slab);
  if (G != 1) transpose(c, x, nullptr, xn, G, N, B, 1, (long long)G * N, 0, (long long)N * G, 0);
  if (F != 1) transpose(c, dy, nullptr, dyn, F, N, B, 1, (long long)F * N, 0, (long long)N * F, 0);
  if (dh) {
    shift_chain(c, xn, chain, E, K, G, B);
    contract_wgrad(c, chain_slabs(xn, chain, E, K, slab), dyn, dh, (long long)B * N, F, E * K, G);
  }
  colsum(c, dyn, dbias, (long long)B * N, F);
  if (dx) {
    contract_bwd_data(c, mut_slabs(dz, E * K, slab), h, dyn, (long long)B * N, F, E * K, G, 0);
    reverse_chain(c, dz, E, K, G, B, dxn, false);
    if (G != 1) transpose(c, dxn, nullptr, dx, N, G, B, 1, (long long)N * G, 0, (long long)G * N, 0);
  }
  return a.off;
}

// ===================================================================================================
// graph attention, reference layout
// ===================================================================================================
size_t gat_forward_f32(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x, float* y,
                       int F, int G, int64_t B, void* ws, size_t wsb, cudaStream_t st) {
  GCRNN_CHECK(g->E == 1, "graph attention is defined for E == 1 (graphML.py:2327), got E=%d", g->E);
  Arena a(ws, wsb);
  Ctx c{g, st, a.dry()};
  const int N = g->N;
  float* xn = a.get<float>((size_t)B * N * G);
  float* yn = a.get<float>((size_t)B * N * F);
  GatBufs b; b.alloc(a, B, N, F, g->nnz_att);
  transpose(c, x, nullptr, xn, G, N, B, 1, (long long)G * N, 0, (long long)N * G, 0);
  gat_fwd_nm(c, mixer, weight, xn, yn, b, F, G, B);
  transpose(c, yn, nullptr, y, N, F, B, 1, (long long)N * F, 0, (long long)F * N, 0);
  return a.off;
}

size_t gat_backward_f32(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x, const float* dy,
                        float* dx, float* dmixer, float* dweight, int F, int G, int64_t B, void* ws, size_t wsb,
                        cudaStream_t st) {
  GCRNN_CHECK(g->E == 1, "graph attention is defined for E == 1 (graphML.py:2327), got E=%d", g->E);
  Arena a(ws, wsb);
  Ctx c{g, st, a.dry()};
  const int N = g->N;
  float* xn = a.get<float>((size_t)B * N * G);
  float* yn = a.get<float>((size_t)B * N * F);
  float* dyn = a.get<float>((size_t)B * N * F);
  float* dxn = a.get<float>((size_t)B * N * G);
  GatBufs b; b.alloc(a, B, N, F, g->nnz_att);
  GatBwdBufs w; w.alloc(a, B, N, F, g->nnz_att);
  transpose(c, x, nullptr, xn, G, N, B, 1, (long long)G * N, 0, (long long)N * G, 0);
  transpose(c, dy, nullptr, dyn, F, N, B, 1, (long long)F * N, 0, (long long)N * F, 0);
  gat_fwd_nm(c, mixer, weight, xn, yn, b, F, G, B);
  gat_bwd_nm(c, mixer, weight, xn, yn, dyn, b, w, dx ? dxn : nullptr, dmixer, dweight, F, G, B);
  if (dx) transpose(c, dxn, nullptr, dx, N, G, B, 1, (long long)N * G, 0, (long long)G * N, 0);
  return a.off;
}

// ===================================================================================================
// the gated GCRNN cell
// ===================================================================================================
namespace {

struct CellDims {
  int N, E, G, F, Kin, Kst;
  long long B, T, TB, NG, NF;
  bool tg, node, edge, gates;
};

CellDims dims_of(const gcrnn_cell* c, int64_t B, int64_t T) {
  CellDims d;
  d.N = c->g->N; d.E = c->d.E; d.G = c->d.G; d.F = c->d.F; d.Kin = c->d.Kin; d.Kst = c->d.Kst;
  d.B = B; d.T = T; d.TB = B * T; d.NG = (long long)d.N * d.G; d.NF = (long long)d.N * d.F;
  d.tg = c->d.time_gating != 0;
  d.node = c->d.spatial_gating == GCRNN_SPATIAL_NODE;
  d.edge = c->d.spatial_gating == GCRNN_SPATIAL_EDGE;
  d.gates = d.tg || d.node;
  GCRNN_CHECK(d.E == c->g->E, "cell E=%d but graph E=%d", d.E, c->g->E);
  GCRNN_CHECK(!d.edge || d.E == 1, "edge gating needs E == 1");
  GCRNN_CHECK(B > 0 && T > 0, "empty batch or sequence (B=%lld, T=%lld)", (long long)B, (long long)T);
  return d;
}

// what forward leaves for backward (all node-major, time-major)
struct Saved {
  float *Hn, *Xn, *zx, *h0n, *zh0, *gt, *qn;
  void layout(Arena& a, const CellDims& d) {
    Hn = a.get<float>(d.TB * d.NF);
    Xn = a.get<float>(d.TB * d.NG);
    zx = a.get<float>((size_t)d.E * (d.Kin - 1) * d.TB * d.NG);
    h0n = a.get<float>(d.B * d.NF);
    zh0 = d.gates ? a.get<float>((size_t)d.E * (d.Kst - 1) * d.B * d.NF) : nullptr;
    gt = d.tg ? a.get<float>(2 * d.TB) : nullptr;
    qn = d.node ? a.get<float>(2 * d.TB * d.N) : nullptr;
  }
};

struct SubCell { const float *A, *Bw, *b; };

// u = tanh( LSIGF(A_s, x_t, b_s) + LSIGF(B_s, h0, b_s) ) for every (t, b): graphML.py:2362 / :2383 with :2417-2423
void subcell_state(const Ctx& c, const CellDims& d, const Saved& s, const SubCell& sc, float* c0, float* ubuf) {
  contract_fwd(c, chain_slabs(s.h0n, s.zh0, d.E, d.Kst, d.B * d.NF), sc.Bw, sc.b, 1.f, nullptr, 1, c0, d.B, d.F, d.E * d.Kst, d.F);
  contract_fwd(c, chain_slabs(s.Xn, s.zx, d.E, d.Kin, d.TB * d.NG), sc.A, sc.b, 1.f, c0, d.B, ubuf, d.TB, d.F, d.E * d.Kin, d.G);
}

// node-gate head on u: q = sigmoid( sum_e Horner_k( p_{e,k} ) + c ), p_{e,k}[rn] = w[e,k,:] . u[rn,:]   (graphML.py:2385-2389)
void node_head_fwd(const Ctx& c, const CellDims& d, const float* ubuf, const float* w, const float* cb, float* pbuf,
                   float* v1, float* v2, float* q) {
  const long long RN = d.TB * d.N;
  const int S = d.E * d.Kst;
  if (!c.dry) {
    node_proj_fwd_k<<<grid1d(RN, 128), 128, (size_t)S * d.F * sizeof(float), c.st>>>(ubuf, w, pbuf, RN, d.F, S);
    check_launch();
  }
  float* lin = nullptr;
  for (int e = 0; e < d.E; ++e) {
    float* cur = pbuf ? pbuf + (long long)(e * d.Kst + d.Kst - 1) * RN : nullptr;
    for (int kk = d.Kst - 2; kk >= 0; --kk) {
      float* pk = pbuf ? pbuf + (long long)(e * d.Kst + kk) * RN : nullptr;
      float* out = (cur == v1) ? v2 : v1;
      spmm(c, c.g->fwd[e], cur, pk, out, 1, d.TB);
      cur = out;
    }
    if (e == 0) lin = cur;
    else add_inplace(c, lin, cur, RN);
    if (e == 0 && d.E > 1 && lin != q) {  // keep the running sum out of the ping-pong buffers
      copy(c, q, lin, RN * sizeof(float));
      lin = q;
    }
  }
  if (!c.dry) {
    sigmoid_bias_k<<<grid1d(RN, TPB), TPB, 0, c.st>>>(lin, cb, q, RN);
    check_launch();
  }
}

}  // namespace

// ===================================================================================================
// fused edge-gated path for F == 32 (sp32_kernels.cuh): 4 kernels per forward step, 5 per backward step
// ===================================================================================================
namespace {

bool edge32_ok(const gcrnn_cell* cell) {
  const gcrnn_cell_desc& d = cell->d;
  const gcrnn_graph* g = cell->g;
  return opt().sparse_fused && d.spatial_gating == GCRNN_SPATIAL_EDGE && !d.time_gating && d.E == 1 && g->E == 1 && d.F == 32 &&
         d.Kst >= 2 && d.Kst <= 4 && d.Kin * d.G <= e32::MAXKG && d.Kin <= e32::MAXK && g->max_row_deg <= 32 && g->N >= 64 &&
         (long long)g->N * 32 < INT_MAX;
}

// extra state the fused forward leaves for the fused backward (appended after the generic `Saved` block)
struct Saved32 {
  float *zc, *wu_a, *wu_r; float4* info; uint2* masks;
  void layout(Arena& a, const CellDims& d) {
    zc = a.get<float>((size_t)(d.Kst - 2) * d.TB * d.NF);     // z_1 .. z_{Kst-2} of every step
    wu_a = a.get<float>(d.TB * d.NF);
    wu_r = a.get<float>(d.TB * d.NF);
    info = a.get<float4>(2 * d.TB * d.N);
    masks = a.get<uint2>(d.TB * d.N);
  }
};

int sm_count() {                       // of the CURRENT device (cached per device: a process may drive several)
  static int cached[64] = {0};
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && cached[dev]) return cached[dev];
  int n = 0;
  CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) cached[dev] = n;
  return n;
}
template <class K>
int persistent_grid(K kernel, int block, long long tasks_per_block_unit, long long tasks, size_t dyn_smem = 0) {
  int occ = 1;
  if (dyn_smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, dyn_smem));
  long long g = (long long)sm_count() * std::max(occ, 1);
  const long long need = (tasks + tasks_per_block_unit - 1) / tasks_per_block_unit;
  return (int)std::max<long long>(1, std::min(g, need));
}

// Grid of a second-generation (sp32_tile.cuh) kernel: `blocks_per_sm` resident blocks per SM, and a shared-memory
// carve-out that covers exactly those blocks so that the rest of the 228 KB stays L1 (the neighbour gathers live on L1 hits).
template <class K>
int tile_grid(K kernel, int block, size_t dyn_smem, int blocks_per_sm, long long units) {
  if (dyn_smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  const size_t want = (dyn_smem + 1024) * (size_t)blocks_per_sm;
  const int carve = dyn_smem == 0 ? 0 : (int)std::min<size_t>(100, (want * 100 + 228 * 1024 - 1) / (228 * 1024));
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  int occ = 1;
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, dyn_smem));
  occ = std::max(1, std::min(occ, blocks_per_sm));
  return (int)std::max<long long>(1, std::min<long long>((long long)sm_count() * occ, units));
}
enum { V2_SPMM = 1, V2_FILTER = 2, V2_AGG = 4, V2_ROWS = 8, V2_NODE = 16, V2_DH = 32 };
// Stage mask in effect for the current call.  The aggregate and bwd_rows stages share the layout of the saved softmax
// statistics, so they switch generation together, and a backward always follows the choice its forward made (recorded in the
// cell) even if the debug option changed in between.
thread_local int t_v2_mask = 63;
inline int normalise_v2(int m) {
  if ((m & (V2_AGG | V2_ROWS)) != (V2_AGG | V2_ROWS)) m &= ~(V2_AGG | V2_ROWS);
  return m;
}
inline bool v2(int bit) { return (t_v2_mask & bit) != 0; }
inline int tile_bps() { return std::max(1, std::min(2, opt().sparse_v2_bps)); }

void spmm32(const Ctx& c, const Gather& op, const float* in, float* out, long long R) {
  const long long RN = R * c.g->N;
  const e32::Gather3 g3{op.ptr, op.idx, op.val};
  if (v2(V2_SPMM)) {
    const long long groups = R * ((c.g->N + 63) / 64);
    e32::spmm32_v2_k<<<tile_grid(e32::spmm32_v2_k, 256, 0, 3, groups), 256, 0, c.st>>>(g3, in, out, c.g->N, R);
  } else {
    e32::spmm32_k<<<persistent_grid(e32::spmm32_k, 256, 8, RN), 256, 0, c.st>>>(g3, in, out, c.g->N, RN);
  }
  check_launch();
}

e32::Chain x_taps(const CellDims& d, const Saved& s, long long t) {
  e32::Chain xs{};
  for (int k = 0; k < d.Kin; ++k)
    xs.p[k] = (k == 0 ? s.Xn : s.zx + (long long)(k - 1) * d.TB * d.NG) + t * d.B * d.NG;
  return xs;
}

template <int KST>
void e32_forward_steps(const Ctx& c, const CellDims& d, const gcrnn_cell_params* p, const Saved& s, const Saved32& x,
                       const float* prep, float4* rc) {
  const gcrnn_graph* g = c.g;
  const long long BN = d.B * d.N, BNF = d.B * d.NF;
  const Gather& fw = g->fwd[0];
  const int g_filter = persistent_grid(e32::filter_fwd_k<KST>, 128, 4, BN);
  const size_t agg_smem = (size_t)4 * 2 * e32::AGG_STAGE * sizeof(float);
  const int g_agg = persistent_grid(e32::aggregate_k, 128, 4, BN, agg_smem);
  const e32::Gather3 gop{fw.ptr, fw.idx, fw.val};
  const bool tc = opt().sparse_v2_tc != 0;        // 3xTF32 mma.sync contraction, else packed FFMA2
  const long long tiles = d.B * ((d.N + 127) / 128), groups = d.B * ((d.N + 31) / 32);      // 256-thread blocks on 128-node tiles
  const size_t gc_smem = (size_t)e32::gc_smem_floats<KST, 256>(e32::GC_FILTER) * sizeof(float);
  auto k_filter = tc ? e32::gather_contract_k<KST, e32::GC_FILTER, 256, true> : e32::gather_contract_k<KST, e32::GC_FILTER, 256, false>;
  const int g_filter2 = v2(V2_FILTER) ? tile_grid(k_filter, 256, gc_smem, tile_bps(), tiles) : 0;
  const int g_agg2 = v2(V2_AGG) ? tile_grid(e32::aggregate_v2_k, 256, 0, 4, groups) : 0;
  for (long long t = 0; t < d.T; ++t) {
    e32::Chain zc{};
    zc.p[0] = t == 0 ? s.h0n : s.Hn + (t - 1) * BNF;
    for (int k = 1; k <= KST - 2; ++k) {
      float* out = x.zc + ((long long)(k - 1) * d.T + t) * BNF;
      spmm32(c, fw, zc.p[k - 1], out, d.B);
      zc.p[k] = out;
    }
    float* wa = x.wu_a + t * BNF; float* wr = x.wu_r + t * BNF;
    float4* info = x.info + 2 * t * BN;
    if (v2(V2_FILTER))
      k_filter<<<g_filter2, 256, gc_smem, c.st>>>(gop, zc, x_taps(d, s, t), d.Kin, d.G, prep, p->e_mixer[0], p->e_mixer[1], wa, wr, rc, d.N, d.B,
                                                  e32::DpreFuse{});
    else
      e32::filter_fwd_k<KST><<<g_filter, 128, 0, c.st>>>(gop, zc, x_taps(d, s, t), d.Kin, d.G, prep,
                                                       p->e_mixer[0], p->e_mixer[1], wa, wr, rc, d.N, BN);
    check_launch();
    if (v2(V2_AGG))               // compact statistics layout inside the same `info` block: BN float4 (cl) + BN float2 (rr)
      e32::rowstats_v2_k<<<grid1d(BN, 256), 256, 0, c.st>>>(g->att_rptr, g->att_col, rc, info, reinterpret_cast<float2*>(info + BN), d.N, BN);
    else
      e32::rowstats_k<<<grid1d(BN, 256), 256, 0, c.st>>>(g->att_rptr, g->att_col, rc, info, d.N, BN);
    check_launch();
    if (v2(V2_AGG))
      e32::aggregate_v2_k<<<g_agg2, 256, 0, c.st>>>(g->att_cptr, g->att_crow, g->att_cval, info, reinterpret_cast<const float2*>(info + BN),
                                                    wa, wr, s.Hn + t * BNF, x.masks + t * BN, d.N, d.B);
    else
      e32::aggregate_k<<<g_agg, 128, agg_smem, c.st>>>(g->att_cptr, g->att_crow, g->att_cval, info, wa, wr, s.Hn + t * BNF,
                                              x.masks + t * BN, d.N, BN);
    check_launch();
  }
}

size_t cell_forward_e32(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                        void* saved, size_t savedb, size_t* saved_used, void* ws, size_t wsb, int64_t B, int64_t T,
                        cudaStream_t st) {
  const CellDims d = dims_of(cell, B, T);
  Arena a(ws, wsb);
  // the graph itself, or its renumbered copy (graph.cu).  A caller that will ask for dX gets the plain order: dX comes from the
  // generic reverse sweep, which reads this forward's saved state in the caller's numbering
  const gcrnn_graph* view = cell->need_dx ? cell->g : locality_view(cell->g);
  const NodeMap perm = node_map(cell->g, view);
  Ctx c{view, st, a.dry()};
  t_v2_mask = normalise_v2(opt().sparse_v2);
  if (ws != nullptr) { cell->fwd_v2_mask = t_v2_mask; cell->fwd_reordered = view != cell->g; }
  Saved s; Saved32 x;
  {
    Arena sa(saved, savedb);
    s.layout(sa, d); x.layout(sa, d);
    if (saved_used) *saved_used = sa.off;
  }
  GCRNN_CHECK(a.dry() || saved, "forward needs the `saved` buffer (see gcrnn_cell_workspace_bytes)");
  float* prep = a.get<float>(e32::PrepLayout::TOTAL);
  float4* rc = a.get<float4>(d.B * d.N);
  // x-tap chain through a node-major detour ([T*B][N*G] -> [N][G*T*B]): a neighbour is G*T*B contiguous floats there, so the
  // shift is a coalesced row gather instead of 4-byte gathers (L1-bound: 3.5 ms per tap at cfg5)
  const bool detour = d.Kin > 1 && d.TB * d.G >= 32;
  float* xa = detour ? a.get<float>(d.TB * d.NG) : nullptr;
  float* xb = detour ? a.get<float>(d.TB * d.NG) : nullptr;
  if (a.dry()) return a.off;
  if (perm) {
    permute_nodes(c, true, X, s.Xn, perm, d.G, d.N, d.B, d.T, d.T * d.NG, d.NG, d.NG, d.B * d.NG);
    permute_nodes(c, true, h0, s.h0n, perm, d.F, d.N, d.B, 1, d.NF, 0, d.NF, 0);
  } else {
    transpose(c, X, nullptr, s.Xn, d.G, d.N, d.B, d.T, d.T * d.NG, d.NG, d.NG, d.B * d.NG);        // [B,T,G,N] -> [T,B,N,G]
    transpose(c, h0, nullptr, s.h0n, d.F, d.N, d.B, 1, d.NF, 0, d.NF, 0);                          // [B,F,N]  -> [B,N,F]
  }
  if (detour) {
    transpose(c, s.Xn, nullptr, xa, (int)d.TB, (int)d.NG, 1, 1, 0, 0, 0, 0);                       // [TB][N*G] -> [N*G][TB]
    for (int k = 1; k < d.Kin; ++k) {
      spmm(c, view->fwd[0], xa, nullptr, xb, (int)(d.G * d.TB), 1);
      transpose(c, xb, nullptr, s.zx + (long long)(k - 1) * d.TB * d.NG, (int)d.NG, (int)d.TB, 1, 1, 0, 0, 0, 0);
      std::swap(xa, xb);
    }
  } else {
    shift_chain(c, s.Xn, s.zx, d.E, d.Kin, d.G, d.TB);
  }
  e32::prep_k<<<1, 1024, 0, st>>>(p->weight_A, p->weight_B, p->bias, p->e_weight[0], p->e_weight[1], prep, d.Kin * d.G, d.Kst);
  check_launch();
  switch (d.Kst) {
    case 2: e32_forward_steps<2>(c, d, p, s, x, prep, rc); break;
    case 3: e32_forward_steps<3>(c, d, p, s, x, prep, rc); break;
    default: e32_forward_steps<4>(c, d, p, s, x, prep, rc); break;
  }
  if (perm) permute_nodes(c, false, s.Hn, H, perm, d.F, d.N, d.T, d.B, d.B * d.NF, d.NF, d.NF, d.T * d.NF);
  else transpose(c, s.Hn, nullptr, H, d.N, d.F, d.T, d.B, d.B * d.NF, d.NF, d.NF, d.T * d.NF);      // [T,B,N,F] -> [B,T,F,N]
  return a.off;
}

struct Bwd32Bufs { float *dya, *dyr, *pa, *pr, *dd, *wch, *dhrec, *acc, *dhn; float2* dr; };

template <int KST>
void e32_backward_steps(const Ctx& c, const CellDims& d, const gcrnn_cell_params* p, const Saved& s, const Saved32& x,
                        const DhView& dv, const Bwd32Bufs& b, NodeMap perm) {
  const gcrnn_graph* g = c.g;
  const long long BN = d.B * d.N, BNF = d.B * d.NF;
  const Gather& fw = g->fwd[0];
  const Gather& bw = g->bwd[0];
  const int g_dpre = (int)std::min<long long>(d.B * ((d.N + 31) / 32), (long long)sm_count() * 8);
  const int rows_cap = (g->max_row_deg + 3) & ~3;                    // staged rows per gate in bwd_rows_k
  const size_t rows_smem = (size_t)4 * 2 * e32::rows_stage_floats(rows_cap) * sizeof(float);
  const int g_rows = persistent_grid(e32::bwd_rows_k, 128, 4, BN, rows_smem);
  const e32::Gather3 gfw{fw.ptr, fw.idx, fw.val}, gbw{bw.ptr, bw.idx, bw.val};
  const int g_node = persistent_grid(e32::bwd_node_k<KST>, 128, 4, BN);
  const int g_dh = persistent_grid(e32::dh_k<KST>, 128, 4, BN);
  const bool tc = opt().sparse_v2_tc != 0;
  const long long tiles = d.B * ((d.N + 127) / 128), groups = d.B * ((d.N + 31) / 32);
  const size_t dh_smem = (size_t)e32::gc_smem_floats<KST, 256>(e32::GC_DH) * sizeof(float);
  const size_t node_smem = (size_t)e32::bwd_node_smem_floats<KST, 256>() * sizeof(float);
  auto k_dh = tc ? e32::gather_contract_k<KST, e32::GC_DH, 256, true> : e32::gather_contract_k<KST, e32::GC_DH, 256, false>;
  auto k_node = tc ? e32::bwd_node_v2_k<KST, 256, true> : e32::bwd_node_v2_k<KST, 256, false>;
  const int g_dh2 = v2(V2_DH) ? tile_grid(k_dh, 256, dh_smem, tile_bps(), tiles) : 0;
  const bool fuse_dpre = tc && v2(V2_DH) && opt().sparse_v2_fuse_dpre;     // dh epilogue finishes the next step's dpre (t = 0 still writes dh0)
  const int g_node2 = v2(V2_NODE) ? tile_grid(k_node, 256, node_smem, tile_bps(), tiles) : 0;
  auto k_rows = opt().sparse_v2_rows_bps == 3 ? e32::bwd_rows_v2_k<3> : e32::bwd_rows_v2_k<2>;     // 3: more warps, some spills
  const int g_rows2 = v2(V2_ROWS) ? tile_grid(k_rows, 256, 0, opt().sparse_v2_rows_bps == 3 ? 3 : 2, groups) : 0;
  // dH of step t as the dpre code reads it: the caller's reference layout, or (renumbered graph) converted into b.dhn first
  struct DhStep { const float* p; long long bs; int node_major; };
  auto dh_step = [&](long long t) {
    if (!perm || (dv.last_only && t < d.T - 1)) return DhStep{dv.ptr(t), dv.bstride(t), 0};      // the zero slab needs no renumbering
    permute_nodes(c, true, dv.ptr(t), b.dhn, perm, d.F, d.N, d.B, 1, dv.bstride(t), 0, d.NF, 0);
    return DhStep{b.dhn, d.NF, 1};
  };
  for (long long t = d.T - 1; t >= 0; --t) {
    const float* hn = s.Hn + t * BNF;
    const float* wa = x.wu_a + t * BNF; const float* wr = x.wu_r + t * BNF;
    const float4* info = x.info + 2 * t * BN;
    if (!(fuse_dpre && t < d.T - 1)) {            // otherwise the dh kernel of step t+1 already left dya / dyr of this step
      const DhStep g = dh_step(t);
      e32::dpre_k<<<g_dpre, 256, 0, c.st>>>(g.p, g.bs, t == d.T - 1 ? nullptr : b.dhrec, hn, x.masks + t * BN, b.dya, b.dyr,
                                            d.N, d.B, g.node_major);
      check_launch();
    }
    zero(c, b.dr, BN * sizeof(float2));
    if (v2(V2_ROWS))
      k_rows<<<g_rows2, 256, 0, c.st>>>(g->att_rptr, g->att_col, g->att_val, info, reinterpret_cast<const float2*>(info + BN), wa, wr, b.dya, b.dyr,
                                                    p->e_mixer[0], p->e_mixer[1], b.pa, b.pr, b.dr, b.acc, d.N, d.B);
    else
      e32::bwd_rows_k<<<g_rows, 128, rows_smem, c.st>>>(g->att_rptr, g->att_col, g->att_val, info, wa, wr, b.dya, b.dyr,
                                              p->e_mixer[0], p->e_mixer[1], b.pa, b.pr, reinterpret_cast<float*>(b.dr), b.acc, rows_cap, d.N, BN);
    check_launch();
    e32::Chain zc{};
    zc.p[0] = t == 0 ? s.h0n : s.Hn + (t - 1) * BNF;
    for (int k = 1; k <= KST - 2; ++k) zc.p[k] = x.zc + ((long long)(k - 1) * d.T + t) * BNF;
    if (v2(V2_NODE))
      k_node<<<g_node2, 256, node_smem, c.st>>>(gfw, zc, x_taps(d, s, t), d.Kin, d.G, b.pa, b.pr, b.dr, wa, wr,
                                                           p->e_mixer[0], p->e_mixer[1], p->e_weight[1], b.dd, b.acc, d.N, d.B);
    else
      e32::bwd_node_k<KST><<<g_node, 128, 0, c.st>>>(gfw, zc, x_taps(d, s, t), d.Kin, d.G, b.pa, b.pr, b.dr, wa, wr,
                                                   p->e_mixer[0], p->e_mixer[1], p->e_weight[1], b.dd, b.acc, d.N, BN);
    check_launch();
    e32::Chain wc{};
    wc.p[0] = b.dd;
    for (int k = 1; k <= KST - 2; ++k) {
      float* out = b.wch + (long long)(k - 1) * BNF;
      spmm32(c, bw, wc.p[k - 1], out, d.B);
      wc.p[k] = out;
    }
    if (v2(V2_DH)) {
      e32::DpreFuse fz{};
      if (fuse_dpre && t > 0) {
        const DhStep g = dh_step(t - 1);
        fz = e32::DpreFuse{g.p, g.bs, s.Hn + (t - 1) * BNF, x.masks + (t - 1) * BN, b.dya, b.dyr, g.node_major};
      }
      k_dh<<<g_dh2, 256, dh_smem, c.st>>>(gbw, wc, e32::Chain{}, 0, 1, p->weight_B, nullptr, nullptr, nullptr, b.dhrec, nullptr, d.N, d.B, fz);
    }
    else
      e32::dh_k<KST><<<g_dh, 128, 0, c.st>>>(gbw, wc, p->weight_B, b.dhrec, d.N, BN);
    check_launch();
  }
}

size_t cell_backward_e32(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* dH, const void* saved, size_t savedb,
                         const gcrnn_cell_params* gr, float* dh0, void* ws, size_t wsb, int64_t B, int64_t T, cudaStream_t st) {
  const CellDims d = dims_of(cell, B, T);
  Arena a(ws, wsb);
  const gcrnn_graph* view = cell->fwd_reordered ? locality_view(cell->g) : cell->g;      // the numbering the forward's saved state is in
  const NodeMap perm = node_map(cell->g, view);
  Ctx c{view, st, a.dry()};
  t_v2_mask = (normalise_v2(opt().sparse_v2) & ~(V2_AGG | V2_ROWS)) | (cell->fwd_v2_mask & (V2_AGG | V2_ROWS));
  Saved s; Saved32 x;
  { Arena sa(const_cast<void*>(saved), savedb); s.layout(sa, d); x.layout(sa, d); }
  Bwd32Bufs b;
  b.dya = a.get<float>(d.B * d.NF); b.dyr = a.get<float>(d.B * d.NF); b.pa = a.get<float>(d.B * d.NF); b.pr = a.get<float>(d.B * d.NF);
  b.dd = a.get<float>(d.B * d.NF); b.wch = a.get<float>((size_t)(d.Kst - 2) * d.B * d.NF); b.dhrec = a.get<float>(d.B * d.NF);
  b.dr = a.get<float2>(d.B * d.N); b.acc = a.get<float>(e32::AccLayout::TOTAL);
  float* zslab = cell->dh_last_only ? a.get<float>(d.NF) : nullptr;
  b.dhn = perm ? a.get<float>(d.B * d.NF) : nullptr;      // one step's dH converted to the renumbered node-major layout
  if (a.dry()) return a.off;
  if (zslab) zero(c, zslab, d.NF * sizeof(float));
  const DhView dv{dH, zslab, d.T, d.NF, cell->dh_last_only != 0};
  zero(c, b.acc, e32::AccLayout::TOTAL * sizeof(float));
  switch (d.Kst) {
    case 2: e32_backward_steps<2>(c, d, p, s, x, dv, b, perm); break;
    case 3: e32_backward_steps<3>(c, d, p, s, x, dv, b, perm); break;
    default: e32_backward_steps<4>(c, d, p, s, x, dv, b, perm); break;
  }
  e32::finalize_k<<<1, 1024, 0, st>>>(b.acc, p->weight_A, p->weight_B, p->bias, p->e_weight[0], p->e_weight[1], gr->weight_A,
                                      gr->weight_B, gr->bias, gr->e_mixer[0], gr->e_weight[0], gr->e_mixer[1], gr->e_weight[1],
                                      d.Kin * d.G, d.Kst);
  check_launch();
  if (dh0 && perm) permute_nodes(c, false, b.dhrec, dh0, perm, d.F, d.N, d.B, 1, d.NF, 0, d.NF, 0);
  else if (dh0) transpose(c, b.dhrec, nullptr, dh0, d.N, d.F, d.B, 1, d.NF, 0, d.NF, 0);
  return a.off;
}

// test aid: the ReLU decisions the fused forward saved, in the caller's layout and node order, out[gate][b][t][f][n] (1 = passed)
__global__ void e32_decode_masks_k(const uint2* __restrict__ masks, const int* __restrict__ perm, unsigned char* __restrict__ out,
                                   int N, long long B, long long T) {
  const long long total = T * B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long tb = i / N, b = tb % B, t = tb / B;
    const uint2 mk = masks[i];
    const int no = perm ? perm[n] : n;
    for (int f = 0; f < 32; ++f) {
      const int bit = e32::lane_of_feat(f);
      out[(((0 * B + b) * T + t) * 32 + f) * N + no] = (mk.x >> bit) & 1u;
      out[(((1 * B + b) * T + t) * 32 + f) * N + no] = (mk.y >> bit) & 1u;
    }
  }
}
}  // namespace

void debug_edge_relu_masks(const gcrnn_cell* cell, const void* saved, size_t savedb, int64_t B, int64_t T, unsigned char* out, cudaStream_t st) {
  GCRNN_CHECK(cell->last_path == GCRNN_PATH_NODE32, "the last forward of this cell did not take the fused edge-gated path");
  const CellDims d = dims_of(cell, B, T);
  Saved s; Saved32 x;
  { Arena sa(const_cast<void*>(saved), savedb); s.layout(sa, d); x.layout(sa, d); }
  const gcrnn_graph* view = cell->fwd_reordered ? locality_view(cell->g) : cell->g;
  e32_decode_masks_k<<<grid1d(d.TB * d.N, 256), 256, 0, st>>>(x.masks, view != cell->g ? cell->g->perm : nullptr, out, d.N, B, T);
  check_launch();
}

namespace {
// ===================================================================================================
// persistent fused recurrence for small graphs (persist_f32.cuh): one launch forward, one launch backward
// ===================================================================================================
constexpr long long PERSIST_SMEM_MAX = 220 * 1024;
persist::Shape persist_shape(const gcrnn_cell* cell) {
  const gcrnn_cell_desc& d = cell->d;
  return persist::Shape{d.F, d.G, d.Kin, d.Kst, cell->g->N, d.time_gating != 0, d.spatial_gating == GCRNN_SPATIAL_NODE,
                        d.spatial_gating == GCRNN_SPATIAL_EDGE, (int)cell->g->nnz_att};
}
// bytes that must sit in shared memory besides the float buffers: the attention pattern of an edge-gated cell
long long persist_fixed_bytes(const gcrnn_cell* cell) {
  return cell->d.spatial_gating == GCRNN_SPATIAL_EDGE ? persist::att_bytes(cell->g->N, (int)cell->g->nnz_att) : 0;
}
bool persist_ok(const gcrnn_cell* cell) {
  const gcrnn_cell_desc& d = cell->d;
  const gcrnn_graph* g = cell->g;
  if (!opt().persist || d.E != 1 || g->E != 1 || cell->need_dx || g->N >= 65536 || g->nnz_att >= 65536) return false;
  return persist::bwd_floats(persist_shape(cell)) * 4 + persist_fixed_bytes(cell) <= PERSIST_SMEM_MAX;
}
// stage the gather lists in shared memory when both fit next to the rest
bool persist_lists_fit(const gcrnn_cell* cell) {
  const gcrnn_graph* g = cell->g;
  return persist::bwd_floats(persist_shape(cell)) * 4 + persist_fixed_bytes(cell) + 2 * persist::list_bytes(g->N, (int)g->fwd[0].nnz) <= 226 * 1024;
}
persist::Args persist_args(const gcrnn_cell* cell, const gcrnn_cell_params* p, int64_t B, int64_t T) {
  const gcrnn_graph* g = cell->g;
  persist::Args a{};
  a.N = g->N; a.F = cell->d.F; a.G = cell->d.G; a.Kin = cell->d.Kin; a.Kst = cell->d.Kst;
  a.tg = cell->d.time_gating != 0; a.node = cell->d.spatial_gating == GCRNN_SPATIAL_NODE; a.edge = cell->d.spatial_gating == GCRNN_SPATIAL_EDGE;
  a.has_bias = cell->d.bias != 0; a.B = B; a.T = T;
  a.arptr = g->att_rptr; a.acol = g->att_col; a.aval = g->att_val; a.acptr = g->att_cptr; a.acrow = g->att_crow; a.aceid = g->att_ceid;
  a.annz = (int)g->nnz_att;
  a.cptr = g->fwd[0].ptr; a.cidx = g->fwd[0].idx; a.cval = g->fwd[0].val;
  a.rptr = g->bwd[0].ptr; a.ridx = g->bwd[0].idx; a.rval = g->bwd[0].val;
  a.nnz = (int)g->fwd[0].nnz; a.lists_smem = persist_lists_fit(cell);
  if (p) {
    a.A = p->weight_A; a.Bw = p->weight_B; a.bias = p->bias;
    for (int i = 0; i < 2; ++i) {
      a.tA[i] = p->t_weight_A[i]; a.tB[i] = p->t_weight_B[i]; a.tb[i] = p->t_bias[i]; a.tW[i] = p->t_mlp_w[i]; a.tc[i] = p->t_mlp_b[i];
      a.nA[i] = p->n_weight_A[i]; a.nB[i] = p->n_weight_B[i]; a.nb[i] = p->n_bias[i]; a.nhw[i] = p->n_head_w[i]; a.nhb[i] = p->n_head_b[i];
      a.eW[i] = p->e_weight[i]; a.em[i] = p->e_mixer[i];
    }
  }
  return a;
}
// kernel variants: NB = 4 / 1, spatial gating none / node / edge, quad layout on / off
template <int NB, int SG, bool QZ>
void persist_launch_v(bool bwd, const persist::Args& a, unsigned B, size_t smem, cudaStream_t st) {
  static DeviceOnce once;
  if (once.first()) {
    CUDA_OK(cudaFuncSetAttribute(persist::persist_fwd_k<NB, SG, QZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(persist::persist_bwd_k<NB, SG, QZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  if (bwd) persist::persist_bwd_k<NB, SG, QZ><<<B, persist::PT, smem, st>>>(a);
  else persist::persist_fwd_k<NB, SG, QZ><<<B, persist::PT, smem, st>>>(a);
  check_launch();
}
template <int SG>
void persist_launch_sg(bool bwd, const persist::Args& a, unsigned B, size_t smem, cudaStream_t st) {
  // NB = 4: blocked loops over 4 consecutive nodes with 16-byte loads along the nodes (N % 4 == 0 and F % 4 == 0: cfg1).  Otherwise, with
  // F % 4 == 0 (cfg2: N = 59), the state-side slabs take the quad layout (QZ): +21 % / +12 % at cfg2-node / cfg2-edge.  The two do not
  // combine as written: four consecutive nodes of a quad-layout slab are 64 bytes apart per thread (4-way bank conflicts, cfg1 205k ->
  // 182k), and NB = 1 with the quad layout measures the same as NB = 4 without it at cfg1 (200k vs 205k).
  if (a.N % 4 == 0 && a.F % 4 == 0) persist_launch_v<4, SG, false>(bwd, a, B, smem, st);
  else if (a.F % 4 == 0) persist_launch_v<1, SG, true>(bwd, a, B, smem, st);
  else persist_launch_v<1, SG, false>(bwd, a, B, smem, st);
}
void persist_launch(bool bwd, const persist::Args& a, unsigned B, size_t smem, cudaStream_t st) {
  if (a.node) persist_launch_sg<1>(bwd, a, B, smem, st);
  else if (a.edge) persist_launch_sg<2>(bwd, a, B, smem, st);
  else persist_launch_sg<0>(bwd, a, B, smem, st);
}
size_t cell_forward_persist(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                            void* saved, size_t savedb, size_t* saved_used, void* ws, int64_t B, int64_t T, cudaStream_t st) {
  Arena sa(saved, savedb);
  float* gt = sa.get<float>(2 * B * T);
  float* qn = sa.get<float>(cell->d.spatial_gating == GCRNN_SPATIAL_NODE ? (size_t)2 * B * T * cell->g->N : 0);
  if (saved_used) *saved_used = sa.off;
  if (ws == nullptr) return 256;
  GCRNN_CHECK(saved != nullptr, "forward needs the `saved` buffer");
  persist::Args a = persist_args(cell, p, B, T);
  a.X = X; a.h0 = h0; a.H = H; a.gt = gt; a.qn = qn;
  const size_t smem = (size_t)persist::fwd_floats(persist_shape(cell)) * sizeof(float) + persist_fixed_bytes(cell) + (a.lists_smem ? persist::list_bytes(a.N, a.nnz) : 0);
  persist_launch(false, a, (unsigned)B, smem, st);
  return 256;
}
size_t cell_backward_persist(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0, const float* H,
                             const float* dH, const void* saved, size_t savedb, const gcrnn_cell_params* gr, float* dh0, void* ws,
                             int64_t B, int64_t T, cudaStream_t st) {
  Arena sa(const_cast<void*>(saved), savedb);
  float* gt = sa.get<float>(2 * B * T);
  float* qn = sa.get<float>(cell->d.spatial_gating == GCRNN_SPATIAL_NODE ? (size_t)2 * B * T * cell->g->N : 0);
  if (ws == nullptr) return 256;
  persist::Args a = persist_args(cell, p, B, T);
  const long long FN = (long long)a.F * a.N;
  a.X = X; a.h0 = h0; a.H = const_cast<float*>(H); a.gt = gt; a.qn = qn;
  a.dH = dH; a.dh_last_only = cell->dh_last_only;
  a.dH_bstride = cell->dh_last_only ? FN : T * FN; a.dH_tstride = FN;
  a.dA = gr->weight_A; a.dBw = gr->weight_B; a.dbias = gr->bias;
  for (int i = 0; i < 2; ++i) {
    a.dtA[i] = gr->t_weight_A[i]; a.dtB[i] = gr->t_weight_B[i]; a.dtb[i] = gr->t_bias[i]; a.dtW[i] = gr->t_mlp_w[i]; a.dtc[i] = gr->t_mlp_b[i];
    a.dnA[i] = gr->n_weight_A[i]; a.dnB[i] = gr->n_weight_B[i]; a.dnb[i] = gr->n_bias[i]; a.dnhw[i] = gr->n_head_w[i]; a.dnhb[i] = gr->n_head_b[i];
    a.deW[i] = gr->e_weight[i]; a.dem[i] = gr->e_mixer[i];
  }
  a.dh0 = dh0;
  const size_t smem = (size_t)persist::bwd_floats(persist_shape(cell)) * sizeof(float) + persist_fixed_bytes(cell) + (a.lists_smem ? 2 * persist::list_bytes(a.N, a.nnz) : 0);
  persist_launch(true, a, (unsigned)B, smem, st);
  return 256;
}

// which path a forward takes (GCRNN_PATH_*): the persistent kernel for small graphs (any gating, one edge feature), the fused per-node
// kernels for F == 32 edge gating when the shape allows, else the generic kernels
int pick_path(const gcrnn_cell* cell) {
  if (cell->forced_path >= 0) {
    if (cell->forced_path == GCRNN_PATH_NODE32) GCRNN_CHECK(edge32_ok(cell), "path NODE32 does not support this cell");
    if (cell->forced_path == GCRNN_PATH_PERSIST) {
      GCRNN_CHECK(cell->d.E == 1 && cell->g->nnz_att < 65536 &&
                  persist::bwd_floats(persist_shape(cell)) * 4 + persist_fixed_bytes(cell) <= PERSIST_SMEM_MAX,
                  "path PERSIST does not support this cell");
    }
    return cell->forced_path;
  }
  if (persist_ok(cell)) return GCRNN_PATH_PERSIST;
  return edge32_ok(cell) ? GCRNN_PATH_NODE32 : GCRNN_PATH_GENERIC;
}
}  // namespace

size_t cell_forward_f32(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                        void* saved, size_t savedb, size_t* saved_used, void* ws, size_t wsb, int64_t B, int64_t T,
                        cudaStream_t st) {
  {
    const int path = pick_path(cell);
    if (ws != nullptr) cell->last_path = path;
    if (path == GCRNN_PATH_NODE32) return cell_forward_e32(cell, p, X, h0, H, saved, savedb, saved_used, ws, wsb, B, T, st);
    if (path == GCRNN_PATH_PERSIST) return cell_forward_persist(cell, p, X, h0, H, saved, savedb, saved_used, ws, B, T, st);
  }
  const CellDims d = dims_of(cell, B, T);
  Arena a(ws, wsb);
  Ctx c{cell->g, st, a.dry()};
  Saved s;
  {
    Arena sa(saved, savedb);
    s.layout(sa, d);
    if (saved_used) *saved_used = sa.off;
  }
  GCRNN_CHECK(a.dry() || saved, "forward needs the `saved` buffer (see gcrnn_cell_workspace_bytes)");
  const gcrnn_graph* g = cell->g;
  // scratch
  float* zh = a.get<float>((size_t)d.E * (d.Kst - 1) * d.B * d.NF);
  float* ua = a.get<float>(d.B * d.NF);
  float* ur = a.get<float>(d.B * d.NF);
  GatBufs gb; float *qa = nullptr, *qr = nullptr;
  if (d.edge) { gb.alloc(a, d.B, d.N, d.F, g->nnz_att); qa = a.get<float>(d.B * d.NF); qr = a.get<float>(d.B * d.NF); }
  float *ubuf = nullptr, *c0 = nullptr, *wg = nullptr, *pbuf = nullptr, *v1 = nullptr, *v2 = nullptr;
  if (d.gates) { ubuf = a.get<float>(d.TB * d.NF); c0 = a.get<float>(d.B * d.NF); }
  if (d.tg) wg = a.get<float>(d.NF);
  if (d.node) { pbuf = a.get<float>((size_t)d.E * d.Kst * d.TB * d.N); v1 = a.get<float>(d.TB * d.N); v2 = a.get<float>(d.TB * d.N); }
  if (a.dry()) return a.off;

  // ---- layout change + the non-recurrent part ---------------------------------------------------------
  transpose(c, X, nullptr, s.Xn, d.G, d.N, d.B, d.T, d.T * d.NG, d.NG, d.NG, d.B * d.NG);          // [B,T,G,N] -> [T,B,N,G]
  transpose(c, h0, nullptr, s.h0n, d.F, d.N, d.B, 1, d.NF, 0, d.NF, 0);                            // [B,F,N]  -> [B,N,F]
  shift_chain(c, s.Xn, s.zx, d.E, d.Kin, d.G, d.TB);
  if (d.gates) shift_chain(c, s.h0n, s.zh0, d.E, d.Kst, d.F, d.B);
  if (d.tg)
    for (int gi = 0; gi < 2; ++gi) {                                                             // graphML.py:2357-2374
      subcell_state(c, d, s, SubCell{p->t_weight_A[gi], p->t_weight_B[gi], p->t_bias[gi]}, c0, ubuf);
      transpose(c, p->t_mlp_w[gi], nullptr, wg, d.F, d.N, 1, 1, 0, 0, 0, 0);                       // index f*N+n -> n*F+f
      gate_logit_k<<<(unsigned)std::min<long long>(d.TB, 148 * 8), TPB, 0, st>>>(ubuf, wg, p->t_mlp_b[gi], s.gt + gi * d.TB, d.TB, d.NF);
      check_launch();
    }
  if (d.node)
    for (int gi = 0; gi < 2; ++gi) {                                                             // graphML.py:2379-2399
      subcell_state(c, d, s, SubCell{p->n_weight_A[gi], p->n_weight_B[gi], p->n_bias[gi]}, c0, ubuf);
      node_head_fwd(c, d, ubuf, p->n_head_w[gi], p->n_head_b[gi], pbuf, v1, v2, s.qn + gi * d.TB * d.N);
    }
  // ---- the recurrence (graphML.py:2351-2427) ------------------------------------------------------------
  for (long long t = 0; t < d.T; ++t) {
    const float* hprev = t == 0 ? s.h0n : s.Hn + (t - 1) * d.B * d.NF;
    shift_chain(c, hprev, zh, d.E, d.Kst, d.F, d.B);
    contract_fwd(c, chain_slabs(hprev, zh, d.E, d.Kst, d.B * d.NF), p->weight_B, p->bias, 1.f, nullptr, 1, ur, d.B, d.F, d.E * d.Kst, d.F);
    contract_fwd(c, chain_slabs(s.Xn, s.zx, d.E, d.Kin, d.TB * d.NG, t * d.B * d.NG), p->weight_A, p->bias, 1.f, nullptr, 1, ua,
                 d.B, d.F, d.E * d.Kin, d.G);
    const float *va = ua, *vr = ur;
    if (d.edge) {
      gat_fwd_nm(c, p->e_mixer[0], p->e_weight[0], ua, qa, gb, d.F, d.F, d.B);
      gat_fwd_nm(c, p->e_mixer[1], p->e_weight[1], ur, qr, gb, d.F, d.F, d.B);
      va = qa; vr = qr;
    }
    combine_fwd_k<<<grid1d(d.B * d.NF, TPB), TPB, 0, st>>>(
        va, vr, d.tg ? s.gt + t * d.B : nullptr, d.tg ? s.gt + d.TB + t * d.B : nullptr,
        d.node ? s.qn + t * d.B * d.N : nullptr, d.node ? s.qn + d.TB * d.N + t * d.B * d.N : nullptr,
        s.Hn + t * d.B * d.NF, d.B, d.N, d.F);
    check_launch();
  }
  transpose(c, s.Hn, nullptr, H, d.N, d.F, d.T, d.B, d.B * d.NF, d.NF, d.NF, d.T * d.NF);           // [T,B,N,F] -> [B,T,F,N]
  return a.off;
}

size_t cell_backward_f32(const gcrnn_cell* cell, const gcrnn_cell_params* p, const float* X, const float* h0,
                         const float* H, const float* dH, const void* saved, size_t savedb,
                         const gcrnn_cell_params* gr, float* dX, float* dh0, void* ws, size_t wsb, int64_t B,
                         int64_t T, cudaStream_t st) {
  (void)X; (void)h0; (void)H;
  const CellDims d = dims_of(cell, B, T);
  {
    // backward runs on the path whose saved state the forward wrote (the autograd glue forces it).  NODE32 has no dX: the
    // generic sweep then runs on the generic prefix of what NODE32's forward saved.
    const int path = pick_path(cell);
    if (path == GCRNN_PATH_NODE32 && !dX) return cell_backward_e32(cell, p, dH, saved, savedb, gr, dh0, ws, wsb, B, T, st);
    GCRNN_CHECK(!(path == GCRNN_PATH_NODE32 && ws != nullptr && cell->fwd_reordered),
                "dX requested from a forward that ran on the renumbered graph: set the cell option \"need_dx\" before the forward");
    if (path == GCRNN_PATH_PERSIST) {
      GCRNN_CHECK(ws == nullptr || dX == nullptr, "the persistent small-graph path does not produce dX");
      return cell_backward_persist(cell, p, X, h0, H, dH, saved, savedb, gr, dh0, ws, B, T, st);
    }
  }
  Arena a(ws, wsb);
  Ctx c{cell->g, st, a.dry()};
  const gcrnn_graph* g = cell->g;
  Saved s;
  { Arena sa(const_cast<void*>(saved), savedb); s.layout(sa, d); }
  GCRNN_CHECK(a.dry() || saved, "backward needs the buffer written by forward");
  const int SA = d.E * d.Kin, SB = d.E * d.Kst;
  const bool need_dx = dX != nullptr;
  // scratch
  float* zh = a.get<float>((size_t)d.E * (d.Kst - 1) * d.B * d.NF);
  float* ua = a.get<float>(d.B * d.NF);
  float* ur = a.get<float>(d.B * d.NF);
  float* dht = a.get<float>(d.B * d.NF);
  float* dhn = a.get<float>(d.B * d.NF);
  float* da = a.get<float>(d.B * d.NF);
  float* dr = a.get<float>(d.B * d.NF);
  float* dzh = a.get<float>((size_t)SB * d.B * d.NF);
  GatBufs gba, gbr; GatBwdBufs gw; float *qa = nullptr, *qr = nullptr, *da2 = nullptr, *dr2 = nullptr;
  if (d.edge) {
    gba.alloc(a, d.B, d.N, d.F, g->nnz_att); gbr.alloc(a, d.B, d.N, d.F, g->nnz_att); gw.alloc(a, d.B, d.N, d.F, g->nnz_att);
    qa = a.get<float>(d.B * d.NF); qr = a.get<float>(d.B * d.NF); da2 = a.get<float>(d.B * d.NF); dr2 = a.get<float>(d.B * d.NF);
  }
  float* dzx = need_dx ? a.get<float>((size_t)SA * d.TB * d.NG) : nullptr;
  float* dxn = need_dx ? a.get<float>(d.TB * d.NG) : nullptr;
  float *dgt = nullptr, *dqn = nullptr, *ubuf = nullptr, *c0 = nullptr, *dc0 = nullptr, *wg = nullptr, *dwg = nullptr,
        *dl = nullptr, *dzh0 = nullptr, *pbuf = nullptr;
  if (d.tg) { dgt = a.get<float>(2 * d.TB); wg = a.get<float>(d.NF); dwg = a.get<float>(d.NF); dl = a.get<float>(d.TB); }
  if (d.node) { dqn = a.get<float>(2 * d.TB * d.N); pbuf = a.get<float>((size_t)SB * d.TB * d.N); }
  if (d.gates) {
    ubuf = a.get<float>(d.TB * d.NF); c0 = a.get<float>(d.B * d.NF); dc0 = a.get<float>(d.B * d.NF);
    dzh0 = a.get<float>((size_t)SB * d.B * d.NF);
  }
  float* zslab = cell->dh_last_only ? a.get<float>(d.NF) : nullptr;
  if (a.dry()) return a.off;
  if (zslab) zero(c, zslab, d.NF * sizeof(float));
  const DhView dv{dH, zslab, d.T, d.NF, cell->dh_last_only != 0};

  zero(c, dhn, d.B * d.NF * sizeof(float));
  if (d.tg) zero(c, dgt, 2 * d.TB * sizeof(float));

  // ---- reverse-time sweep -----------------------------------------------------------------------------
  for (long long t = d.T - 1; t >= 0; --t) {
    const float* hprev = t == 0 ? s.h0n : s.Hn + (t - 1) * d.B * d.NF;
    const Slabs zhs = chain_slabs(hprev, zh, d.E, d.Kst, d.B * d.NF);
    const Slabs zxs = chain_slabs(s.Xn, s.zx, d.E, d.Kin, d.TB * d.NG, t * d.B * d.NG);
    // recompute the step's filter outputs from the stored states
    shift_chain(c, hprev, zh, d.E, d.Kst, d.F, d.B);
    contract_fwd(c, zhs, p->weight_B, p->bias, 1.f, nullptr, 1, ur, d.B, d.F, SB, d.F);
    contract_fwd(c, zxs, p->weight_A, p->bias, 1.f, nullptr, 1, ua, d.B, d.F, SA, d.G);
    const float *va = ua, *vr = ur;
    if (d.edge) {
      gat_fwd_nm(c, p->e_mixer[0], p->e_weight[0], ua, qa, gba, d.F, d.F, d.B);
      gat_fwd_nm(c, p->e_mixer[1], p->e_weight[1], ur, qr, gbr, d.F, d.F, d.B);
      va = qa; vr = qr;
    }
    // dh_t = dH[:, t] (reference layout) + what flowed back from step t+1
    transpose(c, dv.ptr(t), dhn, dht, d.F, d.N, d.B, 1, dv.bstride(t), 0, d.NF, 0);
    {
      dim3 grid((d.N + 127) / 128, (unsigned)std::min<long long>(d.B, 32768));
      combine_bwd_k<<<grid, 128, 0, st>>>(dht, s.Hn + t * d.B * d.NF, va, vr,
                                          d.tg ? s.gt + t * d.B : nullptr, d.tg ? s.gt + d.TB + t * d.B : nullptr,
                                          d.node ? s.qn + t * d.B * d.N : nullptr, d.node ? s.qn + d.TB * d.N + t * d.B * d.N : nullptr,
                                          da, dr, d.tg ? dgt + t * d.B : nullptr, d.tg ? dgt + d.TB + t * d.B : nullptr,
                                          d.node ? dqn + t * d.B * d.N : nullptr, d.node ? dqn + d.TB * d.N + t * d.B * d.N : nullptr,
                                          d.B, d.N, d.F);
      check_launch();
    }
    const float *dua = da, *dur = dr;
    if (d.edge) {
      gat_bwd_nm(c, p->e_mixer[0], p->e_weight[0], ua, qa, da, gba, gw, da2, gr->e_mixer[0], gr->e_weight[0], d.F, d.F, d.B);
      gat_bwd_nm(c, p->e_mixer[1], p->e_weight[1], ur, qr, dr, gbr, gw, dr2, gr->e_mixer[1], gr->e_weight[1], d.F, d.F, d.B);
      dua = da2; dur = dr2;
    }
    contract_wgrad(c, zxs, dua, gr->weight_A, d.B * d.N, d.F, SA, d.G);
    contract_wgrad(c, zhs, dur, gr->weight_B, d.B * d.N, d.F, SB, d.F);
    colsum(c, dua, gr->bias, d.B * d.N, d.F);
    colsum(c, dur, gr->bias, d.B * d.N, d.F);
    contract_bwd_data(c, mut_slabs(dzh, SB, d.B * d.NF), p->weight_B, dur, d.B * d.N, d.F, SB, d.F, 0);
    reverse_chain(c, dzh, d.E, d.Kst, d.F, d.B, dhn, false);
    if (need_dx) contract_bwd_data(c, mut_slabs(dzx, SA, d.TB * d.NG, t * d.B * d.NG), p->weight_A, dua, d.B * d.N, d.F, SA, d.G, 0);
  }

  // ---- gates: batched over all (t, b); they depend on (x_t, h0) only ------------------------------------
  bool first_h0 = true;
  auto subcell_bwd = [&](const SubCell& sc, float* gA, float* gB, float* gb) {
    // ubuf holds d(pre-activation) of the sub-cell state for every (t, b)
    contract_wgrad(c, chain_slabs(s.Xn, s.zx, d.E, d.Kin, d.TB * d.NG), ubuf, gA, d.TB * d.N, d.F, SA, d.G);
    colsum(c, ubuf, gb, d.TB * d.N, d.F, 2.f);                    // the sub-cell adds its bias twice (graphML.py:2421-2422)
    reduce_t_k<<<grid1d(d.B * d.NF, TPB), TPB, 0, st>>>(ubuf, dc0, d.T, d.B, d.NF);
    check_launch();
    contract_wgrad(c, chain_slabs(s.h0n, s.zh0, d.E, d.Kst, d.B * d.NF), dc0, gB, d.B * d.N, d.F, SB, d.F);
    if (need_dx) contract_bwd_data(c, mut_slabs(dzx, SA, d.TB * d.NG), sc.A, ubuf, d.TB * d.N, d.F, SA, d.G, 1);
    if (dh0) { contract_bwd_data(c, mut_slabs(dzh0, SB, d.B * d.NF), sc.Bw, dc0, d.B * d.N, d.F, SB, d.F, first_h0 ? 0 : 1); first_h0 = false; }
  };
  if (d.tg)
    for (int gi = 0; gi < 2; ++gi) {
      const SubCell sc{p->t_weight_A[gi], p->t_weight_B[gi], p->t_bias[gi]};
      subcell_state(c, d, s, sc, c0, ubuf);
      transpose(c, p->t_mlp_w[gi], nullptr, wg, d.F, d.N, 1, 1, 0, 0, 0, 0);
      gate_dlogit_k<<<1, 1024, 0, st>>>(dgt + gi * d.TB, s.gt + gi * d.TB, dl, gr->t_mlp_b[gi], d.TB);
      check_launch();
      zero(c, dwg, d.NF * sizeof(float));
      {
        dim3 grid((unsigned)std::min<long long>((d.NF + TPB - 1) / TPB, 148 * 8), (unsigned)std::max<long long>(1, std::min<long long>(d.TB / 64, 64)));
        gate_du_k<<<grid, TPB, 0, st>>>(ubuf, wg, dl, dwg, d.TB, d.NF);
        check_launch();
      }
      if (gr->t_mlp_w[gi]) transpose(c, dwg, gr->t_mlp_w[gi], gr->t_mlp_w[gi], d.N, d.F, 1, 1, 0, 0, 0, 0);
      subcell_bwd(sc, gr->t_weight_A[gi], gr->t_weight_B[gi], gr->t_bias[gi]);
    }
  if (d.node)
    for (int gi = 0; gi < 2; ++gi) {
      const SubCell sc{p->n_weight_A[gi], p->n_weight_B[gi], p->n_bias[gi]};
      const long long RN = d.TB * d.N;
      subcell_state(c, d, s, sc, c0, ubuf);
      // d(lin) = dq q (1-q); dp_{e,0} = dlin, dp_{e,k} = dp_{e,k-1} @ S_e^T
      dsigmoid_k<<<grid1d(RN, TPB), TPB, 0, st>>>(dqn + gi * RN, s.qn + gi * RN, pbuf, gr->n_head_b[gi], RN);
      check_launch();
      for (int e = 0; e < d.E; ++e) {
        float* p0 = pbuf + (long long)(e * d.Kst) * RN;
        if (e > 0) copy(c, p0, pbuf, RN * sizeof(float));
        for (int kk = 1; kk < d.Kst; ++kk) spmm(c, g->bwd[e], p0 + (long long)(kk - 1) * RN, nullptr, p0 + (long long)kk * RN, 1, d.TB);
      }
      {
        float* dw = gr->n_head_w[gi];
        GCRNN_CHECK(dw, "node-gate head gradient buffer missing");
        node_proj_bwd_k<<<grid1d(RN, 128, 148 * 4), 128, (size_t)2 * SB * d.F * sizeof(float), st>>>(ubuf, p->n_head_w[gi], pbuf, dw, RN, d.F, SB);
        check_launch();
      }
      subcell_bwd(sc, gr->n_weight_A[gi], gr->n_weight_B[gi], gr->n_bias[gi]);
    }
  // ---- input gradients ------------------------------------------------------------------------------------
  if (dh0) {
    if (d.gates) reverse_chain(c, dzh0, d.E, d.Kst, d.F, d.B, dhn, true);
    transpose(c, dhn, nullptr, dh0, d.N, d.F, d.B, 1, d.NF, 0, d.NF, 0);
  }
  if (need_dx) {
    reverse_chain(c, dzx, d.E, d.Kin, d.G, d.TB, dxn, false);
    transpose(c, dxn, nullptr, dX, d.N, d.G, d.T, d.B, d.B * d.NG, d.NG, d.NG, d.T * d.NG);          // [T,B,N,G] -> [B,T,G,N]
  }
  return a.off;
}

}  // namespace gcrnn
