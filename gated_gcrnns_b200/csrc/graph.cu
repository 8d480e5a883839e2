// Graph shift operator preprocessing: host CSR -> device gather forms (CSC for z@S, CSR for g@S^T),
// attention pattern of S+I, optional dense bf16 copies for the tensor-core path.
// Replaces what the reference keeps as the dense `S` attribute set by addGSO (Utils/graphML.py:1166-1173,
// :2074-2081, :2237-2244) and the `S + I` / mask construction of graphAttention (:577, :611-613).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace gcrnn {

template <class T>
static T* upload(gcrnn_graph* g, const std::vector<T>& v) {
  T* d = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CUDA_OK(cudaMalloc(&d, bytes));
  g->owned.push_back(d);
  if (!v.empty()) CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct HostCsr { std::vector<int> ptr, idx; std::vector<float> val; };
constexpr int REORDER_MIN_N = 4096;     // below this the whole signal sits in L1 / shared memory anyway

static HostCsr transpose_csr(int N, const HostCsr& a) {
  HostCsr t;
  size_t nnz = a.idx.size();
  t.ptr.assign(N + 1, 0); t.idx.resize(nnz); t.val.resize(nnz);
  for (size_t p = 0; p < nnz; ++p) t.ptr[a.idx[p] + 1]++;
  for (int i = 0; i < N; ++i) t.ptr[i + 1] += t.ptr[i];
  std::vector<int> cur(t.ptr.begin(), t.ptr.end() - 1);
  for (int i = 0; i < N; ++i)
    for (int p = a.ptr[i]; p < a.ptr[i + 1]; ++p) {
      int q = cur[a.idx[p]]++;
      t.idx[q] = i; t.val[q] = a.val[p];
    }
  return t;
}

static Gather to_device(gcrnn_graph* g, const HostCsr& c) {
  Gather d;
  d.nnz = (int64_t)c.idx.size();
  d.ptr = upload(g, c.ptr); d.idx = upload(g, c.idx); d.val = upload(g, c.val);
  return d;
}

static void build_attention_pattern(gcrnn_graph* g, const HostCsr& s) {
  // S' = S + I, keep |S'| > 1e-9; edges numbered in row order (graphML.py:577, 611-613)
  const int N = g->N;
  HostCsr a; a.ptr.assign(N + 1, 0);
  int maxdeg = 0;
  for (int i = 0; i < N; ++i) {
    std::vector<std::pair<int, float>> row;
    bool diag = false;
    for (int p = s.ptr[i]; p < s.ptr[i + 1]; ++p) {
      float v = s.val[p];
      if (s.idx[p] == i) { v += 1.0f; diag = true; }
      row.emplace_back(s.idx[p], v);
    }
    if (!diag) row.emplace_back(i, 1.0f);
    std::sort(row.begin(), row.end());
    // merge duplicates (a CSR input may repeat an entry)
    std::vector<std::pair<int, float>> m;
    for (auto& e : row) { if (!m.empty() && m.back().first == e.first) m.back().second += e.second; else m.push_back(e); }
    for (auto& e : m) if (std::fabs(e.second) > 1e-9f) { a.idx.push_back(e.first); a.val.push_back(e.second); }
    a.ptr[i + 1] = (int)a.idx.size();
    maxdeg = std::max(maxdeg, a.ptr[i + 1] - a.ptr[i]);
  }
  g->nnz_att = (int64_t)a.idx.size();
  g->max_row_deg = maxdeg;
  g->att_rptr = upload(g, a.ptr); g->att_col = upload(g, a.idx); g->att_val = upload(g, a.val);
  // column view carrying (row, edge id)
  std::vector<int> cptr(N + 1, 0), crow(a.idx.size()), ceid(a.idx.size());
  std::vector<float> cval(a.idx.size());
  for (size_t p = 0; p < a.idx.size(); ++p) cptr[a.idx[p] + 1]++;
  for (int i = 0; i < N; ++i) cptr[i + 1] += cptr[i];
  std::vector<int> cur(cptr.begin(), cptr.end() - 1);
  for (int i = 0; i < N; ++i)
    for (int p = a.ptr[i]; p < a.ptr[i + 1]; ++p) { int q = cur[a.idx[p]]++; crow[q] = i; ceid[q] = p; cval[q] = a.val[p]; }
  g->att_cptr = upload(g, cptr); g->att_crow = upload(g, crow); g->att_ceid = upload(g, ceid);
  g->att_cval = upload(g, cval);
}

gcrnn_graph* graph_from_host_csr(int N, int E, const std::vector<HostCsr>& ops, int device, bool keep_host = true) {
  GCRNN_CHECK(N > 0 && E > 0, "bad graph size N=%d E=%d", N, E);
  DeviceScope dev_scope(device);
  auto* g = new gcrnn_graph();
  g->N = N; g->E = E; g->device = device;
  try {
    if (E == 1 && keep_host) { g->h_ptr = ops[0].ptr; g->h_idx = ops[0].idx; g->h_val = ops[0].val; }
    for (int e = 0; e < E; ++e) {
      GCRNN_CHECK(ops[e].idx.size() < (size_t)INT32_MAX, "nnz too large");
      g->bwd.push_back(to_device(g, ops[e]));                    // CSR: rows i gather over j
      g->fwd.push_back(to_device(g, transpose_csr(N, ops[e])));  // CSC: columns j gather over i
      g->nnz_total += (int64_t)ops[e].idx.size();
    }
    if (E == 1) build_attention_pattern(g, ops[0]);
  } catch (...) {
    for (void* p : g->owned) cudaFree(p);
    delete g;
    throw;
  }
  return g;
}

gcrnn_graph* graph_create_csr(int N, int E, const int64_t* const* rowptr, const int32_t* const* colidx,
                              const float* const* vals, int device) {
  std::vector<HostCsr> ops(E);
  for (int e = 0; e < E; ++e) {
    int64_t nnz = rowptr[e][N];
    GCRNN_CHECK(nnz >= 0 && nnz < INT32_MAX, "nnz out of range");
    ops[e].ptr.resize(N + 1);
    for (int i = 0; i <= N; ++i) ops[e].ptr[i] = (int)rowptr[e][i];
    ops[e].idx.assign(colidx[e], colidx[e] + nnz);
    ops[e].val.assign(vals[e], vals[e] + nnz);
    for (int64_t p = 0; p < nnz; ++p) GCRNN_CHECK(ops[e].idx[p] >= 0 && ops[e].idx[p] < N, "column index out of range");
  }
  return graph_from_host_csr(N, E, ops, device);
}

gcrnn_graph* graph_create_dense(int N, int E, const float* S, int keep_dense, int device) {
  std::vector<HostCsr> ops(E);
  for (int e = 0; e < E; ++e) {
    HostCsr& c = ops[e];
    c.ptr.assign(N + 1, 0);
    const float* Se = S + (size_t)e * N * N;
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < N; ++j) {
        float v = Se[(size_t)i * N + j];
        if (v != 0.0f) { c.idx.push_back(j); c.val.push_back(v); }
      }
      c.ptr[i + 1] = (int)c.idx.size();
    }
  }
  gcrnn_graph* g = graph_from_host_csr(N, E, ops, device);
  if (keep_dense) {
    DeviceScope dev_scope(device);
    try {
      GCRNN_CHECK(E == 1, "the tensor-core path needs E == 1 (got %d)", E);
      tc_prepare_graph(g, S);
    } catch (...) {
      for (void* p : g->owned) cudaFree(p);
      delete g;
      throw;
    }
  }
  return g;
}

void graph_destroy(gcrnn_graph* g) {
  if (!g) return;
  DeviceScope dev_scope(g->device);
  if (g->reordered) graph_destroy(g->reordered);
  for (void* p : g->owned) cudaFree(p);
  delete g;
}

// ---- library-owned node reordering -------------------------------------------------------------------------------------------
// The fused sparse kernels process tiles of 128 consecutive nodes and live on L1 hits of the gathered neighbour rows
// (DESIGN.md 4a): what matters is how many DISTINCT neighbour rows a tile touches.  The reference fixes no node order
// (Utils/graphTools.py builds S in whatever order the data came), so for a graph whose numbering has no locality the library
// renumbers the nodes itself: tiles are grown as breadth-first balls of 128 nodes (seeded in the order a Cuthill-McKee sweep
// discovers them), which makes every tile a compact patch of the graph.  X / h0 / dH are gathered and H / dh0 scattered through
// `perm` at the boundary kernels that convert layouts anyway, so callers never see the internal numbering.
namespace {

// distinct neighbour rows per tile of `tile` consecutive rows, relative to the tile size (1 = perfect reuse)
float tile_rows_metric(int N, const std::vector<int>& ptr, const std::vector<int>& idx, int tile) {
  std::vector<int> stamp(N, -1);
  long long distinct = 0;
  for (int i = 0; i < N; ++i) {
    const int tl = i / tile;
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) if (stamp[idx[p]] != tl) { stamp[idx[p]] = tl; ++distinct; }
  }
  return (float)distinct / (float)N;
}

// new -> old order: breadth-first balls of `tile` nodes over the symmetrised pattern
std::vector<int> ball_order(int N, const HostCsr& a, const HostCsr& at, int tile) {
  std::vector<int> order; order.reserve(N);
  std::vector<char> taken(N, 0);
  std::vector<int> seeds; seeds.reserve(N);      // FIFO of nodes seen next to a finished ball (may contain taken nodes)
  size_t seed_head = 0;
  int scan = 0, fill = 0;
  std::vector<int> q; q.reserve(tile); size_t qh = 0;
  auto visit = [&](int u, auto&& on_neighbour) {
    for (int p = a.ptr[u]; p < a.ptr[u + 1]; ++p) on_neighbour(a.idx[p]);
    for (int p = at.ptr[u]; p < at.ptr[u + 1]; ++p) on_neighbour(at.idx[p]);
  };
  while ((int)order.size() < N) {
    int seed = -1;
    while (seed_head < seeds.size()) { const int c = seeds[seed_head++]; if (!taken[c]) { seed = c; break; } }
    if (seed < 0) { while (taken[scan]) ++scan; seed = scan; }
    taken[seed] = 1; q.clear(); qh = 0; q.push_back(seed); ++fill;
    while (qh < q.size()) {
      const int u = q[qh++];
      order.push_back(u);
      visit(u, [&](int v) {
        if (taken[v]) return;
        if (fill < tile) { taken[v] = 1; q.push_back(v); ++fill; } else seeds.push_back(v);
      });
    }
    if (fill >= tile) fill = 0;                    // ball complete; otherwise the pocket was smaller: next seed continues this tile
  }
  return order;
}

}  // namespace

const gcrnn_graph* locality_view(const gcrnn_graph* g) {
  if (g->reorder_mode == 0 || g->E != 1 || g->h_ptr.empty() || (g->reorder_mode == 1 && g->N < REORDER_MIN_N)) return g;
  if (g->reorder_state == 0) {
    const int N = g->N, TILE = 128;
    HostCsr a{g->h_ptr, g->h_idx, g->h_val};
    const HostCsr at = transpose_csr(N, a);
    const std::vector<int> order = ball_order(N, a, at, TILE);
    std::vector<int> inv(N);
    for (int i = 0; i < N; ++i) inv[order[i]] = i;
    HostCsr b; b.ptr.assign(N + 1, 0); b.idx.reserve(a.idx.size()); b.val.reserve(a.idx.size());
    std::vector<std::pair<int, float>> row;
    for (int i = 0; i < N; ++i) {
      const int o = order[i];
      row.clear();
      for (int p = a.ptr[o]; p < a.ptr[o + 1]; ++p) row.emplace_back(inv[a.idx[p]], a.val[p]);
      std::sort(row.begin(), row.end());
      for (auto& e : row) { b.idx.push_back(e.first); b.val.push_back(e.second); }
      b.ptr[i + 1] = (int)b.idx.size();
    }
    g->tile_rows[0] = tile_rows_metric(N, a.ptr, a.idx, TILE);
    g->tile_rows[1] = tile_rows_metric(N, b.ptr, b.idx, TILE);
    const bool pays = g->tile_rows[0] > 1.5f * g->tile_rows[1];
    if (g->reorder_mode == 2 || pays) {
      DeviceScope dev_scope(g->device);
      g->reordered = graph_from_host_csr(N, 1, std::vector<HostCsr>{b}, g->device, false);
      int* d = nullptr;
      CUDA_OK(cudaMalloc(&d, (size_t)2 * N * sizeof(int)));
      const_cast<gcrnn_graph*>(g)->owned.push_back(d);
      CUDA_OK(cudaMemcpy(d, order.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(d + N, inv.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice));
      g->perm = d; g->iperm = d + N;
      g->reordered->opt = g->opt;
      g->reorder_state = 1;
    } else {
      g->reorder_state = 2;
    }
  }
  return g->reorder_state == 1 ? g->reordered : g;
}

}  // namespace gcrnn
