// extern "C" surface of libgcrnn_b200.so (see include/gcrnn_b200.h).  Exceptions stop here.
#include "common.cuh"
#include <dlfcn.h>
#include <cstring>

#include <atomic>
#include <string>

namespace gcrnn {
static thread_local char g_err[1024] = "";
// monotonic launch statistic (gcrnn_debug_launch_count); internal linkage, atomic
static std::atomic<unsigned long long> s_launches{0};
unsigned long long launch_count() { return s_launches.load(std::memory_order_relaxed); }
void count_launch(unsigned long long n) { s_launches.fetch_add(n, std::memory_order_relaxed); }
void launch_count_rewind(unsigned long long value) { s_launches.store(value, std::memory_order_relaxed); }
// options of the handle whose API call is executing on this thread (call-scoped; the defaults are immutable)
static const Options k_default_options{};
static thread_local const Options* t_opt = &k_default_options;
const Options& opt() { return *t_opt; }
OptScope::OptScope(const Options* o) : prev(t_opt) { t_opt = o ? o : &k_default_options; }
OptScope::~OptScope() { t_opt = prev; }
int* option_field(Options& o, const char* name) {
  const std::string n(name ? name : "");
  if (n == "bwd_fused") return &o.bwd_fused;
  if (n == "sparse_fused") return &o.sparse_fused;
  if (n == "sparse_v2") return &o.sparse_v2;
  if (n == "sparse_v2_rows_bps") return &o.sparse_v2_rows_bps;
  if (n == "sparse_v2_fuse_dpre") return &o.sparse_v2_fuse_dpre;
  if (n == "sparse_v2_tc") return &o.sparse_v2_tc;
  if (n == "sparse_v2_bps") return &o.sparse_v2_bps;
  if (n == "graph_capture") return &o.graph_capture;
  if (n == "gate_fq8") return &o.gate_fq8;
  if (n == "gemm_pair") return &o.gemm_pair;
  if (n == "fwd_fused") return &o.fwd_fused;
  if (n == "persist") return &o.persist;
  return nullptr;
}
void set_last_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
gcrnn_graph* graph_create_csr(int N, int E, const int64_t* const* rowptr, const int32_t* const* colidx,
                              const float* const* vals, int device);
gcrnn_graph* graph_create_dense(int N, int E, const float* S, int keep_dense, int device);
void graph_destroy(gcrnn_graph* g);
}  // namespace gcrnn

using namespace gcrnn;

#define API_BEGIN try {
#define API_END                                                                     \
  } catch (const gcrnn::Error& e) { set_last_error("%s", e.what()); return e.code; } \
    catch (const std::exception& e) { set_last_error("%s", e.what()); return -1; }   \
    catch (...) { set_last_error("unknown error"); return -1; }                      \
  return 0;

// ---- CUDA-graph replay of small fp32 cell calls ------------------------------------------------------------------------
// The reference's own configurations (N = 59 / 80 nodes) are launch-bound: a forward+backward is
// 150-1000 tiny kernels.  A forward (or backward) call is a pure function of its pointer set and sizes, and PyTorch's caching
// allocator hands a training loop the same pointers step after step, so the launch sequence is captured once per key into a
// CUDA graph and replayed afterwards.  Capture runs on a side stream (torch's default stream is the legacy stream, which cannot
// be captured) fenced by events against the caller's stream; a key miss simply captures again (8 entries, LRU).
namespace {
constexpr int GK_PTRS = 72;
struct GraphKey {
  const void* p[GK_PTRS]; int64_t B, T; int kind, path, need_dx, pad; int64_t epoch;   // epoch: bumped by every option change on the cell
  bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; unsigned long long launches; uint64_t stamp; int last_path; };
struct GraphCache {
  std::vector<GraphEntry> entries;
  cudaStream_t side = nullptr; cudaEvent_t ev_in = nullptr, ev_out = nullptr; uint64_t clock = 0;
  ~GraphCache() {
    for (auto& e : entries) cudaGraphExecDestroy(e.exec);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (side) cudaStreamDestroy(side);
  }
};
void key_params(GraphKey& k, int& n, const gcrnn_cell_params* p) {
  static_assert(sizeof(gcrnn_cell_params) % sizeof(void*) == 0, "parameter block is an array of pointers");
  const void* const* q = reinterpret_cast<const void* const*>(p);
  for (size_t i = 0; i < sizeof(gcrnn_cell_params) / sizeof(void*); ++i) { GCRNN_CHECK(n < GK_PTRS, "graph key overflow"); k.p[n++] = q[i]; }
}
bool graph_eligible(const gcrnn_cell* c, int64_t B, int64_t T, cudaStream_t user) {
  // a caller that is itself capturing (whole-step CUDA graph: gated_gcrnns_b200/train.py) gets plain launches on its stream
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(user, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cs != cudaStreamCaptureStatusNone) return false;
  return c->opt.graph_capture && c->d.precision == GCRNN_PREC_FP32 &&
         (long long)B * T * c->g->N * c->d.F <= (1ll << 22);
}
// run `body(stream)` through the cache; body only enqueues work on the stream it is given
template <class Body>
void run_graphed(const gcrnn_cell* c, const GraphKey& key, cudaStream_t user, Body&& body) {
  if (!c->graph_cache) c->graph_cache = new GraphCache();
  GraphCache& gc = *static_cast<GraphCache*>(c->graph_cache);
  if (!gc.side) {
    CUDA_OK(cudaStreamCreateWithFlags(&gc.side, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&gc.ev_in, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&gc.ev_out, cudaEventDisableTiming));
  }
  CUDA_OK(cudaEventRecord(gc.ev_in, user));
  CUDA_OK(cudaStreamWaitEvent(gc.side, gc.ev_in, 0));
  GraphEntry* hit = nullptr;
  for (auto& e : gc.entries) if (e.key == key) { hit = &e; break; }
  if (!hit) {
    const unsigned long long l0 = launch_count();
    CUDA_OK(cudaStreamBeginCapture(gc.side, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    try {
      body(gc.side);
    } catch (...) {
      cudaStreamEndCapture(gc.side, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      throw;
    }
    CUDA_OK(cudaStreamEndCapture(gc.side, &graph));
    GraphEntry e{key, nullptr, launch_count() - l0, 0, c->last_path};
    launch_count_rewind(l0);                          // counted again at every replay, including the first one below
    cudaError_t st = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OK(st);
    if (gc.entries.size() >= 8) {
      size_t old = 0;
      for (size_t i = 1; i < gc.entries.size(); ++i) if (gc.entries[i].stamp < gc.entries[old].stamp) old = i;
      cudaGraphExecDestroy(gc.entries[old].exec);
      gc.entries.erase(gc.entries.begin() + old);
    }
    gc.entries.push_back(e);
    hit = &gc.entries.back();
  }
  hit->stamp = ++gc.clock;
  if (key.kind == 0) c->last_path = hit->last_path;
  CUDA_OK(cudaGraphLaunch(hit->exec, gc.side));
  count_launch(hit->launches);
  CUDA_OK(cudaEventRecord(gc.ev_out, gc.side));
  CUDA_OK(cudaStreamWaitEvent(user, gc.ev_out, 0));
}
}  // namespace

extern "C" {

int gcrnn_abi_version(void) { return GCRNN_ABI_VERSION; }
const char* gcrnn_last_error(void) { return g_err; }
uint64_t gcrnn_debug_launch_count(void) { return launch_count(); }
int gcrnn_graph_set_option(gcrnn_graph* g, const char* name, int32_t value) {
  API_BEGIN
  GCRNN_CHECK(g && name, "null argument");
  if (std::strcmp(name, "reorder") == 0) {          // node renumbering for the fused sparse path (graph.cu: locality_view)
    GCRNN_CHECK(value >= 0 && value <= 2, "reorder: 0 never, 1 when it pays, 2 always");
    g->reorder_mode = value;
    if (g->reorder_state == 2) g->reorder_state = 0;
    return 0;
  }
  int* f = option_field(g->opt, name);
  GCRNN_CHECK(f != nullptr, "unknown option '%s'", name);
  *f = value; ++g->opt.epoch;
  API_END
}
int gcrnn_debug_edge_relu_masks(const gcrnn_cell* cell, const void* saved, size_t saved_bytes, int64_t B, int64_t T, uint8_t* out, void* stream) {
  API_BEGIN
  GCRNN_CHECK(cell && saved && out, "null argument");
  DeviceScope dev(cell->g->device);
  debug_edge_relu_masks(cell, saved, saved_bytes, B, T, out, (cudaStream_t)stream);
  API_END
}
int gcrnn_graph_get_option(const gcrnn_graph* g, const char* name, int32_t* value) {
  API_BEGIN
  GCRNN_CHECK(g && name && value, "null argument");
  if (std::strcmp(name, "reorder") == 0) *value = g->reorder_mode;
  else if (std::strcmp(name, "reordered") == 0) *value = locality_view(g) != g;      // evaluates the ordering if not done yet
  else if (std::strcmp(name, "tile_rows_before_x100") == 0) { locality_view(g); *value = (int32_t)std::lround(100.f * g->tile_rows[0]); }
  else if (std::strcmp(name, "tile_rows_after_x100") == 0) { locality_view(g); *value = (int32_t)std::lround(100.f * g->tile_rows[1]); }
  else {
    const int* f = option_field(const_cast<gcrnn::Options&>(g->opt), name);
    GCRNN_CHECK(f != nullptr, "unknown option '%s'", name);
    *value = *f;
  }
  API_END
}

int gcrnn_graph_create_csr(gcrnn_graph** out, int32_t N, int32_t E, const int64_t* const* rowptr,
                           const int32_t* const* colidx, const float* const* vals, int32_t device) {
  API_BEGIN
  GCRNN_CHECK(out && rowptr && colidx && vals, "null argument");
  *out = graph_create_csr(N, E, rowptr, colidx, vals, device);
  API_END
}
int gcrnn_graph_create_dense(gcrnn_graph** out, int32_t N, int32_t E, const float* S, int32_t keep_dense, int32_t device) {
  API_BEGIN
  GCRNN_CHECK(out && S, "null argument");
  *out = graph_create_dense(N, E, S, keep_dense, device);
  API_END
}
int gcrnn_graph_destroy(gcrnn_graph* g) {
  API_BEGIN
  graph_destroy(g);
  API_END
}
int gcrnn_graph_info(const gcrnn_graph* g, int32_t* N, int32_t* E, int64_t* nnz, int64_t* nnz_att) {
  API_BEGIN
  GCRNN_CHECK(g, "null graph");
  if (N) *N = g->N;
  if (E) *E = g->E;
  if (nnz) *nnz = g->nnz_total;
  if (nnz_att) *nnz_att = g->nnz_att;
  API_END
}

// ---- LSIGF -------------------------------------------------------------------------------------------
size_t gcrnn_lsigf_workspace_bytes(const gcrnn_graph* g, int32_t F, int32_t K, int32_t G, int64_t B) {
  try {
    size_t f = lsigf_forward_f32(g, nullptr, nullptr, nullptr, nullptr, F, K, G, B, nullptr, 0, nullptr);
    size_t b = lsigf_backward_f32(g, nullptr, nullptr, nullptr, (float*)1, (float*)1, (float*)1, F, K, G, B, nullptr, 0, nullptr);
    return (f > b ? f : b) + 256;
  } catch (const std::exception& e) { set_last_error("%s", e.what()); return 0; }
}
int gcrnn_lsigf_forward(const gcrnn_graph* g, const float* h, const float* bias, const float* x, float* y, int32_t F,
                        int32_t K, int32_t G, int64_t B, void* ws, size_t wsb, void* stream) {
  API_BEGIN
  GCRNN_CHECK(g && h && x && y && ws, "null argument");
  GCRNN_CHECK(F > 0 && K > 0 && G > 0 && B > 0, "bad sizes F=%d K=%d G=%d B=%lld", F, K, G, (long long)B);
  DeviceScope dev_scope(g->device); OptScope opt_scope(&g->opt);
  lsigf_forward_f32(g, h, bias, x, y, F, K, G, B, ws, wsb, (cudaStream_t)stream);
  API_END
}
int gcrnn_lsigf_backward(const gcrnn_graph* g, const float* h, const float* x, const float* dy, float* dx, float* dh,
                         float* dbias, int32_t F, int32_t K, int32_t G, int64_t B, void* ws, size_t wsb, void* stream) {
  API_BEGIN
  GCRNN_CHECK(g && h && x && dy && ws, "null argument");
  DeviceScope dev_scope(g->device); OptScope opt_scope(&g->opt);
  lsigf_backward_f32(g, h, x, dy, dx, dh, dbias, F, K, G, B, ws, wsb, (cudaStream_t)stream);
  API_END
}

// ---- attention ---------------------------------------------------------------------------------------
size_t gcrnn_gat_workspace_bytes(const gcrnn_graph* g, int32_t F, int32_t G, int64_t B) {
  try {
    return gat_backward_f32(g, nullptr, nullptr, nullptr, nullptr, (float*)1, nullptr, nullptr, F, G, B, nullptr, 0, nullptr) + 256;
  } catch (const std::exception& e) { set_last_error("%s", e.what()); return 0; }
}
int gcrnn_gat_forward(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x, float* y, int32_t F,
                      int32_t G, int64_t B, void* ws, size_t wsb, void* stream) {
  API_BEGIN
  GCRNN_CHECK(g && mixer && weight && x && y && ws, "null argument");
  DeviceScope dev_scope(g->device); OptScope opt_scope(&g->opt);
  gat_forward_f32(g, mixer, weight, x, y, F, G, B, ws, wsb, (cudaStream_t)stream);
  API_END
}
int gcrnn_gat_backward(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x, const float* dy,
                       float* dx, float* dmixer, float* dweight, int32_t F, int32_t G, int64_t B, void* ws, size_t wsb,
                       void* stream) {
  API_BEGIN
  GCRNN_CHECK(g && mixer && weight && x && dy && ws, "null argument");
  DeviceScope dev_scope(g->device); OptScope opt_scope(&g->opt);
  gat_backward_f32(g, mixer, weight, x, dy, dx, dmixer, dweight, F, G, B, ws, wsb, (cudaStream_t)stream);
  API_END
}

// ---- cell ----------------------------------------------------------------------------------------------
int gcrnn_cell_create(gcrnn_cell** out, const gcrnn_cell_desc* d, const gcrnn_graph* g) {
  API_BEGIN
  GCRNN_CHECK(out && d && g, "null argument");
  GCRNN_CHECK(d->G > 0 && d->F > 0 && d->Kin > 0 && d->Kst > 0, "bad cell sizes");
  GCRNN_CHECK(d->E == g->E, "cell E=%d does not match graph E=%d", d->E, g->E);
  GCRNN_CHECK(d->spatial_gating >= 0 && d->spatial_gating <= 2, "bad spatial_gating %d", d->spatial_gating);
  if (d->precision == GCRNN_PREC_BF16_TC || d->precision == GCRNN_PREC_BF16X2_TC) {
    GCRNN_CHECK(g->S_bf16 != nullptr, "tensor-core precision needs a graph created with keep_dense != 0");
    GCRNN_CHECK(d->spatial_gating != GCRNN_SPATIAL_EDGE, "tensor-core path: time and node gating only; use fp32 for edge gating");
  } else {
    GCRNN_CHECK(d->precision == GCRNN_PREC_FP32, "unknown precision %d", d->precision);
  }
  auto* c = new gcrnn_cell();
  c->d = *d; c->g = g;
  *out = c;
  API_END
}
int gcrnn_cell_destroy(gcrnn_cell* c) {
  API_BEGIN
  if (c) { DeviceScope dev_scope(c->g->device); delete static_cast<GraphCache*>(c->graph_cache); }
  delete c;
  API_END
}
int gcrnn_cell_set_option(gcrnn_cell* c, const char* name, int32_t value) {
  API_BEGIN
  GCRNN_CHECK(c && name, "null argument");
  const std::string n(name);
  if (n == "need_dx") c->need_dx = value != 0;
  else if (n == "dh_last_only") c->dh_last_only = value != 0;
  else if (n == "path") { GCRNN_CHECK(value >= -1 && value <= GCRNN_PATH_PERSIST, "bad path %d", value); c->forced_path = value; }
  else {
    int* f = option_field(c->opt, name);
    GCRNN_CHECK(f != nullptr, "unknown cell option '%s'", name);
    if (*f != value) { *f = value; ++c->opt.epoch; }       // captured CUDA graphs were built under the old options
  }
  API_END
}
int gcrnn_cell_get_option(const gcrnn_cell* c, const char* name, int32_t* value) {
  API_BEGIN
  GCRNN_CHECK(c && name && value, "null argument");
  const std::string n(name);
  if (n == "need_dx") *value = c->need_dx;
  else if (n == "dh_last_only") *value = c->dh_last_only;
  else if (n == "path") *value = c->forced_path;
  else if (n == "last_path") *value = c->last_path;
  else {
    const int* f = option_field(const_cast<gcrnn_cell*>(c)->opt, name);
    GCRNN_CHECK(f != nullptr, "unknown cell option '%s'", name);
    *value = *f;
  }
  API_END
}
int gcrnn_cell_workspace_bytes(const gcrnn_cell* c, int64_t B, int64_t T, int32_t need_input_grads, size_t* saved_bytes,
                               size_t* fwd_bytes, size_t* bwd_bytes) {
  API_BEGIN
  GCRNN_CHECK(c, "null cell");
  size_t su = 0, f, b;
  float* flag = need_input_grads ? (float*)1 : nullptr;
  OptScope opt_scope(&c->opt);
  if (c->d.precision != GCRNN_PREC_FP32) {
    f = cell_forward_tc(c, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &su, nullptr, 0, B, T, nullptr);
    b = cell_backward_tc(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, flag, flag, nullptr, 0, B, T, nullptr);
  } else {
    f = cell_forward_f32(c, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &su, nullptr, 0, B, T, nullptr);
    b = cell_backward_f32(c, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, flag, flag, nullptr, 0, B, T, nullptr);
  }
  if (saved_bytes) *saved_bytes = su + 256;
  if (fwd_bytes) *fwd_bytes = f + 256;
  if (bwd_bytes) *bwd_bytes = b + 256;
  API_END
}
int gcrnn_cell_forward(gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0, float* H, void* saved,
                       size_t savedb, void* ws, size_t wsb, int64_t B, int64_t T, void* stream) {
  API_BEGIN
  GCRNN_CHECK(c && p && X && h0 && H && ws && saved, "null argument");
  DeviceScope dev_scope(c->g->device); OptScope opt_scope(&c->opt);
  if (c->d.precision != GCRNN_PREC_FP32) cell_forward_tc(c, p, X, h0, H, saved, savedb, nullptr, ws, wsb, B, T, (cudaStream_t)stream);
  else if (graph_eligible(c, B, T, (cudaStream_t)stream)) {
    GraphKey key; memset(&key, 0, sizeof key);
    int n = 0;
    key_params(key, n, p);
    key.p[n++] = X; key.p[n++] = h0; key.p[n++] = H; key.p[n++] = saved; key.p[n++] = ws;
    key.p[n++] = (const void*)savedb; key.p[n++] = (const void*)wsb;
    key.B = B; key.T = T; key.kind = 0; key.path = c->forced_path; key.need_dx = c->need_dx; key.epoch = c->opt.epoch;
    run_graphed(c, key, (cudaStream_t)stream, [&](cudaStream_t st) { cell_forward_f32(c, p, X, h0, H, saved, savedb, nullptr, ws, wsb, B, T, st); });
  } else cell_forward_f32(c, p, X, h0, H, saved, savedb, nullptr, ws, wsb, B, T, (cudaStream_t)stream);
  API_END
}
int gcrnn_cell_backward(gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0, const float* H,
                        const float* dH, const void* saved, size_t savedb, const gcrnn_cell_params* grads, float* dX,
                        float* dh0, void* ws, size_t wsb, int64_t B, int64_t T, void* stream) {
  API_BEGIN
  GCRNN_CHECK(c && p && X && h0 && H && dH && saved && grads && ws, "null argument");
  DeviceScope dev_scope(c->g->device); OptScope opt_scope(&c->opt);
  if (c->d.precision != GCRNN_PREC_FP32) cell_backward_tc(c, p, X, h0, H, dH, saved, savedb, grads, dX, dh0, ws, wsb, B, T, (cudaStream_t)stream);
  else if (graph_eligible(c, B, T, (cudaStream_t)stream)) {
    GraphKey key; memset(&key, 0, sizeof key);
    int n = 0;
    key_params(key, n, p);
    key_params(key, n, grads);
    key.p[n++] = X; key.p[n++] = h0; key.p[n++] = H; key.p[n++] = dH; key.p[n++] = saved; key.p[n++] = dX; key.p[n++] = dh0; key.p[n++] = ws;
    key.p[n++] = (const void*)savedb; key.p[n++] = (const void*)wsb;
    key.B = B; key.T = T; key.kind = 1; key.path = c->forced_path; key.need_dx = dX != nullptr; key.pad = c->dh_last_only; key.epoch = c->opt.epoch;
    run_graphed(c, key, (cudaStream_t)stream,
                [&](cudaStream_t st) { cell_backward_f32(c, p, X, h0, H, dH, saved, savedb, grads, dX, dh0, ws, wsb, B, T, st); });
  } else cell_backward_f32(c, p, X, h0, H, dH, saved, savedb, grads, dX, dh0, ws, wsb, B, T, (cudaStream_t)stream);
  API_END
}

// ---- NCCL (resolved at run time so that the library shares the process's libnccl with torch) ------------
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (!api.lib) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) throw gcrnn::Error(-5, std::string("cannot load libnccl: ") + dlerror());
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce)
      throw gcrnn::Error(-5, "libnccl is missing symbols");
  }
  return api;
}
#define NCCL_OK(expr) do { int _r = (expr); if (_r != 0) throw gcrnn::Error(-6, std::string(#expr ": ") + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error")); } while (0)

struct gcrnn_comm { nccl_comm_t comm; int rank, world, device; };

int gcrnn_comm_unique_id(void* id128) {
  API_BEGIN
  GCRNN_CHECK(id128, "null argument");
  nccl_uid id; NCCL_OK(nccl().GetUniqueId(&id)); memcpy(id128, &id, 128);
  API_END
}
int gcrnn_comm_create(gcrnn_comm** out, const void* id128, int32_t rank, int32_t world, int32_t device) {
  API_BEGIN
  GCRNN_CHECK(out && id128, "null argument");
  DeviceScope dev_scope(device);
  nccl_uid id; memcpy(&id, id128, 128);
  auto* c = new gcrnn_comm{nullptr, rank, world, device};
  int r = nccl().CommInitRank(&c->comm, world, id, rank);
  if (r != 0) { delete c; NCCL_OK(r); }
  *out = c;
  API_END
}
int gcrnn_comm_destroy(gcrnn_comm* c) {
  API_BEGIN
  if (c) { nccl().CommDestroy(c->comm); delete c; }
  API_END
}
int gcrnn_allreduce_sum(gcrnn_comm* c, float* bucket, int64_t count, void* stream) {
  API_BEGIN
  GCRNN_CHECK(c && bucket && count >= 0, "bad argument");
  DeviceScope dev_scope(c->device);
  NCCL_OK(nccl().AllReduce(bucket, bucket, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->comm, (cudaStream_t)stream));
  API_END
}

}  // extern "C"
