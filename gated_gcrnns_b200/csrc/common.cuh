// Shared declarations of the gcrnn_b200 library (internal; the public surface is include/gcrnn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/gcrnn_b200.h"

namespace gcrnn {

// ---- error handling: C++ exceptions never cross the ABI; api.cu converts them to codes ----------------
void set_last_error(const char* fmt, ...);
// ---- tuning switches: they live ON THE HANDLE (gcrnn_cell_set_option / gcrnn_graph_set_option); the library keeps no
// process-wide mutable state.  While an API call executes, opt() returns the options of the handle it was made on.
struct Options {
  int bwd_fused = 1;            // fused reverse-time step kernel (tc_bwd.cuh) when the shape allows
  int sparse_fused = 1;         // fused F == 32 edge-gated sparse kernels (sp32_kernels.cuh) when the shape allows
  int sparse_v2 = 63;           // bit mask: which fused sparse stages run their second-generation kernel (sp32_tile.cuh)
  int sparse_v2_rows_bps = 2;   // resident blocks per SM of bwd_rows_v2_k (2 or 3)
  int sparse_v2_fuse_dpre = 1;  // the dh kernel finishes the next reverse step's dpre in its epilogue
  int sparse_v2_tc = 1;         // tile contractions on tensor cores (3xTF32 mma.sync), 0: packed FFMA2
  int sparse_v2_bps = 2;        // resident blocks per SM of the tile kernels (the rest of the 228 KB stays L1)
  int graph_capture = 1;        // replay small fp32 cell calls as CUDA graphs keyed by their pointer set (api.cu)
  int gate_fq8 = 2;             // time-gate kernels with 8 feature groups (512 threads): 1 = forward, 2 = forward + backward
  int gemm_pair = 1;            // CTA-pair (cta_group::2) shift GEMM when the shape allows
  int fwd_fused = 0;            // 1: forward tap contraction fused into the shift GEMMs (Horner form, tc_hshift.cuh); measured
                                // slower than chain + tap kernel in bf16 and 2 % faster in bf16x2, hence off by default
  int persist = 1;              // persistent fused recurrence (persist_f32.cuh) for small ungated / time-gated fp32 cells
  long long epoch = 0;          // bumped by every option change on the handle: captured CUDA graphs are keyed by it
};
const Options& opt();
struct OptScope {
  const Options* prev;
  explicit OptScope(const Options* o);
  ~OptScope();
};
int* option_field(Options& o, const char* name);     // nullptr for an unknown name

// kernels launched by this library (gcrnn_debug_launch_count): a monotonic statistic, atomically updated
unsigned long long launch_count();
void count_launch(unsigned long long n = 1);
void launch_count_rewind(unsigned long long value);   // CUDA-graph capture: captured launches are counted at replay time

// cudaFuncSetAttribute is per DEVICE: one-time configuration keyed by the current device
struct DeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    const unsigned long long old = __atomic_fetch_or(&mask, bit, __ATOMIC_RELAXED);
    return (old & bit) == 0;
  }
};
// the caller's current device is restored when an API call returns
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int device) {
    cudaGetDevice(&prev);
    if (prev != device) { cudaError_t e = cudaSetDevice(device); if (e != cudaSuccess) throw std::runtime_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    else prev = -1;
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define GCRNN_CHECK(cond, ...)                                                   \
  do { if (!(cond)) { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__);        \
       throw ::gcrnn::Error(-2, std::string(_b) + " [" #cond "]"); } } while (0)

#define CUDA_OK(expr)                                                            \
  do { cudaError_t _e = (expr); if (_e != cudaSuccess) {                         \
       throw ::gcrnn::Error(-3, std::string(#expr ": ") + cudaGetErrorString(_e)); } } while (0)

// ---- sparse operator on the device ------------------------------------------------------------------
// "gather form": out[dst] = sum_{p in ptr[dst]..ptr[dst+1]} val[p] * in[idx[p]]
struct Gather {
  int64_t nnz = 0;
  int* ptr = nullptr;   // [N+1]
  int* idx = nullptr;   // [nnz]
  float* val = nullptr; // [nnz]
};

}  // namespace gcrnn

struct gcrnn_graph {
  int N = 0, E = 0, device = 0;
  int64_t nnz_total = 0;
  std::vector<gcrnn::Gather> fwd;   // per e: CSC of S_e  (z @ S_e : gather sources i for each destination j)
  std::vector<gcrnn::Gather> bwd;   // per e: CSR of S_e  (g @ S_e^T: gather j for each destination i)
  // attention pattern of S' = S + I with |S'| > 1e-9 (E == 1 only), edges numbered in CSR (row i) order
  int64_t nnz_att = 0;
  int *att_rptr = nullptr, *att_col = nullptr;          // row i -> columns j
  float* att_val = nullptr;                             // S'_ij
  int *att_cptr = nullptr, *att_crow = nullptr, *att_ceid = nullptr;  // column j -> (row i, edge id)
  float* att_cval = nullptr;                            // S'_ij in column order (same order as att_crow)
  int max_row_deg = 0;
  // dense copies for the tensor-core path (E == 1): row-major [Npad, Npad] bf16, zero padded
  int Npad = 0;
  float dense_scale = 1.f;            // S_bf16 = bf16(S / dense_scale), dense_scale = max|S|
  __nv_bfloat16* S_bf16 = nullptr;    // S    (K-major B operand of the backward shift  g @ S^T)
  __nv_bfloat16* St_bf16 = nullptr;   // S^T  (K-major B operand of the forward shift   z @ S)
  int s_planes = 1;                   // operator planes kept, stacked [s_planes * N][N]: 2 when S / dense_scale is not exact in bf16
  std::vector<void*> owned;           // every device allocation, for destroy
  gcrnn::Options opt;                 // tuning switches used by calls made on the graph handle itself (gcrnn_graph_set_option)
  // Library-owned node reordering for the fused sparse path (graph.cu: locality_view).  The host CSR of a one-operator graph is
  // kept so that the renumbered copy can be built on first use; `perm` maps new -> old node index (device).
  std::vector<int> h_ptr, h_idx; std::vector<float> h_val;
  int reorder_mode = 1;               // graph option "reorder": 0 never, 1 when it pays (default), 2 always
  mutable int reorder_state = 0;      // 0 not evaluated yet, 1 renumbered copy in use, 2 evaluated and rejected
  mutable gcrnn_graph* reordered = nullptr;
  mutable int* perm = nullptr;        // new -> old
  mutable int* iperm = nullptr;       // old -> new
  mutable float tile_rows[2] = {0.f, 0.f};   // distinct neighbour rows per 128-node tile / 128, before and after
};

struct gcrnn_cell {
  gcrnn_cell_desc d;
  const gcrnn_graph* g = nullptr;
  // execution-path selection of the fp32 sparse path (gcrnn_cell_set_option / gcrnn_cell_get_option)
  mutable int need_dx = 0;        // hint for the next forward: the caller will ask backward for dX
  mutable int dh_last_only = 0;   // backward: `dH` is the gradient of the LAST state only, [B,F,N] (classification readout)
  mutable int forced_path = -1;   // -1: choose automatically; otherwise GCRNN_PATH_*
  mutable int last_path = 0;      // path taken by the last forward
  mutable int fwd_v2_mask = 63;   // fused sparse path: stage generations the last forward used (its backward follows them)
  mutable int fwd_reordered = 0;  // fused sparse path: the last forward ran on the graph's renumbered copy (its saved state is in that order)
  gcrnn::Options opt;             // tuning switches of this handle (gcrnn_cell_set_option)
  mutable void* graph_cache = nullptr;   // api.cu: CUDA graphs of small (launch-bound) fp32 forward / backward calls
};

namespace gcrnn {

// Gradient of the output at step t.  Dense dH[B,T,F,N]; or, with the cell option "dh_last_only", `dH` holds only the gradient of
// H[:, T-1] as [B,F,N] and every earlier step reads ONE shared zero slab [F,N] with sample stride 0 (L2-resident: no HBM
// traffic and no [B,T,F,N] gradient tensor at all).  Modules/architectures.py:1841-1850 uses only H.select(1, -1).
struct DhView {
  const float* dH; const float* zero; long long T, FN; bool last_only;
  const float* ptr(long long t) const { return last_only ? (t == T - 1 ? dH : zero) : dH + t * FN; }
  long long bstride(long long t) const { return last_only ? (t == T - 1 ? FN : 0) : T * FN; }
};

// Bump allocator over a caller-provided workspace.  With base == nullptr it only counts, which is how the
// *_workspace_bytes entry points are implemented (same code path as the real run).
struct Arena {
  char* base; size_t cap; size_t off = 0;
  Arena(void* b, size_t c) : base((char*)b), cap(c) {}
  bool dry() const { return base == nullptr; }
  template <class T> T* get(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    size_t o = off; off += bytes;
    if (!base) return nullptr;
    if (off > cap) throw Error(-4, "workspace too small: need > " + std::to_string(off) + " bytes, have " + std::to_string(cap));
    return (T*)(base + o);
  }
};

// fp32 path (gcrnn_f32.cu)
size_t lsigf_forward_f32(const gcrnn_graph* g, const float* h, const float* bias, const float* x, float* y,
                         int F, int K, int G, int64_t B, void* ws, size_t wsb, cudaStream_t st);
size_t lsigf_backward_f32(const gcrnn_graph* g, const float* h, const float* x, const float* dy, float* dx,
                          float* dh, float* dbias, int F, int K, int G, int64_t B, void* ws, size_t wsb,
                          cudaStream_t st);
size_t gat_forward_f32(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x, float* y,
                       int F, int G, int64_t B, void* ws, size_t wsb, cudaStream_t st);
size_t gat_backward_f32(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x,
                        const float* dy, float* dx, float* dmixer, float* dweight, int F, int G, int64_t B,
                        void* ws, size_t wsb, cudaStream_t st);
// returns scratch bytes used; *saved_used gets the bytes of `saved` used
size_t cell_forward_f32(const gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                        void* saved, size_t savedb, size_t* saved_used, void* ws, size_t wsb, int64_t B, int64_t T,
                        cudaStream_t st);
size_t cell_backward_f32(const gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0,
                         const float* H, const float* dH, const void* saved, size_t savedb,
                         const gcrnn_cell_params* gr, float* dX, float* dh0, void* ws, size_t wsb, int64_t B,
                         int64_t T, cudaStream_t st);

// test aid (gcrnn_debug_edge_relu_masks)
void debug_edge_relu_masks(const gcrnn_cell* cell, const void* saved, size_t savedb, int64_t B, int64_t T, unsigned char* out, cudaStream_t st);

// tensor-core path (gcrnn_tc.cu)
size_t cell_forward_tc(const gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0, float* H,
                       void* saved, size_t savedb, size_t* saved_used, void* ws, size_t wsb, int64_t B, int64_t T,
                       cudaStream_t st);
size_t cell_backward_tc(const gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0,
                        const float* H, const float* dH, const void* saved, size_t savedb,
                        const gcrnn_cell_params* gr, float* dX, float* dh0, void* ws, size_t wsb, int64_t B,
                        int64_t T, cudaStream_t st);
void tc_prepare_graph(gcrnn_graph* g, const float* S_host);

// graph.cu: the graph whose node numbering the fused sparse kernels should run on (g itself, or its renumbered copy)
const gcrnn_graph* locality_view(const gcrnn_graph* g);

}  // namespace gcrnn
