/* gcrnn_b200.h — C ABI of the B200-native gated-GCRNN hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * Python host layer (gated_gcrnns_b200/) binds it with ctypes and mirrors the
 * reference's module interface (Utils/graphML.py) on top of it.
 *
 * Every entry point replaces one piece of /root/reference/Utils/graphML.py:
 *   gcrnn_graph_*            <- the `S` tensor handed to `addGSO`            (:1166-1173, :2074-2081, :2237-2244)
 *   gcrnn_lsigf_forward/...  <- `LSIGF(h, S, x, b)` + its autograd backward   (:47-140)
 *   gcrnn_gat_forward/...    <- `graphAttention` + `GraphAttentional.forward` (:521-627, :2084-2116)
 *   gcrnn_cell_forward/...   <- `GGCRNNCell.forward` + its autograd backward  (:2336-2428)
 *   gcrnn_allreduce_*        <- (new) the data-parallel gradient sum; the reference is single-process.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 row-major arrays in the
 *     reference's own layouts (x:[B,G,N], X:[B,T,G,N], H:[B,T,F,N], weights as in the
 *     reference's Parameters); graph construction takes HOST pointers;
 *   - work is enqueued on the `cudaStream_t` passed as `void* stream`; nothing synchronises;
 *   - return value 0 = OK, negative = error; `gcrnn_last_error()` returns a thread-local message;
 *   - the caller owns every tensor and the workspace; a graph/cell handle owns only its
 *     pre-processed copies of the shift operator (CSR/CSC, S+I pattern, bf16 tiles, TMA maps);
 *   - a handle is used from one host thread at a time; one handle per device.
 */
#ifndef GCRNN_B200_H
#define GCRNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCRNN_ABI_VERSION 2

typedef struct gcrnn_graph gcrnn_graph;   /* shift operator S (E edge features, N nodes) on one device */
typedef struct gcrnn_cell  gcrnn_cell;    /* one GGCRNNCell configuration bound to a graph            */

enum { GCRNN_SPATIAL_NONE = 0, GCRNN_SPATIAL_NODE = 1, GCRNN_SPATIAL_EDGE = 2 };
enum { GCRNN_PREC_FP32 = 0,        /* CUDA-core fp32 everywhere, sparse (CSR) shift — exact path                                  */
       GCRNN_PREC_BF16_TC = 1,     /* dense shift on tcgen05 tensor cores, bf16 operands (8-bit mantissa), fp32 accumulate         */
       GCRNN_PREC_BF16X2_TC = 2 }; /* same kernels, every operand split into bf16 hi + lo planes (16-bit mantissa), products as    *
                                    * K-concatenated MMAs x0 y0 + x1 y0 + x0 y1 into one fp32 accumulator; accurate tanh          */

/* GGCRNNCell(G, F, Kin, Kst, sigma=tanh, time_gating, spatial_gating, E, bias)  graphML.py:2196 */
typedef struct {
  int32_t G, F, Kin, Kst, E;
  int32_t time_gating;       /* 0 / 1 */
  int32_t spatial_gating;    /* GCRNN_SPATIAL_* */
  int32_t bias;              /* 0 / 1 */
  int32_t precision;         /* GCRNN_PREC_* */
} gcrnn_cell_desc;

/* Parameter (and gradient) pointers, named after the reference's state_dict keys.
 * Index 0 = input gate, 1 = forget gate.  Unused groups are NULL.  In a gradient block a NULL pointer
 * means "do not compute"; gradients are ACCUMULATED (+=) into the given buffers. */
typedef struct {
  float *weight_A, *weight_B, *bias;                /* [F,E,Kin,G] [F,E,Kst,F] [F]                         */
  float *t_weight_A[2], *t_weight_B[2], *t_bias[2]; /* GFL_in / GFL_forget sub-cells                       */
  float *t_mlp_w[2], *t_mlp_b[2];                   /* MLP_in.0 / MLP_forget.0: [F*N] (index f*N+n), [1]   */
  float *n_weight_A[2], *n_weight_B[2], *n_bias[2]; /* GRNN_node_in / GRNN_node_forget sub-cells           */
  float *n_head_w[2], *n_head_b[2];                 /* GFL_node_*.0: [1,E,Kst,F], [1]                      */
  float *e_mixer[2], *e_weight[2];                  /* input_attention / forget_attention: [2F], [F,F]     */
} gcrnn_cell_params;

/* the library is built with -fvisibility=hidden: only the functions declared here are exported (no data symbols) */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

int         gcrnn_abi_version(void);
const char* gcrnn_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's `gpu_launches`) */
uint64_t    gcrnn_debug_launch_count(void);
/* the dominant kernel on its own (unit tests, roofline timing): out = A @ S (backward=0) or A @ S^T (backward=1),
 * A: device bf16 [M, planes_in * N] row-major (plane q of row r at columns [q*N, (q+1)*N): plane 0 = bf16(x), plane 1 = bf16 of the
 * rounding residual); out_bf16: device [M, planes_out * N], out_f32: device [M, N]; either may be NULL.  Needs keep_dense. */
int         gcrnn_debug_shift_gemm(const gcrnn_graph* g, int32_t backward, const void* A_bf16, int64_t M, int32_t planes_in,
                                   void* out_bf16, int32_t planes_out, float* out_f32, void* stream);

/* Test aid: the ReLU decisions (1 = passed) of the two attention layers that the last fused edge-gated forward (GCRNN_PATH_NODE32)
 * left in `saved`, decoded into the caller's layout and node order: out = device uint8 [2 gates][B][T][32][N].  Lets a test evaluate
 * the fp64 oracle with exactly the one-sided derivatives the GPU picked at the kinks (tests/test_gpu_sparse_fused.py). */
int         gcrnn_debug_edge_relu_masks(const gcrnn_cell* cell, const void* saved, size_t saved_bytes, int64_t B, int64_t T,
                                        uint8_t* out, void* stream);

/* Tuning switches for tests and A/B measurements live on the HANDLE (the library keeps no process-wide mutable state):
 * gcrnn_cell_set_option(cell, name, value) / gcrnn_graph_set_option(graph, name, value) with name =
 *   "gemm_pair"     1: CTA-pair cta_group::2 shift GEMM when the shape allows, 0: single-CTA kernel  (cell and graph handles)
 *   "bwd_fused"     fused reverse-time step kernel of the tensor-core path
 *   "sparse_fused"  fused F == 32 edge-gated kernels of the sparse fp32 path; 0 = the generic per-op kernels
 *   "graph_capture" small fp32 cell calls are captured once per pointer set into a CUDA graph and replayed; 0 = direct launches
 *   "gate_fq8"      time-gate kernels with 8 feature groups per CTA: 0 off, 1 forward, 2 forward + backward
 *   "sparse_v2"     bit mask of the fused sparse stages that run their second-generation kernel: 1 shift, 2 gather-contract,
 *                   4|8 aggregate + bwd_rows, 16 bwd_node, 32 dh; 0 = all first generation
 *   "persist"       persistent small-graph kernels (GCRNN_PATH_PERSIST) when the cell allows; 0 = per-op kernels
 *   "fwd_fused"     1 = Horner-form forward of the tensor-core path (csrc/tc_hshift.cuh; off by default, see its header)
 *   "sparse_v2_tc"  tile contractions: 1 = 3xTF32 mma.sync, 0 = packed FFMA2;  "sparse_v2_fuse_dpre", "sparse_v2_bps",
 *                   "sparse_v2_rows_bps"
 * A backward always follows the stage generations its forward used; every change invalidates the handle's captured CUDA graphs.
 *
 * Graph handles only: "reorder" = 0 never, 1 when it pays (default), 2 always.  The reference fixes no node order
 * (Utils/graphTools.py builds S in whatever order the data came in); the fused sparse kernels live on L1 reuse of neighbour rows
 * inside tiles of 128 consecutive nodes.  On first use of a one-operator CSR graph with N >= 4096 the library builds a renumbered
 * copy (breadth-first balls of 128 nodes) and keeps it when it lowers the distinct neighbour rows per tile by >= 1.5x; X / h0 / dH
 * are gathered and H / dh0 scattered through the permutation inside the layout-conversion kernels, so callers never see the
 * internal numbering.  Set it before the first forward on the graph.  gcrnn_graph_get_option also reads "reordered" (0 / 1),
 * "tile_rows_before_x100" and "tile_rows_after_x100" (distinct neighbour rows per 128-node tile / 128, x 100). */
int         gcrnn_graph_set_option(gcrnn_graph* g, const char* name, int32_t value);
int         gcrnn_graph_get_option(const gcrnn_graph* g, const char* name, int32_t* value);

/* ---- graph -------------------------------------------------------------------------------------- */
/* E operators in CSR, HOST arrays: rowptr[e] has N+1 entries, entry (i, colidx[p]) = S_e[i, j] = vals[p].
 * Row-vector convention of the reference: shift is z <- z @ S_e (graphML.py:123). */
int gcrnn_graph_create_csr(gcrnn_graph** out, int32_t N, int32_t E,
                           const int64_t* const* rowptr, const int32_t* const* colidx,
                           const float* const* vals, int32_t device);
/* Dense HOST array S[E,N,N] (what `addGSO` receives); exact zeros are dropped from the CSR form.
 * keep_dense != 0 additionally keeps bf16 copies of S and S^T for the tensor-core path (E must be 1). */
int gcrnn_graph_create_dense(gcrnn_graph** out, int32_t N, int32_t E, const float* S,
                             int32_t keep_dense, int32_t device);
int gcrnn_graph_destroy(gcrnn_graph* g);
int gcrnn_graph_info(const gcrnn_graph* g, int32_t* N, int32_t* E, int64_t* nnz, int64_t* nnz_att);

/* ---- LSIGF: y[b,f,n] = sum_{e,k,g} h[f,e,k,g] (x S_e^k)[b,g,n] + bias[f] -------------------------- */
size_t gcrnn_lsigf_workspace_bytes(const gcrnn_graph* g, int32_t F, int32_t K, int32_t G, int64_t B);
int gcrnn_lsigf_forward(const gcrnn_graph* g, const float* h, const float* bias /*NULL ok*/,
                        const float* x, float* y, int32_t F, int32_t K, int32_t G, int64_t B,
                        void* workspace, size_t workspace_bytes, void* stream);
/* dx / dh / dbias may be NULL; dh and dbias are accumulated (+=), dx is overwritten. */
int gcrnn_lsigf_backward(const gcrnn_graph* g, const float* h, const float* x, const float* dy,
                         float* dx, float* dh, float* dbias, int32_t F, int32_t K, int32_t G, int64_t B,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- graph attention (edge gate): y = relu(GAT_{1 head}(x)); mixer[2F], weight[F,G] ---------------- */
size_t gcrnn_gat_workspace_bytes(const gcrnn_graph* g, int32_t F, int32_t G, int64_t B);
int gcrnn_gat_forward(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x,
                      float* y, int32_t F, int32_t G, int64_t B,
                      void* workspace, size_t workspace_bytes, void* stream);
int gcrnn_gat_backward(const gcrnn_graph* g, const float* mixer, const float* weight, const float* x,
                       const float* dy, float* dx, float* dmixer, float* dweight,
                       int32_t F, int32_t G, int64_t B,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- the gated GCRNN cell --------------------------------------------------------------------------- */
int gcrnn_cell_create(gcrnn_cell** out, const gcrnn_cell_desc* desc, const gcrnn_graph* g);
int gcrnn_cell_destroy(gcrnn_cell* c);
/* Execution paths of the fp32 sparse precision (same results within the stated fp32 tolerance):
 *   GENERIC  per-op kernels, any shape / gating mode, supports dX;
 *   PERSIST  persistent fused recurrence for small graphs (csrc/persist_f32.cuh): ONE launch runs the whole sequence of every sample
 *            with the shift operator, the taps, every gate's weights (time, node or edge gates; for edge gates also the attention
 *            pattern of S + I) and h_t on chip across all T steps, ONE launch the reverse sweep; E == 1, no dX, sizes that fit one
 *            SM's shared memory (the reference's own N = 80 and N = 59 configurations);
 *   NODE32   fused edge-gated kernels for F == 32, one warp per (sample, node) (csrc/sp32_kernels.cuh); a dX request makes
 *            backward run the generic sweep on the generic prefix of the saved state.
 * Options (per cell handle, used from one host thread at a time; the tuning switches listed above are set the same way):
 *   "path"      (set)  -1 = automatic (default), otherwise force GCRNN_PATH_* — the autograd glue forces backward onto the path
 *                      its forward took;
 *   "need_dx"   (set)  hint for the next forward: backward will be asked for dX (reserved for paths that cannot serve it);
 *   "dh_last_only" (set) 1: the next backward's `dH` is the gradient of the LAST state only, dH[:, T-1] as [B,F,N]; the gradient of
 *                      every earlier output is zero and no [B,T,F,N] gradient tensor exists (classification readout,
 *                      Modules/architectures.py:1841-1850 uses only H.select(1, -1));
 *   "last_path" (get)  path taken by the last forward on this handle. */
enum { GCRNN_PATH_GENERIC = 0, GCRNN_PATH_NODE32 = 1, GCRNN_PATH_PERSIST = 2 };
int gcrnn_cell_set_option(gcrnn_cell* c, const char* name, int32_t value);
int gcrnn_cell_get_option(const gcrnn_cell* c, const char* name, int32_t* value);
/* saved_bytes: buffer written by forward and read by backward; fwd/bwd_bytes: scratch. */
int gcrnn_cell_workspace_bytes(const gcrnn_cell* c, int64_t B, int64_t T, int32_t need_input_grads,
                               size_t* saved_bytes, size_t* fwd_bytes, size_t* bwd_bytes);
/* X:[B,T,G,N], h0:[B,F,N] -> H:[B,T,F,N].  `saved` (saved_bytes from gcrnn_cell_workspace_bytes) is required. */
int gcrnn_cell_forward(gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0,
                       float* H, void* saved, size_t saved_bytes,
                       void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);
/* dH: [B,T,F,N], or [B,F,N] with the cell option "dh_last_only".  grads: accumulated (+=).  dX / dh0 may be NULL.  Parameters the reference never uses
 * (GFL_out.*, MLP_out.*) do not appear here at all: their gradient stays None, as in the reference. */
int gcrnn_cell_backward(gcrnn_cell* c, const gcrnn_cell_params* p, const float* X, const float* h0,
                        const float* H, const float* dH, const void* saved, size_t saved_bytes,
                        const gcrnn_cell_params* grads, float* dX, float* dh0,
                        void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);

/* ---- data-parallel gradient sum (one NCCL all-reduce per step on a flat fp32 bucket) ----------------- */
typedef struct gcrnn_comm gcrnn_comm;
int gcrnn_comm_unique_id(void* id128 /* 128 bytes out */);
int gcrnn_comm_create(gcrnn_comm** out, const void* id128, int32_t rank, int32_t world, int32_t device);
int gcrnn_comm_destroy(gcrnn_comm* c);
int gcrnn_allreduce_sum(gcrnn_comm* c, float* bucket, int64_t count, void* stream);

/* ---- on-device builders of the synthetic benchmark inputs ------------------------------------------
 * Directed kNN graph of N points of the unit square (xy_dev: DEVICE fp32 [N][2]): row i holds its k nearest neighbours (i itself
 * excluded, ties broken by index) with weights exp(-d^2 / sigma2) (sigma2 <= 0: the mean squared distance to the k-th neighbour),
 * divided by the spectral-norm estimate of `power_iters` power iterations on A^T A (0: no normalisation) — what the reference does
 * with dense matrices and numpy eig (Utils/graphTools.py:516-634, kStepPredGRNNs.py:768) and gated_gcrnns_b200/graphs.py did with
 * scipy on the host.  Uniform-grid search and power iteration run on the GPU; the CSR (N*k entries, ascending columns per row)
 * is returned in HOST arrays rowptr[N+1], colidx[N*k], vals[N*k].  reorder != 0 renumbers the nodes along a Hilbert curve of the
 * grid cells (neighbour rows then share cache lines in the gather kernels); perm[new] = old (may be NULL). */
int gcrnn_build_knn_csr(int32_t N, int32_t k, const float* xy_dev, float sigma2, int32_t power_iters, int32_t reorder, int32_t device,
                        int64_t* rowptr, int32_t* colidx, float* vals, int32_t* perm, float* sigma2_out, float* lambda_out);
/* Diffusion-process signals of the k-step prediction task (Utils/dataTools.py:1290-1297): out[0] = x0, out[t+1] = out[t] S + noise[t].
 * x0: device [R,N]; noise: device [T,R,N] or NULL; out: device [T+1,R,N]. */
int gcrnn_data_diffusion(const gcrnn_graph* g, const float* x0, const float* noise, float* out, int64_t R, int32_t T, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* GCRNN_B200_H */
