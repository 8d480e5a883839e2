// Dense tensor-core path (bf16 operands, fp32 accumulate) of the gated GCRNN recurrence for sm_100a.
// The graph shift z @ S (Utils/graphML.py:123) runs on tcgen05 (tc_gemm.cuh); see DESIGN.md §TC-path.
#include "tc_gemm.cuh"
#include "tc_cell.cuh"

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

int num_sms(int device) {
  static int cached[64] = {0};
  if (device < 64 && cached[device]) return cached[device];
  int n = 0;
  CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  if (device < 64) cached[device] = n;
  return n;
}

// out = A @ (transposed ? S^T... see below).  forward shift z @ S uses Bop = S^T (stored K-major);
// backward shift g @ S^T uses Bop = S.
void shift_gemm(const gcrnn_graph* g, bool backward, const __nv_bfloat16* A, long long M, __nv_bfloat16* out_bf16,
                float* out_f32, cudaStream_t st) {
  const int N = g->N;
  GCRNN_CHECK(g->S_bf16 && g->St_bf16, "graph has no dense bf16 operator");
  GCRNN_CHECK(M > 0 && M < (1ll << 31), "row count out of range");
  const __nv_bfloat16* Bop = backward ? g->S_bf16 : g->St_bf16;
  EpiStore epi{out_bf16, out_f32, (long long)N};
  const int sms = num_sms(g->device);
  if (N % 256 == 0) {
    CUtensorMap tmA = make_tmap_bf16(A, M, N, BM), tmB = make_tmap_bf16(Bop, N, N, 256);
    launch_shift_gemm<256, EpiStore>(tmA, tmB, epi, (int)M, N, sms, st);
  } else {
    CUtensorMap tmA = make_tmap_bf16(A, M, N, BM), tmB = make_tmap_bf16(Bop, N, N, 128);
    launch_shift_gemm<128, EpiStore>(tmA, tmB, epi, (int)M, N, sms, st);
  }
}

}  // namespace tc

void tc_prepare_graph(gcrnn_graph* g, const float* S) {
  const int N = g->N;
  GCRNN_CHECK(N % 128 == 0, "the tensor-core path needs N %% 128 == 0 (N=%d)", N);
  std::vector<__nv_bfloat16> s((size_t)N * N), st((size_t)N * N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      __nv_bfloat16 v = __float2bfloat16(S[(size_t)i * N + j]);
      s[(size_t)i * N + j] = v;
      st[(size_t)j * N + i] = v;
    }
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat16* d = nullptr;
    CUDA_OK(cudaMalloc(&d, (size_t)N * N * sizeof(__nv_bfloat16)));
    g->owned.push_back(d);
    CUDA_OK(cudaMemcpy(d, which ? st.data() : s.data(), (size_t)N * N * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    (which ? g->St_bf16 : g->S_bf16) = d;
  }
  g->Npad = N;
}

}  // namespace gcrnn

extern "C" int gcrnn_debug_shift_gemm(const gcrnn_graph* g, int32_t backward, const void* A_bf16, int64_t M, void* out_bf16,
                                      float* out_f32, void* stream) {
  try {
    if (!g || !A_bf16) throw gcrnn::Error(-2, "null argument");
    CUDA_OK(cudaSetDevice(g->device));
    gcrnn::tc::shift_gemm(g, backward != 0, (const __nv_bfloat16*)A_bf16, M, (__nv_bfloat16*)out_bf16, out_f32, (cudaStream_t)stream);
  } catch (const std::exception& e) {
    gcrnn::set_last_error("%s", e.what());
    return -1;
  }
  return 0;
}
