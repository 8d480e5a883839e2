// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace gcrnn {
size_t cell_forward_tc(const gcrnn_cell*, const gcrnn_cell_params*, const float*, const float*, float*, void*, size_t, size_t*, void*, size_t, int64_t, int64_t, cudaStream_t) { throw Error(-7, "tensor-core path not built"); }
size_t cell_backward_tc(const gcrnn_cell*, const gcrnn_cell_params*, const float*, const float*, const float*, const float*, const void*, size_t, const gcrnn_cell_params*, float*, float*, void*, size_t, int64_t, int64_t, cudaStream_t) { throw Error(-7, "tensor-core path not built"); }
void tc_prepare_graph(gcrnn_graph*, const float*) { throw Error(-7, "tensor-core path not built"); }
}
