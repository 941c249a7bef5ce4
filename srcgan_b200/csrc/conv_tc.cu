// placeholder until the tcgen05 engine lands
#include "common.cuh"
namespace srcgan {
bool conv_tc_supported(const srcgan_conv_params*) { return false; }
int conv_fprop_tc(const srcgan_conv_params*, cudaStream_t) { set_error("tcgen05 engine not built"); return SRCGAN_E_INVALID; }
int pack_weights_tc_host(const float*, int, int, int, int, void*, cudaStream_t) { set_error("tcgen05 engine not built"); return SRCGAN_E_INVALID; }
size_t packed_weight_bytes_tc(int, int, int, int) { return 0; }
bool conv_wgrad_tc_supported(const srcgan_conv_params*) { return false; }
size_t conv_wgrad_tc_workspace(const srcgan_conv_params*) { return 0; }
int conv_wgrad_tc(const srcgan_conv_params*, float*, float*, int, void*, size_t, cudaStream_t) { set_error("tcgen05 engine not built"); return SRCGAN_E_INVALID; }
}
