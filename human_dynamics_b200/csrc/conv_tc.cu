// placeholder until the tcgen05 kernel lands
#include "conv_common.cuh"
namespace hd {
int launch_conv_tc(const ConvParams &, const hd_conv_desc *, cudaStream_t) {
  set_last_error_text("tcgen05 path not built");
  return HD_ERR_UNSUPPORTED;
}
}
extern "C" int hd_make_weight_tmap(const float *, int, int, int, void *) { return HD_ERR_UNSUPPORTED; }
