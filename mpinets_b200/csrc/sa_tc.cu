// sa_tc.cu -- bf16 tcgen05 tensor-core path (placeholder until the tcgen05 kernels land).
#include "engine.h"
namespace mpn {
int tc_prepare_weights(mpn_ctx*) { return MPN_OK; }
size_t tc_scratch_bytes(int) { return 0; }
int tc_encoder_forward(mpn_ctx*, cudaStream_t, const float*, int, int, float*, int) {
  set_error("MPN_PREC_BF16: tensor-core path not built");
  return MPN_ERR_STATE;
}
}  // namespace mpn
