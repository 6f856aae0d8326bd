// tcgen05 executor of the UNet layer program (placeholder until the tensor-core path lands).
#include "common.cuh"
#include "unet.cuh"

namespace mmdk {

int unet_forward_tc(UnetImpl*, int, const float*, int, int, float*, cudaStream_t) {
  return fail(MMDK_EINVAL, "tcgen05 UNet executor not built yet; use MMDK_UNET_FP32");
}
void unet_tc_release(UnetImpl*) {}

}  // namespace mmdk
