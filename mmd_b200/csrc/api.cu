// C ABI glue of libmmdk: error reporting, device probe, mmdk_unet_* entry points (see include/mmdk.h).
#include <atomic>
#include <string>

#include "common.cuh"
#include "unet.cuh"

namespace mmdk {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MMDK_OK;
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return MMDK_ECUDA;
}

uint64_t next_uid() {
  static std::atomic<uint64_t> counter{0};
  return ++counter;
}

uint64_t unet_state_uid(const UnetImpl* net, int mode, int B) {
  if (mode == MMDK_UNET_F16X3) return unet_fused_state_uid(net, B);
  if (mode == MMDK_UNET_F16X3_LAYERS) return unet_tc_state_uid(net, B);
  return net->uid;
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace mmdk

using namespace mmdk;

struct mmdk_unet {
  UnetImpl* impl;
};

extern "C" {

const char* mmdk_last_error(void) { return g_last_error.c_str(); }

int mmdk_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device) return fail(MMDK_ECUDA, "no CUDA device visible");
  cudaDeviceProp p{};
  MMDK_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) return fail(MMDK_ECUDA, "libmmdk is built for sm_100a only; found compute capability " +
                                                 std::to_string(p.major) + "." + std::to_string(p.minor));
  return MMDK_OK;
}

int mmdk_unet_create(const mmdk_unet_config* cfg, int n_tensors, const char* const* names,
                     const float* const* tensors_dev, const int64_t* numels, void* stream, mmdk_unet** out) {
  if (!out) return fail(MMDK_EINVAL, "null out");
  UnetImpl* impl = nullptr;
  int rc = unet_create(cfg, n_tensors, names, tensors_dev, numels, (cudaStream_t)stream, &impl);
  if (rc != MMDK_OK) return rc;
  impl->uid = next_uid();
  *out = new mmdk_unet{impl};
  return MMDK_OK;
}

/* internal (chain.cu): ids that make a captured graph's key unique to the handle and executor state it baked in */
uint64_t mmdk_internal_unet_uid(const mmdk_unet* net) { return net ? net->impl->uid : 0; }
uint64_t mmdk_internal_state_uid(const mmdk_unet* net, int mode, int B) { return net ? unet_state_uid(net->impl, mode, B) : 0; }

void mmdk_unet_destroy(mmdk_unet* net) {
  if (!net) return;
  unet_destroy(net->impl);
  delete net;
}

int mmdk_unet_forward(const mmdk_unet* net, int mode, const float* x_dev, int B, int t, float* eps_dev, void* stream) {
  if (!net || !x_dev || !eps_dev) return fail(MMDK_EINVAL, "null argument");
  if (B <= 0) return MMDK_OK;
  if (t < 0 || t >= net->impl->cfg.n_diffusion_steps) return fail(MMDK_EINVAL, "timestep out of range");
  if (mode == MMDK_UNET_FP32) return unet_forward_ffma(net->impl, x_dev, B, t, eps_dev, (cudaStream_t)stream);
  net->impl->last_mode = mode;
  if (mode == MMDK_UNET_F16X3) return unet_forward_fused(net->impl, x_dev, B, t, eps_dev, (cudaStream_t)stream);
  if (mode == MMDK_UNET_F16X3_LAYERS) return unet_forward_tc(net->impl, mode, x_dev, B, t, eps_dev, (cudaStream_t)stream);
  return fail(MMDK_EINVAL, "unknown UNet mode");
}

int mmdk_unet_cond_row(const mmdk_unet* net, int t, float* out_dev, int* n_cond, void* stream) {
  if (!net) return fail(MMDK_EINVAL, "null argument");
  if (n_cond) *n_cond = net->impl->n_cond;
  if (!out_dev) return MMDK_OK;
  if (t < 0 || t >= net->impl->cfg.n_diffusion_steps) return fail(MMDK_EINVAL, "timestep out of range");
  int n = net->impl->n_cond;
  copy_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(net->impl->cond_table + (size_t)t * n, out_dev, n);
  return check_cuda(cudaGetLastError(), "copy_kernel");
}

int mmdk_unet_debug_tap(const mmdk_unet* net, int op_index, float* out_dev, int* c_out, int* l_out, int* n_ops_out,
                        void* stream) {
  if (!net) return fail(MMDK_EINVAL, "null argument");
  if (n_ops_out) *n_ops_out = (int)net->impl->ops.size();
  if (net->impl->last_mode == MMDK_UNET_F16X3_LAYERS) return unet_tc_tap(net->impl, op_index, out_dev, c_out, l_out, (cudaStream_t)stream);
  return unet_fused_tap(net->impl, op_index, out_dev, c_out, l_out, (cudaStream_t)stream);
}

int mmdk_unet_debug_timeline(const mmdk_unet* net, int op_index, long long* dbg_dev, void* stream) {
  if (!net) return fail(MMDK_EINVAL, "null argument");
  if (op_index == -2) { net->impl->fused_dbg = dbg_dev; return MMDK_OK; }   // fused executor: per-item timeline of CTA 0
  return unet_tc_timeline(net->impl, op_index, dbg_dev, (cudaStream_t)stream);
}

int mmdk_unet_debug_keep_activations(const mmdk_unet* net, int on) {
  if (!net) return fail(MMDK_EINVAL, "null argument");
  net->impl->fused_keep = on != 0;
  return MMDK_OK;
}

int mmdk_unet_debug_stamps(const mmdk_unet* net, unsigned long long* stamps_dev, int n_slots, int max_ctas) {
  if (!net) return fail(MMDK_EINVAL, "null argument");
  if (stamps_dev && (n_slots < 1 || max_ctas < 1)) return fail(MMDK_EINVAL, "stamps: n_slots and max_ctas must be >= 1");
  net->impl->stamps = stamps_dev;
  net->impl->stamp_slots = stamps_dev ? n_slots : 0;
  net->impl->stamp_ctas = stamps_dev ? max_ctas : 0;
  net->impl->stamp_next = 0;
  return MMDK_OK;
}

int mmdk_debug_mma_calibrate(int N, int n_iters, int n_ctas, int n_acc, long long* out_dev, void* stream) {
  if (!out_dev) return fail(MMDK_EINVAL, "null argument");
  return mma_calibrate(N, n_iters, n_ctas, n_acc, out_dev, (cudaStream_t)stream);
}

}  // extern "C"
