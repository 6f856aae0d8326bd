#pragma once
#include <map>
#include <vector>

#include "common.cuh"

namespace mmdk {

struct TcState;     // per-layer tcgen05 executor state (unet_tc.cu)
struct FusedState;  // persistent whole-forward tcgen05 executor state (unet_fused.cu)

// Process-wide unique ids: a captured CUDA graph (chain.cu) bakes in the device pointers of a network handle AND of the executor
// state (activation images, packed weights) of one batch size; both can be destroyed and re-created -- possibly at the same host
// address -- so graph keys carry these ids instead of trusting pointers.
uint64_t next_uid();

struct UnetImpl {
  uint64_t uid = 0;
  mmdk_unet_config cfg{};
  std::vector<Op> ops;       // host copy of the layer program
  Op* ops_dev = nullptr;
  float* blob = nullptr;     // packed fp32 weights: conv [cin][k][cout], vectors as-is
  size_t blob_floats = 0;
  float* cond_table = nullptr;  // [T][n_cond]
  int n_cond = 0;
  int per_sample_floats = 0;    // activation floats per sample for the fp32 executor
  int ffma_S = 1;               // samples per CTA of the fp32 executor
  int attn_scratch_floats = 0;  // shared scratch of the LinearAttention op (qkv + context + stats), 0 without attention
  TcState* tc = nullptr;              // per-layer executor state of the last batch size used
  std::map<int, TcState*> tc_cache;   // per batch size (bounded; see unet_forward_tc)
  std::map<int, FusedState*> fused;   // per batch size (bounded; see unet_forward_fused)
  FusedState* fused_last = nullptr;
  int last_mode = 0;
  bool fused_keep = false;          // fused executor writes every activation image to global memory (mmdk_unet_debug_keep_activations)
  long long* fused_dbg = nullptr;   // debug timeline buffer of the fused executor (mmdk_unet_debug_timeline)
  unsigned long long* stamps = nullptr;   // launch-duration probe of the fused executor (mmdk_unet_debug_stamps)
  int stamp_slots = 0, stamp_ctas = 0;
  unsigned stamp_next = 0;
};

int unet_create(const mmdk_unet_config* cfg, int n_tensors, const char* const* names, const float* const* tensors,
                const int64_t* numels, cudaStream_t stream, UnetImpl** out);
void unet_destroy(UnetImpl* net);
// id of the executor state that mode `mode` would use for batch size B right now (0: not built yet; fp32: the handle's id)
uint64_t unet_state_uid(const UnetImpl* net, int mode, int B);
uint64_t unet_fused_state_uid(const UnetImpl* net, int B);
uint64_t unet_tc_state_uid(const UnetImpl* net, int B);
int unet_forward_ffma(const UnetImpl* net, const float* x, int B, int t, float* eps, cudaStream_t stream);

// persistent whole-forward tcgen05 executor (unet_fused.cu)
int unet_forward_fused(UnetImpl* net, const float* x, int B, int t, float* eps, cudaStream_t stream);
void unet_fused_release(UnetImpl* net);
int unet_fused_tap(UnetImpl* net, int op_index, float* out, int* c_out, int* l_out, cudaStream_t stream);

// per-layer tcgen05 executor (unet_tc.cu)
int unet_forward_tc(UnetImpl* net, int mode, const float* x, int B, int t, float* eps, cudaStream_t stream);
void unet_tc_release(UnetImpl* net);
int unet_tc_timeline(UnetImpl* net, int op_index, long long* dbg_dev, cudaStream_t stream);
int mma_calibrate(int N, int n_iters, int n_ctas, int n_acc, long long* out_dev, cudaStream_t stream);
int unet_tc_tap(UnetImpl* net, int op_index, float* out, int* c_out, int* l_out, cudaStream_t stream);

}  // namespace mmdk
