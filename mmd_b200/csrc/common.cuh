// Shared host/device definitions of libmmdk (B200 / sm_100a guided-diffusion trajectory sampler).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/mmdk.h"

namespace mmdk {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);

#define MMDK_CUDA(call)                                   \
  do {                                                    \
    int _rc = ::mmdk::check_cuda((call), #call);          \
    if (_rc != MMDK_OK) return _rc;                       \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// UNet layer program: the network is lowered once (mmdk_unet_create) into a flat list of ops over
// per-sample activation buffers; both the FFMA and the tcgen05 executors interpret the same list.
// ---------------------------------------------------------------------------------------------------
enum OpType : int {
  OP_CONVBLOCK = 0,  // Conv1d(k5,p2) -> GroupNorm -> Mish [-> + cond] [-> + residual] (layers.py:279-296,346-358)
  OP_DOWN = 1,       // Conv1d(k3,s2,p1)                                               (layers.py:261-267)
  OP_UP = 2,         // ConvTranspose1d(k4,s2,p1)                                      (layers.py:270-276)
  OP_FINAL = 3,      // Conv1d(k1) -> eps [B,H,D]                                      (temporal_unet.py:116-119)
  OP_ATTN = 4,       // Residual(PreNorm(LinearAttention)), in place                    (layers.py:177-229)
};

struct Op {
  int type;
  int cin, cout;
  int lin, lout;
  int src;       // activation buffer offset (floats, per sample) of the input, layout [C][L+4] (2 zero halo columns each side)
  int dst;       // raw conv output buffer offset (GroupNorm is applied in place; final value lands here too)
  int w, b;      // packed weight [cin][k][cout] / bias offsets in the weight blob
  int gn_w, gn_b, n_groups;
  int cond;      // offset into the cond row of this timestep, or -1
  int res_src;   // residual input buffer offset or -1
  int res_cin;   // channels of the residual input
  int res_w, res_b;  // 1x1 residual conv [cin][cout] or -1 (identity)
  // dataflow for executors that keep activations outside shared memory: index of the op producing each input
  // (-1 = the network input x); p_src1 = second half of a channel concat (temporal_unet.py:162) or -2 if none
  int p_src0, p_src1, p_res, p_res1;
  int c_src0;        // channels coming from p_src0 (cin - c_src0 come from p_src1)
  int c_res0;        // channels of the residual input coming from p_res (res_cin - c_res0 from p_res1)
};

static constexpr int kMaxOps = 64;

}  // namespace mmdk
