// TemporalUnet on B200: network lowering (state_dict -> packed weights + layer program + cond table) and the
// fp32 CUDA-core executor (exact-parity mode).  The tcgen05 executor lives in unet_tc.cu and runs the same program.
//
// Reference: mmd/models/diffusion_models/temporal_unet.py:23-174, mmd/models/layers/layers.py:232-398.
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "unet.cuh"

namespace mmdk {

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mish_accurate(float x) {
  // nn.Mish: x * tanh(softplus(x)) (layers.py:292); softplus with the usual overflow guard
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}

// [cout][cin][k] (Conv1d) or [cin][cout][k] (ConvTranspose1d) -> [cin][k][cout]
__global__ void repack_conv_kernel(const float* __restrict__ src, float* __restrict__ dst, int cin, int cout, int k,
                                   int transposed) {
  int n = cin * cout * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int co = i % cout;
    int kk = (i / cout) % k;
    int ci = i / (cout * k);
    int s = transposed ? ((ci * cout + co) * k + kk) : ((co * cin + ci) * k + kk);
    dst[i] = src[s];
  }
}

struct CondLayer {
  const float* w;  // [C][temb]
  const float* b;  // [C]
  int C, off;
};

// U0: TimeEncoder (layers.py:232-258) + every ResidualTemporalBlock.cond_mlp (layers.py:337-341), one block per t.
// sincos: [T][32] sinusoidal embedding evaluated by the host with the reference's own expression.
__global__ void time_embed_kernel(const float* __restrict__ sincos, const float* __restrict__ w1,
                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                  const float* __restrict__ b2, const CondLayer* __restrict__ layers, int n_layers,
                                  int temb_dim, int n_cond, float* __restrict__ out) {
  __shared__ float emb[32];
  __shared__ float h1[128];
  __shared__ float m[64];
  int t = blockIdx.x;
  if (threadIdx.x < 32) emb[threadIdx.x] = sincos[t * 32 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x < 128) {
    float a = b1[threadIdx.x];
    for (int i = 0; i < 32; ++i) a = fmaf(w1[threadIdx.x * 32 + i], emb[i], a);
    h1[threadIdx.x] = mish_accurate(a);
  }
  __syncthreads();
  if (threadIdx.x < temb_dim) {
    float a = b2[threadIdx.x];
    for (int i = 0; i < 128; ++i) a = fmaf(w2[threadIdx.x * 128 + i], h1[i], a);
    m[threadIdx.x] = mish_accurate(a);
  }
  __syncthreads();
  for (int l = 0; l < n_layers; ++l) {
    CondLayer L = layers[l];
    for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
      float a = L.b[c];
      for (int i = 0; i < temb_dim; ++i) a = fmaf(L.w[c * temb_dim + i], m[i], a);
      out[(size_t)t * n_cond + L.off + c] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// fp32 executor: one CTA owns S whole samples; activations stay in shared memory across all layers.
// ---------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void conv5_raw(const float* __restrict__ w, const float* __restrict__ bias,
                                          const float* src, float* dst, int cin, int cout, int L, int sstride) {
  const int pitch = L + 4;
  const int units = cout * (L >> 3);
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    const int co = u % cout;
    const int p0 = (u / cout) << 3;
    float acc[S][8];
    const float bv = __ldg(bias + co);
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[s][j] = bv;
    const float* wp = w + co;
#pragma unroll 2
    for (int ci = 0; ci < cin; ++ci) {
      float wk[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) wk[k] = __ldg(wp + (ci * 5 + k) * cout);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float4* a4 = reinterpret_cast<const float4*>(src + s * sstride + ci * pitch + p0);
        float4 v0 = a4[0], v1 = a4[1], v2 = a4[2];
        float a[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int k = 0; k < 5; ++k) acc[s][j] = fmaf(wk[k], a[j + k], acc[s][j]);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float* d = dst + s * sstride + co * pitch + 2 + p0;
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = acc[s][j];
    }
  }
}

// GroupNorm (biased variance, eps 1e-5) -> Mish -> (+cond[c]) -> (+residual), in place on `buf`.
template <int S>
__device__ __forceinline__ void gn_finalize(float* buf, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, int C, int L, int n_groups,
                                            const float* __restrict__ cond, const float* res_src, int res_cin,
                                            const float* __restrict__ res_w, const float* __restrict__ res_b,
                                            int sstride) {
  const int pitch = L + 4;
  const int cpg = C / n_groups;
  const int n = cpg * L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int pair = warp; pair < S * n_groups; pair += nwarps) {
    const int s = pair / n_groups, g = pair % n_groups;
    float* base = buf + s * sstride + (g * cpg) * pitch + 2;
    float sum = 0.f;
    for (int e = lane; e < n; e += 32) sum += base[(e / L) * pitch + (e % L)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)n;
    float sq = 0.f;
    for (int e = lane; e < n; e += 32) {
      float d = base[(e / L) * pitch + (e % L)] - mean;
      sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.f / sqrtf(sq / (float)n + 1e-5f);
    for (int e = lane; e < n; e += 32) {
      const int cl = e / L, pos = e % L;
      const int c = g * cpg + cl;
      float v = base[cl * pitch + pos];
      v = (v - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      v = mish_accurate(v);
      if (cond) v += __ldg(cond + c);
      if (res_src) {
        const float* rs = res_src + s * sstride + 2 + pos;
        if (res_w) {
          float r = __ldg(res_b + c);
          const int rp = pitch;
          for (int ci = 0; ci < res_cin; ++ci) r = fmaf(__ldg(res_w + ci * C + c), rs[ci * rp], r);
          v += r;
        } else {
          v += rs[c * pitch];
        }
      }
      base[cl * pitch + pos] = v;
    }
  }
}

template <int S>
__device__ __forceinline__ void zero_halo(float* buf, int C, int L, int sstride) {
  const int pitch = L + 4;
  for (int i = threadIdx.x; i < S * C * 4; i += blockDim.x) {
    int s = i / (C * 4), r = i % (C * 4);
    int c = r >> 2, h = r & 3;
    buf[s * sstride + c * pitch + (h < 2 ? h : L + h)] = 0.f;
  }
}

template <int S>
__global__ void __launch_bounds__(256, 1)
unet_ffma_kernel(const Op* __restrict__ ops, int n_ops, const float* __restrict__ W,
                 const float* __restrict__ cond_row, const float* __restrict__ x, float* __restrict__ eps, int B,
                 int H, int D, int sstride) {
  extern __shared__ __align__(16) float smem[];
  const int b0 = blockIdx.x * S;
  // load x [B][H][D] -> buffer 0 as [D][H+4]
  {
    const int pitch = H + 4;
    for (int i = threadIdx.x; i < S * H * D; i += blockDim.x) {
      int s = i / (H * D), r = i % (H * D);
      int h = r / D, d = r % D;
      int b = b0 + s;
      float v = (b < B) ? x[((size_t)b * H + h) * D + d] : 0.f;
      smem[s * sstride + d * pitch + 2 + h] = v;
    }
    zero_halo<S>(smem, D, H, sstride);
  }
  __syncthreads();
  for (int oi = 0; oi < n_ops; ++oi) {
    const Op op = ops[oi];
    const float* src = smem + op.src;
    float* dst = smem + op.dst;
    if (op.type == OP_CONVBLOCK) {
      conv5_raw<S>(W + op.w, W + op.b, src, dst, op.cin, op.cout, op.lin, sstride);
      __syncthreads();
      gn_finalize<S>(dst, W + op.gn_w, W + op.gn_b, op.cout, op.lout, op.n_groups,
                     op.cond >= 0 ? cond_row + op.cond : nullptr, op.res_src >= 0 ? smem + op.res_src : nullptr,
                     op.res_cin, op.res_w >= 0 ? W + op.res_w : nullptr, op.res_w >= 0 ? W + op.res_b : nullptr,
                     sstride);
      zero_halo<S>(dst, op.cout, op.lout, sstride);
    } else if (op.type == OP_DOWN) {
      // out[co][j] = b + sum_ci sum_k w[ci][k][co] * x[ci][2j + k - 1]
      const int pin = op.lin + 4, pout = op.lout + 4;
      const int units = op.cout * op.lout;
      for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int co = u % op.cout, j = u / op.cout;
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = __ldg(W + op.b + co);
        for (int ci = 0; ci < op.cin; ++ci) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float wv = __ldg(W + op.w + (ci * 3 + k) * op.cout + co);
#pragma unroll
            for (int s = 0; s < S; ++s) acc[s] = fmaf(wv, src[s * sstride + ci * pin + 2 + 2 * j + k - 1], acc[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s * sstride + co * pout + 2 + j] = acc[s];
      }
      zero_halo<S>(dst, op.cout, op.lout, sstride);
    } else if (op.type == OP_UP) {
      // ConvTranspose1d(k4,s2,p1): out[j] += x[i] * w[k], j = 2i - 1 + k
      const int pin = op.lin + 4, pout = op.lout + 4;
      const int units = op.cout * op.lout;
      for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int co = u % op.cout, j = u / op.cout;
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = __ldg(W + op.b + co);
        const int kA = (j & 1) ? 0 : 1;         // taps with (j + 1 - k) even
        const int iA = (j + 1 - kA) >> 1;       // k = kA
        const int iB = iA - 1;                  // k = kA + 2
        for (int ci = 0; ci < op.cin; ++ci) {
          const float wA = __ldg(W + op.w + (ci * 4 + kA) * op.cout + co);
          const float wB = __ldg(W + op.w + (ci * 4 + kA + 2) * op.cout + co);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float* xr = src + s * sstride + ci * pin + 2;  // halo columns are zero, so i = -1 / lin read 0
            acc[s] = fmaf(wA, xr[iA], acc[s]);
            acc[s] = fmaf(wB, xr[iB], acc[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s * sstride + co * pout + 2 + j] = acc[s];
      }
      zero_halo<S>(dst, op.cout, op.lout, sstride);
    } else if (op.type == OP_ATTN) {
      // Residual(PreNorm(LinearAttention)) (layers.py:177-229), heads=4, dim_head=32, one sample at a time:
      // LayerNorm over channels -> qkv = W x (1x1, no bias) -> q *= 32^-1/2, k = softmax over positions ->
      // context[h] = k[h] v[h]^T (32x32) -> out[h] = context[h]^T q[h] -> W_out out + b + x
      const int C = op.cin, L = op.lin, pitch = L + 4;
      float* scr = smem + S * sstride;          // qkv [384][L] | context [4][32][32] | mean [L] | rstd [L]
      float* ctx = scr + 384 * L;
      float* mu = ctx + 4 * 32 * 32;
      float* rs = mu + L;
      for (int s = 0; s < S; ++s) {
        float* xs = smem + s * sstride + op.src + 2;
        for (int n = threadIdx.x; n < L; n += blockDim.x) {
          float m = 0.f;
          for (int c = 0; c < C; ++c) m += xs[c * pitch + n];
          m /= (float)C;
          float v = 0.f;
          for (int c = 0; c < C; ++c) { float d = xs[c * pitch + n] - m; v = fmaf(d, d, v); }
          mu[n] = m;
          rs[n] = 1.f / sqrtf(v / (float)C + 1e-5f);
        }
        __syncthreads();
        for (int u = threadIdx.x; u < 384 * L; u += blockDim.x) {
          const int o = u % 384, n = u / 384;
          float acc = 0.f;
          for (int c = 0; c < C; ++c) {
            const float xn = (xs[c * pitch + n] - mu[n]) * rs[n] * __ldg(W + op.gn_w + c) + __ldg(W + op.gn_b + c);
            acc = fmaf(__ldg(W + op.w + c * 384 + o), xn, acc);
          }
          scr[o * L + n] = (o < 128) ? acc * 0.17677669529663687f : acc;   // q * dim_head^-0.5
        }
        __syncthreads();
        {  // softmax over positions of every k row (rows 128..255)
          const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
          for (int r = warp; r < 128; r += nw) {
            float* kr = scr + (128 + r) * L;
            float mx = -INFINITY;
            for (int n = lane; n < L; n += 32) mx = fmaxf(mx, kr[n]);
            for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
            float sum = 0.f;
            for (int n = lane; n < L; n += 32) { float e = expf(kr[n] - mx); kr[n] = e; sum += e; }
            for (int o2 = 16; o2 > 0; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
            const float inv = 1.f / sum;
            for (int n = lane; n < L; n += 32) kr[n] *= inv;
          }
        }
        __syncthreads();
        for (int u = threadIdx.x; u < 4 * 32 * 32; u += blockDim.x) {   // context[h][d][e] = sum_n k[h][d][n] v[h][e][n]
          const int e = u & 31, d = (u >> 5) & 31, h = u >> 10;
          const float* kr = scr + (128 + h * 32 + d) * L;
          const float* vr = scr + (256 + h * 32 + e) * L;
          float acc = 0.f;
          for (int n = 0; n < L; ++n) acc = fmaf(kr[n], vr[n], acc);
          ctx[u] = acc;
        }
        __syncthreads();
        for (int u = threadIdx.x; u < 128 * L; u += blockDim.x) {       // out[h][e][n] -> stored over the k rows
          const int n = u % L, e = (u / L) & 31, h = u / (L * 32);
          float acc = 0.f;
          for (int d = 0; d < 32; ++d) acc = fmaf(ctx[(h * 32 + d) * 32 + e], scr[(h * 32 + d) * L + n], acc);
          scr[(128 + h * 32 + e) * L + n] = acc;
        }
        __syncthreads();
        for (int u = threadIdx.x; u < C * L; u += blockDim.x) {          // to_out (1x1) + bias + residual, in place
          const int c = u % C, n = u / C;
          float acc = __ldg(W + op.b + c);
          for (int j = 0; j < 128; ++j) acc = fmaf(__ldg(W + op.res_w + j * C + c), scr[(128 + j) * L + n], acc);
          xs[c * pitch + n] += acc;
        }
        __syncthreads();
      }
    } else {  // OP_FINAL: 1x1 conv to state_dim channels, written straight to eps [B][H][D]
      const int pin = op.lin + 4;
      for (int u = threadIdx.x; u < S * op.lin * op.cout; u += blockDim.x) {
        const int d = u % op.cout, h = (u / op.cout) % op.lin, s = u / (op.cout * op.lin);
        float acc = __ldg(W + op.b + d);
        for (int ci = 0; ci < op.cin; ++ci)
          acc = fmaf(__ldg(W + op.w + ci * op.cout + d), src[s * sstride + ci * pin + 2 + h], acc);
        if (b0 + s < B) eps[((size_t)(b0 + s) * op.lin + h) * op.cout + d] = acc;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// host: lowering
// ---------------------------------------------------------------------------------------------------
namespace {

struct Builder {
  const std::map<std::string, std::pair<const float*, int64_t>>& T;
  std::vector<float*> dev_srcs;
  cudaStream_t stream;
  float* blob = nullptr;     // device
  size_t blob_cap = 0, blob_used = 0;
  std::string err;

  const float* find(const std::string& k, int64_t numel) {
    auto it = T.find(k);
    if (it == T.end()) { err = "missing tensor '" + k + "'"; return nullptr; }
    if (it->second.second != numel) { err = "tensor '" + k + "' has wrong size"; return nullptr; }
    return it->second.first;
  }
  int alloc(size_t n) {
    size_t off = (blob_used + 3) & ~size_t(3);
    blob_used = off + n;
    return (int)off;
  }
  // returns blob offset of packed conv weight [cin][k][cout]
  int conv(const std::string& key, int cin, int cout, int k, bool transposed) {
    const float* src = find(key, (int64_t)cin * cout * k);
    if (!src) return -1;
    int off = alloc((size_t)cin * cout * k);
    if (blob) repack_conv_kernel<<<64, 256, 0, stream>>>(src, blob + off, cin, cout, k, transposed ? 1 : 0);
    return off;
  }
  int vec(const std::string& key, int n) {
    const float* src = find(key, n);
    if (!src) return -1;
    int off = alloc(n);
    if (blob) cudaMemcpyAsync(blob + off, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream);
    return off;
  }
};

int gn_groups(int c) {  // layers.py:392-398
  if (c < 8) return 1;
  for (int n = 8; n < 18; ++n)
    if (c % n == 0) return n;
  return 1;
}

}  // namespace

struct CondSpec { std::string prefix; int C; int off; };

static int lower(Builder& bld, const mmdk_unet_config& cfg, std::vector<Op>& ops, std::vector<CondSpec>& conds,
                 int& per_sample_floats, int& n_cond) {
  const int n = cfg.n_levels;
  std::vector<int> dims(n + 1);
  dims[0] = cfg.state_dim;
  for (int i = 0; i < n; ++i) dims[i + 1] = cfg.unet_input_dim * cfg.dim_mults[i];
  std::vector<int> L(n);
  for (int i = 0; i < n; ++i) L[i] = cfg.horizon >> i;

  // activation buffers (floats per sample)
  int rot = dims[0] * (cfg.horizon + 4);
  for (int i = 0; i < n; ++i) rot = std::max(rot, dims[i + 1] * (L[i] + 4));
  for (int i = 1; i < n; ++i) rot = std::max(rot, dims[i] * (L[i - 1] + 4));  // upsample outputs
  rot = (rot + 3) & ~3;
  int R[3] = {0, rot, 2 * rot};
  int used = 3 * rot;
  std::vector<int> cat(n, -1);
  for (int i = 1; i < n; ++i) {
    cat[i] = used;
    used += (2 * dims[i + 1] * (L[i] + 4) + 3) & ~3;
  }
  per_sample_floats = used;
  n_cond = 0;
  ops.clear();
  conds.clear();
  std::map<int, int> prod;   // activation buffer offset -> index of the op that last wrote it (-1: network input)
  prod[R[0]] = -1;
  auto wire = [&](Op& o, int pitch) {
    o.p_src0 = prod.count(o.src) ? prod[o.src] : -1;
    o.p_src1 = -2;
    o.c_src0 = o.cin;
    if (o.p_src0 >= 0 && ops[o.p_src0].cout < o.cin) {   // channel concat: second producer wrote the upper half
      o.c_src0 = ops[o.p_src0].cout;
      o.p_src1 = prod[o.src + o.c_src0 * pitch];
    }
    o.p_res = (o.res_src >= 0) ? (prod.count(o.res_src) ? prod[o.res_src] : -1) : -2;
    o.p_res1 = -2;
    o.c_res0 = o.res_cin;
    if (o.p_res >= 0 && ops[o.p_res].cout < o.res_cin) {
      o.c_res0 = ops[o.p_res].cout;
      o.p_res1 = prod[o.res_src + o.c_res0 * pitch];
    }
  };

  auto convblock = [&](const std::string& pre, int cin, int cout, int len, int src, int dst, int cond_off, int res_src,
                       int res_cin, const std::string& res_key) -> bool {
    Op o{};
    o.type = OP_CONVBLOCK; o.cin = cin; o.cout = cout; o.lin = len; o.lout = len; o.src = src; o.dst = dst;
    o.w = bld.conv(pre + ".block.0.weight", cin, cout, 5, false);
    o.b = bld.vec(pre + ".block.0.bias", cout);
    o.gn_w = bld.vec(pre + ".block.2.weight", cout);
    o.gn_b = bld.vec(pre + ".block.2.bias", cout);
    o.n_groups = gn_groups(cout);
    o.cond = cond_off; o.res_src = res_src; o.res_cin = res_cin; o.res_w = -1; o.res_b = -1;
    if (!res_key.empty()) {
      o.res_w = bld.conv(res_key + ".weight", res_cin, cout, 1, false);
      o.res_b = bld.vec(res_key + ".bias", cout);
    }
    wire(o, len + 4);
    prod[dst] = (int)ops.size();
    ops.push_back(o);
    return bld.err.empty();
  };
  // ResidualTemporalBlock (layers.py:326-358): x in `xin`, scratch `tmp`, result in `out`
  auto rtb = [&](const std::string& pre, int cin, int cout, int len, int xin, int tmp, int out) -> bool {
    int coff = n_cond;
    conds.push_back({pre + ".cond_mlp.1", cout, coff});
    n_cond += cout;
    if (!convblock(pre + ".blocks.0", cin, cout, len, xin, tmp, coff, -1, 0, "")) return false;
    return convblock(pre + ".blocks.1", cout, cout, len, tmp, out, -1, xin, cin,
                     cin != cout ? pre + ".residual_conv" : "");
  };

  // Residual(PreNorm(LinearAttention(dim))) in place on `buf` (keys: <pre>.fn.norm.{g,b}, <pre>.fn.fn.to_qkv.weight,
  // <pre>.fn.fn.to_out.{weight,bias})
  auto attn = [&](const std::string& pre, int C, int len, int buf) -> bool {
    Op o{};
    o.type = OP_ATTN; o.cin = C; o.cout = C; o.lin = len; o.lout = len; o.src = buf; o.dst = buf;
    o.w = bld.conv(pre + ".fn.fn.to_qkv.weight", C, 384, 1, false);        // [C][384]
    o.res_w = bld.conv(pre + ".fn.fn.to_out.weight", 128, C, 1, false);    // [128][C]
    o.b = bld.vec(pre + ".fn.fn.to_out.bias", C);
    o.gn_w = bld.vec(pre + ".fn.norm.g", C);
    o.gn_b = bld.vec(pre + ".fn.norm.b", C);
    o.res_src = -1; o.cond = -1; o.res_b = -1;
    wire(o, len + 4);
    prod[buf] = (int)ops.size();
    ops.push_back(o);
    return bld.err.empty();
  };
  int cur = R[0];  // input lives in R[0]
  int ri = 0;      // index of rotating buffer holding `cur` (when cur is a rotating buffer)
  auto other = [&](int a, int b) { for (int k = 0; k < 3; ++k) if (R[k] != a && R[k] != b) return R[k]; return -1; };

  for (int i = 0; i < n; ++i) {
    const std::string p = "downs." + std::to_string(i);
    int t1 = other(cur, -1);
    int o1 = other(cur, t1);
    if (!rtb(p + ".0", dims[i], dims[i + 1], L[i], cur, t1, o1)) return MMDK_EINVAL;
    cur = o1;
    int t2 = other(cur, -1);
    int o2 = (i >= 1) ? cat[i] + dims[i + 1] * (L[i] + 4) : other(cur, t2);  // skip goes to the upper half of CAT[i]
    if (!rtb(p + ".1", dims[i + 1], dims[i + 1], L[i], cur, t2, o2)) return MMDK_EINVAL;
    cur = o2;
    if (cfg.self_attention && !attn("downs." + std::to_string(i) + ".2", dims[i + 1], L[i], cur)) return MMDK_EINVAL;
    if (i < n - 1) {
      Op o{};
      o.type = OP_DOWN; o.cin = dims[i + 1]; o.cout = dims[i + 1]; o.lin = L[i]; o.lout = L[i + 1];
      o.src = cur; o.dst = other(cur, -1);
      o.w = bld.conv(p + ".4.conv.weight", dims[i + 1], dims[i + 1], 3, false);
      o.b = bld.vec(p + ".4.conv.bias", dims[i + 1]);
      o.res_src = -1;
      wire(o, L[i] + 4);
      prod[o.dst] = (int)ops.size();
      ops.push_back(o);
      cur = o.dst;
    }
  }
  (void)ri;
  {
    const int C = dims[n], len = L[n - 1];
    int t1 = other(cur, -1), o1 = other(cur, t1);
    if (!rtb("mid_block1", C, C, len, cur, t1, o1)) return MMDK_EINVAL;
    cur = o1;
    if (cfg.self_attention && !attn("mid_attn", C, len, cur)) return MMDK_EINVAL;
    int t2 = other(cur, -1);
    int o2 = (n >= 2) ? cat[n - 1] : other(cur, t2);  // lower half of CAT[n-1]
    if (!rtb("mid_block2", C, C, len, cur, t2, o2)) return MMDK_EINVAL;
    cur = o2;
  }
  for (int j = 0; j < n - 1; ++j) {
    const int lvl = n - 1 - j;            // resolution level of this up stage
    const int cin2 = 2 * dims[lvl + 1];   // concat(x, skip)
    const int cmid = dims[lvl];           // dim_in of in_out[lvl]
    const std::string p = "ups." + std::to_string(j);
    // cur == cat[lvl] (x in the lower half, skip in the upper half)
    int t1 = R[0], o1 = R[1];
    if (!rtb(p + ".0", cin2, cmid, L[lvl], cat[lvl], t1, o1)) return MMDK_EINVAL;
    cur = o1;
    int t2 = other(cur, -1), o2 = other(cur, t2);
    if (!rtb(p + ".1", cmid, cmid, L[lvl], cur, t2, o2)) return MMDK_EINVAL;
    cur = o2;
    if (cfg.self_attention && !attn(p + ".2", cmid, L[lvl], cur)) return MMDK_EINVAL;
    Op o{};
    o.type = OP_UP; o.cin = cmid; o.cout = cmid; o.lin = L[lvl]; o.lout = L[lvl - 1];
    o.src = cur;
    o.dst = (lvl - 1 >= 1) ? cat[lvl - 1] : other(cur, -1);
    o.w = bld.conv(p + ".4.conv.weight", cmid, cmid, 4, true);
    o.b = bld.vec(p + ".4.conv.bias", cmid);
    o.res_src = -1;
    wire(o, L[lvl] + 4);
    prod[o.dst] = (int)ops.size();
    ops.push_back(o);
    cur = o.dst;
  }
  {
    int t1 = other(cur, -1);
    if (!convblock("final_conv.0", cfg.unet_input_dim, cfg.unet_input_dim, cfg.horizon, cur, t1, -1, -1, 0, ""))
      return MMDK_EINVAL;
    Op o{};
    o.type = OP_FINAL; o.cin = cfg.unet_input_dim; o.cout = cfg.state_dim; o.lin = cfg.horizon; o.lout = cfg.horizon;
    o.src = t1; o.dst = 0;
    o.w = bld.conv("final_conv.1.weight", cfg.unet_input_dim, cfg.state_dim, 1, false);
    o.b = bld.vec("final_conv.1.bias", cfg.state_dim);
    o.res_src = -1;
    wire(o, cfg.horizon + 4);
    ops.push_back(o);
  }
  if (!bld.err.empty()) return MMDK_EINVAL;
  if ((int)ops.size() > kMaxOps) { bld.err = "too many ops"; return MMDK_EINVAL; }
  return MMDK_OK;
}

int unet_create(const mmdk_unet_config* cfg, int n_tensors, const char* const* names, const float* const* tensors,
                const int64_t* numels, cudaStream_t stream, UnetImpl** out) {
  if (!cfg || !out) return fail(MMDK_EINVAL, "null argument");
  if (cfg->n_levels < 1 || cfg->n_levels > MMDK_MAX_LEVELS) return fail(MMDK_EINVAL, "n_levels out of range");
  if (cfg->horizon % (8 << (cfg->n_levels - 1)) != 0)
    return fail(MMDK_EINVAL, "horizon must be a multiple of 8 * 2^(n_levels-1)");
  if (cfg->time_emb_dim > 64 || cfg->time_emb_dim < 1) return fail(MMDK_EINVAL, "time_emb_dim out of range");
  std::map<std::string, std::pair<const float*, int64_t>> T;
  for (int i = 0; i < n_tensors; ++i) T[names[i]] = {tensors[i], numels[i]};

  auto* net = new UnetImpl();
  net->cfg = *cfg;
  // pass 1: sizes only
  Builder dry{T, {}, stream};
  std::vector<CondSpec> conds;
  int rc = lower(dry, *cfg, net->ops, conds, net->per_sample_floats, net->n_cond);
  if (rc != MMDK_OK) { std::string e = dry.err; delete net; return fail(rc, e); }
  net->blob_floats = dry.blob_used;
  if (check_cuda(cudaMalloc(&net->blob, sizeof(float) * net->blob_floats), "cudaMalloc(weights)") != MMDK_OK) {
    delete net; return MMDK_ECUDA;
  }
  // pass 2: repack into the blob
  Builder real{T, {}, stream};
  real.blob = net->blob;
  rc = lower(real, *cfg, net->ops, conds, net->per_sample_floats, net->n_cond);
  if (rc != MMDK_OK) { std::string e = real.err; unet_destroy(net); return fail(rc, e); }

#define MMDK_CUDA_NET(call)                                                     \
  do {                                                                          \
    int _rc = ::mmdk::check_cuda((call), #call);                                \
    if (_rc != MMDK_OK) { unet_destroy(net); return _rc; }                      \
  } while (0)
  MMDK_CUDA_NET(cudaMalloc(&net->ops_dev, sizeof(Op) * net->ops.size()));
  MMDK_CUDA_NET(cudaMemcpyAsync(net->ops_dev, net->ops.data(), sizeof(Op) * net->ops.size(), cudaMemcpyHostToDevice, stream));

  // cond table for all t
  const int Tn = cfg->n_diffusion_steps;
  if (Tn < 1) { unet_destroy(net); return fail(MMDK_EINVAL, "n_diffusion_steps must be >= 1"); }
  auto sc = T.find("__sincos_table__");
  if (sc == T.end() || sc->second.second != (int64_t)Tn * 32) {
    unet_destroy(net);
    return fail(MMDK_EINVAL, "missing '__sincos_table__' [T,32] (SinusoidalPosEmb evaluated by the host, layers.py:246-258)");
  }
  std::vector<CondLayer> cl;
  for (auto& c : conds) {
    auto w = T.find(c.prefix + ".weight");
    auto b = T.find(c.prefix + ".bias");
    if (w == T.end() || b == T.end() || w->second.second != (int64_t)c.C * cfg->time_emb_dim || b->second.second != c.C) {
      unet_destroy(net);
      return fail(MMDK_EINVAL, "missing/wrong cond_mlp tensor " + c.prefix);
    }
    cl.push_back({w->second.first, b->second.first, c.C, c.off});
  }
  const char* tk[4] = {"time_mlp.encoder.1.weight", "time_mlp.encoder.1.bias", "time_mlp.encoder.3.weight",
                       "time_mlp.encoder.3.bias"};
  const int64_t tn[4] = {128 * 32, 128, (int64_t)cfg->time_emb_dim * 128, cfg->time_emb_dim};
  const float* tp[4];
  for (int i = 0; i < 4; ++i) {
    auto it = T.find(tk[i]);
    if (it == T.end() || it->second.second != tn[i]) { unet_destroy(net); return fail(MMDK_EINVAL, std::string("missing/wrong ") + tk[i]); }
    tp[i] = it->second.first;
  }
  CondLayer* cl_dev = nullptr;
  MMDK_CUDA_NET(cudaMalloc(&cl_dev, sizeof(CondLayer) * cl.size()));
  if (check_cuda(cudaMemcpyAsync(cl_dev, cl.data(), sizeof(CondLayer) * cl.size(), cudaMemcpyHostToDevice, stream),
                 "cudaMemcpyAsync(cond layers)") != MMDK_OK ||
      check_cuda(cudaMalloc(&net->cond_table, sizeof(float) * (size_t)Tn * net->n_cond), "cudaMalloc(cond table)") != MMDK_OK) {
    cudaFree(cl_dev);
    unet_destroy(net);
    return MMDK_ECUDA;
  }
  time_embed_kernel<<<Tn, 128, 0, stream>>>(sc->second.first, tp[0], tp[1], tp[2], tp[3], cl_dev, (int)cl.size(),
                                            cfg->time_emb_dim, net->n_cond, net->cond_table);
  {
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(stream);  // cl (host vector) and cl_dev must outlive the kernel
    cudaFree(cl_dev);
    if (check_cuda(e1, "time_embed_kernel") != MMDK_OK || check_cuda(e2, "cudaStreamSynchronize") != MMDK_OK) {
      unet_destroy(net);
      return MMDK_ECUDA;
    }
  }

  // executor configuration: as many samples per CTA as shared memory allows (<= 4)
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  net->attn_scratch_floats = cfg->self_attention ? (384 * cfg->horizon + 4 * 32 * 32 + 2 * cfg->horizon) : 0;
  int S = (int)(((size_t)max_smem - sizeof(float) * net->attn_scratch_floats) / (sizeof(float) * net->per_sample_floats));
  if (S < 1) { unet_destroy(net); return fail(MMDK_EINVAL, "network activations do not fit in shared memory"); }
  net->ffma_S = S > 3 ? 3 : S;
  *out = net;
  return MMDK_OK;
}

void unet_destroy(UnetImpl* net) {
  if (!net) return;
  cudaFree(net->blob);
  cudaFree(net->ops_dev);
  cudaFree(net->cond_table);
  unet_tc_release(net);
  unet_fused_release(net);
  delete net;
}

template <int S>
static int launch_ffma(const UnetImpl* net, const float* x, int B, int t, float* eps, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)S * net->per_sample_floats + net->attn_scratch_floats);
  {  // function attributes are per device: opt in to the device maximum once per device
    int dev = 0;
    MMDK_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return fail(MMDK_EINVAL, "device index out of range");
    if (!configured[dev]) {
      int max_optin = 0;
      MMDK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      MMDK_CUDA(cudaFuncSetAttribute(unet_ffma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
      configured[dev] = true;
    }
  }
  const int grid = (B + S - 1) / S;
  unet_ffma_kernel<S><<<grid, 256, smem, stream>>>(net->ops_dev, (int)net->ops.size(), net->blob,
                                                    net->cond_table + (size_t)t * net->n_cond, x, eps, B,
                                                    net->cfg.horizon, net->cfg.state_dim, net->per_sample_floats);
  return check_cuda(cudaGetLastError(), "unet_ffma_kernel launch");
}

int unet_forward_ffma(const UnetImpl* net, const float* x, int B, int t, float* eps, cudaStream_t stream) {
  switch (net->ffma_S) {
    case 1: return launch_ffma<1>(net, x, B, t, eps, stream);
    case 2: return launch_ffma<2>(net, x, B, t, eps, stream);
    default: return launch_ffma<3>(net, x, B, t, eps, stream);
  }
}

}  // namespace mmdk
