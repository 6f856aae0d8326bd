// Shared device helpers of the tcgen05 executors (unet_tc.cu: one launch per layer; unet_fused.cu: one persistent launch
// per forward): PTX wrappers (mbarrier, bulk async copy, tcgen05 mma / ld / commit), the "tile image" activation format
// and the weight-chunk packing.  See unet_tc.cu for the layout description.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "unet.cuh"

namespace mmdk {
namespace {

constexpr int ST = 7;          // samples per tile (per-layer executor; the persistent executor picks 7 / 3 / 1 per batch size)
constexpr int MAX_W_STAGES = 3;
constexpr int TC_THREADS = 320;
constexpr int TMEM_COLS = 256;
constexpr int MAX_SRC = 4;
constexpr int MAX_CHUNKS = 96;   // 5 taps x 16 (32-channel) chunks + 16 residual chunks: the 512-channel concat block of the 4-level net
constexpr int TC_CLUSTER = 1;    // >1: CTAs per cluster sharing every weight chunk through TMA multicast.  Measured on B200 (r01): the main
                                 // loop is tensor-pipe bound (two resident CTAs share the pipe), not L2 bound, and 4-CTA clusters cost
                                 // scheduling tails (1.85 vs 1.29 ms per forward) -> off by default, code path kept.

struct ChunkDesc {
  uint32_t a_off;    // smem byte offset of the chunk's first A panel (plane hi, buffer row 0)
  uint32_t a_plane;  // bytes from the hi plane to the lo plane of that source image
  uint32_t a_lbo;    // bytes between K panels of that source image (ROWS * 16)
  uint32_t w_off;    // byte offset of the packed weights of this chunk
  uint32_t w_bytes;
  int d;             // row shift of this tap
  int acc;           // accumulator region (0: columns [0,128), 1: [128,256))
  int first;         // first chunk accumulated into this region
  int k16;           // 16-channel K steps in this chunk (1 or 2)
  int phase;         // input phase this chunk reads (two-phase ops reload the input buffer between phases)
};

enum TcKind : int { TC_CONVBLOCK = 0, TC_DOWN = 1, TC_UP = 2, TC_FINAL = 3, TC_ATTN_QKV = 4, TC_ATTN_CORE = 5, TC_ATTN_OUT = 6 };

struct TcOpParams {
  int kind, L, P, n_mt, rows, N, cout, B, n_tiles;
  const uint8_t* src[MAX_SRC];
  uint32_t src_tile_bytes[MAX_SRC];
  uint32_t src_smem_off[MAX_SRC];
  int n_src;
  int src_phase[MAX_SRC];   // input phase that loads this source (all 0 unless the sources do not fit shared memory together)
  int n_phase;              // 1, or 2: phase 1's sources overwrite phase 0's once its MMAs have completed
  int n_split;              // output channels cut into n_split slices of N (gridDim.y): slice s owns channels [s N, (s + 1) N)
  uint32_t w_split_bytes;   // packed weight bytes per slice
  const uint8_t* wchunks;
  int n_chunks;
  uint32_t w_stage_bytes;
  int w_stages;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* cond;      // row of this timestep or nullptr
  const float* res_bias;  // residual 1x1 conv bias (region 1) or nullptr
  const uint8_t* res_id;  // identity residual image (global) or nullptr
  uint32_t res_id_tile_bytes;
  int res_id_rows, res_id_C;
  uint8_t* out;
  uint32_t out_tile_bytes;
  int out_L, out_rows, out_C;
  float* eps;
  uint32_t smem_w_off, smem_scratch_off, smem_bar_off;
  int cluster;      // CTAs per cluster sharing every weight chunk through TMA multicast (1 = off)
  long long* dbg;   // optional per-CTA timeline [n_tiles][16] (clock64), nullptr in production
  ChunkDesc chunks[MAX_CHUNKS];   // in the kernel parameter (constant) bank: no dependent global loads per chunk
};

// ------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: park, do not spin
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {   // long waits: do not hog issue slots
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, no swizzle: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// tcgen05.mma with the 64-bit shared-memory descriptors given as (low word, shared high word): K-major, no swizzle,
// SBO = 128 B and descriptor version 1 live in the high word (0x4008), start address >> 4 and LBO >> 4 in the low word,
// so walking taps / m-tiles / K steps / hi-lo planes is one 32-bit add on the low word.  Executed by the whole (converged)
// warp with uniform operands; only the elected lane issues.
constexpr uint32_t F_DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ void mma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(leader), "r"(F_DESC_HI) : "memory");
}

__device__ __forceinline__ float mish_fast(float y) {
  // y * tanh(softplus(y)) = y * n / (n + 2), n = e^y (e^y + 2); for y >= 20 the ratio is exactly 1 in fp32, so clamping
  // the exponent's argument replaces the overflow guard
  float e = __expf(fminf(y, 20.f));
  float n = e * (e + 2.f);
  return y * __fdividef(n, n + 2.f);
}

// 8 fp32 -> 8 fp16 hi (uint4) + 8 fp16 lo (uint4)
__device__ __forceinline__ void split8(const float* y, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half a = __float2half_rn(y[2 * i]), b = __float2half_rn(y[2 * i + 1]);
    h[i] = __halves2half2(a, b);
    l[i] = __halves2half2(__float2half_rn(y[2 * i] - __half2float(a)), __float2half_rn(y[2 * i + 1] - __half2float(b)));
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                  *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                  *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}
__device__ __forceinline__ void add8(const uint4& hi, const uint4& lo, float* y) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 a = __half22float2(h[i]), b = __half22float2(l[i]);
    y[2 * i] += a.x + b.x;
    y[2 * i + 1] += a.y + b.y;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// setup kernels
// ------------------------------------------------------------------------------------------------------------------
// one weight chunk: W [cin][ktaps][cout] fp32 -> [plane][CK/8][N][8] fp16 (hi | lo)
static __global__ void pack_wchunk_kernel(const float* __restrict__ W, int cin, int ktaps, int cout, int tap, int ci0, int CK,
                                   int N, __half* __restrict__ dst, int n0 = 0, const float* __restrict__ scale = nullptr) {
  // n0: first output channel of this N-column slice (1x1 convs wider than one MMA N are cut into slices); scale: optional
  // per-input-channel factor folded into the weights (LayerNorm gain of the attention block's PreNorm)
  const int n_el = CK * N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_el; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, n = (idx >> 3) % N, kp = idx / (8 * N);
    const int ci = ci0 + kp * 8 + e;
    float w = (ci < cin && n0 + n < cout) ? W[((size_t)ci * ktaps + tap) * cout + n0 + n] : 0.f;
    if (scale != nullptr && ci < cin) w *= scale[ci];
    __half h = __float2half_rn(w);
    dst[idx] = h;
    dst[n_el + idx] = __float2half_rn(w - __half2float(h));
  }
}

// LinearAttention PreNorm folding (layers.py:186-229): qkv = W^T LN(x) = rstd (x^T G - mean s) + u with G[c][n] = g[c] W[c][n],
// s[n] = sum_c G[c][n], u[n] = sum_c b[c] W[c][n].  W: [C][n_out] fp32 (packed 1x1 conv), g / b: [C].
static __global__ void attn_fold_kernel(const float* __restrict__ W, const float* __restrict__ g, const float* __restrict__ b, int C,
                                        int n_out, float* __restrict__ s_out, float* __restrict__ u_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_out) return;
  float s = 0.f, u = 0.f;
  for (int c = 0; c < C; ++c) {
    const float w = W[(size_t)c * n_out + n];
    s = fmaf(g[c], w, s);
    u = fmaf(b[c], w, u);
  }
  s_out[n] = s;
  u_out[n] = u;
}

// network input x [B][L][D] fp32 -> level-0 image with C = 16 (channels D..15 stay zero)
static __global__ void pack_input_kernel(const float* __restrict__ x, int B, int L, int D, int rows, uint8_t* __restrict__ img,
                                  uint32_t tile_bytes, int st = ST) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * L) return;
  const int b = idx / L, pos = idx - b * L;
  const int tile = b / st, s = b - tile * st;
  const int r = 2 + s * (L + 2) + pos;
  float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int d = 0; d < D && d < 8; ++d) y[d] = x[(size_t)idx * D + d];
  uint4 hi, lo;
  split8(y, hi, lo);
  uint8_t* ob = img + (size_t)tile * tile_bytes + (size_t)r * 16;   // panel 0
  *reinterpret_cast<uint4*>(ob) = hi;
  *reinterpret_cast<uint4*>(ob + (size_t)2 * rows * 16) = lo;        // plane stride = (16/8) panels * rows * 16
}

// debug tap: image -> fp32 [B][C][L]
static __global__ void unpack_image_kernel(const uint8_t* __restrict__ img, uint32_t tile_bytes, int B, int C, int L, int rows,
                                    float* __restrict__ out, int st = ST) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C * L) return;
  const int pos = idx % L, c = (idx / L) % C, b = idx / (L * C);
  const int tile = b / st, s = b - tile * st;
  const int r = 2 + s * (L + 2) + pos;
  const __half* base = reinterpret_cast<const __half*>(img + (size_t)tile * tile_bytes);
  const size_t o = ((size_t)(c >> 3) * rows + r) * 8 + (c & 7);
  const size_t plane = (size_t)(C / 8) * rows * 8;
  out[idx] = __half2float(base[o]) + __half2float(base[plane + o]);
}



inline int level_rows(int L, int st = ST) { return 2 + 128 * ((st * (L + 2) + 127) / 128) + 2; }

}  // namespace

struct TcImage {
  uint8_t* dev = nullptr;
  int C = 0, L = 0, rows = 0;
  uint32_t tile_bytes = 0;
};

}  // namespace mmdk
