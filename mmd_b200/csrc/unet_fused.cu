// Persistent whole-forward tcgen05 executor of the TemporalUnet: ONE launch per forward.
//
// Reference semantics: mmd/models/layers/layers.py:261-358, temporal_unet.py:121-174 (the same Op program and the same
// "tile image" activation format as unet_tc.cu, which runs one launch per layer and is kept as the A/B baseline).
//
// Why: the per-layer executor (round 1) serialises load -> MMA -> epilogue inside a CTA (tensor pipe idle ~45% of a CTA's
// lifetime) and pays 34 launches per forward.  Here every CTA owns up to two tiles (2 x 7 samples) at a time and walks
// the whole layer program tile-major:
//     item = (tile parity, op);  items alternate between the two tiles
// so the tensor pipe works on tile Y's op j while the 16 epilogue warps finish tile X's op j (and vice versa): a two-stage
// software pipeline with in-order queues between four warp roles.  Layer-to-layer dependencies are tile-local, so no
// grid-wide synchronisation exists: the epilogue stores an op's output image to global memory (it stays in the 126 MB
// L2), fences it for the async proxy and arrives on an mbarrier; the input producer then bulk-copies (TMA, UBLKCP) it
// back as the next op's A operand.
//
//   warps 0-15  epilogue: thread = (TMEM lane = image row, quarter of the columns); the 32 (or 16) accumulator values of
//               a thread are read ONCE into registers (TMEM is released immediately), GroupNorm statistics via per-row
//               (sum, M2) partials in shared memory + an 8-thread-per-(sample, group) Chan combination, then
//               normalise -> Mish -> +cond -> +residual -> FP16 hi/lo split -> 16-byte stores
//   warp 16     tcgen05.mma issuer (one thread), accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 17     weight producer: streams every op's packed weight chunks through a 3-stage ring, continuously across ops
//   warp 18     input producer: bulk-copies the input image(s) of the next item into one of two shared-memory buffers
//
// The residual 1x1 convolution of a ResidualTemporalBlock (layers.py:353-356) is evaluated by the block's FIRST conv
// op (same input x) into a second accumulator region and stored as its own image r; the second conv op then adds r in
// its epilogue like an identity residual, so no op ever needs more than one input image set in shared memory.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "tc_common.cuh"

namespace mmdk {

namespace {

constexpr int FE_WARPS = 16;
constexpr int F_WARP_MMA = 16, F_WARP_W = 17, F_WARP_IN = 18;
constexpr int F_THREADS = 19 * 32;
constexpr int F_STAGES = 3;
constexpr uint32_t F_STAGE_BYTES = 16384;
constexpr int F_PART_ROWS = 512;
constexpr int F_MAX_N = 128;

// epilogue variants of the conv block: (m-tiles per image, columns per m-tile)
enum FVariant : int { FV_4_32 = 0, FV_2_64 = 1, FV_1_128 = 2, FV_2_32 = 3, FV_1_64 = 4, FV_GENERIC = 5, FV_1_32 = 6 };

struct FOp {
  int kind, L, P, n_mt, N, cout, variant;
  int n_src;
  const uint8_t* src[MAX_SRC];
  uint32_t src_tile_bytes[MAX_SRC];
  uint32_t src_smem_off[MAX_SRC];
  uint32_t src_lbo[MAX_SRC];    // bytes between K panels of the source image (rows * 16)
  uint32_t src_plane[MAX_SRC];  // bytes from its hi plane to its lo plane
  int src_by_tile[MAX_SRC];   // 1: image indexed by tile (the packed network input), 0: by the CTA's slot
  uint32_t in_bytes;
  int big;                    // inputs need both buffers
  const uint8_t* wchunks;
  int chunk_base, n_chunks;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* res_bias;      // bias of the residual 1x1 conv evaluated by this op into region 1 (-> out2), or nullptr
  int cond_off;               // offset into the cond row of the timestep, or -1
  const uint8_t* res_id;      // image added in the epilogue (identity residual x, or the r image), or nullptr
  uint32_t res_id_tile_bytes;
  int res_id_rows, res_id_C;
  uint8_t* out;
  uint32_t out_tile_bytes;
  int out_L, out_rows, out_C;
  uint8_t* out2;
  uint32_t out2_tile_bytes;
  int out2_rows, out2_C;
  // in-place activation hand-over inside shared memory: this op's epilogue writes its output image straight into the tile's
  // (now free) input buffer, where the next op of the same tile reads it as its A operand -- no L2 round trip on the tile's
  // critical path; the global image is written only when a later op needs it (identity residual, skip concat, debug taps)
  // LinearAttention (layers.py:210-229) on this executor: TC_ATTN_QKV ops (one N-slice of the to_qkv 1x1 conv each; the PreNorm
  // LayerNorm is folded into the weights and undone per row in the epilogue) write fp32 q / k / v rows into per-slot scratch,
  // TC_ATTN_CORE (no MMA) does softmax over positions, the 32x32 context per (sample, head) and context^T q, TC_ATTN_OUT is the
  // to_out 1x1 conv + bias + residual.
  float* aux0;      // QKV: destination scratch of this slice (already offset by the slice's first column); CORE: q
  float* aux1;      // CORE: k
  float* aux2;      // CORE: v
  int aux_ld;       // floats per scratch row (128) ; QKV: unused
  int src_C;        // channels of the (single) input image: QKV needs them for the LayerNorm statistics
  int q_base;       // ATTN_OUT over a half-height `out` image (4 m-tile levels): first tile row this op covers (0 / 256)
  int in_smem;      // input already in the parity's buffer (placed by the previous op's epilogue): the producer copies nothing
  int out_smem;     // write the output image into the parity's buffer
  int out_global;   // write the output image to global memory
};

// weight chunk, 16 bytes (one uniform constant load per chunk); everything the MMA issuer needs is precomputed on the host
struct __align__(16) FChunk {
  uint32_t a_w;      // low word of the A descriptor at (m-tile 0, K step 0, hi plane), RELATIVE to the item's input buffer:
                     // ((a_off + (2 + tap shift) * 16) >> 4) | (LBO >> 4) << 16; the issuer adds (buffer base >> 4)
  uint32_t a_ks_pl;  // (bytes per K step of A) >> 4  |  ((hi -> lo plane distance) >> 4) << 16
  uint32_t w_off;    // byte offset of the packed weights relative to the op's wchunks
  uint32_t meta;     // acc | first << 1 | k16 << 2
};
constexpr int F_MAX_OPS = 64, F_MAX_CHUNKS = 512;

// The op and chunk tables live in the kernel PARAMETER bank (16 KB of the 32 KB CUDA 12 allows): indexed by warp-uniform
// loop counters they are read with uniform constant loads, so the MMA issuer's descriptor arithmetic stays in uniform
// registers (round-2 timeline: with the tables in global memory the issue loop, not the tensor pipe, set the pace:
// ~120 cycles per tcgen05.mma).
struct FParams {
  int n_ops, n_tiles, B;
  int st;                     // samples per tile: 7, or 3 / 1 for batches that then still fit one tile per SM (build_fused)
  int by_slot;                // 1: intermediate images are indexed by (CTA, parity) slot (L2-resident footprint), 0: by tile
  const float* cond_row;
  float* eps;
  uint32_t buf_bytes;
  uint32_t off_ring, off_prm, off_part, off_stat, off_bar;
  int dbg_flags;              // timing experiments only (MMDK_FUSED_DEBUG): 1 = no weight copies after the first ring fill, 2 = hi*hi term only
  long long* dbg;             // optional timeline of CTA 0: [item][16] clock64 stamps (see tests/tc_timeline.py), nullptr in production
  unsigned long long* stamp;  // optional launch-duration probe: [grid][2] %globaltimer (ns) at CTA start / end (mmdk_unet_debug_stamps)
  FOp ops[F_MAX_OPS];
  FChunk chunks[F_MAX_CHUNKS];
};

static_assert(sizeof(FParams) <= 32764, "kernel parameter bank (CUDA 12: 32,764 bytes)");

// barrier slots
constexpr int B_IN_FULL = 0, B_IN_EMPTY = 2, B_ACC_FULL = 4, B_ACC_EMPTY = 6, B_OUT_DONE = 8, B_W_FULL = 10,
              B_W_EMPTY = 10 + F_STAGES, B_PRM_FULL = 10 + 2 * F_STAGES, B_COUNT = 12 + 2 * F_STAGES;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar16() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int NH>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float* v) {
  if constexpr (NH == 32) tmem_ld32(taddr, v);
  else if constexpr (NH == 16) tmem_ld16(taddr, v);
  else if constexpr (NH == 8) tmem_ld8(taddr, v);
  else tmem_ld4(taddr, v);
}

__device__ __forceinline__ uint4 ld_cg_u4(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

struct EpiCtx {
  int tid, warp, lane, q4, cb, row;
  // per-channel parameters of this item, structure of arrays (each [F_MAX_N] floats): one LDS.128 serves 4 channels
  float* p_bias;
  float* p_gamma;
  float* p_beta;
  float* p_cond;
  float* p_rb;       // residual-conv bias
  float2* part;      // [8][F_PART_ROWS] (sum, M2) per group and image row
  float2* stat;      // [ST * 8] (-mean * rstd, rstd) of this item
  uint32_t tmem;     // TMEM address of this item's accumulators (lane 0, first column)
  uint8_t* sbuf;     // the parity's input buffer (shared memory), target of the in-place output
  const uint8_t* inbuf;   // shared-memory buffer holding THIS item's input image(s)
  int slot;          // (CTA, parity) index of the per-slot scratch
  int tile, img, B, st;   // st = samples per tile of this launch
  long long* dbg;    // this item's stamp row (CTA 0, thread 0 only) or nullptr
};

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2: two lanes per issue slot) and approx transcendental wrappers -----
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Mish on two values: y * tanh(softplus(y)) = y * n / (n + 2), n = e^y (e^y + 2); for y >= 20 the ratio is exactly 1 in
// fp32, so clamping the exponent's argument replaces the overflow guard
__device__ __forceinline__ float2 mish2(float2 y) {
  const float2 t = f2mul(make_float2(fminf(y.x, 20.f), fminf(y.y, 20.f)), make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 e = make_float2(ex2_ftz(t.x), ex2_ftz(t.y));
  const float2 n = f2mul(e, f2add(e, make_float2(2.f, 2.f)));
  const float2 d = f2add(n, make_float2(2.f, 2.f));
  const float2 r = make_float2(rcp_ftz(d.x), rcp_ftz(d.y));
  return f2mul(y, f2mul(n, r));
}

// 8 fp32 (as 4 float2) -> 8 fp16 hi (uint4) + 8 fp16 lo (uint4)
__device__ __forceinline__ void split8p(const float2* y, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __float22half2_rn(y[i]);
    const float2 back = __half22float2(hh);
    const float2 rem = f2add(y[i], make_float2(-back.x, -back.y));
    const __half2 ll = __float22half2_rn(rem);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void add8p(const uint4& hi, const uint4& lo, float2* y) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = f2add(y[i], f2add(__half22float2(h[i]), __half22float2(l[i])));
}

// Conv1dBlock epilogue (+cond, +residual image), NMT m-tiles of N columns.
template <int NMT, int N>
__device__ __forceinline__ void epi_convblock(const FOp* __restrict__ op, const EpiCtx& c, uint32_t bar_acc_empty) {
  constexpr int NH = N / 4;          // columns per thread and m-tile
  constexpr int CPG = N / 8;         // channels per GroupNorm group
  constexpr int NP = NH / 8;         // 8-channel panels per thread and m-tile
  static_assert(NH == 2 * CPG, "a thread covers exactly two groups");
  const int L = op->L, Pp = op->P;
  const int c0 = c.cb * NH;
  const uint32_t lane_base = c.tmem + ((uint32_t)(c.q4 * 32) << 16);
  const uint32_t pinv = (65536u + (uint32_t)Pp - 1u) / (uint32_t)Pp;   // q / Pp for q < 512, Pp >= 10

  // ---- residual 1x1 conv evaluated by this op (region 1): r = acc1 + bias -> image out2 -------------------------
  if (op->out2 != nullptr) {
    const size_t oplane = (size_t)(op->out2_C / 8) * op->out2_rows * 16;
#pragma unroll 1
    for (int i = 0; i < NMT; ++i) {
      const int q = 128 * i + c.row;
      const int si = (int)(((uint32_t)q * pinv) >> 16), pi = q - si * Pp;
      const bool ok = (si < c.st) && (pi < L) && (c.tile * c.st + si < c.B);
      uint8_t* obase = op->out2 + (size_t)c.img * op->out2_tile_bytes + (size_t)(2 + q) * 16;
#pragma unroll 1
      for (int pc = 0; pc < NP; ++pc) {
        float2 y[4];
        tmem_ld8(lane_base + 128 + i * N + c0 + pc * 8, reinterpret_cast<float*>(y));
        tmem_wait_ld();
        if (ok) {
          const float4 b0 = *reinterpret_cast<const float4*>(c.p_rb + c0 + pc * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(c.p_rb + c0 + pc * 8 + 4);
          y[0] = f2add(y[0], make_float2(b0.x, b0.y)); y[1] = f2add(y[1], make_float2(b0.z, b0.w));
          y[2] = f2add(y[2], make_float2(b1.x, b1.y)); y[3] = f2add(y[3], make_float2(b1.z, b1.w));
          uint4 hi, lo;
          split8p(y, hi, lo);
          uint8_t* ob = obase + (size_t)((c0 >> 3) + pc) * op->out2_rows * 16;
          *reinterpret_cast<uint4*>(ob) = hi;
          *reinterpret_cast<uint4*>(ob + oplane) = lo;
        }
      }
    }
  }
  // ---- accumulators -> registers in NLD load steps; step s + 1 is in flight while the statistics of step s are computed ------
  float2 v[NMT][NH / 2];
  constexpr int NLD = (NMT == 1) ? 2 : NMT;   // one m-tile per step, or the two groups (column halves) of the single m-tile
  auto tmem_issue = [&](int sidx) {
    if constexpr (NMT == 1) tmem_ldn<NH / 2>(lane_base + c0 + sidx * (NH / 2), reinterpret_cast<float*>(&v[0][sidx * (NH / 4)]));
    else tmem_ldn<NH>(lane_base + sidx * N + c0, reinterpret_cast<float*>(v[sidx]));
  };
#ifdef F_PIPE_LD
  tmem_issue(0);
#else
#pragma unroll
  for (int sidx = 0; sidx < NLD; ++sidx) tmem_issue(sidx);
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(bar_acc_empty);
  if (c.dbg) c.dbg[8] = clock64();
#endif

  // rows of this thread (one per m-tile) and the first residual panel: its L2-latency load overlaps the statistics
  const size_t oplane = (size_t)(op->out_C / 8) * op->out_rows * 16;
  const size_t rplane = (size_t)(op->res_id_C / 8) * op->res_id_rows * 16;
  const uint8_t* rimg = op->res_id ? op->res_id + (size_t)c.img * op->res_id_tile_bytes + (size_t)(c0 >> 3) * op->res_id_rows * 16 : nullptr;
  bool okv[NMT];
  int siv[NMT];
#pragma unroll
  for (int i = 0; i < NMT; ++i) {
    const int q = 128 * i + c.row;
    siv[i] = (int)(((uint32_t)q * pinv) >> 16);
    const int pi = q - siv[i] * Pp;
    okv[i] = (siv[i] < c.st) && (pi < L) && (c.tile * c.st + siv[i] < c.B);
  }
  uint4 rh = make_uint4(0, 0, 0, 0), rl = rh;
  if (rimg && okv[0]) {
    rh = ld_cg_u4(rimg + (size_t)(2 + c.row) * 16);
    rl = ld_cg_u4(rimg + (size_t)(2 + c.row) * 16 + rplane);
  }

  // ---- + bias, per-row (sum, M2) of both groups (packed adds; the bias of 4 channels per LDS.128) ----------------------
  {
    float2 b2[NH / 2];
#pragma unroll
    for (int k = 0; k < NH / 4; ++k) {
      const float4 bq = *reinterpret_cast<const float4*>(c.p_bias + c0 + 4 * k);
      b2[2 * k] = make_float2(bq.x, bq.y);
      b2[2 * k + 1] = make_float2(bq.z, bq.w);
    }
    auto group_stats = [&](int i, int g) {
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CPG / 2; ++k) {
        const float2 t = f2add(v[i][g * (CPG / 2) + k], b2[g * (CPG / 2) + k]);
        v[i][g * (CPG / 2) + k] = t;
        s2 = f2add(s2, t);
      }
      const float sm = s2.x + s2.y;
      const float nm = sm * (-1.f / CPG);
      const float2 nm2 = make_float2(nm, nm);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CPG / 2; ++k) {
        const float2 d = f2add(v[i][g * (CPG / 2) + k], nm2);
        q2 = f2fma(d, d, q2);
      }
      c.part[(c.cb * 2 + g) * F_PART_ROWS + 128 * i + c.row] = make_float2(sm, q2.x + q2.y);
    };
#pragma unroll
    for (int sidx = 0; sidx < NLD; ++sidx) {
#ifdef F_PIPE_LD
      tmem_wait_ld();
      if (sidx + 1 < NLD) {
        tmem_issue(sidx + 1);
      } else {   // everything is in registers: hand the accumulator columns back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(bar_acc_empty);
        if (c.dbg) c.dbg[8] = clock64();
      }
#endif
      if constexpr (NMT == 1) {
        group_stats(0, sidx);
      } else {
        group_stats(sidx, 0);
        group_stats(sidx, 1);
      }
    }
  }
  if (c.dbg) c.dbg[9] = clock64();
  epi_bar16();
  if (c.dbg) c.dbg[10] = clock64();
  // ---- Chan combination: 8 threads per (sample, group), partials read ONCE into registers --------------------------------
  if (c.tid < c.st * 8 * 8) {
    const int pair = c.tid >> 3, sub = c.tid & 7;
    const int s = pair >> 3, g = pair & 7;
    const float2* pg = c.part + g * F_PART_ROWS + s * Pp;
    const float inv_n = 1.f / (float)(CPG * L);
    float2 e[8];                       // L <= 64 (checked by the host)
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pp = sub + 8 * j;
      e[j] = (pp < L) ? pg[pp] : make_float2(0.f, 0.f);
      sum += e[j].x;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float mean = sum * inv_n;
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = e[j].x * (1.f / CPG) - mean;
      m2 += (sub + 8 * j < L) ? e[j].y + (float)CPG * d * d : 0.f;
    }
    m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
    if (sub == 0) {
      const float rstd = rsqrtf(m2 * inv_n + 1e-5f);
      c.stat[pair] = make_float2(-mean * rstd, rstd);
    }
  }
  epi_bar16();
  if (c.dbg) c.dbg[11] = clock64();
  const bool out_smem = op->out_smem != 0, out_global = op->out_global != 0;
  if (out_smem) {   // the two rows above and below the m-tiles (conv halos of the first / last sample)
    const int n_pan = op->out_C / 4;   // hi + lo panels
    if (c.tid < n_pan * 4) {
      const int pan = c.tid >> 2, e = c.tid & 3;
      const int r = (e < 2) ? e : op->out_rows - 4 + e;
      *reinterpret_cast<uint4*>(c.sbuf + ((size_t)pan * op->out_rows + r) * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  // ---- normalise, Mish, +cond, +residual image, split, store: one flat sequence of (panel, m-tile) steps -- the per-channel
  // parameters of a panel are loaded once for all m-tiles -- with the residual of the NEXT step already in flight ------------
#pragma unroll
  for (int step = 0; step < NMT * NP; ++step) {
    const int pc = step / NMT, i = step % NMT;
    const uint4 ch = rh, cl = rl;
    if (step + 1 < NMT * NP) {
      const int pc2 = (step + 1) / NMT, i2 = (step + 1) % NMT;
      if (rimg && okv[i2]) {
        const uint8_t* rp = rimg + (size_t)(2 + 128 * i2 + c.row) * 16 + (size_t)pc2 * op->res_id_rows * 16;
        rh = ld_cg_u4(rp);
        rl = ld_cg_u4(rp + rplane);
      }
    }
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (okv[i]) {
      const int si = siv[i];
      const float4 ga = *reinterpret_cast<const float4*>(c.p_gamma + c0 + pc * 8), gb = *reinterpret_cast<const float4*>(c.p_gamma + c0 + pc * 8 + 4);
      const float4 ba = *reinterpret_cast<const float4*>(c.p_beta + c0 + pc * 8), bb = *reinterpret_cast<const float4*>(c.p_beta + c0 + pc * 8 + 4);
      const float4 ca = *reinterpret_cast<const float4*>(c.p_cond + c0 + pc * 8), cbv = *reinterpret_cast<const float4*>(c.p_cond + c0 + pc * 8 + 4);
      const float2 gam[4] = {make_float2(ga.x, ga.y), make_float2(ga.z, ga.w), make_float2(gb.x, gb.y), make_float2(gb.z, gb.w)};
      const float2 bet[4] = {make_float2(ba.x, ba.y), make_float2(ba.z, ba.w), make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};
      const float2 cnd[4] = {make_float2(ca.x, ca.y), make_float2(ca.z, ca.w), make_float2(cbv.x, cbv.y), make_float2(cbv.z, cbv.w)};
      const float2 st0 = c.stat[si * 8 + c.cb * 2], st1 = c.stat[si * 8 + c.cb * 2 + 1];
      float2 y[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = pc * 8 + 2 * e;                   // first of the two channels (both in the same group: CPG is even)
        const float2 st = (k < CPG) ? st0 : st1;
        float2 t = f2fma(v[i][pc * 4 + e], make_float2(st.y, st.y), make_float2(st.x, st.x));
        t = f2fma(t, gam[e], bet[e]);
        y[e] = f2add(mish2(t), cnd[e]);
      }
      if (rimg) add8p(ch, cl, y);
      split8p(y, hi, lo);
    }
    const size_t ooff = (size_t)(2 + 128 * i + c.row) * 16 + (size_t)((c0 >> 3) + pc) * op->out_rows * 16;
    if (okv[i] && out_global) {
      uint8_t* ob = op->out + (size_t)c.img * op->out_tile_bytes + ooff;
      *reinterpret_cast<uint4*>(ob) = hi;
      *reinterpret_cast<uint4*>(ob + oplane) = lo;
    }
    if (out_smem) {   // every image row of the m-tiles is written: rows outside a sample (conv halos, unused samples) as zeros
      *reinterpret_cast<uint4*>(c.sbuf + ooff) = hi;
      *reinterpret_cast<uint4*>(c.sbuf + ooff + oplane) = lo;
    }
  }
  if (c.dbg) c.dbg[12] = clock64();
}

// Down / up resampling convs and the final 1x1 conv: bias only, no normalisation (layers.py:261-276, temporal_unet.py:116-119)
template <bool ATTN>
__device__ __forceinline__ void epi_plain(const FOp* __restrict__ op, const EpiCtx& c, uint32_t bar_acc_empty, float* eps) {
  const int N = op->N, NMT = op->n_mt, L = op->L, Pp = op->P;
  const uint32_t lane_base = c.tmem + ((uint32_t)(c.q4 * 32) << 16);
  const uint32_t pinv = (65536u + (uint32_t)Pp - 1u) / (uint32_t)Pp;
  if (op->kind == TC_FINAL) {
    if (c.cb == 0) {
#pragma unroll 1
      for (int i = 0; i < NMT; ++i) {
        const int q = 128 * i + c.row;
        const int si = (int)(((uint32_t)q * pinv) >> 16), pi = q - si * Pp;
        const bool ok = (si < c.st) && (pi < L) && (c.tile * c.st + si < c.B);
        float y[8];
        tmem_ld8(lane_base + i * N, y);
        tmem_wait_ld();
        if (ok) {
          const size_t b = (size_t)c.tile * c.st + si;
          *reinterpret_cast<float4*>(eps + (b * L + pi) * 4) =
              make_float4(y[0] + c.p_bias[0], y[1] + c.p_bias[1], y[2] + c.p_bias[2], y[3] + c.p_bias[3]);
        }
      }
    }
  } else {
    const int NH = N / 4, c0 = c.cb * NH;
    const int Po = op->out_L + 2;
    const size_t oplane = (size_t)(op->out_C / 8) * op->out_rows * 16;
    const int n_regions = (op->kind == TC_UP) ? 2 : 1;
#pragma unroll 1
    for (int region = 0; region < n_regions; ++region) {
#pragma unroll 1
      for (int i = 0; i < NMT; ++i) {
        const int q = ((ATTN && op->kind == TC_ATTN_OUT) ? op->q_base : 0) + 128 * i + c.row;
        const int si = (int)(((uint32_t)q * pinv) >> 16), pi = q - si * Pp;
        bool ok = (si < c.st) && (pi < L) && (c.tile * c.st + si < c.B);
        int ro;
        if (op->kind == TC_DOWN) {   // stride-2 conv evaluated at every position; keep the even ones
          ok = ok && ((pi & 1) == 0);
          ro = 2 + si * Po + (pi >> 1);
        } else if (ATTN && op->kind == TC_ATTN_OUT) {   // to_out 1x1 conv at the same resolution (+ bias + residual x below)
          ro = 2 + q;
        } else {                      // transposed conv: region 0 -> output 2p, region 1 -> 2p + 1
          ro = 2 + si * Po + 2 * pi + region;
        }
        uint8_t* obase = op->out + (size_t)c.img * op->out_tile_bytes + (size_t)ro * 16;
#pragma unroll 1
        for (int pc = 0; pc < NH / 8; ++pc) {
          const int cbase = c0 + pc * 8;
          float y[8];
          tmem_ld8(lane_base + region * 128 + i * N + cbase, y);
          tmem_wait_ld();
          if (ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] += c.p_bias[cbase + e];
            if (ATTN && op->kind == TC_ATTN_OUT && op->res_id != nullptr) {   // Residual(...): + x
              const uint8_t* rp = op->res_id + (size_t)c.img * op->res_id_tile_bytes + ((size_t)(cbase / 8) * op->res_id_rows + 2 + q) * 16;
              const uint4 rh = ld_cg_u4(rp), rl = ld_cg_u4(rp + (size_t)(op->res_id_C / 8) * op->res_id_rows * 16);
              add8(rh, rl, y);
            }
            uint4 hi, lo;
            split8(y, hi, lo);
            uint8_t* ob = obase + (size_t)(cbase / 8) * op->out_rows * 16;
            *reinterpret_cast<uint4*>(ob) = hi;
            *reinterpret_cast<uint4*>(ob + oplane) = lo;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(bar_acc_empty);
}

// ---- LinearAttention epilogues ------------------------------------------------------------------------------------------
// TC_ATTN_QKV: acc[row][n] = sum_c x[row][c] g[c] W[c][n] (raw x, gain folded into the weights).  Per row: LayerNorm statistics
// of x over the C channels from the input image in shared memory, then y = rstd (acc - mean s[n]) + u[n] (s = p_gamma,
// u = p_bias), stored as fp32 into the slot's scratch.
__device__ __forceinline__ void epi_attn_qkv(const FOp* __restrict__ op, const EpiCtx& c, uint32_t bar_acc_empty) {
  const int N = op->N, NMT = op->n_mt, L = op->L, Pp = op->P, C = op->src_C;
  const int NH = N / 4, c0 = c.cb * NH;
  const uint32_t lane_base = c.tmem + ((uint32_t)(c.q4 * 32) << 16);
  const uint32_t pinv = (65536u + (uint32_t)Pp - 1u) / (uint32_t)Pp;
  const int rows = (int)(op->src_lbo[0] >> 4);
  const uint8_t* img = c.inbuf + op->src_smem_off[0];
  const size_t plane = op->src_plane[0];
  const int ppq = C / 32;                      // 8-channel panels per column-block thread (the 4 threads of a row split the channels)
  // partial (sum, sum of squares) of this thread's channels, every m-tile
  for (int i = 0; i < NMT; ++i) {
    const int q = 128 * i + c.row;
    float sm = 0.f, sq = 0.f;
    for (int pp = 0; pp < ppq; ++pp) {
      const uint8_t* pr = img + ((size_t)(c.cb * ppq + pp) * rows + 2 + q) * 16;
      const uint4 hi = *reinterpret_cast<const uint4*>(pr), lo = *reinterpret_cast<const uint4*>(pr + plane);
      const __half2* h = reinterpret_cast<const __half2*>(&hi);
      const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(h[e]), b = __half22float2(l[e]);
        const float x0 = a.x + b.x, x1 = a.y + b.y;
        sm += x0 + x1;
        sq = fmaf(x0, x0, fmaf(x1, x1, sq));
      }
    }
    c.part[c.cb * F_PART_ROWS + q] = make_float2(sm, sq);
  }
  epi_bar16();
  const float inv_c = 1.f / (float)C;
  for (int i = 0; i < NMT; ++i) {
    const int q = 128 * i + c.row;
    const int si = (int)(((uint32_t)q * pinv) >> 16), pi = q - si * Pp;
    const bool ok = (si < c.st) && (pi < L) && (c.tile * c.st + si < c.B);
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 e = c.part[j * F_PART_ROWS + q]; sm += e.x; sq += e.y; }
    const float mean = sm * inv_c;
    const float rstd = rsqrtf(fmaxf(sq * inv_c - mean * mean, 0.f) + 1e-5f);
    float* dst = op->aux0 + ((size_t)c.slot * (128 * NMT) + q) * op->aux_ld + c0;
    for (int pc = 0; pc < NH / 8; ++pc) {
      float y[8];
      tmem_ld8(lane_base + i * N + c0 + pc * 8, y);
      tmem_wait_ld();
      if (ok) {
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = fmaf(rstd, y[e] - mean * c.p_gamma[c0 + pc * 8 + e], c.p_bias[c0 + pc * 8 + e]);
        *reinterpret_cast<float4*>(dst + pc * 8) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(dst + pc * 8 + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) mbar_arrive(bar_acc_empty);
}

// TC_ATTN_CORE: per (sample, head) one warp, lane = channel of the head.  k is soft-maxed over the L positions, context^T[e][d] =
// sum_n v[n][e] p[n][d] lives in registers (lane e holds row e), out[n][e] = scale sum_d context[d][e] q[n][d]; `out` is stored
// as the hi/lo tile image the to_out GEMM reads.  heads = 4, dim_head = 32 (layers.py:211).
__device__ __forceinline__ void epi_attn_core(const FOp* __restrict__ op, const EpiCtx& c, uint32_t bar_acc_empty) {
  if (c.lane == 0) mbar_arrive(bar_acc_empty);   // no accumulators in this op
  epi_bar16();                                   // every thread's q / k / v stores of the previous items are visible
  const int L = op->L, Pp = op->P, ld = op->aux_ld;
  const int rows_q = 128 * op->n_mt;
  const float* qb = op->aux0 + (size_t)c.slot * rows_q * ld;
  const float* kb = op->aux1 + (size_t)c.slot * rows_q * ld;
  const float* vb = op->aux2 + (size_t)c.slot * rows_q * ld;
  const size_t oplane = (size_t)(op->out_C / 8) * op->out_rows * 16;
  uint8_t* oimg = op->out + (size_t)c.img * op->out_tile_bytes;
  uint8_t* oimg2 = op->out2 ? op->out2 + (size_t)c.img * op->out2_tile_bytes : nullptr;   // rows >= 256 (half-height images)
  const int lane = c.lane;
  for (int pair = c.warp; pair < c.st * 4; pair += FE_WARPS) {
    const int s = pair >> 2, h = pair & 3;
    if (c.tile * c.st + s >= c.B) continue;
    const int col = h * 32 + lane;
    const int q0 = s * Pp;
    // softmax statistics of k[:, col] over the positions
    float mx = -INFINITY;
    for (int n = 0; n < L; ++n) mx = fmaxf(mx, kb[(size_t)(q0 + n) * ld + col]);
    float sum = 0.f;
    for (int n = 0; n < L; ++n) sum += expf(kb[(size_t)(q0 + n) * ld + col] - mx);
    const float inv = 1.f / sum;
    float ctx[32];                               // lane e: context[d][e] for d = 0 .. 31
#pragma unroll
    for (int d = 0; d < 32; ++d) ctx[d] = 0.f;
    for (int n = 0; n < L; ++n) {
      const float pk = expf(kb[(size_t)(q0 + n) * ld + col] - mx) * inv;   // lane d: softmax(k)[n][d]
      const float ve = vb[(size_t)(q0 + n) * ld + col];                    // lane e: v[n][e]
#pragma unroll
      for (int d = 0; d < 32; ++d) ctx[d] = fmaf(ve, __shfl_sync(0xffffffffu, pk, d), ctx[d]);
    }
    for (int n = 0; n < L; ++n) {
      const float qd = qb[(size_t)(q0 + n) * ld + col] * 0.17677669529663687f;   // q * dim_head^-0.5
      float o = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) o = fmaf(ctx[d], __shfl_sync(0xffffffffu, qd, d), o);
      const __half oh = __float2half_rn(o);
      const __half ol = __float2half_rn(o - __half2float(oh));
      const int qr = q0 + n;
      uint8_t* ob = (oimg2 != nullptr && qr >= 256) ? oimg2 + ((size_t)(col >> 3) * op->out_rows + 2 + qr - 256) * 16
                                                    : oimg + ((size_t)(col >> 3) * op->out_rows + 2 + qr) * 16;
      __half* dsth = reinterpret_cast<__half*>(ob) + (col & 7);
      *dsth = oh;
      *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(dsth) + oplane) = ol;
    }
  }
}

// ATTN: the LinearAttention epilogues are compiled in (a separate instantiation keeps them out of the register allocation and
// the instruction stream of the attention-free networks every shipped config uses)
template <bool ATTN>
__global__ void __launch_bounds__(F_THREADS, 1) unet_fused_kernel(const __grid_constant__ FParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar0 = smem_base + P.off_bar;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + P.off_bar + 8 * B_COUNT);
#define FBAR(i) (bar0 + 8u * (uint32_t)(i))
#define FSTAMP(item, slot) do { if (P.dbg && blockIdx.x == 0) P.dbg[(size_t)(item) * 16 + (slot)] = clock64(); } while (0)

  if (tid == 0) {
    if (P.stamp) P.stamp[2 * blockIdx.x] = globaltimer_ns();
    for (int i = 0; i < 2; ++i) {
      mbar_init(FBAR(B_IN_FULL + i), 1);
      mbar_init(FBAR(B_IN_EMPTY + i), 1);
      mbar_init(FBAR(B_ACC_FULL + i), 1);
      mbar_init(FBAR(B_ACC_EMPTY + i), FE_WARPS);
      mbar_init(FBAR(B_OUT_DONE + i), FE_WARPS);
      mbar_init(FBAR(B_PRM_FULL + i), 1);
    }
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(FBAR(B_W_FULL + s), 1); mbar_init(FBAR(B_W_EMPTY + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == F_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                 "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int G = (int)gridDim.x;
  const int n_my = ((int)blockIdx.x < P.n_tiles) ? (P.n_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
  const int n_rounds = (n_my + 1) / 2;
  // every role walks the same item sequence: for round: for op: for parity (if that tile exists)
#define F_FOR_ITEMS                                        \
  for (int r_ = 0; r_ < n_rounds; ++r_)                    \
    for (int j = 0; j < P.n_ops; ++j)                      \
      for (int par = 0; par < 2; ++par)                    \
        if (2 * r_ + par < n_my)

  if (warp == F_WARP_IN) {
    // ================= input producer (lane 0) + per-channel parameter staging (all lanes) =================
    {
      int uses0 = 0, uses1 = 0, cpar0 = 0, cpar1 = 0, k = 0;
      F_FOR_ITEMS {
        const FOp* op = &P.ops[j];
        const int tile = (int)blockIdx.x + (2 * r_ + par) * G;
        const int slot = P.by_slot ? (int)blockIdx.x * 2 + par : tile;
        const int p = k & 1;
        if (lane == 0) {
          // the previous item of this tile parity must have stored its output (also keeps the barrier phases in step)
          const int cpar = par ? cpar1 : cpar0;
          if (cpar > 0) mbar_wait(FBAR(B_OUT_DONE + par), (uint32_t)((cpar - 1) & 1));
          const int big = op->big;
          if ((big || p == 0) && uses0 > 0) mbar_wait(FBAR(B_IN_EMPTY + 0), (uint32_t)((uses0 - 1) & 1));
          if ((big || p == 1) && uses1 > 0) mbar_wait(FBAR(B_IN_EMPTY + 1), (uint32_t)((uses1 - 1) & 1));
          FSTAMP(k, 0);
          const uint32_t base = smem_base + (big ? 0u : (uint32_t)p * P.buf_bytes);
          uint32_t total = 0;
          const int ns = op->in_smem ? 0 : op->n_src;
          for (int s = 0; s < ns; ++s) total += op->src_tile_bytes[s];
          if (op->in_smem) mbar_arrive(FBAR(B_IN_FULL + p));   // the previous op's epilogue left the image in the buffer
          else mbar_expect_tx(FBAR(B_IN_FULL + p), total);
          for (int s = 0; s < ns; ++s) {
            const uint32_t tb = op->src_tile_bytes[s];
            const uint8_t* g = op->src[s] + (size_t)(op->src_by_tile[s] ? tile : slot) * tb;
            uint32_t off = 0;
            while (off < tb) {
              const uint32_t n = min(tb - off, 32768u);
              bulk_g2s(base + op->src_smem_off[s] + off, g + off, n, FBAR(B_IN_FULL + p));
              off += n;
            }
          }
        }
        __syncwarp();
        // Parameters of this item -> bank k & 1.  The bank was last read by item k - 2, whose epilogue has finished: lane 0
        // waited for the OUT_DONE of an item >= k - 2 above, and epilogues complete in item order.
        {
          float* bank = reinterpret_cast<float*>(smem + P.off_prm + (uint32_t)p * (F_MAX_N * 20));
          const float* cond = (op->cond_off >= 0) ? P.cond_row + op->cond_off : nullptr;
          const int Nn = op->N;
          for (int ch = lane; ch < Nn; ch += 32) {
            const float vb = (ch < op->cout) ? __ldg(op->bias + ch) : 0.f;
            const float vg = op->gamma ? __ldg(op->gamma + ch) : 1.f;
            const float ve = op->beta ? __ldg(op->beta + ch) : 0.f;
            const float vc = cond ? __ldg(cond + ch) : 0.f;
            const float vr = op->res_bias ? __ldg(op->res_bias + ch) : 0.f;
            bank[ch] = vb;
            bank[F_MAX_N + ch] = vg;
            bank[2 * F_MAX_N + ch] = ve;
            bank[3 * F_MAX_N + ch] = vc;
            bank[4 * F_MAX_N + ch] = vr;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(FBAR(B_PRM_FULL + p));
        }
        if (op->big || p == 0) uses0++;
        if (op->big || p == 1) uses1++;
        if (par) cpar1++; else cpar0++;
        k++;
      }
    }
  } else if (warp == F_WARP_W) {
    // ================= weight producer =================
    if (lane == 0) {
      int st = 0, use = 0;
      F_FOR_ITEMS {
        const FOp* op = &P.ops[j];
        const int nc = op->n_chunks, cb0 = op->chunk_base;
        const uint32_t Nw = (uint32_t)op->N;
        for (int ci = 0; ci < nc; ++ci, st = (st + 1 == F_STAGES) ? 0 : st + 1, use += (st == 0)) {
          if (use > 0) mbar_wait(FBAR(B_W_EMPTY + st), (uint32_t)((use - 1) & 1));
          const FChunk cd = P.chunks[cb0 + ci];
          const uint32_t wb = ((cd.meta >> 2) & 3u) * 16u * Nw * 4u;
          if ((P.dbg_flags & 1) && use > 0) { mbar_arrive(FBAR(B_W_FULL + st)); continue; }
          mbar_expect_tx(FBAR(B_W_FULL + st), wb);
          bulk_g2s(smem_base + P.off_ring + (uint32_t)st * F_STAGE_BYTES, op->wchunks + cd.w_off, wb, FBAR(B_W_FULL + st));
        }
      }
    }
  } else if (warp == F_WARP_MMA) {
    // ================= MMA issuer: the whole warp walks the loop (uniform), one elected lane issues =================
    // Per chunk the issuer does: one prefetched 16-byte constant load, one barrier wait, a handful of uniform adds, the
    // MMAs and one commit -- the round-2 timeline showed the tensor pipe itself runs at its floor (bytes / 128 per MMA)
    // and that a fat per-chunk prologue (~300 cycles of dependent uniform ALU) was what stretched the MMA phases.
    {
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const bool dbg_on = (P.dbg != nullptr);
      const uint32_t ring_w0 = ((smem_base + P.off_ring) >> 4) & 0x3FFFu;
      uint32_t st = 0, ph = 0;                   // ring stage and its phase, carried across items
      int k = 0, nuse0 = 0, nuse1 = 0;           // items that used accumulator / in_full parity 0 / 1
      F_FOR_ITEMS {
        const FOp* op = &P.ops[j];
        const int p = k & 1;
        const uint32_t N = (uint32_t)op->N;
        const int NMT = op->n_mt, big = op->big;
        const uint32_t idesc = (1u << 4) | ((N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t abase16 = (smem_base + (big ? 0u : (uint32_t)p * P.buf_bytes)) >> 4;
        const uint32_t dbase = tmem_u + (uint32_t)p * 256u;
        const uint32_t b_hi = N << 16, b_kstep = 2u * N;      // LBO field of B; (2 panels * N rows * 16 B) >> 4 per K step
        const int nuse = p ? nuse1 : nuse0;
        const int nc = op->n_chunks, cb0 = op->chunk_base;
        FChunk cd = P.chunks[cb0];
        if (nuse > 0) mbar_wait(FBAR(B_ACC_EMPTY + p), (uint32_t)((nuse - 1) & 1));
        mbar_wait(FBAR(B_IN_FULL + p), (uint32_t)(nuse & 1));
        tc_fence_after();
        if (dbg_on) FSTAMP(k, 1);
        long long wstall = 0;
        for (int ci = 0; ci < nc; ++ci) {
          const FChunk nxt = P.chunks[cb0 + ci + 1];           // table is padded by one entry
          if (dbg_on) { const long long w0 = clock64(); mbar_wait(FBAR(B_W_FULL + st), ph); wstall += clock64() - w0; }
          else mbar_wait(FBAR(B_W_FULL + st), ph);
          tc_fence_after();
          const uint32_t acc = cd.meta & 1u, first = (cd.meta >> 1) & 1u, k16 = (cd.meta >> 2) & 3u;
          uint32_t a_w = cd.a_w + abase16;
          uint32_t b_w = (ring_w0 + st * (F_STAGE_BYTES >> 4)) | b_hi;
          const uint32_t a_kstep = cd.a_ks_pl & 0xFFFFu, a_pl = cd.a_ks_pl >> 16;
          const uint32_t b_pl = k16 * b_kstep;
          const uint32_t dt = dbase + acc * 128u;
          for (uint32_t kk = 0; kk < k16; ++kk) {
            const uint32_t acc0 = (first && kk == 0) ? 0u : 1u;
            // the three terms of the hi/lo split, m-tiles innermost: consecutive MMAs write different accumulators
            if (NMT == 1) {
              mma_lohi(dt, a_w, b_w, idesc, acc0, leader);
              if (!(P.dbg_flags & 2)) {
                mma_lohi(dt, a_w + a_pl, b_w, idesc, 1u, leader);
                mma_lohi(dt, a_w, b_w + b_pl, idesc, 1u, leader);
              }
            } else if (NMT == 2) {
              mma_lohi(dt, a_w, b_w, idesc, acc0, leader);
              mma_lohi(dt + N, a_w + 128u, b_w, idesc, acc0, leader);
              if (!(P.dbg_flags & 2)) {
                mma_lohi(dt, a_w + a_pl, b_w, idesc, 1u, leader);
                mma_lohi(dt + N, a_w + 128u + a_pl, b_w, idesc, 1u, leader);
                mma_lohi(dt, a_w, b_w + b_pl, idesc, 1u, leader);
                mma_lohi(dt + N, a_w + 128u, b_w + b_pl, idesc, 1u, leader);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) mma_lohi(dt + (uint32_t)i * N, a_w + (uint32_t)i * 128u, b_w, idesc, acc0, leader);
              if (!(P.dbg_flags & 2)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mma_lohi(dt + (uint32_t)i * N, a_w + (uint32_t)i * 128u + a_pl, b_w, idesc, 1u, leader);
#pragma unroll
                for (int i = 0; i < 4; ++i) mma_lohi(dt + (uint32_t)i * N, a_w + (uint32_t)i * 128u, b_w + b_pl, idesc, 1u, leader);
              }
            }
            a_w += a_kstep;
            b_w += b_kstep;
          }
          if (leader) tc_commit(FBAR(B_W_EMPTY + st));
          if (++st == F_STAGES) { st = 0; ph ^= 1u; }
          cd = nxt;
        }
        if (leader) {
          tc_commit(FBAR(B_ACC_FULL + p));
          if (big) { tc_commit(FBAR(B_IN_EMPTY + 0)); tc_commit(FBAR(B_IN_EMPTY + 1)); }
          else tc_commit(FBAR(B_IN_EMPTY + p));
        }
        if (dbg_on && lane == 0) {
          FSTAMP(k, 2);
          if (blockIdx.x == 0) P.dbg[(size_t)k * 16 + 6] = wstall;
        }
        if (p) nuse1++; else nuse0++;
        k++;
      }
    }
  } else {
    // ================= epilogue: 16 warps =================
    EpiCtx c;
    c.tid = tid; c.warp = warp; c.lane = lane; c.q4 = warp & 3; c.cb = warp >> 2; c.row = c.q4 * 32 + lane;
    c.part = reinterpret_cast<float2*>(smem + P.off_part);
    c.B = P.B;
    c.st = P.st;
    int k = 0, nuse0 = 0, nuse1 = 0;
    F_FOR_ITEMS {
      const FOp* op = &P.ops[j];
      const int p = k & 1;
      c.tile = (int)blockIdx.x + (2 * r_ + par) * G;
      c.img = P.by_slot ? (int)blockIdx.x * 2 + par : c.tile;
      c.p_bias = reinterpret_cast<float*>(smem + P.off_prm + (uint32_t)p * (F_MAX_N * 20));
      c.p_gamma = c.p_bias + F_MAX_N;
      c.p_beta = c.p_bias + 2 * F_MAX_N;
      c.p_cond = c.p_bias + 3 * F_MAX_N;
      c.p_rb = c.p_bias + 4 * F_MAX_N;
      c.stat = reinterpret_cast<float2*>(smem + P.off_stat + (uint32_t)p * 512);
      c.tmem = tmem_base + (uint32_t)p * 256u;
      // in-place hand-over target = the buffer the NEXT op of this tile will use: items alternate buffers, so it is this
      // item's own (free) buffer while the round has two tiles, and the other buffer in a single-tile round
      c.sbuf = smem + (size_t)((2 * r_ + 1 < n_my) ? p : (p ^ 1)) * P.buf_bytes;
      c.inbuf = smem + (op->big ? 0u : (uint32_t)p * P.buf_bytes);
      c.slot = (int)blockIdx.x * 2 + par;
      c.dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg + (size_t)k * 16 : nullptr;
      if (tid == 0) FSTAMP(k, 7);
      // per-channel parameters of this item: staged by the producer warp into bank p
      mbar_wait(FBAR(B_PRM_FULL + p), (uint32_t)((p ? nuse1 : nuse0) & 1));
      mbar_wait(FBAR(B_ACC_FULL + p), (uint32_t)((p ? nuse1 : nuse0) & 1));
      tc_fence_after();
      if (tid == 0) FSTAMP(k, 3);
      const uint32_t bae = FBAR(B_ACC_EMPTY + p);
      if (op->kind == TC_CONVBLOCK) {
        switch (op->variant) {
          case FV_4_32: epi_convblock<4, 32>(op, c, bae); break;
          case FV_2_64: epi_convblock<2, 64>(op, c, bae); break;
          case FV_1_128: epi_convblock<1, 128>(op, c, bae); break;
          case FV_2_32: epi_convblock<2, 32>(op, c, bae); break;
          case FV_1_32: epi_convblock<1, 32>(op, c, bae); break;
          default: epi_convblock<1, 64>(op, c, bae); break;
        }
      } else if (ATTN && op->kind == TC_ATTN_QKV) {
        if constexpr (ATTN) epi_attn_qkv(op, c, bae);
      } else if (ATTN && op->kind == TC_ATTN_CORE) {
        if constexpr (ATTN) epi_attn_core(op, c, bae);
      } else {
        epi_plain<ATTN>(op, c, bae, P.eps);
      }
      // output image visible to the async proxy (the input producer's bulk copies / the next op's MMAs) before the arrival
      fence_proxy_async_global();
      if (op->out_smem) fence_proxy_async_smem();
      if (c.dbg) c.dbg[13] = clock64();
      __syncwarp();
      if (lane == 0) mbar_arrive(FBAR(B_OUT_DONE + par));
      if (tid == 0) FSTAMP(k, 5);
      if (p) nuse1++; else nuse0++;
      k++;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (P.stamp && tid == 0) P.stamp[2 * blockIdx.x + 1] = globaltimer_ns();
  if (warp == F_WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
#undef FBAR
#undef FSTAMP
#undef F_FOR_ITEMS
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
struct FusedState {
  uint64_t uid = 0;
  int B = 0, n_tiles = 0, grid = 0, by_slot = 0, st = ST;
  bool keep_all = false;           // every activation image also goes to global memory (debug taps)
  std::vector<TcImage> images;     // [0] = packed network input, [1 + j] = output of op j
  std::vector<TcImage> images_r;   // [j] = residual-conv image written by op j (dev == nullptr if none)
  std::vector<uint8_t*> owned;     // distinct device allocations behind the images (buffers may be shared)
  uint8_t* w_all = nullptr;
  std::vector<FOp> ops;
  FParams prm{};
  size_t smem = 0;
  size_t activation_bytes = 0;
};

static void fused_free(FusedState* s) {
  if (!s) return;
  for (auto* p : s->owned) cudaFree(p);
  cudaFree(s->w_all);
  delete s;
}

void unet_fused_release(UnetImpl* net) {
  if (!net) return;
  for (auto& kv : net->fused) fused_free(kv.second);
  net->fused.clear();
}

static int build_fused(UnetImpl* net, int B, int by_slot, bool keep_all, cudaStream_t stream, FusedState** out) {
  const auto& cfg = net->cfg;
  if (cfg.state_dim > 8) return fail(MMDK_EINVAL, "tensor-core executor: state_dim > 8 not supported");
  int dev = 0, n_sm = 0, max_smem = 0;
  MMDK_CUDA(cudaGetDevice(&dev));
  MMDK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  MMDK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  auto* st = new FusedState();
  st->uid = next_uid();
  st->B = B;
  // samples per tile.  Measured kernel durations on B200 (tests/tc_stamps.py): a lone tile's 30-op chain costs 196 us with 1
  // sample, 204 us with 3, 255 us with 7 (the C=128 / L=16 layers sweep a full M=128 tile whatever it holds), so smaller tiles
  // only pay while they still fit one per SM.  MMDK_FUSED_ST overrides (A/B runs).
  {
    int spt = (B <= n_sm) ? 1 : ((B <= 3 * n_sm) ? 3 : 7);
    if (const char* e = getenv("MMDK_FUSED_ST")) { const int v = atoi(e); if (v == 7 || v == 3 || v == 1) spt = v; }
    st->st = spt;
  }
  st->n_tiles = (B + st->st - 1) / st->st;
  st->grid = std::min(st->n_tiles, n_sm);
  st->by_slot = by_slot;
  st->keep_all = keep_all;
  const int n_img = by_slot ? 2 * st->grid : st->n_tiles;   // entries of every intermediate image
  const int n_ops = (int)net->ops.size();
  st->images.reserve(n_ops + 1 + 16);    // [n_ops + 1 ..]: internal images of the attention blocks (references stay valid)
  st->images.resize(n_ops + 1);
  st->images_r.resize(n_ops);
  st->ops.reserve(F_MAX_OPS + 8);        // FOp list: one per layer op, several per LinearAttention block
  auto fail_free = [&](const std::string& m) { fused_free(st); return fail(MMDK_EINVAL, m); };

  auto make_image = [&](TcImage& im, int C, int L, int entries, int rows_override = 0) -> int {
    im.C = C; im.L = L; im.rows = rows_override ? rows_override : level_rows(L, st->st);
    im.tile_bytes = (uint32_t)C * im.rows * 4;
    const size_t bytes = (size_t)im.tile_bytes * entries;
    if (cudaMalloc(&im.dev, bytes) != cudaSuccess) return MMDK_ENOMEM;
    cudaMemsetAsync(im.dev, 0, bytes, stream);   // halo rows / unused samples stay zero for ever
    st->owned.push_back(im.dev);
    st->activation_bytes += bytes;
    return MMDK_OK;
  };
  if (make_image(st->images[0], 16, cfg.horizon, st->n_tiles) != MMDK_OK) return fail_free("out of memory (input image)");

  std::vector<FChunk> all_chunks;
  struct WSrc { const float* W; int cin, ktaps, tap, ci0, CK, cout, N; uint32_t dst_off; int n0; const float* scale; };
  std::vector<WSrc> wsrc;
  uint32_t w_total = 0;
  std::vector<uint32_t> op_w_off;     // per FOp
  uint32_t max_in = 0;
  std::vector<uint32_t> in_bytes;     // per FOp
  struct HChunk { uint32_t a_off, w_off, w_bytes; int d, acc, first, k16, slot; };
  std::string ferr;
  // chunk list of one FOp -> packed descriptors; bookkeeping shared by layer ops and attention ops
  auto finalize_op = [&](FOp& p, const std::vector<HChunk>& chunks, uint32_t off) -> bool {
    p.in_bytes = off;
    in_bytes.push_back(off);
    max_in = std::max(max_in, off);
    p.chunk_base = (int)all_chunks.size();
    p.n_chunks = (int)chunks.size();
    op_w_off.push_back(chunks.empty() ? w_total : chunks[0].w_off);   // chunk weight offsets are relative to the op's base
    for (const auto& cd : chunks) {
      FChunk fc{};
      const uint32_t lbo = p.src_lbo[cd.slot], plane = p.src_plane[cd.slot];
      fc.a_w = (((cd.a_off + (uint32_t)((2 + cd.d) * 16)) >> 4) & 0x3FFFu) | ((lbo >> 4) << 16);
      if (((2 * lbo) >> 4) > 0xFFFFu || (plane >> 4) > 0xFFFFu) { ferr = "tensor-core executor: image too large for the packed chunk descriptor"; return false; }
      fc.a_ks_pl = ((2 * lbo) >> 4) | ((plane >> 4) << 16);
      fc.w_off = cd.w_off - op_w_off.back();
      fc.meta = (uint32_t)cd.acc | ((uint32_t)cd.first << 1) | ((uint32_t)cd.k16 << 2);
      all_chunks.push_back(fc);
    }
    return true;
  };
  // one 1x1-conv GEMM over a single input image (attention projections): source slot 0, tap 0, region 0
  auto gemm_1x1 = [&](FOp& p, std::vector<HChunk>& chunks, uint32_t& off, const TcImage& im, const float* W, int cin, int cout_total,
                      int n0, const float* scale) {
    p.n_src = 1;
    p.src[0] = im.dev; p.src_by_tile[0] = 0; p.src_tile_bytes[0] = im.tile_bytes; p.src_smem_off[0] = 0;
    p.src_lbo[0] = (uint32_t)im.rows * 16;
    p.src_plane[0] = (uint32_t)(im.C / 8) * im.rows * 16;
    off = (im.tile_bytes + 127) & ~127u;
    for (int c0 = 0; c0 < im.C; c0 += 32) {
      const int CK = std::min(32, im.C - c0);
      HChunk cd{};
      cd.a_off = (uint32_t)(c0 / 8) * im.rows * 16;
      cd.slot = 0;
      cd.w_bytes = (uint32_t)CK * p.N * 4;
      cd.w_off = w_total;
      wsrc.push_back({W, cin, 1, 0, c0, CK, cout_total, p.N, w_total, n0, scale});
      w_total += cd.w_bytes;
      cd.d = 0; cd.acc = 0; cd.first = (c0 == 0) ? 1 : 0; cd.k16 = CK / 16;
      chunks.push_back(cd);
    }
  };

  for (int j = 0; j < n_ops; ++j) {
    const Op& op = net->ops[j];
    if (op.type == OP_ATTN) {
      // ---- Residual(PreNorm(LinearAttention)) (layers.py:177-229) -> QKV slices, core, to_out ------------------------------
      const int C = op.cin, Lx = op.lin;
      const int rows = level_rows(Lx, st->st), n_mt = (rows - 4) / 128;
      if ((C != 32 && C != 64 && C != 128) || (n_mt != 1 && n_mt != 2 && n_mt != 4) || Lx > 64 || op.p_src0 < 0)
        return fail_free("tensor-core executor: unsupported LinearAttention shape");
      const int xi = op.p_src0 + 1;                       // image of x
      const int Nq = (n_mt == 4) ? 64 : 128;              // columns per QKV slice: n_mt * Nq <= 256 accumulator columns
      const size_t n_slots = 2 * (size_t)st->grid, rows_q = 128 * (size_t)n_mt;
      float* scr[3] = {nullptr, nullptr, nullptr};
      float* fold = nullptr;
      for (int w = 0; w < 3; ++w) {
        if (cudaMalloc(&scr[w], n_slots * rows_q * 128 * sizeof(float)) != cudaSuccess) return fail_free("out of memory (attention scratch)");
        st->owned.push_back(reinterpret_cast<uint8_t*>(scr[w]));
        st->activation_bytes += n_slots * rows_q * 128 * sizeof(float);
      }
      if (cudaMalloc(&fold, 768 * sizeof(float)) != cudaSuccess) return fail_free("out of memory (attention fold)");
      st->owned.push_back(reinterpret_cast<uint8_t*>(fold));
      attn_fold_kernel<<<3, 128, 0, stream>>>(net->blob + op.w, net->blob + op.gn_w, net->blob + op.gn_b, C, 384, fold, fold + 384);
      // `out` of the core: 128 channels at this resolution.  With 4 m-tiles (7 samples at L = 64) the image would be 264 KB, more
      // than the input buffers hold: two half-height images (rows 0..255 / 256..511 of the tile) and two to_out ops instead
      const int n_half = (n_mt == 4) ? 2 : 1;
      int oi[2] = {-1, -1};
      for (int hf = 0; hf < n_half; ++hf) {
        st->images.emplace_back();
        oi[hf] = (int)st->images.size() - 1;
        if (make_image(st->images[oi[hf]], 128, Lx, n_img, n_half == 2 ? 2 + 256 + 2 : 0) != MMDK_OK) return fail_free("out of memory (attention image)");
      }
      if (make_image(st->images[j + 1], C, Lx, n_img) != MMDK_OK) return fail_free("out of memory (activations)");
      for (int sl = 0; sl < 384 / Nq; ++sl) {
        st->ops.emplace_back();
        FOp& p = st->ops.back();
        p = FOp{};
        p.kind = TC_ATTN_QKV; p.L = Lx; p.P = Lx + 2; p.n_mt = n_mt; p.N = Nq; p.cout = Nq; p.variant = FV_GENERIC; p.cond_off = -1;
        p.src_C = C;
        std::vector<HChunk> chunks;
        uint32_t off = 0;
        gemm_1x1(p, chunks, off, st->images[xi], net->blob + op.w, C, 384, sl * Nq, net->blob + op.gn_w);
        p.bias = fold + 384 + sl * Nq;                    // u
        p.gamma = fold + sl * Nq;                         // s
        p.aux0 = scr[(sl * Nq) / 128] + (sl * Nq) % 128;
        p.aux_ld = 128;
        if (!finalize_op(p, chunks, off)) return fail_free(ferr);
      }
      {
        st->ops.emplace_back();
        FOp& p = st->ops.back();
        p = FOp{};
        p.kind = TC_ATTN_CORE; p.L = Lx; p.P = Lx + 2; p.n_mt = n_mt; p.N = 32; p.cout = 0; p.variant = FV_GENERIC; p.cond_off = -1;
        p.aux0 = scr[0]; p.aux1 = scr[1]; p.aux2 = scr[2]; p.aux_ld = 128;
        const TcImage& im = st->images[oi[0]];
        p.out = im.dev; p.out_tile_bytes = im.tile_bytes; p.out_L = im.L; p.out_rows = im.rows; p.out_C = im.C;
        if (n_half == 2) {
          const TcImage& im2 = st->images[oi[1]];
          p.out2 = im2.dev; p.out2_tile_bytes = im2.tile_bytes; p.out2_rows = im2.rows; p.out2_C = im2.C;
        }
        std::vector<HChunk> none;
        if (!finalize_op(p, none, 0)) return fail_free(ferr);
      }
      for (int hf = 0; hf < n_half; ++hf) {
        st->ops.emplace_back();
        FOp& p = st->ops.back();
        p = FOp{};
        p.kind = TC_ATTN_OUT; p.L = Lx; p.P = Lx + 2; p.n_mt = n_mt / n_half; p.N = C; p.cout = C; p.variant = FV_GENERIC; p.cond_off = -1;
        p.q_base = 256 * hf;
        std::vector<HChunk> chunks;
        uint32_t off = 0;
        gemm_1x1(p, chunks, off, st->images[oi[hf]], net->blob + op.res_w, 128, C, 0, nullptr);
        p.bias = net->blob + op.b;
        const TcImage& ri = st->images[xi];
        p.res_id = ri.dev; p.res_id_tile_bytes = ri.tile_bytes; p.res_id_rows = ri.rows; p.res_id_C = ri.C;
        const TcImage& im = st->images[j + 1];
        p.out = im.dev; p.out_tile_bytes = im.tile_bytes; p.out_L = im.L; p.out_rows = im.rows; p.out_C = im.C;
        if (!finalize_op(p, chunks, off)) return fail_free(ferr);
      }
      continue;
    }
    st->ops.emplace_back();
    FOp& p = st->ops.back();
    p = FOp{};
    p.kind = op.type == OP_CONVBLOCK ? TC_CONVBLOCK : op.type == OP_DOWN ? TC_DOWN : op.type == OP_UP ? TC_UP : TC_FINAL;
    p.L = op.lin; p.P = op.lin + 2;
    const int rows = level_rows(op.lin, st->st);
    p.n_mt = (rows - 4) / 128;
    p.cout = op.cout;
    p.N = op.cout < 16 ? 16 : op.cout;
    if (p.N != 16 && p.N != 32 && p.N != 64 && p.N != 128) return fail_free("tensor-core executor: unsupported channel count");
    const int NV = p.n_mt * p.N;
    if (NV > 128 || p.n_mt * 128 > F_PART_ROWS || (p.n_mt != 1 && p.n_mt != 2 && p.n_mt != 4)) return fail_free("tensor-core executor: unsupported (rows, channels) combination for this network shape");
    p.variant = FV_GENERIC;
    if (p.kind == TC_CONVBLOCK) {
      if (op.n_groups != 8) return fail_free("tensor-core executor: GroupNorm needs 8 groups");
      if (op.lin > 64) return fail_free("tensor-core executor: horizon > 64 not supported");
      if (p.n_mt == 4 && p.N == 32) p.variant = FV_4_32;
      else if (p.n_mt == 2 && p.N == 64) p.variant = FV_2_64;
      else if (p.n_mt == 1 && p.N == 128) p.variant = FV_1_128;
      else if (p.n_mt == 2 && p.N == 32) p.variant = FV_2_32;
      else if (p.n_mt == 1 && p.N == 64) p.variant = FV_1_64;
      else if (p.n_mt == 1 && p.N == 32) p.variant = FV_1_32;
      else return fail_free("tensor-core executor: unsupported conv block shape");
    } else if (p.kind != TC_FINAL && p.N < 32) {
      return fail_free("tensor-core executor: resampling convs need >= 32 channels");
    }
    // output image
    if (p.kind != TC_FINAL) {
      if (make_image(st->images[j + 1], op.cout, op.lout, n_img) != MMDK_OK) return fail_free("out of memory (activations)");
      const TcImage& oi = st->images[j + 1];
      p.out = oi.dev; p.out_tile_bytes = oi.tile_bytes;
      p.out_L = oi.L; p.out_rows = oi.rows; p.out_C = oi.C;
    }
    auto img_of = [&](int prod) { return prod + 1; };   // -1 (network input) -> image 0
    std::vector<int> main_imgs = {img_of(op.p_src0)};
    if (op.p_src1 > -2) main_imgs.push_back(img_of(op.p_src1));
    // residual handling
    const bool is_cb = (p.kind == TC_CONVBLOCK);
    const bool res_id_x = is_cb && op.res_src >= 0 && op.res_w < 0;                 // identity residual: add x
    const bool res_from_r = is_cb && op.res_w >= 0;                                   // r image written by op j-1
    bool eval_res = false;                                                            // this op evaluates the NEXT op's residual conv
    if (is_cb && j + 1 < n_ops) {
      const Op& nx = net->ops[j + 1];
      eval_res = nx.type == OP_CONVBLOCK && nx.res_w >= 0 && nx.p_src0 == j && nx.p_res == op.p_src0 && nx.p_res1 == op.p_src1 &&
                 nx.res_cin == op.cin && nx.cout == op.cout;
    }
    if (res_from_r) {
      if (j == 0 || !st->images_r[j - 1].dev) return fail_free("tensor-core executor: residual conv without a producing block");
      const TcImage& ri = st->images_r[j - 1];
      p.res_id = ri.dev; p.res_id_tile_bytes = ri.tile_bytes; p.res_id_rows = ri.rows; p.res_id_C = ri.C;
    } else if (res_id_x) {
      const TcImage& ri = st->images[img_of(op.p_res)];
      p.res_id = ri.dev;
      p.res_id_tile_bytes = ri.tile_bytes;
      p.res_id_rows = ri.rows; p.res_id_C = ri.C;
      if (op.p_res < 0) return fail_free("tensor-core executor: identity residual of the network input is not supported");
    }
    uint32_t off = 0;
    auto add_src = [&](int img) -> int {
      for (int k = 0; k < p.n_src; ++k) if (p.src[k] == st->images[img].dev) return k;
      if (p.n_src >= MAX_SRC) return -1;
      int k = p.n_src++;
      p.src[k] = st->images[img].dev;
      p.src_by_tile[k] = (img == 0) ? 1 : 0;
      p.src_tile_bytes[k] = st->images[img].tile_bytes;
      p.src_smem_off[k] = off;
      p.src_lbo[k] = (uint32_t)st->images[img].rows * 16;
      p.src_plane[k] = (uint32_t)(st->images[img].C / 8) * st->images[img].rows * 16;
      off += (st->images[img].tile_bytes + 127) & ~127u;
      return k;
    };
    std::vector<HChunk> chunks;
    auto add_chunks = [&](const std::vector<int>& imgs, int cin_total, int w_off_blob, int ktaps, int tap, int d, int acc,
                          bool first_in_acc) -> bool {
      int ci = 0;
      bool first = first_in_acc;
      for (size_t si = 0; si < imgs.size(); ++si) {
        const TcImage& im = st->images[imgs[si]];
        const int slot = add_src(imgs[si]);
        if (slot < 0) return false;
        const int c_here = im.C;
        for (int c0 = 0; c0 < c_here; c0 += 32) {
          const int CK = std::min(32, c_here - c0);
          HChunk cd{};
          cd.a_off = p.src_smem_off[slot] + (uint32_t)(c0 / 8) * im.rows * 16;
          cd.slot = slot;
          cd.w_bytes = (uint32_t)CK * p.N * 4;
          cd.w_off = w_total;
          wsrc.push_back({net->blob + w_off_blob, cin_total, ktaps, tap, ci + c0, CK, op.cout, p.N, w_total, 0, nullptr});
          w_total += cd.w_bytes;
          cd.d = d; cd.acc = acc; cd.first = first ? 1 : 0; cd.k16 = CK / 16;
          first = false;
          chunks.push_back(cd);
        }
        ci += (imgs[si] == 0) ? cfg.state_dim : c_here;
      }
      return true;
    };
    bool okc = true;
    if (p.kind == TC_CONVBLOCK) {
      for (int tap = 0; tap < 5; ++tap) okc = okc && add_chunks(main_imgs, op.cin, op.w, 5, tap, tap - 2, 0, tap == 0);
      if (eval_res) {
        const Op& nx = net->ops[j + 1];
        okc = okc && add_chunks(main_imgs, nx.res_cin, nx.res_w, 1, 0, 0, 1, true);
        if (make_image(st->images_r[j], op.cout, op.lout, n_img) != MMDK_OK) return fail_free("out of memory (residual image)");
        TcImage& ri = st->images_r[j];
        p.out2 = ri.dev; p.out2_tile_bytes = ri.tile_bytes; p.out2_rows = ri.rows; p.out2_C = ri.C;
        p.res_bias = net->blob + nx.res_b;
      }
    } else if (p.kind == TC_DOWN) {
      for (int tap = 0; tap < 3; ++tap) okc = okc && add_chunks(main_imgs, op.cin, op.w, 3, tap, tap - 1, 0, tap == 0);
    } else if (p.kind == TC_UP) {
      // out[2m] = x[m] w1 + x[m-1] w3 ; out[2m+1] = x[m+1] w0 + x[m] w2   (ConvTranspose1d k4 s2 p1)
      okc = okc && add_chunks(main_imgs, op.cin, op.w, 4, 1, 0, 0, true);
      okc = okc && add_chunks(main_imgs, op.cin, op.w, 4, 3, -1, 0, false);
      okc = okc && add_chunks(main_imgs, op.cin, op.w, 4, 0, 1, 1, true);
      okc = okc && add_chunks(main_imgs, op.cin, op.w, 4, 2, 0, 1, false);
    } else {
      okc = okc && add_chunks(main_imgs, op.cin, op.w, 1, 0, 0, 0, true);
    }
    if (!okc) return fail_free("tensor-core executor: too many input images in one op");
    if (!finalize_op(p, chunks, off)) return fail_free(ferr);
    p.bias = net->blob + op.b;
    p.cond_off = -1;
    if (p.kind == TC_CONVBLOCK) {
      p.gamma = net->blob + op.gn_w;
      p.beta = net->blob + op.gn_b;
      p.cond_off = op.cond;
    }
  }
  const int n_f = (int)st->ops.size();   // FOps (>= n_ops when the network has attention blocks)
  // shared-memory plan
  const uint32_t scratch = 2u * (F_MAX_N * 20) + 8u * F_PART_ROWS * 8u + 1024u + 8u * B_COUNT + 16u;
  const uint32_t avail = (uint32_t)max_smem - F_STAGES * F_STAGE_BYTES - scratch - 256u;
  uint32_t buf = 0;
  for (int j = 0; j < n_f; ++j) if (in_bytes[j] * 2 <= avail) buf = std::max(buf, in_bytes[j]);
  buf = (buf + 127) & ~127u;
  if (buf == 0 || max_in > 2 * buf) {
    // the largest input needs both buffers: grow them as far as shared memory allows
    buf = std::max(buf, ((max_in + 1) / 2 + 127) & ~127u);
  }
  if (2 * buf > avail) return fail_free("tensor-core executor: an op's input images do not fit in shared memory");
  for (int j = 0; j < n_f; ++j) st->ops[j].big = in_bytes[j] > buf ? 1 : 0;
  // in-place hand-over (see FOp): conv block j -> op j + 1 when j + 1 reads nothing but j's output at the same resolution
  for (int j = 0; j < n_f; ++j) { st->ops[j].in_smem = 0; st->ops[j].out_smem = 0; st->ops[j].out_global = 1; }
  const bool inplace_on = getenv("MMDK_FUSED_NO_INPLACE") == nullptr;
  for (int j = 0; inplace_on && j + 1 < n_f; ++j) {
    FOp& a = st->ops[j];
    FOp& b = st->ops[j + 1];
    if (a.kind != TC_CONVBLOCK || a.big || b.big || b.n_src != 1 || b.src[0] != a.out || b.L != a.out_L) continue;
    if (a.out_tile_bytes > buf || b.src_smem_off[0] != 0) continue;
    a.out_smem = 1;
    b.in_smem = 1;
  }
  for (int j = 0; j < n_f; ++j) {
    FOp& a = st->ops[j];
    if (a.kind == TC_FINAL || !a.out) continue;
    bool needed = keep_all;
    for (int m = j + 1; m < n_f && !needed; ++m) {
      const FOp& b = st->ops[m];
      if (b.res_id == a.out) needed = true;
      if (!b.in_smem) for (int k = 0; k < b.n_src; ++k) if (b.src[k] == a.out) needed = true;
    }
    a.out_global = needed ? 1 : 0;
  }
  FParams& P = st->prm;
  P.buf_bytes = buf;
  P.off_ring = 2 * buf;
  P.off_prm = P.off_ring + F_STAGES * F_STAGE_BYTES;
  P.off_part = P.off_prm + 2u * (F_MAX_N * 20);
  P.off_stat = P.off_part + 8u * F_PART_ROWS * 8u;
  P.off_bar = P.off_stat + 1024u;
  st->smem = P.off_bar + 8u * B_COUNT + 16u;
  if (st->smem > (size_t)max_smem) return fail_free("tensor-core executor: shared-memory plan exceeds the device limit");

  // weights
  if (cudaMalloc(&st->w_all, std::max<uint32_t>(w_total, 16)) != cudaSuccess) return fail_free("out of memory (packed weights)");
  for (const auto& w : wsrc)
    pack_wchunk_kernel<<<8, 256, 0, stream>>>(w.W, w.cin, w.ktaps, w.cout, w.tap, w.ci0, w.CK, w.N,
                                              reinterpret_cast<__half*>(st->w_all + w.dst_off), w.n0, w.scale);
  for (int j = 0; j < n_f; ++j) st->ops[j].wchunks = st->w_all + op_w_off[j];
  if (n_f > F_MAX_OPS || all_chunks.size() + 1 > (size_t)F_MAX_CHUNKS) return fail_free("tensor-core executor: layer program too large for the parameter bank");
  if (check_cuda(cudaGetLastError(), "tensor-core executor setup") != MMDK_OK) { fused_free(st); return MMDK_ECUDA; }
  for (int j = 0; j < n_f; ++j) P.ops[j] = st->ops[j];
  for (size_t c = 0; c < all_chunks.size(); ++c) P.chunks[c] = all_chunks[c];
  P.n_ops = n_f;
  P.n_tiles = st->n_tiles;
  P.B = B;
  P.st = st->st;
  P.by_slot = by_slot;
  *out = st;
  return MMDK_OK;
}

uint64_t unet_fused_state_uid(const UnetImpl* net, int B) {
  auto it = net->fused.find(B);
  if (it == net->fused.end() || it->second->keep_all != net->fused_keep) return 0;   // not built, or about to be rebuilt
  return it->second->uid;
}

int unet_forward_fused(UnetImpl* net, const float* x, int B, int t, float* eps, cudaStream_t stream) {
  FusedState* st = nullptr;
  auto it = net->fused.find(B);
  if (it != net->fused.end()) st = it->second;
  if (st && st->keep_all != net->fused_keep) {   // debug-tap mode changed: rebuild this batch size's state
    fused_free(st);
    net->fused.erase(it);
    st = nullptr;
  }
  if (!st) {
    // bounded cache of per-batch-size states (planner warm-up at B=2, sampling at B=K, batched sampling at B=R*K)
    if (net->fused.size() >= 4) {
      auto victim = net->fused.begin();
      for (auto i2 = net->fused.begin(); i2 != net->fused.end(); ++i2)
        if (i2->second->activation_bytes > victim->second->activation_bytes) victim = i2;
      fused_free(victim->second);
      net->fused.erase(victim);
    }
    int rc = build_fused(net, B, 0, net->fused_keep, stream, &st);
    if (rc != MMDK_OK) return rc;
    net->fused[B] = st;
  }
  net->fused_last = st;
  {
    int dev = 0;
    MMDK_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return fail(MMDK_EINVAL, "device index out of range");
    if (!configured[dev]) {
      int max_optin = 0;
      MMDK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      MMDK_CUDA(cudaFuncSetAttribute(unet_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
      MMDK_CUDA(cudaFuncSetAttribute(unet_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
      configured[dev] = true;
    }
  }
  const auto& cfg = net->cfg;
  {
    const TcImage& im = st->images[0];
    const int n = B * cfg.horizon;
    pack_input_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x, B, cfg.horizon, cfg.state_dim, im.rows, im.dev, im.tile_bytes, st->st);
  }
  FParams& P = st->prm;   // ~16 KB: patched in place, copied once by the launch
  P.cond_row = net->cond_table + (size_t)t * net->n_cond;
  P.eps = eps;
  P.dbg = net->fused_dbg;
  P.stamp = nullptr;
  if (net->stamps && net->stamp_slots > 0 && st->grid <= net->stamp_ctas) {
    // slot chosen at launch (= capture) time: a replayed graph keeps rewriting the slots of its own forwards
    P.stamp = net->stamps + (size_t)(net->stamp_next++ % net->stamp_slots) * net->stamp_ctas * 2;
  }
  { const char* e = getenv("MMDK_FUSED_DEBUG"); P.dbg_flags = e ? atoi(e) : 0; }
  if (cfg.self_attention) unet_fused_kernel<true><<<st->grid, F_THREADS, st->smem, stream>>>(P);
  else unet_fused_kernel<false><<<st->grid, F_THREADS, st->smem, stream>>>(P);
  return check_cuda(cudaGetLastError(), "unet_fused_kernel launch");
}

int unet_fused_tap(UnetImpl* net, int op_index, float* out, int* c_out, int* l_out, cudaStream_t stream) {
  FusedState* st = net->fused_last;
  if (!st) return fail(MMDK_EINVAL, "run a tensor-core forward first");
  if (op_index < -1 || op_index + 1 >= (int)st->images.size() || !st->images[op_index + 1].dev)
    return fail(MMDK_EINVAL, "op index out of range (or op has no activation image)");
  if (st->by_slot) return fail(MMDK_EINVAL, "activation taps need the per-tile image mode");
  if (!st->keep_all) return fail(MMDK_EINVAL, "activation taps need mmdk_unet_debug_keep_activations(net, 1) before the forward");
  const TcImage& im = st->images[op_index + 1];
  if (c_out) *c_out = im.C;
  if (l_out) *l_out = im.L;
  if (!out) return MMDK_OK;
  const int n = st->B * im.C * im.L;
  unpack_image_kernel<<<(n + 255) / 256, 256, 0, stream>>>(im.dev, im.tile_bytes, st->B, im.C, im.L, im.rows, out, st->st);
  return check_cuda(cudaGetLastError(), "unpack_image_kernel");
}

}  // namespace mmdk
