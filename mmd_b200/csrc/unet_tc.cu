// tcgen05 executor of the UNet layer program: every convolution of the TemporalUnet as an implicit GEMM on the
// 5th-gen tensor cores, one fused kernel per layer op (conv -> +bias -> GroupNorm -> Mish -> +cond -> +residual).
//
// Reference semantics: mmd/models/layers/layers.py:261-358, temporal_unet.py:121-174 (same Op program as unet.cu).
//
// Numerics: operands are split x = hi + lo in FP16, D += A_hi B_hi + A_lo B_hi + A_hi B_lo with
// fp32 accumulation in TMEM (kind::f16): ~2e-6 relative error per forward at the full 16-bit MMA rate.
//
// Layout ("tile image"): a tile = 7 whole samples.  For a level of length L the image is
//     [plane hi|lo][C/8 panels][ROWS][8 fp16]   ROWS = 2 + 128*n_mt + 2, sample s position p at row 2 + s*(L+2) + p,
// every other row zero (shared conv halo).  A row is 16 bytes inside a panel, so the image IS the UMMA no-swizzle
// K-major canonical layout (core matrix = 8 rows x 16 B, SBO = 128 B, LBO = ROWS*16 B) and a conv tap is nothing but
// a 16-byte shift of the A descriptor's start address: no im2col, no shifted copies.  The GEMM is
//     D[row, cout] = sum_tap sum_cin  A[row + tap - pad, cin] * W_tap[cout, cin]
// with M = 128 rows per MMA (n_mt MMAs tiles per image), N = cout, K = 16 per instruction.
// Images move global<->shared with bulk async copies (TMA, UBLKCP) completing on mbarriers; weights stream through a
// 3-stage ring; one elected thread issues tcgen05.mma; four warps run the epilogue straight out of TMEM.
//
// 4-level networks (dim_mults (1,2,4,8), temporal_unet.py:17-20): 256-channel conv blocks run as gridDim.y = 2 output-channel
// slices of N = 128 (GroupNorm groups are channel-contiguous: a slice owns 4 whole groups, template parameter NG = 4), and ops
// whose source images do not fit shared memory together (the 512-channel concat at L = 8 and the block that adds its 1x1
// residual conv) STREAM them: one image per input phase through the same buffer, accumulating in TMEM across phases.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "tc_common.cuh"

namespace mmdk {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// the layer kernel.  NV = accumulator values per image row and region (n_mt * N), N = columns per m-tile.
// Warps 0-7: epilogue (warp w owns TMEM lanes 32*(w&3).. and the column half w>>2), warp 8: TMEM alloc + MMA issue,
// warp 9: bulk-copy producer.  __launch_bounds__(320, 2): two CTAs per SM so that one CTA's epilogue (CUDA cores)
// overlaps the other's loads + MMAs (tensor pipe).
// ------------------------------------------------------------------------------------------------------------------
template <int NV, int N, int NG>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const __grid_constant__ TcOpParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NMT = NV / N;
  constexpr int NH = (N >= 32) ? N / 2 : N;      // columns per thread and m-tile
  constexpr int NVH = NMT * NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  // blockIdx.y = output-channel slice (256-channel layers of the 4-level net: two slices of N = 128, four GroupNorm groups each --
  // groups are channel-contiguous, so a slice is a self-contained conv block over the same input)
  const int col0 = (int)blockIdx.y * N;
  const uint8_t* const wchunks = p.wchunks + (size_t)blockIdx.y * p.w_split_bytes;
  const int CS = p.cluster;
  const uint32_t crank = (CS > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << CS) - 1u);
  const int tile_ld = tile < p.n_tiles ? tile : p.n_tiles - 1;   // padding CTAs of the last cluster read a valid image
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + p.smem_bar_off;
  // barriers: [0] in_full, [1..3] w_full, [4..6] w_empty, [7] acc_full ; then the TMEM base pointer
  const uint32_t bar_in = bar_base, bar_wfull = bar_base + 8, bar_wempty = bar_base + 8 + 8 * MAX_W_STAGES,
                 bar_acc = bar_base + 8 + 16 * MAX_W_STAGES, bar_in_empty = bar_acc + 16;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + p.smem_bar_off + 8 + 16 * MAX_W_STAGES + 8);
  const int WS = p.w_stages;

  if (tid == 0) {
    mbar_init(bar_in, 1);
    for (int s = 0; s < MAX_W_STAGES; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, (uint32_t)CS); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_in_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                 "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CS > 1) cluster_sync_all();   // barrier inits visible to the peers before any multicast copy / remote arrive
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
#define MMDK_STAMP(k) do { if (p.dbg && tile < p.n_tiles) p.dbg[(size_t)tile * 16 + (k)] = clock64(); } while (0)
  if (tid == 0) { MMDK_STAMP(0); if (p.dbg && tile < p.n_tiles) { unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); p.dbg[(size_t)tile * 16 + 15] = smid; } }

  // programmatic dependent launch: everything above (and the weight prefetch below) overlaps the previous layer's
  // tail; activations written by earlier kernels are only touched after griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 9) {
    // ================= producer: input images, then the weight ring =================
    if (lane == 0) {
      // weights do not depend on the previous kernel: fill the ring first
      int st = 0, ph = 0, c_pre = 0;
      for (; c_pre < p.n_chunks && c_pre < WS; ++c_pre) {
        const ChunkDesc cd = p.chunks[c_pre];
        mbar_expect_tx(bar_wfull + 8 * st, cd.w_bytes);
        if (CS > 1) {
          const uint32_t part = cd.w_bytes / (uint32_t)CS, po = crank * part;
          bulk_g2s_mc(smem_base + p.smem_w_off + st * p.w_stage_bytes + po, wchunks + cd.w_off + po, part, bar_wfull + 8 * st, cmask);
        } else {
          bulk_g2s(smem_base + p.smem_w_off + st * p.w_stage_bytes, wchunks + cd.w_off, cd.w_bytes, bar_wfull + 8 * st);
        }
        if (++st == WS) { st = 0; ph ^= 1; }
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      MMDK_STAMP(1);
      auto load_inputs = [&](int phase) {
        uint32_t total = 0;
        for (int s = 0; s < p.n_src; ++s) if (p.src_phase[s] == phase) total += p.src_tile_bytes[s];
        mbar_expect_tx(bar_in, total);
        for (int s = 0; s < p.n_src; ++s) {
          if (p.src_phase[s] != phase) continue;
          const uint8_t* g = p.src[s] + (size_t)tile_ld * p.src_tile_bytes[s];
          uint32_t off = 0;
          while (off < p.src_tile_bytes[s]) {
            uint32_t n = min(p.src_tile_bytes[s] - off, 32768u);
            bulk_g2s(smem_base + p.src_smem_off[s] + off, g + off, n, bar_in);
            off += n;
          }
        }
      };
      load_inputs(0);
      int in_phase = 0;
      // chunks already in the ring belong to phase 0 (the builder keeps >= MAX_W_STAGES chunks per phase)
      for (int c = c_pre; c < p.n_chunks; ++c) {
        mbar_wait(bar_wempty + 8 * st, ph ^ 1);
        const ChunkDesc cd = p.chunks[c];
        if (cd.phase != in_phase) {
          // the MMA warp commits bar_in_empty after the last MMA that reads the previous phase's input buffer
          mbar_wait(bar_in_empty, (uint32_t)(in_phase & 1));
          in_phase = cd.phase;
          load_inputs(in_phase);
        }
        mbar_expect_tx(bar_wfull + 8 * st, cd.w_bytes);
        if (CS > 1) {
          const uint32_t part = cd.w_bytes / (uint32_t)CS, po = crank * part;
          bulk_g2s_mc(smem_base + p.smem_w_off + st * p.w_stage_bytes + po, wchunks + cd.w_off + po, part, bar_wfull + 8 * st, cmask);
        } else {
          bulk_g2s(smem_base + p.smem_w_off + st * p.w_stage_bytes, wchunks + cd.w_off, cd.w_bytes, bar_wfull + 8 * st);
        }
        if (++st == WS) { st = 0; ph ^= 1; }
      }
      MMDK_STAMP(2);
    }
  } else if (warp == 8) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      // instruction descriptor: D fp32 (bit 4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(bar_in, 0);
      MMDK_STAMP(3);
      int st = 0, ph = 0, in_phase = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        const ChunkDesc cd = p.chunks[c];
        if (cd.phase != in_phase) {
          tc_commit(bar_in_empty);                 // arrives once every MMA issued so far has read its operands
          in_phase = cd.phase;
          mbar_wait(bar_in, (uint32_t)(in_phase & 1));
          tc_fence_after();
        }
        mbar_wait(bar_wfull + 8 * st, ph);
        tc_fence_after();
        const uint32_t wbase = smem_base + p.smem_w_off + st * p.w_stage_bytes;
        const uint32_t b_plane = (uint32_t)cd.k16 * 2u * N * 16u;   // (CK/8) panels * N rows * 16 B
        for (int i = 0; i < NMT; ++i) {
          const uint32_t d_tmem = tmem_base + (uint32_t)cd.acc * 128u + (uint32_t)i * N;
          const uint32_t a_row = smem_base + cd.a_off + (uint32_t)(2 + 128 * i + cd.d) * 16u;
          for (int k = 0; k < cd.k16; ++k) {
            const uint32_t a_hi = a_row + (uint32_t)k * 2u * cd.a_lbo, a_lo = a_hi + cd.a_plane;
            const uint32_t b_hi = wbase + (uint32_t)k * 2u * N * 16u, b_lo = b_hi + b_plane;
            const uint64_t da_hi = make_desc(a_hi, cd.a_lbo, 128), da_lo = make_desc(a_lo, cd.a_lbo, 128);
            const uint64_t db_hi = make_desc(b_hi, N * 16u, 128), db_lo = make_desc(b_lo, N * 16u, 128);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (cd.first && k == 0) ? 0u : 1u);
            tc_mma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
            tc_mma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
          }
        }
        if (CS > 1) tc_commit_mc(bar_wempty + 8 * st, cmask);   // every CTA of the cluster must be done with the stage
        else tc_commit(bar_wempty + 8 * st);                       // frees the weight stage once these MMAs have read it
        if (++st == WS) { st = 0; ph ^= 1; }
      }
      tc_commit(bar_acc);
      MMDK_STAMP(4);
    }
  } else {
    // ================= epilogue: 8 warps; thread = (TMEM lane = image row, column half) =================
    // The accumulators are streamed out of TMEM eight (or sixteen) columns at a time inside rolled loops: TMEM reads
    // are cheap, while a fully unrolled 64-values-in-registers epilogue is ~200 KB of SASS and runs at I-cache-miss
    // speed (ncu: 22% of stall samples "no_instructions" on the unrolled version).
    float4* prm4 = reinterpret_cast<float4*>(smem + p.smem_scratch_off);   // per channel (bias, gamma, beta, cond)
    float* prm_rb = reinterpret_cast<float*>(prm4 + N);                    // residual-conv bias
    float2* part = reinterpret_cast<float2*>(prm_rb + N);               // [NMT*128][8] (sum, M2) per row and group
    float* stat_mean = reinterpret_cast<float*>(part + NMT * 128 * 8);  // [ST][8]
    float* stat_rstd = stat_mean + ST * 8;
    const int P = p.P, L = p.L;
    const int q4 = warp & 3, half = (N >= 32) ? (warp >> 2) : 0;
    const bool active = (N >= 32) || (warp < 4);
    const int row = q4 * 32 + lane;
    const int c0 = half * NH;                                            // first column of this thread inside an m-tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // stage the per-channel parameters while the MMAs run
    if (tid < N) {
      const int cg = col0 + tid;   // global output channel
      prm4[tid] = make_float4((cg < p.cout) ? __ldg(p.bias + cg) : 0.f, p.gamma ? __ldg(p.gamma + cg) : 1.f,
                              p.beta ? __ldg(p.beta + cg) : 0.f, p.cond ? __ldg(p.cond + cg) : 0.f);
      prm_rb[tid] = p.res_bias ? __ldg(p.res_bias + cg) : 0.f;
    }
    epi_bar();
    asm volatile("griddepcontrol.wait;" ::: "memory");   // outputs / residual reads only after the previous grid is complete
    if (tid == 0) MMDK_STAMP(5);
    mbar_wait_sleep(bar_acc, 0);
    tc_fence_after();
    if (tid == 0) MMDK_STAMP(6);

    if (p.kind == TC_CONVBLOCK) {
      if constexpr (N >= 32) {
        constexpr int CPG = N / NG;                // channels per group (NG groups inside this CTA's N columns)
        constexpr int GB = (CPG > 8) ? CPG : 8;    // columns per statistics block
        // ---- GroupNorm statistics: per-row (sum, M2) -> one exchange -> Chan's combination (stable, 2 barriers)
#pragma unroll 1
        for (int i = 0; i < NMT; ++i) {
          const int q = 128 * i + row;
          const int si = q / P, pi = q - si * P;
          const bool ok = (si < ST) && (pi < L) && (tile * ST + si < p.B);
#pragma unroll 1
          for (int gb = 0; gb < NH / GB; ++gb) {
            float w[GB];
            const int cb = c0 + gb * GB;
            if constexpr (GB == 32) tmem_ld32(lane_base + i * N + cb, w);
            else if constexpr (GB == 16) tmem_ld16(lane_base + i * N + cb, w);
            else tmem_ld8(lane_base + i * N + cb, w);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < GB; ++e) w[e] += prm4[cb + e].x;
#pragma unroll
            for (int g = 0; g < GB / CPG; ++g) {
              float sm = 0.f;
#pragma unroll
              for (int c = 0; c < CPG; ++c) sm += w[g * CPG + c];
              const float m = sm * (1.f / CPG);
              float m2 = 0.f;
#pragma unroll
              for (int c = 0; c < CPG; ++c) { const float d = w[g * CPG + c] - m; m2 = fmaf(d, d, m2); }
              part[q * 8 + cb / CPG + g] = ok ? make_float2(sm, m2) : make_float2(0.f, 0.f);
            }
          }
        }
        if (tid == 0) MMDK_STAMP(7);
        epi_bar();
        if (tid == 0) MMDK_STAMP(8);
        if (tid < ST * 8 && (tid & 7) < NG) {
          const int sidx = tid >> 3, g = tid & 7;
          const float inv_n = 1.f / (float)(CPG * L);
          float sum = 0.f;
          for (int pp = 0; pp < L; ++pp) sum += part[(sidx * P + pp) * 8 + g].x;
          const float mean = sum * inv_n;
          float m2 = 0.f;
          for (int pp = 0; pp < L; ++pp) {
            const float2 e = part[(sidx * P + pp) * 8 + g];
            const float d = e.x * (1.f / CPG) - mean;
            m2 += e.y + (float)CPG * d * d;
          }
          stat_mean[tid] = mean;
          stat_rstd[tid] = rsqrtf(m2 * inv_n + 1e-5f);
        }
        epi_bar();
        if (tid == 0) MMDK_STAMP(9);
        // ---- normalise, Mish, +cond, +residual, split, store ----
        const size_t rplane = (size_t)(p.res_id_C / 8) * p.res_id_rows * 16;
        const size_t oplane = (size_t)(p.out_C / 8) * p.out_rows * 16;
#pragma unroll 1
        for (int i = 0; i < NMT; ++i) {
          const int q = 128 * i + row;
          const int si = q / P, pi = q - si * P;
          const bool ok = (si < ST) && (pi < L) && (tile * ST + si < p.B);
          const int r = 2 + q;
          // identity residual: issue the (L2-latency) loads of panel pc+1 before the arithmetic of panel pc
          const uint8_t* rbase = (p.res_id && ok) ? p.res_id + (size_t)tile * p.res_id_tile_bytes + (size_t)r * 16 : nullptr;
          uint4 rh = make_uint4(0, 0, 0, 0), rl = rh;
          if (rbase) {
            rh = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)((col0 + c0) / 8) * p.res_id_rows * 16));
            rl = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)((col0 + c0) / 8) * p.res_id_rows * 16 + rplane));
          }
          uint8_t* obase = p.out + (size_t)tile * p.out_tile_bytes + (size_t)r * 16;
#pragma unroll 1
          for (int pc = 0; pc < NH / 8; ++pc) {
            const int cb = c0 + pc * 8;          // first channel of this panel
            const uint4 ch = rh, cl = rl;
            if (rbase && pc + 1 < NH / 8) {
              rh = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)((col0 + cb) / 8 + 1) * p.res_id_rows * 16));
              rl = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)((col0 + cb) / 8 + 1) * p.res_id_rows * 16 + rplane));
            }
            float y[8], r1[8];
            tmem_ld8(lane_base + i * N + cb, y);
            if (p.res_bias != nullptr) tmem_ld8(lane_base + 128 + i * N + cb, r1);   // residual 1x1 conv: region 1
            tmem_wait_ld();
            if (ok) {
              const int g = cb / CPG;            // CPG >= 4 and 8 | cb: a panel spans 8/CPG groups
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = cb + e;
                const int ge = (CPG >= 8) ? g : g + e / CPG;
                const float4 pr = prm4[c];
                float t = (y[e] + pr.x - stat_mean[si * 8 + ge]) * stat_rstd[si * 8 + ge];
                t = fmaf(t, pr.y, pr.z);
                t = mish_fast(t) + pr.w;
                if (p.res_bias) t += r1[e] + prm_rb[c];
                y[e] = t;
              }
              if (rbase) add8(ch, cl, y);
              uint4 hi, lo;
              split8(y, hi, lo);
              uint8_t* ob = obase + (size_t)((col0 + cb) / 8) * p.out_rows * 16;
              *reinterpret_cast<uint4*>(ob) = hi;
              *reinterpret_cast<uint4*>(ob + oplane) = lo;
            }
          }
        }
      }
    } else if (p.kind == TC_DOWN || p.kind == TC_UP) {
      if constexpr (N >= 32) {
        const int Po = p.out_L + 2;
        const size_t oplane = (size_t)(p.out_C / 8) * p.out_rows * 16;
        const int n_regions = (p.kind == TC_UP) ? 2 : 1;
#pragma unroll 1
        for (int region = 0; region < n_regions; ++region) {
#pragma unroll 1
          for (int i = 0; i < NMT; ++i) {
            const int q = 128 * i + row;
            const int si = q / P, pi = q - si * P;
            bool ok = (si < ST) && (pi < L) && (tile * ST + si < p.B);
            int ro;
            if (p.kind == TC_DOWN) {   // stride-2 conv evaluated at every position; keep the even ones
              ok = ok && ((pi & 1) == 0);
              ro = 2 + si * Po + (pi >> 1);
            } else {                    // transposed conv: region 0 -> output 2p, region 1 -> 2p + 1
              ro = 2 + si * Po + 2 * pi + region;
            }
            uint8_t* obase = p.out + (size_t)tile * p.out_tile_bytes + (size_t)ro * 16;
#pragma unroll 1
            for (int pc = 0; pc < NH / 8; ++pc) {
              const int cb = c0 + pc * 8;
              float y[8];
              tmem_ld8(lane_base + region * 128 + i * N + cb, y);
              tmem_wait_ld();
              if (ok) {
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] += prm4[cb + e].x;
                uint4 hi, lo;
                split8(y, hi, lo);
                uint8_t* ob = obase + (size_t)(cb / 8) * p.out_rows * 16;
                *reinterpret_cast<uint4*>(ob) = hi;
                *reinterpret_cast<uint4*>(ob + oplane) = lo;
              }
            }
          }
        }
      }
    } else if (active) {  // TC_FINAL: eps [B][L][4]
#pragma unroll 1
      for (int i = 0; i < NMT; ++i) {
        const int q = 128 * i + row;
        const int si = q / P, pi = q - si * P;
        const bool ok = (si < ST) && (pi < L) && (tile * ST + si < p.B);
        float y[8];
        tmem_ld8(lane_base + i * N, y);
        tmem_wait_ld();
        if (ok) {
          const size_t b = (size_t)tile * ST + si;
          *reinterpret_cast<float4*>(p.eps + (b * L + pi) * 4) =
              make_float4(y[0] + prm4[0].x, y[1] + prm4[1].x, y[2] + prm4[2].x, y[3] + prm4[3].x);
        }
      }
    }
  }
  if (tid == 0) MMDK_STAMP(10);
  tc_fence_before();
  if (CS > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into / arrive on this CTA's smem
  else __syncthreads();
  if (tid == 0) MMDK_STAMP(11);
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// calibration: a pure tcgen05.mma loop (M=128, N, K=16, kind::f16, SS mode, operands = whatever is in shared memory) to
// calibrate ncu's sm__pipe_tensor_cycles_active against the issue model (N/2 cycles per instruction at M=128) and to
// measure the real cycles per MMA for every N the UNet uses.  out[blockIdx.x] = {cycles, n_mma}.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_calibrate_kernel(int N, int n_iters, int a_rows, int n_acc, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (a_rows * 16 * 2 + N * 16 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  if (warp == 0) {
    // the whole warp walks the loop with uniform operands, the elected lane issues (same scheme as unet_fused.cu):
    // ~2 instructions per MMA, so the measurement is the tensor pipe, not the issue loop
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + (uint32_t)a_rows * 32u;
    const uint32_t a_w = ((a0 >> 4) & 0x3FFFu) | (((uint32_t)a_rows * 16u >> 4) << 16);
    const uint32_t b_w = ((b0 >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
    const uint32_t dstep = (n_acc > 1) ? 256u / (uint32_t)(n_acc > 2 ? 2 : 1) : 0u;
    const long long t0 = clock64();
    for (int it = 0; it < n_iters; it += 4) {
      // walk the A start address like the conv taps do (16-byte row shifts); n_acc accumulators in rotation
      mma_lohi(tmem_base + (n_acc > 1 ? 0u : 0u), a_w + 0, b_w, idesc, it >= 4 ? 1u : 0u, leader);
      mma_lohi(tmem_base + (n_acc > 1 ? dstep : 0u), a_w + 1, b_w, idesc, it >= 4 ? 1u : 0u, leader);
      mma_lohi(tmem_base + (n_acc > 2 ? 2u * dstep : 0u), a_w + 2, b_w, idesc, (it >= 4 || n_acc <= 2) ? 1u : 0u, leader);
      mma_lohi(tmem_base + (n_acc > 2 ? 3u * dstep : (n_acc > 1 ? dstep : 0u)), a_w + 3, b_w, idesc, (it >= 4 || n_acc <= 2) ? 1u : 0u, leader);
    }
    if (leader) tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (tid == 0) {
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = n_iters;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
struct TcOpHost {
  TcOpParams prm{};
  int NV = 0, N = 0, NG = 8;
  size_t smem = 0;
  uint8_t* w_dev = nullptr;
  int cond_off = -1;
};

struct TcState {
  uint64_t uid = 0;
  int n_tiles = 0, B = 0;
  std::vector<TcImage> images;   // [0] = network input image, [1 + j] = output of op j
  std::vector<TcOpHost> ops;
  bool weights_ready = false;
};

static void tc_free(TcState* s) {
  if (!s) return;
  for (auto& im : s->images) cudaFree(im.dev);
  for (auto& o : s->ops) cudaFree(o.w_dev);
  delete s;
}

void unet_tc_release(UnetImpl* net) {
  if (!net) return;
  for (auto& kv : net->tc_cache) tc_free(kv.second);
  net->tc_cache.clear();
  net->tc = nullptr;
}

template <int NV, int N, int NG>
static int launch_tc(const TcOpHost& o, cudaStream_t stream) {
  static bool configured[64] = {};   // the attribute is per device (and per template instantiation)
  int dev = 0;
  MMDK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(MMDK_EINVAL, "device index out of range");
  if (!configured[dev]) {
    MMDK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NV, N, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  const int cs = o.prm.cluster;
  cfg.gridDim = dim3((unsigned)(((o.prm.n_tiles + cs - 1) / cs) * cs), (unsigned)o.prm.n_split);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = o.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  int na = 1;
  if (cs > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cs;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    na = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return check_cuda(cudaLaunchKernelEx(&cfg, conv_tc_kernel<NV, N, NG>, o.prm), "conv_tc_kernel launch");
}

static int build_tc(UnetImpl* net, int B, cudaStream_t stream) {
  const auto& cfg = net->cfg;
  if (cfg.self_attention) return fail(MMDK_EINVAL, "tensor-core executor: LinearAttention not supported");
  auto* st = new TcState();
  st->uid = next_uid();
  st->B = B;
  st->n_tiles = (B + ST - 1) / ST;
  const int n_ops = (int)net->ops.size();
  st->images.resize(n_ops + 1);
  st->ops.resize(n_ops);
  auto fail_free = [&](const std::string& m) { tc_free(st); return fail(MMDK_EINVAL, m); };

  auto make_image = [&](TcImage& im, int C, int L) -> int {
    im.C = C; im.L = L; im.rows = level_rows(L);
    im.tile_bytes = (uint32_t)C * im.rows * 4;
    size_t bytes = (size_t)im.tile_bytes * st->n_tiles;
    if (cudaMalloc(&im.dev, bytes) != cudaSuccess) return MMDK_ENOMEM;
    cudaMemsetAsync(im.dev, 0, bytes, stream);   // halo rows / unused samples stay zero for ever
    return MMDK_OK;
  };
  if (cfg.state_dim > 8) return fail_free("tensor-core executor: state_dim > 8 not supported");
  if (make_image(st->images[0], 16, cfg.horizon) != MMDK_OK) return fail_free("out of memory (input image)");

  for (int j = 0; j < n_ops; ++j) {
    const Op& op = net->ops[j];
    TcOpHost& h = st->ops[j];
    TcOpParams& p = h.prm;
    p.kind = op.type == OP_CONVBLOCK ? TC_CONVBLOCK : op.type == OP_DOWN ? TC_DOWN : op.type == OP_UP ? TC_UP : TC_FINAL;
    p.L = op.lin; p.P = op.lin + 2; p.rows = level_rows(op.lin);
    p.n_mt = (p.rows - 4) / 128;
    p.cout = op.cout;
    p.N = op.cout < 16 ? 16 : op.cout;
    p.n_split = 1;
    if (op.cout == 256 && op.type == OP_CONVBLOCK) {   // dim_mults (1,2,4,8): two slices of 128 channels = 4 GroupNorm groups each
      p.N = 128;
      p.n_split = 2;
    }
    if (p.N != 16 && p.N != 32 && p.N != 64 && p.N != 128) return fail_free("tensor-core executor: unsupported channel count");
    h.N = p.N;
    h.NG = 8 / p.n_split;
    h.NV = p.n_mt * p.N;
    if (!((h.NV == 128 && (p.N == 32 || p.N == 64 || p.N == 128)) || (h.NV == 64 && (p.N == 64 || p.N == 32 || p.N == 16))))
      return fail_free("tensor-core executor: unsupported (rows, channels) combination for this network shape");
    if (p.kind == TC_CONVBLOCK && op.n_groups != 8) return fail_free("tensor-core executor: GroupNorm needs 8 groups");
    if (p.n_split > 1 && h.NV != 128) return fail_free("tensor-core executor: 256-channel layers need a single 128-row m-tile");
    p.B = B; p.n_tiles = st->n_tiles;
    // output image
    if (p.kind != TC_FINAL) {
      if (make_image(st->images[j + 1], op.cout, op.lout) != MMDK_OK) return fail_free("out of memory (activations)");
      const TcImage& oi = st->images[j + 1];
      p.out = oi.dev; p.out_tile_bytes = oi.tile_bytes; p.out_L = oi.L; p.out_rows = oi.rows; p.out_C = oi.C;
    }
    // sources: main (1 or 2), residual conv input (1 or 2)
    struct Src { int prod; };
    std::vector<int> srcs;           // image indices
    auto img_of = [&](int prod) { return prod + 1; };   // -1 (network input) -> image 0
    std::vector<int> main_imgs = {img_of(op.p_src0)};
    if (op.p_src1 > -2) main_imgs.push_back(img_of(op.p_src1));
    std::vector<int> res_imgs;
    const bool res_conv = (p.kind == TC_CONVBLOCK && op.res_w >= 0);
    const bool res_id = (p.kind == TC_CONVBLOCK && op.res_src >= 0 && op.res_w < 0);
    if (res_conv) {
      res_imgs.push_back(img_of(op.p_res));
      if (op.p_res1 > -2) res_imgs.push_back(img_of(op.p_res1));
    }
    // every source image of the op resident at once, or -- when they do not fit (4-level net: the 512-channel concat at L = 8
    // is 2 x 135 KB, and the block that adds its 1x1 residual conv reads a third image) -- STREAMED: one image per input
    // phase through the same buffer (phase k + 1 is loaded once the MMAs of phase k have completed)
    std::vector<int> distinct;
    uint32_t all_src_bytes = 0, max_src_bytes = 0;
    for (const std::vector<int>* v : {&main_imgs, &res_imgs})
      for (int img : *v)
        if (std::find(distinct.begin(), distinct.end(), img) == distinct.end()) {
          distinct.push_back(img);
          const uint32_t b = (st->images[img].tile_bytes + 127) & ~127u;
          all_src_bytes += b;
          max_src_bytes = std::max(max_src_bytes, b);
        }
    const uint32_t scratch = (uint32_t)(5 * p.N + p.n_mt * 128 * 8 * 2 + 2 * ST * 8) * 4;
    const uint32_t tail = ((scratch + 15) & ~15u) + 8 + 16 * MAX_W_STAGES + 8 + 16;
    const uint32_t ring = (uint32_t)MAX_W_STAGES * 32u * (uint32_t)p.N * 4u;
    const bool streamed = all_src_bytes + ring + tail > 232448u;
    if (streamed && (p.kind != TC_CONVBLOCK || max_src_bytes + ring + tail > 232448u || distinct.size() > (size_t)MAX_SRC))
      return fail_free("tensor-core executor: input images do not fit in shared memory");
    p.n_phase = streamed ? (int)distinct.size() : 1;
    uint32_t off = 0, off_phase[MAX_SRC] = {0, 0, 0, 0};
    int cur_phase = 0;
    auto add_src = [&](int img) -> int {
      for (int k = 0; k < p.n_src; ++k) if (p.src[k] == st->images[img].dev) return k;
      int k = p.n_src++;
      p.src[k] = st->images[img].dev;
      p.src_tile_bytes[k] = st->images[img].tile_bytes;
      p.src_phase[k] = cur_phase;
      p.src_smem_off[k] = off_phase[cur_phase];
      off_phase[cur_phase] += (st->images[img].tile_bytes + 127) & ~127u;
      off = std::max(off, off_phase[cur_phase]);
      return k;
    };
    // chunk list
    std::vector<ChunkDesc> chunks;
    struct WSrc { const float* W; int cin, ktaps, tap, ci0, CK; };
    std::vector<WSrc> wsrc;
    uint32_t w_total = 0, w_stage = 0;
    auto add_chunks = [&](const std::vector<int>& imgs, int cin_total, int w_off_blob, int ktaps, int tap, int d, int acc,
                          bool first_in_acc, int only = -1) {
      int ci = 0;
      bool first = first_in_acc;
      for (size_t si = 0; si < imgs.size(); ++si) {
        const TcImage& im = st->images[imgs[si]];
        const int c_here = im.C;                       // channels of this source image (input image: 16, real 4)
        if (only >= 0 && (int)si != only) { ci += (imgs[si] == 0) ? cfg.state_dim : c_here; continue; }
        const int slot = add_src(imgs[si]);
        for (int c0 = 0; c0 < c_here; c0 += 32) {
          const int CK = std::min(32, c_here - c0);
          ChunkDesc cd{};
          cd.a_off = p.src_smem_off[slot] + (uint32_t)(c0 / 8) * im.rows * 16;
          cd.a_plane = (uint32_t)(im.C / 8) * im.rows * 16;
          cd.a_lbo = (uint32_t)im.rows * 16;
          cd.w_bytes = (uint32_t)CK * p.N * 4;
          cd.w_off = w_total;
          w_total += cd.w_bytes;
          w_stage = std::max(w_stage, cd.w_bytes);
          cd.d = d; cd.acc = acc; cd.first = first ? 1 : 0; cd.k16 = CK / 16;
          cd.phase = cur_phase;
          first = false;
          chunks.push_back(cd);
          wsrc.push_back({net->blob + w_off_blob, cin_total, ktaps, tap, ci + c0, CK});
        }
        ci += (imgs[si] == 0) ? cfg.state_dim : c_here;
      }
    };
    if (streamed) {
      bool first0 = true, first1 = true;
      for (cur_phase = 0; cur_phase < (int)distinct.size(); ++cur_phase) {
        const int img = distinct[cur_phase];
        for (size_t k = 0; k < main_imgs.size(); ++k)
          if (main_imgs[k] == img)
            for (int tap = 0; tap < 5; ++tap) { add_chunks(main_imgs, op.cin, op.w, 5, tap, tap - 2, 0, first0, (int)k); first0 = false; }
        if (res_conv)
          for (size_t k = 0; k < res_imgs.size(); ++k)
            if (res_imgs[k] == img) { add_chunks(res_imgs, op.res_cin, op.res_w, 1, 0, 0, 1, first1, (int)k); first1 = false; }
      }
      cur_phase = 0;
    } else if (p.kind == TC_CONVBLOCK) {
      for (int tap = 0; tap < 5; ++tap) add_chunks(main_imgs, op.cin, op.w, 5, tap, tap - 2, 0, tap == 0);
      if (res_conv) add_chunks(res_imgs, op.res_cin, op.res_w, 1, 0, 0, 1, true);
    } else if (p.kind == TC_DOWN) {
      for (int tap = 0; tap < 3; ++tap) add_chunks(main_imgs, op.cin, op.w, 3, tap, tap - 1, 0, tap == 0);
    } else if (p.kind == TC_UP) {
      // out[2m] = x[m] w1 + x[m-1] w3 ; out[2m+1] = x[m+1] w0 + x[m] w2   (ConvTranspose1d k4 s2 p1)
      add_chunks(main_imgs, op.cin, op.w, 4, 1, 0, 0, true);
      add_chunks(main_imgs, op.cin, op.w, 4, 3, -1, 0, false);
      add_chunks(main_imgs, op.cin, op.w, 4, 0, 1, 1, true);
      add_chunks(main_imgs, op.cin, op.w, 4, 2, 0, 1, false);
    } else {
      add_chunks(main_imgs, op.cin, op.w, 1, 0, 0, 0, true);
    }
    p.n_chunks = (int)chunks.size();
    p.w_stage_bytes = (w_stage + 127) & ~127u;
    p.smem_w_off = off;
    // two CTAs per SM (one's epilogue under the other's MMAs) need <= ~112 KB each: shrink the weight ring if that helps
    p.w_stages = MAX_W_STAGES;
    if (off + MAX_W_STAGES * p.w_stage_bytes + tail > 114688u && off + 2 * p.w_stage_bytes + tail <= 114688u) p.w_stages = 2;
    off += p.w_stages * p.w_stage_bytes;
    p.smem_scratch_off = off;
    off += (scratch + 15) & ~15u;
    p.smem_bar_off = off;
    off += 8 + 16 * MAX_W_STAGES + 8 + 16;
    h.smem = off;
    if (h.smem > 232448) return fail_free("tensor-core executor: layer does not fit in shared memory");
    // weights + chunk table to the device
    if (chunks.size() > (size_t)MAX_CHUNKS) return fail_free("tensor-core executor: too many weight chunks in one layer");
    if (cudaMalloc(&h.w_dev, (size_t)w_total * p.n_split) != cudaSuccess) return fail_free("out of memory (packed weights)");
    p.w_split_bytes = w_total;
    for (int sl = 0; sl < p.n_split; ++sl)
      for (size_t c = 0; c < chunks.size(); ++c) {
        const WSrc& w = wsrc[c];
        pack_wchunk_kernel<<<8, 256, 0, stream>>>(w.W, w.cin, w.ktaps, op.cout, w.tap, w.ci0, w.CK, p.N,
                                                  reinterpret_cast<__half*>(h.w_dev + (size_t)sl * w_total + chunks[c].w_off), sl * p.N);
      }
    for (size_t c = 0; c < chunks.size(); ++c) p.chunks[c] = chunks[c];
    p.cluster = TC_CLUSTER;
    p.wchunks = h.w_dev;
    p.bias = net->blob + op.b;
    if (p.kind == TC_CONVBLOCK) {
      p.gamma = net->blob + op.gn_w;
      p.beta = net->blob + op.gn_b;
      h.cond_off = op.cond;
      if (res_conv) p.res_bias = net->blob + op.res_b;
      if (res_id) {
        const TcImage& ri = st->images[img_of(op.p_res)];
        p.res_id = ri.dev; p.res_id_tile_bytes = ri.tile_bytes; p.res_id_rows = ri.rows; p.res_id_C = ri.C;
      }
    }
  }
  if (check_cuda(cudaGetLastError(), "tensor-core executor setup") != MMDK_OK) { tc_free(st); return MMDK_ECUDA; }
  net->tc = st;
  return MMDK_OK;
}

uint64_t unet_tc_state_uid(const UnetImpl* net, int B) {
  auto it = net->tc_cache.find(B);
  return it == net->tc_cache.end() ? 0 : it->second->uid;
}

int unet_forward_tc(UnetImpl* net, int mode, const float* x, int B, int t, float* eps, cudaStream_t stream) {
  (void)mode;
  // one state (activation images + packed weights) per batch size: a planner warms up at B = 2 and then samples at B = K,
  // CBS alternates single calls and batches -- none of that may rebuild (cudaFree / cudaMalloc / weight packing) every time
  if (!net->tc || net->tc->B != B) {
    auto it = net->tc_cache.find(B);
    if (it != net->tc_cache.end()) {
      net->tc = it->second;
    } else {
      if (net->tc_cache.size() >= 4) {   // bounded: drop everything that was cached and start over
        MMDK_CUDA(cudaStreamSynchronize(stream));
        unet_tc_release(net);
      }
      net->tc = nullptr;
      int rc = build_tc(net, B, stream);
      if (rc != MMDK_OK) return rc;
      net->tc_cache[B] = net->tc;
    }
  }
  TcState* st = net->tc;
  const auto& cfg = net->cfg;
  {
    const TcImage& im = st->images[0];
    const int n = B * cfg.horizon;
    pack_input_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x, B, cfg.horizon, cfg.state_dim, im.rows, im.dev, im.tile_bytes);
  }
  const float* cond_row = net->cond_table + (size_t)t * net->n_cond;
  for (auto& h : st->ops) {
    h.prm.cond = h.cond_off >= 0 ? cond_row + h.cond_off : nullptr;
    h.prm.eps = eps;
    int rc;
    if (h.NV == 128 && h.N == 32) rc = launch_tc<128, 32, 8>(h, stream);
    else if (h.NV == 128 && h.N == 64) rc = launch_tc<128, 64, 8>(h, stream);
    else if (h.NV == 128 && h.N == 128 && h.NG == 4) rc = launch_tc<128, 128, 4>(h, stream);
    else if (h.NV == 128 && h.N == 128) rc = launch_tc<128, 128, 8>(h, stream);
    else if (h.NV == 64 && h.N == 64) rc = launch_tc<64, 64, 8>(h, stream);
    else if (h.NV == 64 && h.N == 32) rc = launch_tc<64, 32, 8>(h, stream);
    else rc = launch_tc<64, 16, 8>(h, stream);
    if (rc != MMDK_OK) return rc;
  }
  return MMDK_OK;
}

int unet_tc_timeline(UnetImpl* net, int op_index, long long* dbg_dev, cudaStream_t stream) {
  if (!net->tc) return fail(MMDK_EINVAL, "run a tensor-core forward first");
  TcState* st = net->tc;
  if (op_index >= (int)st->ops.size()) return fail(MMDK_EINVAL, "op index out of range");
  for (size_t j = 0; j < st->ops.size(); ++j) st->ops[j].prm.dbg = ((int)j == op_index) ? dbg_dev : nullptr;
  return MMDK_OK;
}

int unet_tc_tap(UnetImpl* net, int op_index, float* out, int* c_out, int* l_out, cudaStream_t stream) {
  if (!net->tc) return fail(MMDK_EINVAL, "run a tensor-core forward first");
  TcState* st = net->tc;
  if (op_index < -1 || op_index + 1 >= (int)st->images.size() || !st->images[op_index + 1].dev)
    return fail(MMDK_EINVAL, "op index out of range (or op has no activation image)");
  const TcImage& im = st->images[op_index + 1];
  if (c_out) *c_out = im.C;
  if (l_out) *l_out = im.L;
  if (!out) return MMDK_OK;
  const int n = st->B * im.C * im.L;
  unpack_image_kernel<<<(n + 255) / 256, 256, 0, stream>>>(im.dev, im.tile_bytes, st->B, im.C, im.L, im.rows, out);
  return check_cuda(cudaGetLastError(), "unpack_image_kernel");
}

int mma_calibrate(int N, int n_iters, int n_ctas, int n_acc, long long* out_dev, cudaStream_t stream) {
  if (N < 16 || N > 256 || (N & 15) || n_iters < 4 || (n_iters & 3) || n_ctas < 1 || n_acc < 1 || n_acc > 4 || (n_acc > 2 && N > 128) ||
      (n_acc > 1 && N > 256))
    return fail(MMDK_EINVAL, "mma_calibrate: bad arguments");
  const int a_rows = 128 + 8;
  const size_t smem = (size_t)a_rows * 32 + (size_t)N * 32 + 128;
  mma_calibrate_kernel<<<n_ctas, 128, smem, stream>>>(N, n_iters, a_rows, n_acc, out_dev);
  return check_cuda(cudaGetLastError(), "mma_calibrate_kernel");
}

}  // namespace mmdk
