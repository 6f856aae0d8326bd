// Guide + DDPM reverse step, fp32, closed-form gradients (no autograd), one launch per reverse step.
//
// Reference: mmd/models/diffusion_models/guides.py:180-253 (GuideManagerTrajectoriesWithVelocity.forward),
// sample_functions.py:41-107 (ddpm_sample_fn, guide_gradient_steps, apply_hard_conditioning),
// diffusion_model_base.py:126-160 (p_mean_variance), MPB/planners/costs/cost_functions.py:175-193,297-326,532-542,
// MPB/planners/costs/factors/{field,gp}_factor.py, TR/environments/grid_map_sdf.py:84-114,
// TR/torch_planning_objectives/fields/distance_fields.py:110-129,354-367, mmd/datasets/normalization.py:157-168.
//
// Mapping: one thread per (sample, waypoint); a CTA owns `spc` samples of ONE group; the CTAs of a group form a
// thread-block cluster, because LimitsNormalizer.unnormalize takes a data-dependent decision over the whole [K,H,D]
// batch of a planner call (clip everything iff any element leaves [-1-1e-4, 1+1e-4]); the per-iteration flag is
// exchanged through distributed shared memory + one cluster barrier.
//
// This file is compiled with -fmad=false: the reference evaluates these expressions op by op in fp32, and the
// discontinuities downstream (floor to a grid cell, hinge, in/out of radius) make single-rounding FMAs visible.
#include <cooperative_groups.h>

#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mmdk {

constexpr int kMaxIters = 32;   // guide iterations per launch served by the lazy cluster flag exchange

struct PubArgs {
  int world, rank, row_offset, n_rows;   // ranks of the fleet, this rank, first table row of this launch's groups, rows per table
  float* tables[MMDK_MAX_RANKS];         // peer table of every rank as mapped on THIS device: [2][n_rows][H][2] (world == 1: [n_rows][H][2])
  uint32_t* flags[MMDK_MAX_RANKS];       // flag array of every rank: flags[r][q] = last sequence number rank q published to rank r
  uint32_t* state;                       // local: [0] publish sequence, [1] consume sequence, [2] group counter, [3] error flag
};

__device__ __forceinline__ void st_release_cluster(int* p, int v) {
  asm volatile("st.release.cluster.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cluster(const int* p) {
  int v;
  asm volatile("ld.acquire.cluster.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The normaliser's clip decision of a group (any element outside [-1 - 1e-4, 1 + 1e-4] => clamp everything to [-1, 1]) only
// changes a thread's result if one of ITS elements lies outside [-1, 1].  So the CTAs of a group never barrier on it: each
// CTA stores its flag into slot [round][rank] of every CTA of the cluster (release) and only the (rare) threads that hold an
// element in the ambiguous band, with no flagged element in their own CTA, wait for the other CTAs' slots (acquire).
struct GroupFlag {
  int* slots;   // [kMaxIters + 1][16] in this CTA's shared memory, -1 = not yet published
  int cpg, rank, tid;

  __device__ __forceinline__ int decide(int round, int mine, int band) const {
    const int cta_flag = __syncthreads_or(mine);
    if (cpg == 1) return cta_flag;
    if (tid < cpg) st_release_cluster(cg::this_cluster().map_shared_rank(slots + round * 16 + rank, tid), cta_flag);
    int flag = cta_flag;
    if (!flag && band) {
      for (int r = 0; r < cpg; ++r) {
        int f;
        while ((f = ld_acquire_cluster(slots + round * 16 + r)) < 0) {}
        flag |= f;
      }
    }
    return flag;
  }
};

struct StepArgs {
  mmdk_guide_env env;
  mmdk_groups grp;
  mmdk_step_scalars sc;
  int H;
  int spc;  // samples per CTA
  int cpg;  // CTAs per group (cluster size)
  float* x;
  const float* eps;
  const float* noise;
  float* chain;
  float* grad_out;  // taps mode
  float* raw_out;
  int n_costs_max;
  int peers_in_smem;  // the [n_peers, H, 2] table is staged in shared memory once per launch (it is constant during it)
  int hash_in_smem;   // the sorted peer positions of the hash are staged in shared memory once per launch
  int cs_in_smem;     // the hash's cell_start table is staged in shared memory (fits long after the positions stop fitting)
  // publication of every group's representative sample at the END of the step (it is the x the next step starts from):
  // fused all-gather over peer memory -- the row goes straight into the peer table of every rank (NVLink stores), then
  // the last group to finish releases this rank's sequence number into every rank's flag array
  int pub_on;
  int pub_rep;        // representative sample of a group
  PubArgs pub;
};

__device__ __forceinline__ void clip_by_norm(float g[4], float max_norm) {
  // guides.py:247-253: n = ||g + 1e-6||_2 ; g *= clip(n, 0, max)/n
  float a0 = g[0] + 1e-6f, a1 = g[1] + 1e-6f, a2 = g[2] + 1e-6f, a3 = g[3] + 1e-6f;
  float n = sqrtf(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3);
  float ratio = fminf(fmaxf(n, 0.f), max_norm) / n;
#pragma unroll
  for (int d = 0; d < 4; ++d) g[d] = ratio * g[d];
}

__device__ __forceinline__ void cell_index(const mmdk_guide_env& e, float px, float py, int& ix, int& iy) {
  // grid_map_sdf.py:93-97: ((X - lo) / map_dim * cmap_dim).floor().int(), clamped to the grid
  float fx = floorf(((px - e.grid_lo[0]) / e.grid_map_dim[0]) * (float)e.nx);
  float fy = floorf(((py - e.grid_lo[1]) / e.grid_map_dim[1]) * (float)e.ny);
  fx = fminf(fmaxf(fx, 0.f), (float)(e.nx - 1));
  fy = fminf(fmaxf(fy, 0.f), (float)(e.ny - 1));
  ix = (int)fx;
  iy = (int)fy;
}

// cell of the lock-step peer hash (uniform grid over [lo, lo + G / inv_cell)^2, clamped): used by the build kernel and
// by the query with the SAME fp32 expressions, so a point always lands in the cell the builder put it in
__device__ __forceinline__ void peer_cell(float lo, float inv_cell, int G, float px, float py, int& cx, int& cy) {
  float fx = floorf((px - lo) * inv_cell), fy = floorf((py - lo) * inv_cell);
  fx = fminf(fmaxf(fx, 0.f), (float)(G - 1));
  fy = fminf(fmaxf(fy, 0.f), (float)(G - 1));
  cx = (int)fx;
  cy = (int)fy;
}

template <bool TAPS>
__global__ void __launch_bounds__(512, 2) ddpm_step_kernel(const StepArgs a) {
  // 512 threads (8 samples at H = 64), two CTAs per SM, clusters of up to 16: 512 CTAs of the 32 x 128 bench batch are
  // exactly two waves of 16 clusters x 16 CTAs (round 2: 1024-thread CTAs in clusters of 8 kept only 15 clusters = 120 SMs
  // resident and needed three waves for 2.13 waves of work)
  extern __shared__ float4 s_xu_all[];   // [2][spc][H] unnormalised states of this CTA's samples, double buffered
  __shared__ int s_flags[(kMaxIters + 1) * 16];   // clip flags of the cluster's CTAs, one row per guide iteration + publication

  const int H = a.H;
  const int tid = threadIdx.x;
  const int sl = tid / H;            // local sample
  const int h = tid - sl * H;        // waypoint
  const int rank = (a.cpg > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int g = blockIdx.x / a.cpg;
  const int sidx = rank * a.spc + sl;
  const bool valid = sidx < a.grp.K;
  const size_t b = (size_t)g * a.grp.K + (valid ? sidx : 0);
  const size_t off = (b * H + h) * 4;
  const mmdk_guide_env& E = a.env;

  for (int i = tid; i < (kMaxIters + 1) * 16; i += blockDim.x) s_flags[i] = -1;
  const GroupFlag gflag{s_flags, a.cpg, rank, tid};

  // hard conditions of this group that hit my waypoint (last one wins, dict order)
  bool hc = false;
  float hv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < MMDK_MAX_HARD_ROWS; ++r) {
    int row = a.grp.hard_rows_dev[g * MMDK_MAX_HARD_ROWS + r];
    if (row == h) {
      hc = true;
      const float* v = a.grp.hard_vals_dev + ((size_t)g * MMDK_MAX_HARD_ROWS + r) * 4;
#pragma unroll
      for (int d = 0; d < 4; ++d) hv[d] = v[d];
    }
  }

  float x[4];
  {
    float4 v = *reinterpret_cast<const float4*>(a.x + off);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
  if (a.sc.do_posterior) {
    // diffusion_model_base.py:126-160: x0 = c1 x - c2 eps ; clamp ; mean = k1 x0 + k2 x
    float4 ev = *reinterpret_cast<const float4*>(a.eps + off);
    float e[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float x0 = a.sc.predict_epsilon
                     ? (a.sc.sqrt_recip_alphas_cumprod * x[d] - a.sc.sqrt_recipm1_alphas_cumprod * e[d])
                     : e[d];
      if (a.sc.clip_denoised) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      x[d] = a.sc.posterior_mean_coef1 * x0 + a.sc.posterior_mean_coef2 * x[d];
    }
  }
  if (a.cpg > 1) cg::this_cluster().sync();  // flags zeroed everywhere before any remote write
  else __syncthreads();

  // lock-step peers: constant for the whole launch -> one coalesced copy into shared memory, 20 x reuse
  float2* s_peers = reinterpret_cast<float2*>(s_xu_all + 2 * (size_t)a.spc * H);
  float4* s_sorted = reinterpret_cast<float4*>(s_peers);                                  // hash staging aliases the
  unsigned short* s_cs = reinterpret_cast<unsigned short*>(s_sorted + (a.hash_in_smem ? (size_t)H * a.grp.n_peers : 0));   // brute-force staging
  // double-buffered table: readers use the half of the current sequence number, the publication at the end of this launch
  // writes the other half
  const size_t peer_half = (a.grp.peers_dev && a.grp.peer_seq_dev) ? (size_t)(*a.grp.peer_seq_dev & 1u) * a.grp.n_peers * H : 0;
  if (a.grp.peers_dev && a.peers_in_smem && !a.grp.peer_cell_start_dev) {
    const float2* gp = reinterpret_cast<const float2*>(a.grp.peers_dev) + peer_half;
    for (int i = tid; i < a.grp.n_peers * H; i += blockDim.x) s_peers[i] = __ldg(gp + i);
    __syncthreads();
  }
  if (a.hash_in_smem || a.cs_in_smem) {
    const int G2 = a.grp.peer_grid * a.grp.peer_grid + 1;
    if (a.hash_in_smem) {
      const float4* gs = reinterpret_cast<const float4*>(a.grp.peer_sorted_dev);
      for (int i = tid; i < H * a.grp.n_peers; i += blockDim.x) s_sorted[i] = __ldg(gs + i);
    }
    for (int i = tid; i < H * G2; i += blockDim.x) s_cs[i] = a.grp.peer_cell_start_dev[i];
    __syncthreads();
  }
  const int n_iter = TAPS ? 1 : a.sc.n_guide_steps;
  int o_beg = 0, o_end = 0;
  if (a.grp.obj_ptr_dev) { o_beg = a.grp.obj_ptr_dev[g]; o_end = a.grp.obj_ptr_dev[g + 1]; }
  const int self_peer = a.grp.peers_dev ? a.grp.peer_self_dev[g] : -1;

  for (int it = 0; it < n_iter; ++it) {
    // ---- G3: LimitsNormalizer.unnormalize with its global, data-dependent clip ------------------------------
    int mine = 0, band = 0;
    if (valid) {
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        mine |= (x[d] > 1.0001f) | (x[d] < -1.0001f);
        band |= (x[d] > 1.f) | (x[d] < -1.f);
      }
    }
    const int flag = gflag.decide(it, mine, band);
    float xu[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float v = flag ? fminf(fmaxf(x[d], -1.f), 1.f) : x[d];
      v = (v + 1.f) / 2.f;
      xu[d] = v * E.norm_range[d] + E.norm_min[d];
    }
    // double buffered: the barrier at the top of the NEXT iteration (clip flag) separates this iteration's neighbour
    // reads from the writes two iterations later, so no barrier is needed at the end of an iteration
    float4* s_xu = s_xu_all + (size_t)(it & 1) * a.spc * H;
    s_xu[sl * H + h] = make_float4(xu[0], xu[1], xu[2], xu[3]);
    __syncthreads();

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool interior = (h != 0) && (h != H - 1);
    int tap = 0;
    auto emit = [&](float gk[4], float w) {
      if (TAPS && a.raw_out && valid && tap < a.n_costs_max) {
        float* r = a.raw_out + ((size_t)tap * a.grp.n_groups * a.grp.K) * H * 4 + off;
        *reinterpret_cast<float4*>(r) = make_float4(gk[0], gk[1], gk[2], gk[3]);
      }
      ++tap;
      // clip_by_norm of an all-zero gradient is exactly zero (ratio = clip(2e-6, 0, max) / 2e-6 is finite and positive):
      // most waypoints are neither in collision nor near a constraint, so skip the sqrt / division for them
      if (gk[0] != 0.f || gk[1] != 0.f || gk[2] != 0.f || gk[3] != 0.f) clip_by_norm(gk, E.max_grad_norm);
      if (!interior) { gk[0] = gk[1] = gk[2] = gk[3] = 0.f; }   // guides.py:216-218
#pragma unroll
      for (int d = 0; d < 4; ++d) acc[d] = acc[d] + w * gk[d];
    };

    // ---- G5/G6: object collision through the SDF grid (waypoints 1..H-1) -----------------------------------
    if (E.grid_dev) {
      float gk[4] = {0.f, 0.f, 0.f, 0.f};
      if (h >= 1) {
        int ix, iy;
        cell_index(E, xu[0], xu[1], ix, iy);
        float4 c = __ldg(reinterpret_cast<const float4*>(E.grid_dev) + (size_t)ix * E.ny + iy);
        float v = -(c.x - E.margin);
        if (v > 0.f) { gk[0] = -(E.coll_inv_sigma2 * c.y); gk[1] = -(E.coll_inv_sigma2 * c.z); }
      }
      emit(gk, E.w_collision);
    }
    // ---- G7: workspace boundaries ------------------------------------------------------------------------------
    {
      float gk[4] = {0.f, 0.f, 0.f, 0.f};
      if (h >= 1) {
        float s0 = xu[0] - E.ws_min[0], s1 = xu[1] - E.ws_min[1];
        float s2 = E.ws_max[0] - xu[0], s3 = E.ws_max[1] - xu[1];
        float v0 = fmaxf(-(s0 - E.margin), 0.f), v1 = fmaxf(-(s1 - E.margin), 0.f);
        float v2 = fmaxf(-(s2 - E.margin), 0.f), v3 = fmaxf(-(s3 - E.margin), 0.f);
        int k = 0; float vm = v0;
        if (v1 > vm) { vm = v1; k = 1; }
        if (v2 > vm) { vm = v2; k = 2; }
        if (v3 > vm) { vm = v3; k = 3; }
        if (vm > 0.f) {
          const float K = E.coll_inv_sigma2;
          if (k == 0) gk[0] = -K; else if (k == 1) gk[1] = -K; else if (k == 2) gk[0] = K; else gk[1] = K;
        }
      }
      emit(gk, E.w_border);
    }
    // ---- G8: GP prior, two-waypoint stencil -------------------------------------------------------------------
    {
      float gk[4] = {0.f, 0.f, 0.f, 0.f};
      const float dt = E.dt, q11 = E.gp_q11, q12 = E.gp_q12, q22 = E.gp_q22;
      if (h >= 1) {  // factor h-1: d/dx_h = 2 Q^-1 e_{h-1}
        float4 p = s_xu[sl * H + h - 1];
        float ep0 = xu[0] - (p.x + dt * p.z), ep1 = xu[1] - (p.y + dt * p.w);
        float ev0 = xu[2] - p.z, ev1 = xu[3] - p.w;
        gk[0] += 2.f * (q11 * ep0 + q12 * ev0);
        gk[1] += 2.f * (q11 * ep1 + q12 * ev1);
        gk[2] += 2.f * (q12 * ep0 + q22 * ev0);
        gk[3] += 2.f * (q12 * ep1 + q22 * ev1);
      }
      if (h <= H - 2) {  // factor h: d/dx_h = -2 Phi^T Q^-1 e_h
        float4 nx = s_xu[sl * H + h + 1];
        float ep0 = nx.x - (xu[0] + dt * xu[2]), ep1 = nx.y - (xu[1] + dt * xu[3]);
        float ev0 = nx.z - xu[2], ev1 = nx.w - xu[3];
        float qp0 = q11 * ep0 + q12 * ev0, qp1 = q11 * ep1 + q12 * ev1;
        float qv0 = q12 * ep0 + q22 * ev0, qv1 = q12 * ep1 + q22 * ev1;
        gk[0] -= 2.f * qp0;
        gk[1] -= 2.f * qp1;
        gk[2] -= 2.f * (dt * qp0 + qv0);
        gk[3] -= 2.f * (dt * qp1 + qv1);
      }
      emit(gk, E.w_smooth);
    }
    // ---- G9: CostConstraint objects (each clipped and weighted separately) ------------------------------------
    for (int o = o_beg; o < o_end; ++o) {
      float gk[4] = {0.f, 0.f, 0.f, 0.f};
      const int* bp = a.grp.bucket_ptr_dev + (size_t)o * (H + 1);
      const int e0 = bp[h], e1 = bp[h + 1];
      for (int e = e0; e < e1; ++e) {
        float4 c = __ldg(reinterpret_cast<const float4*>(a.grp.cons_dev) + e);
        float dx = xu[0] - c.x, dy = xu[1] - c.y;
        float d2 = dx * dx + dy * dy;
        // conservative prefilter: sqrt(d2) > r for sure (the exact reference test below decides the rest)
        if (d2 > c.z * c.z * 1.0001f) continue;
        float dist = sqrtf(d2);
        if (!(dist > c.z) && dist > 0.f) { gk[0] -= dx / dist; gk[1] -= dy / dist; }
      }
      emit(gk, a.grp.obj_weight_dev[o]);
    }
    // ---- lock-step peers: one implicit soft CostConstraint with ranges (h, h+1) ----------------------------------
    // The in/out-of-radius decision is the reference's exact fp32 test (cost_functions.py:316-318) in both paths; the
    // hashed path only restricts WHICH peers are tested: per waypoint the [n_peers] table is bucketed into a uniform grid
    // with cell >= radius (mmdk_build_peer_hash), so the 3x3 cells around the waypoint hold every peer within the radius.
    if (a.grp.peers_dev) {
      float gk[4] = {0.f, 0.f, 0.f, 0.f};
      const float r = a.grp.peer_radius;
      const float r2_far = r * r * 1.0001f;
      if (a.grp.peer_cell_start_dev) {
        const int G = a.grp.peer_grid;
        const unsigned short* cs = (a.cs_in_smem ? s_cs : a.grp.peer_cell_start_dev) + (size_t)h * (G * G + 1);
        const float4* sp = (a.hash_in_smem ? s_sorted : reinterpret_cast<const float4*>(a.grp.peer_sorted_dev)) + (size_t)h * a.grp.n_peers;
        int cx, cy;
        peer_cell(a.grp.peer_grid_lo, a.grp.peer_grid_inv_cell, G, xu[0], xu[1], cx, cy);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G - 1);
        const int y0 = max(cy - 1, 0), y1 = min(cy + 1, G - 1);
        for (int yy = y0; yy <= y1; ++yy) {
          const int e0 = cs[yy * G + x0], e1 = cs[yy * G + x1 + 1];
          for (int e = e0; e < e1; ++e) {
            const float4 q = sp[e];
            if (__float_as_int(q.z) == self_peer) continue;
            float dx = xu[0] - q.x, dy = xu[1] - q.y;
            float d2 = dx * dx + dy * dy;
            if (d2 > r2_far) continue;
            float dist = sqrtf(d2);   // exact in/out-of-radius decision; the unit vector uses one reciprocal (<= 1 ulp)
            if (!(dist > r) && dist > 0.f) { const float inv = 1.f / dist; gk[0] -= dx * inv; gk[1] -= dy * inv; }
          }
        }
      } else {
        const float2* pq = (a.peers_in_smem ? s_peers : reinterpret_cast<const float2*>(a.grp.peers_dev) + peer_half) + h;
        // two passes over blocks of 32 peers: (1) branch-free squared-distance prefilter into a bit mask (in-radius peers are
        // rare), (2) the exact reference test (sqrt, dist <= r) and the accumulation for the set bits only, in ascending peer
        // order -- the same peers contribute in the same order as a plain scan
        for (int j0 = 0; j0 < a.grp.n_peers; j0 += 32) {
          const int nb = min(32, a.grp.n_peers - j0);
          unsigned mask = 0u;
#pragma unroll 8
          for (int jj = 0; jj < nb; ++jj) {
            const float2 q = pq[(size_t)(j0 + jj) * H];
            const float dx = xu[0] - q.x, dy = xu[1] - q.y;
            const float d2 = dx * dx + dy * dy;
            mask |= (d2 > r2_far) ? 0u : (1u << jj);   // surely outside the radius otherwise
          }
          if (self_peer >= j0 && self_peer < j0 + 32) mask &= ~(1u << (self_peer - j0));
          while (mask) {
            const int jj = __ffs(mask) - 1;
            mask &= mask - 1u;
            const float2 q = pq[(size_t)(j0 + jj) * H];
            const float dx = xu[0] - q.x, dy = xu[1] - q.y;
            const float dist = sqrtf(dx * dx + dy * dy);
            if (!(dist > r) && dist > 0.f) { const float inv = 1.f / dist; gk[0] -= dx * inv; gk[1] -= dy * inv; }
          }
        }
      }
      emit(gk, a.grp.peer_weight);
    }

    if (TAPS) {
      if (valid) *reinterpret_cast<float4*>(a.grad_out + off) = make_float4(-acc[0], -acc[1], -acc[2], -acc[3]);
    } else {
      // sample_functions.py:104-105: x = x + (-1 * grad) ; apply_hard_conditioning
#pragma unroll
      for (int d = 0; d < 4; ++d) x[d] = x[d] + (-1.f * acc[d]);
      if (hc) {
#pragma unroll
        for (int d = 0; d < 4; ++d) x[d] = hv[d];
      }
    }
  }

  if (!TAPS) {
    if (a.sc.add_noise && a.noise) {
      float4 nv = *reinterpret_cast<const float4*>(a.noise + off);
      float n[4] = {nv.x, nv.y, nv.z, nv.w};
#pragma unroll
      for (int d = 0; d < 4; ++d) x[d] = x[d] + (a.sc.model_std * n[d]) * a.sc.noise_std;
    }
    if (hc && a.sc.final_hard_conds) {
#pragma unroll
      for (int d = 0; d < 4; ++d) x[d] = hv[d];
    }
    if (valid) {
      float4 v = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(a.x + off) = v;
      if (a.chain) *reinterpret_cast<float4*>(a.chain + off) = v;
    }
    if (a.pub_on) {
      // ---- lock-step publication (mmdk_publish_peers semantics) of the x the NEXT step starts from ---------------------
      int mine = 0, band = 0;
      if (valid) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          mine |= (x[d] > 1.0001f) | (x[d] < -1.0001f);
          band |= (x[d] > 1.f) | (x[d] < -1.f);
        }
      }
      const bool is_rep = valid && (sidx == a.pub_rep);
      const int flag = gflag.decide(kMaxIters, mine, is_rep ? band : 0);
      const PubArgs& pb = a.pub;
      if (is_rep) {
        float q[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          float v = flag ? fminf(fmaxf(x[d], -1.f), 1.f) : x[d];
          v = (v + 1.f) / 2.f;
          q[d] = v * E.norm_range[d] + E.norm_min[d];
        }
        // table half of the sequence number this publication will get (read before anybody can advance it: the advance
        // happens after every group of this launch has arrived at the counter below)
        const uint32_t half = (pb.world > 1) ? ((pb.state[0] + 1u) & 1u) : 0u;
        const size_t row = ((size_t)half * pb.n_rows + pb.row_offset + g) * H + h;
        for (int r = 0; r < pb.world; ++r) reinterpret_cast<float2*>(pb.tables[r])[row] = make_float2(q[0], q[1]);
        if (pb.world > 1) __threadfence_system();
      }
      if (pb.world > 1) {
        __syncthreads();   // the representative's H threads have fenced their remote stores
        if (is_rep && h == 0) {
          const uint32_t prev = atomicAdd(pb.state + 2, 1u);
          if (prev == (uint32_t)a.grp.n_groups - 1u) {   // last group of this rank: release the sequence number everywhere
            pb.state[2] = 0u;
            const uint32_t seq = pb.state[0] + 1u;
            pb.state[0] = seq;
            __threadfence_system();
            for (int r = 0; r < pb.world; ++r) st_release_sys(pb.flags[r] + pb.rank, seq);
          }
        }
      }
    }
  }
  if (a.cpg > 1) cg::this_cluster().sync();  // nobody exits while a peer may still write its flags
}

static int launch_step(const StepArgs& args, bool taps, cudaStream_t stream) {
  const int threads = args.H * args.spc;
  size_t smem = 2 * sizeof(float4) * (size_t)args.H * args.spc;
  if (args.peers_in_smem) smem += sizeof(float2) * (size_t)args.grp.n_peers * args.H;
  if (args.hash_in_smem) smem += sizeof(float4) * (size_t)args.grp.n_peers * args.H;
  if (args.cs_in_smem) smem += sizeof(unsigned short) * (size_t)args.H * (args.grp.peer_grid * args.grp.peer_grid + 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(args.grp.n_groups * args.cpg));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (args.cpg > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)args.cpg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    na = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // Function attributes are per device: configure once per (device) and always opt in to the device maximum, so
  // that static + dynamic shared memory can never straddle the 48 KiB default (round-1 N=2 crash: exactly 48 KiB
  // dynamic + 128 B static) and a second device in the same process gets its own opt-in.
  {
    int dev = 0;
    MMDK_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64) return fail(MMDK_EINVAL, "device index out of range");
    if (!configured[dev]) {
      int max_optin = 0;
      MMDK_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      cudaFuncAttributes fa{};
      MMDK_CUDA(cudaFuncGetAttributes(&fa, ddpm_step_kernel<false>));
      const int dyn_max = max_optin - (int)fa.sharedSizeBytes;
      MMDK_CUDA(cudaFuncSetAttribute(ddpm_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
      MMDK_CUDA(cudaFuncGetAttributes(&fa, ddpm_step_kernel<true>));
      MMDK_CUDA(cudaFuncSetAttribute(ddpm_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     max_optin - (int)fa.sharedSizeBytes));
      MMDK_CUDA(cudaFuncSetAttribute(ddpm_step_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      MMDK_CUDA(cudaFuncSetAttribute(ddpm_step_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      configured[dev] = true;
    }
  }
  cudaError_t e = taps ? cudaLaunchKernelEx(&cfg, ddpm_step_kernel<true>, args)
                       : cudaLaunchKernelEx(&cfg, ddpm_step_kernel<false>, args);
  return check_cuda(e, "ddpm_step_kernel launch");
}

static int plan_step(StepArgs& a) {
  const int H = a.H, K = a.grp.K;
  if (H < 2 || H > 256 || (H & 1)) return fail(MMDK_EINVAL, "horizon must be even and in [2, 256]");
  if (H > 512) return fail(MMDK_EINVAL, "horizon too long");
  if (K < 1 || a.grp.n_groups < 1) return fail(MMDK_EINVAL, "n_groups and K must be >= 1");
  if (a.sc.n_guide_steps > kMaxIters) return fail(MMDK_EINVAL, "n_guide_steps > 32 is not supported by the step kernel");
  if (!a.grp.hard_rows_dev || !a.grp.hard_vals_dev) return fail(MMDK_EINVAL, "hard condition arrays are required");
  int max_spc = 512 / H;
  int spc = K < max_spc ? K : max_spc;
  int cpg = (K + spc - 1) / spc;
  if (cpg > 16) return fail(MMDK_EINVAL, "K too large: a group must fit one thread-block cluster (K <= 16 * floor(512/H))");
  // balance the samples over the cluster
  spc = (K + cpg - 1) / cpg;
  a.spc = spc;
  a.cpg = cpg;
  a.peers_in_smem = (a.grp.peers_dev != nullptr && a.grp.peer_cell_start_dev == nullptr &&
                     2 * sizeof(float4) * (size_t)H * spc + sizeof(float2) * (size_t)a.grp.n_peers * H <= 100 * 1024) ? 1 : 0;
  // two CTAs per SM: stage the hash only while both fit comfortably (<= ~100 KB each); the cell_start table alone (46 KB at
  // the default radius) still fits when the sorted positions (16 B x peers x H) no longer do -- then the 3x3-cell walk reads
  // its offsets from shared memory and only the few candidate positions from L2
  const size_t cs_bytes = sizeof(unsigned short) * (size_t)H * (a.grp.peer_grid * a.grp.peer_grid + 1);
  const size_t xu_bytes = 2 * sizeof(float4) * (size_t)H * spc;
  const bool hashed = a.grp.peers_dev != nullptr && a.grp.peer_cell_start_dev != nullptr;
  a.cs_in_smem = (hashed && xu_bytes + cs_bytes <= 100 * 1024) ? 1 : 0;
  a.hash_in_smem = (a.cs_in_smem && xu_bytes + cs_bytes + sizeof(float4) * (size_t)a.grp.n_peers * H <= 100 * 1024) ? 1 : 0;
  if (a.grp.peer_cell_start_dev && (a.grp.peer_grid < 1 || a.grp.peer_grid > MMDK_PEER_GRID_MAX || !a.grp.peer_sorted_dev))
    return fail(MMDK_EINVAL, "peer hash: bad grid size or missing sorted table");
  return MMDK_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void publish_peers_kernel(mmdk_guide_env E, int K, int H, int rep, const float* __restrict__ x,
                                     float* __restrict__ out) {
  // one CTA per group: global clip flag over the group's [K,H,D] block, then unnormalise the representative sample
  const int g = blockIdx.x;
  const float4* xg = reinterpret_cast<const float4*>(x + (size_t)g * K * H * 4);
  int local = 0;
  for (int i = threadIdx.x; i < K * H; i += blockDim.x) {
    const float4 v = xg[i];
    local |= (v.x > 1.0001f) | (v.x < -1.0001f) | (v.y > 1.0001f) | (v.y < -1.0001f) | (v.z > 1.0001f) | (v.z < -1.0001f) |
             (v.w > 1.0001f) | (v.w < -1.0001f);
  }
  int flag = __syncthreads_or(local);
  for (int i = threadIdx.x; i < H * 2; i += blockDim.x) {
    int h = i >> 1, d = i & 1;
    float v = x[((size_t)g * K * H + (size_t)rep * H + h) * 4 + d];
    if (flag) v = fminf(fmaxf(v, -1.f), 1.f);
    v = (v + 1.f) / 2.f;
    out[((size_t)g * H + h) * 2 + d] = v * E.norm_range[d] + E.norm_min[d];
  }
}

// Consumer side of the fused publication: waits until every rank's sequence number in THIS rank's flag array has reached
// the next expected value, then advances the consume sequence (state[1]).  Bounded spin: a dead peer sets state[3].
__global__ void wait_peers_kernel(PubArgs pb) {
  const int r = threadIdx.x;
  const uint32_t expect = pb.state[1] + 1u;
  if (r < pb.world) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(ld_acquire_sys(pb.flags[pb.rank] + r) - expect) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) { pb.state[3] = 1u; break; }   // 4 s
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (r == 0) pb.state[1] = expect;
}

// Lock-step peer hash: one CTA per waypoint h.  Stable counting sort of the [n_peers] positions of waypoint h into a
// G x G uniform grid (cell >= peer radius): cell_start [H][G*G+1] (uint16), sorted [H][n_peers] float4 (x, y, index, 0).
// Stable (peers of a cell stay in index order), so the query visits candidates in a deterministic order.
__global__ void build_peer_hash_kernel(const float* __restrict__ peers_base, const uint32_t* __restrict__ seq, int n_peers, int H,
                                       int G, float lo, float inv_cell, unsigned short* __restrict__ cell_start,
                                       float4* __restrict__ sorted) {
  const float* peers = peers_base + (seq ? (size_t)(*seq & 1u) * n_peers * H * 2 : 0);
  extern __shared__ int s_hash[];            // [G*G+1] counts / offsets, then [n_peers] cell of every peer
  int* s_cnt = s_hash;
  int* s_cell = s_hash + G * G + 1;
  const int h = blockIdx.x, nc = G * G;
  for (int i = threadIdx.x; i <= nc; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n_peers; j += blockDim.x) {
    const float2 q = reinterpret_cast<const float2*>(peers)[(size_t)j * H + h];
    int cx, cy;
    peer_cell(lo, inv_cell, G, q.x, q.y, cx, cy);
    const int c = cy * G + cx;
    s_cell[j] = c;
    atomicAdd(&s_cnt[c + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) for (int i = 0; i < nc; ++i) s_cnt[i + 1] += s_cnt[i];   // <= 1025 entries
  __syncthreads();
  for (int i = threadIdx.x; i <= nc; i += blockDim.x) cell_start[(size_t)h * (nc + 1) + i] = (unsigned short)s_cnt[i];
  for (int j = threadIdx.x; j < n_peers; j += blockDim.x) {
    const int c = s_cell[j];
    int rank = 0;
    for (int k = 0; k < j; ++k) rank += (s_cell[k] == c);
    const float2 q = reinterpret_cast<const float2*>(peers)[(size_t)j * H + h];
    sorted[(size_t)h * n_peers + s_cnt[c] + rank] = make_float4(q.x, q.y, __int_as_float(j), 0.f);
  }
}

__global__ void cross_condition_kernel(float* x1, float* x2, int B, int H, int ind1, int ind2, float4 rel, float4 bnd) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* p1 = x1 + ((size_t)b * H + ind1) * 4;
  float* p2 = x2 + ((size_t)b * H + ind2) * 4;
  const float r[4] = {rel.x, rel.y, rel.z, rel.w}, bd[4] = {bnd.x, bnd.y, bnd.z, bnd.w};
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    float v1 = fminf(p2[d] + r[d], bd[d]);       // sample_functions.py:29
    p1[d] = v1;
    p2[d] = fmaxf(v1 - r[d], -bd[d]);            // sample_functions.py:30
  }
}

__global__ void q_sample_kernel(const float* __restrict__ xs, const float* __restrict__ nz, float a, float b, int64_t n,
                                float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * xs[i] + b * nz[i];
}

__global__ void cell_index_kernel(mmdk_guide_env E, const float* __restrict__ pts, int64_t n, int32_t* __restrict__ idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int ix, iy;
    cell_index(E, pts[2 * i], pts[2 * i + 1], ix, iy);
    idx[2 * i] = ix;
    idx[2 * i + 1] = iy;
  }
}

__global__ void rr_collisions_kernel(const float* __restrict__ pos, int64_t n_batch, int R, float margin,
                                     uint8_t* __restrict__ coll, float* __restrict__ mid) {
  // robot_planar_disk.py:173-203
  const int64_t total = n_batch * R * R;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i % R);
    int r = (int)((i / R) % R);
    int64_t bt = i / ((int64_t)R * R);
    const float* p = pos + (bt * R + r) * 2;
    const float* q = pos + (bt * R + j) * 2;
    float dx = p[0] - q[0], dy = p[1] - q[1];
    float nrm = sqrtf(dx * dx + dy * dy);
    bool c = (nrm < margin) && (r != j);
    coll[i] = c ? 1 : 0;
    if (mid) {
      float nanv = __int_as_float(0x7fc00000);
      mid[2 * i] = c ? (p[0] + q[0]) / 2.f : nanv;
      mid[2 * i + 1] = c ? (p[1] + q[1]) / 2.f : nanv;
    }
  }
}

__global__ void classify_kernel(mmdk_guide_env E, const float* __restrict__ trajs, int B, int H, int n_interp,
                                float radius, float2 qmin, float2 qmax, uint8_t* __restrict__ free_out,
                                float* __restrict__ cost_out, uint8_t* __restrict__ wp_coll) {
  // tasks.py:236-311 + TR/trajectory/utils.py:73-86 + TR/trajectory/metrics.py:7-39; one warp per trajectory
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* tr = trajs + (size_t)warp * H * 4;
  int any_coll = 0, outside = 0;
  float plen = 0.f, smooth = 0.f;
  const int n_pts = (H - 1) * n_interp;
  for (int i = lane; i < n_pts; i += 32) {
    int seg = i / n_interp, k = i % n_interp;
    // alpha = linspace(0, 1, n_interp + 2)[1 + k]; point = x[seg] * alpha + x[seg + 1] * (1 - alpha)
    float step = 1.f / (float)(n_interp + 1);
    float alpha = (k + 1 < (n_interp + 2) / 2) ? (0.f + step * (float)(k + 1)) : (1.f - step * (float)(n_interp - k));
    float px = tr[seg * 4 + 0] * alpha + tr[(seg + 1) * 4 + 0] * (1.f - alpha);
    float py = tr[seg * 4 + 1] * alpha + tr[(seg + 1) * 4 + 1] * (1.f - alpha);
    int c = 0;
    if (E.grid_dev) {
      int ix, iy;
      cell_index(E, px, py, ix, iy);
      float s = __ldg(E.grid_dev + ((size_t)ix * E.ny + iy) * 4);
      c |= (s < radius);
    }
    c |= ((px - E.ws_min[0]) < radius) | ((py - E.ws_min[1]) < radius) | ((E.ws_max[0] - px) < radius) |
         ((E.ws_max[1] - py) < radius);
    if (wp_coll) wp_coll[(size_t)warp * n_pts + i] = (uint8_t)c;
    any_coll |= c;
  }
  for (int hh = lane; hh < H; hh += 32) {
    float px = tr[hh * 4], py = tr[hh * 4 + 1];
    outside |= !((px >= qmin.x) && (px <= qmax.x) && (py >= qmin.y) && (py <= qmax.y));
    if (hh < H - 1) {
      float dx = tr[(hh + 1) * 4] - px, dy = tr[(hh + 1) * 4 + 1] - py;
      plen += sqrtf(dx * dx + dy * dy);
      float dvx = tr[(hh + 1) * 4 + 2] - tr[hh * 4 + 2], dvy = tr[(hh + 1) * 4 + 3] - tr[hh * 4 + 3];
      smooth += sqrtf(dvx * dvx + dvy * dvy);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    any_coll |= __shfl_xor_sync(0xffffffffu, any_coll, o);
    outside |= __shfl_xor_sync(0xffffffffu, outside, o);
    plen += __shfl_xor_sync(0xffffffffu, plen, o);
    smooth += __shfl_xor_sync(0xffffffffu, smooth, o);
  }
  if (lane == 0) {
    free_out[warp] = (!any_coll && !outside) ? 1 : 0;
    cost_out[warp] = plen + smooth;
  }
}

__global__ void unnormalize_kernel(mmdk_guide_env E, const float* __restrict__ x, int64_t n_rows, int clip,
                                   float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows * 4; i += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(i & 3);
    float v = x[i];
    if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
    v = (v + 1.f) / 2.f;
    out[i] = v * E.norm_range[d] + E.norm_min[d];
  }
}

}  // namespace mmdk

using namespace mmdk;

extern "C" {

int mmdk_guide_grad(const mmdk_guide_env* env, const mmdk_groups* groups, int H, const float* x_dev, float* grad_dev,
                    float* raw_dev, int n_costs_max, void* stream) {
  if (!env || !groups || !x_dev || !grad_dev) return fail(MMDK_EINVAL, "null argument");
  StepArgs a{};
  a.env = *env; a.grp = *groups; a.H = H;
  a.sc = mmdk_step_scalars{};
  a.x = const_cast<float*>(x_dev);
  a.grad_out = grad_dev; a.raw_out = raw_dev; a.n_costs_max = n_costs_max;
  int rc = plan_step(a);
  if (rc != MMDK_OK) return rc;
  return launch_step(a, true, (cudaStream_t)stream);
}

int mmdk_ddpm_step(const mmdk_guide_env* env, const mmdk_groups* groups, const mmdk_step_scalars* sc, int H,
                   float* x_dev, const float* eps_dev, const float* noise_dev, float* chain_dev, void* stream) {
  if (!env || !groups || !sc || !x_dev) return fail(MMDK_EINVAL, "null argument");
  if (sc->do_posterior && !eps_dev) return fail(MMDK_EINVAL, "eps_dev required when do_posterior is set");
  StepArgs a{};
  a.env = *env; a.grp = *groups; a.sc = *sc; a.H = H;
  a.x = x_dev; a.eps = eps_dev; a.noise = noise_dev; a.chain = chain_dev;
  int rc = plan_step(a);
  if (rc != MMDK_OK) return rc;
  return launch_step(a, false, (cudaStream_t)stream);
}

static int fill_pub(StepArgs& a, const mmdk_peer_exchange* ex) {
  if (!ex) { a.pub_on = 0; return MMDK_OK; }
  if (ex->world < 1 || ex->world > MMDK_MAX_RANKS || ex->rank < 0 || ex->rank >= ex->world)
    return fail(MMDK_EINVAL, "peer exchange: bad world / rank");
  if (!ex->state_dev) return fail(MMDK_EINVAL, "peer exchange: state_dev is required");
  if (ex->rep_index < 0 || ex->rep_index >= a.grp.K) return fail(MMDK_EINVAL, "peer exchange: rep_index out of range");
  if (ex->row_offset < 0 || ex->row_offset + a.grp.n_groups > ex->n_rows) return fail(MMDK_EINVAL, "peer exchange: rows out of range");
  a.pub_on = 1;
  a.pub_rep = ex->rep_index;
  a.pub.world = ex->world; a.pub.rank = ex->rank; a.pub.row_offset = ex->row_offset; a.pub.n_rows = ex->n_rows;
  a.pub.state = ex->state_dev;
  for (int r = 0; r < ex->world; ++r) {
    if (!ex->tables_dev[r] || (ex->world > 1 && !ex->flags_dev[r])) return fail(MMDK_EINVAL, "peer exchange: null table / flag pointer");
    a.pub.tables[r] = ex->tables_dev[r];
    a.pub.flags[r] = ex->flags_dev[r];
  }
  return MMDK_OK;
}

int mmdk_ddpm_step_publish(const mmdk_guide_env* env, const mmdk_groups* groups, const mmdk_step_scalars* sc, int H,
                           float* x_dev, const float* eps_dev, const float* noise_dev, float* chain_dev,
                           const mmdk_peer_exchange* exchange, void* stream) {
  if (!env || !groups || !sc || !x_dev) return fail(MMDK_EINVAL, "null argument");
  if (sc->do_posterior && !eps_dev) return fail(MMDK_EINVAL, "eps_dev required when do_posterior is set");
  StepArgs a{};
  a.env = *env; a.grp = *groups; a.sc = *sc; a.H = H;
  a.x = x_dev; a.eps = eps_dev; a.noise = noise_dev; a.chain = chain_dev;
  int rc = plan_step(a);
  if (rc != MMDK_OK) return rc;
  rc = fill_pub(a, exchange);
  if (rc != MMDK_OK) return rc;
  return launch_step(a, false, (cudaStream_t)stream);
}

int mmdk_wait_peers(const mmdk_peer_exchange* ex, void* stream) {
  if (!ex || !ex->state_dev) return fail(MMDK_EINVAL, "null argument");
  if (ex->world <= 1) return MMDK_OK;
  if (ex->world > MMDK_MAX_RANKS || !ex->flags_dev[ex->rank]) return fail(MMDK_EINVAL, "peer exchange: bad world / flags");
  PubArgs pb{};
  pb.world = ex->world; pb.rank = ex->rank; pb.state = ex->state_dev;
  for (int r = 0; r < ex->world; ++r) pb.flags[r] = ex->flags_dev[r];
  wait_peers_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pb);
  return check_cuda(cudaGetLastError(), "wait_peers_kernel");
}

// ---- peer-memory plumbing for the exchange: plain cudaMalloc allocations shared through CUDA IPC -----------------------------
int mmdk_p2p_alloc(size_t bytes, void** out_dev) {
  if (!out_dev || bytes == 0) return fail(MMDK_EINVAL, "null argument");
  MMDK_CUDA(cudaMalloc(out_dev, bytes));
  MMDK_CUDA(cudaMemset(*out_dev, 0, bytes));
  return MMDK_OK;
}
int mmdk_p2p_free(void* dev) { return check_cuda(cudaFree(dev), "cudaFree"); }
int mmdk_p2p_export(void* dev, unsigned char handle[64]) {
  if (!dev || !handle) return fail(MMDK_EINVAL, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  MMDK_CUDA(cudaIpcGetMemHandle(&h, dev));
  memcpy(handle, &h, 64);
  return MMDK_OK;
}
int mmdk_p2p_open(const unsigned char handle[64], void** out_dev) {
  if (!handle || !out_dev) return fail(MMDK_EINVAL, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  MMDK_CUDA(cudaIpcOpenMemHandle(out_dev, h, cudaIpcMemLazyEnablePeerAccess));
  return MMDK_OK;
}
int mmdk_p2p_close(void* dev) { return check_cuda(cudaIpcCloseMemHandle(dev), "cudaIpcCloseMemHandle"); }

int mmdk_publish_peers(const mmdk_guide_env* env, int n_groups, int K, int H, int rep_index, const float* x_dev,
                       float* peers_out_dev, void* stream) {
  if (!env || !x_dev || !peers_out_dev) return fail(MMDK_EINVAL, "null argument");
  if (rep_index < 0 || rep_index >= K) return fail(MMDK_EINVAL, "rep_index out of range");
  publish_peers_kernel<<<n_groups, 1024, 0, (cudaStream_t)stream>>>(*env, K, H, rep_index, x_dev, peers_out_dev);
  return check_cuda(cudaGetLastError(), "publish_peers_kernel");
}

int mmdk_build_peer_hash(const float* peers_dev, const uint32_t* seq_dev, int n_peers, int H, int grid, float grid_lo,
                         float grid_inv_cell, uint16_t* cell_start_dev, float* sorted_dev, void* stream) {
  if (!peers_dev || !cell_start_dev || !sorted_dev) return fail(MMDK_EINVAL, "null argument");
  if (grid < 1 || grid > MMDK_PEER_GRID_MAX) return fail(MMDK_EINVAL, "peer hash grid out of range");
  if (n_peers < 1 || n_peers > 65535 || H < 1) return fail(MMDK_EINVAL, "peer hash: n_peers must be in [1, 65535]");
  const size_t smem = sizeof(int) * ((size_t)grid * grid + 1 + n_peers);
  if (smem > 48 * 1024) return fail(MMDK_EINVAL, "peer hash: fleet too large for the single-CTA-per-waypoint builder");
  build_peer_hash_kernel<<<H, 256, smem, (cudaStream_t)stream>>>(peers_dev, seq_dev, n_peers, H, grid, grid_lo, grid_inv_cell,
                                                              cell_start_dev, reinterpret_cast<float4*>(sorted_dev));
  return check_cuda(cudaGetLastError(), "build_peer_hash_kernel");
}

int mmdk_cross_condition(float* x1_dev, float* x2_dev, int B, int H, int ind1, int ind2, const float rel[4],
                         const float bnd[4], void* stream) {
  if (!x1_dev || !x2_dev) return fail(MMDK_EINVAL, "null argument");
  if (ind1 < 0) ind1 += H;
  if (ind2 < 0) ind2 += H;
  if (ind1 < 0 || ind1 >= H || ind2 < 0 || ind2 >= H) return fail(MMDK_EINVAL, "index out of range");
  cross_condition_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      x1_dev, x2_dev, B, H, ind1, ind2, make_float4(rel[0], rel[1], rel[2], rel[3]),
      make_float4(bnd[0], bnd[1], bnd[2], bnd[3]));
  return check_cuda(cudaGetLastError(), "cross_condition_kernel");
}

int mmdk_q_sample(const float* x_start_dev, const float* noise_dev, float a, float b, int64_t n, float* out_dev,
                  void* stream) {
  if (!x_start_dev || !noise_dev || !out_dev) return fail(MMDK_EINVAL, "null argument");
  if (n <= 0) return MMDK_OK;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  q_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_start_dev, noise_dev, a, b, n, out_dev);
  return check_cuda(cudaGetLastError(), "q_sample_kernel");
}

int mmdk_cell_index(const mmdk_guide_env* env, const float* points_dev, int64_t n, int32_t* idx_dev, void* stream) {
  if (!env || !points_dev || !idx_dev) return fail(MMDK_EINVAL, "null argument");
  if (n <= 0) return MMDK_OK;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  cell_index_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*env, points_dev, n, idx_dev);
  return check_cuda(cudaGetLastError(), "cell_index_kernel");
}

int mmdk_check_rr_collisions(const float* pos_dev, int64_t n_batch, int R, float margin, uint8_t* coll_dev,
                             float* mid_dev, void* stream) {
  if (!pos_dev || !coll_dev) return fail(MMDK_EINVAL, "null argument");
  if (n_batch <= 0 || R <= 0) return MMDK_OK;
  int64_t total = n_batch * R * R;
  int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  rr_collisions_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos_dev, n_batch, R, margin, coll_dev, mid_dev);
  return check_cuda(cudaGetLastError(), "rr_collisions_kernel");
}

int mmdk_classify_trajs(const mmdk_guide_env* env, const float* trajs_dev, int B, int H, int n_interp, float radius,
                        const float q_min[2], const float q_max[2], uint8_t* free_dev, float* cost_dev,
                        uint8_t* waypoint_coll_dev, void* stream) {
  if (!env || !trajs_dev || !free_dev || !cost_dev) return fail(MMDK_EINVAL, "null argument");
  if (B <= 0) return MMDK_OK;
  if (n_interp < 1) return fail(MMDK_EINVAL, "n_interp must be >= 1");
  int threads = 256, wpb = threads / 32;
  classify_kernel<<<(B + wpb - 1) / wpb, threads, 0, (cudaStream_t)stream>>>(
      *env, trajs_dev, B, H, n_interp, radius, make_float2(q_min[0], q_min[1]), make_float2(q_max[0], q_max[1]),
      free_dev, cost_dev, waypoint_coll_dev);
  return check_cuda(cudaGetLastError(), "classify_kernel");
}

int mmdk_unnormalize(const mmdk_guide_env* env, const float* x_dev, int64_t n_rows, int clip, float* out_dev,
                     void* stream) {
  if (!env || !x_dev || !out_dev) return fail(MMDK_EINVAL, "null argument");
  if (n_rows <= 0) return MMDK_OK;
  int64_t n = n_rows * 4;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  unnormalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*env, x_dev, n_rows, clip, out_dev);
  return check_cuda(cudaGetLastError(), "unnormalize_kernel");
}

}  // extern "C"
