/*
 * mmdk.h -- C ABI of libmmdk.so: the B200 (sm_100a) guided-diffusion trajectory sampler.
 *
 * The reference (yoraish/mmd) has NO FFI/plugin ABI: its hot path is a Python object protocol
 * (SURVEY.md 8b).  This header is the boundary a maintainer would bind (ctypes stub in INTEGRATION.md)
 * behind the reference's own classes.  Every entry point names the reference code it replaces; paths are
 * relative to the reference root, TR = deps/torch_robotics/torch_robotics, MPB =
 * deps/motion_planning_baselines/mp_baselines.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a DEVICE pointer owned by the caller
 *     (e.g. torch tensors via data_ptr()); nothing is allocated inside the per-step calls.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); no call synchronises the host.
 *   - return value: 0 = MMDK_OK, otherwise an MMDK_E* code; mmdk_last_error() gives the message
 *     (the Python host maps codes to exceptions: ValueError for MMDK_EINVAL, RuntimeError otherwise).
 *   - trajectories are fp32 [B, H, D] row-major, D = 4 (x, y, vx, vy), H = horizon (64), normalised to [-1, 1]
 *     by a LimitsNormalizer (mmd/datasets/normalization.py:150-168).  B = n_groups * K: a "group" is one
 *     (robot, tile) planner call of K samples sharing hard conditions, constraints and the normaliser's
 *     global clip decision.
 */
#ifndef MMDK_H_
#define MMDK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMDK_OK 0
#define MMDK_EINVAL 1   /* bad argument / unsupported shape */
#define MMDK_ECUDA 2    /* CUDA runtime error */
#define MMDK_ENOMEM 3

#define MMDK_MAX_LEVELS 4
#define MMDK_MAX_HARD_ROWS 4
#define MMDK_STATE_DIM 4
#define MMDK_PEER_GRID_MAX 32
#define MMDK_MAX_RANKS 8

/* Thread-local message of the last failing call. */
const char* mmdk_last_error(void);
/* Library / device probe: fills sm count, returns MMDK_ECUDA when no sm_100 device is visible. */
int mmdk_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * TemporalUnet  (mmd/models/diffusion_models/temporal_unet.py:23-174, mmd/models/layers/layers.py:197-398)
 * ------------------------------------------------------------------------------------------------ */
typedef struct mmdk_unet mmdk_unet;

typedef struct {
  int state_dim;       /* 4 */
  int horizon;         /* n_support_points, 64 */
  int unet_input_dim;  /* 32 */
  int n_levels;        /* len(dim_mults) */
  int dim_mults[MMDK_MAX_LEVELS];
  int time_emb_dim;    /* 32 */
  int n_diffusion_steps; /* T: the time-embedding table is precomputed for t in [0, T) */
  int self_attention;  /* LinearAttention blocks present (layers.py:210-229) */
} mmdk_unet_config;

/* Builds the device-resident network from the reference state_dict: `names[i]` is the reference key
 * (e.g. "downs.0.0.blocks.0.block.0.weight"), `tensors_dev[i]` a contiguous fp32 device tensor in the
 * reference layout, `numels[i]` its element count.  Replaces TemporalUnet.__init__ + load_state_dict
 * (mpd.py:154-171).  Weights are re-packed on the device; the time/cond embedding table (TimeEncoder +
 * every ResidualTemporalBlock.cond_mlp, layers.py:232-258,337-341) is evaluated once for all t. */
int mmdk_unet_create(const mmdk_unet_config* cfg, int n_tensors, const char* const* names,
                     const float* const* tensors_dev, const int64_t* numels, void* stream, mmdk_unet** out);
void mmdk_unet_destroy(mmdk_unet* net);

/* Numerical mode of the conv contractions. */
#define MMDK_UNET_FP32 0     /* CUDA-core FFMA, fp32 throughout (exact-parity mode) */
#define MMDK_UNET_F16X3 1    /* tcgen05 kind::f16, operands split hi+lo in FP16 (3 MMAs), fp32 accumulate in TMEM:
                                ONE persistent launch per forward (unet_fused.cu) */
#define MMDK_UNET_F16X3_LAYERS 2 /* same arithmetic, one launch per layer (unet_tc.cu; round-1 executor, A/B baseline) */

/* eps = TemporalUnet.forward(x, t, context=None) for an integer timestep t shared by the batch
 * (temporal_unet.py:121-174; make_timesteps, diffusion_model_base.py:27-29).  x_dev, eps_dev: [B, H, D]. */
/* A handle owns its executor state (activation workspace per batch size, the per-forward patch of the layer program): use
 * a handle from ONE host thread and ONE stream at a time; concurrent streams need one handle each (the weights are small). */
int mmdk_unet_forward(const mmdk_unet* net, int mode, const float* x_dev, int B, int t, float* eps_dev, void* stream);

/* Debug/parity tap: copies the precomputed cond table row of timestep t ([n_cond] floats) to out_dev. */
int mmdk_unet_cond_row(const mmdk_unet* net, int t, float* out_dev, int* n_cond, void* stream);

/* Debug/parity tap of the tensor-core executor: activation written by op `op_index` of the layer program (-1 = the
 * packed network input) during the LAST MMDK_UNET_F16X3 forward, as fp32 [B, C, L].  C/L are returned through
 * c_out/l_out; pass out_dev = NULL to query them.  n_ops_out (optional) receives the number of ops. */
int mmdk_unet_debug_tap(const mmdk_unet* net, int op_index, float* out_dev, int* c_out, int* l_out, int* n_ops_out,
                        void* stream);

/* Debug: the persistent executor (MMDK_UNET_F16X3) hands most activations from one op to the next inside shared memory and
 * never writes them to global memory; on != 0 makes every later forward also store every activation image so that
 * mmdk_unet_debug_tap can read it (same arithmetic, same eps).  Off by default. */
int mmdk_unet_debug_keep_activations(const mmdk_unet* net, int on);

/* Debug: from the next MMDK_UNET_F16X3 forward on, op `op_index` writes a per-CTA clock64 timeline into
 * dbg_dev [n_tiles, 16] int64 (slot 15 = SM id); op_index = -1 / dbg_dev = NULL switches it off (per-layer executor).
 * op_index = -2: the persistent executor (MMDK_UNET_F16X3) writes CTA 0's per-item stamps [n_items, 16] int64 into dbg_dev
 * (0 load issued, 1 input in shared memory, 2 MMAs issued, 3 accumulators complete, 5 output stored, 6 cycles the
 * issuer waited for weights, 7 epilogue idle since); dbg_dev = NULL switches it off. */
int mmdk_unet_debug_timeline(const mmdk_unet* net, int op_index, long long* dbg_dev, void* stream);

/* Launch-duration probe of the persistent executor (bench.py's roofline: works inside replayed CUDA graphs, where host
 * events cannot be placed): from the next MMDK_UNET_F16X3 forward on, every CTA writes its first and last %globaltimer
 * reading (ns) into stamps_dev [n_slots, max_ctas, 2] uint64; forward number i (counted from this call, or baked at capture
 * time) uses slot i mod n_slots.  Launch duration = max(end) - min(start) over the CTAs of a slot.  NULL switches it off. */
int mmdk_unet_debug_stamps(const mmdk_unet* net, unsigned long long* stamps_dev, int n_slots, int max_ctas);

/* Calibration (profiles/): n_ctas CTAs each issue n_iters back-to-back tcgen05.mma (M=128, N, K=16, kind::f16, SS mode)
 * rotating over n_acc (1, 2 or 4) TMEM accumulators, and write {elapsed clock64 cycles, n_iters} to out_dev
 * [n_ctas, 2] int64.  Run under ncu to read what sm__pipe_tensor_cycles_active reports for a loop that is 100% MMA. */
int mmdk_debug_mma_calibrate(int N, int n_iters, int n_ctas, int n_acc, long long* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Guide + DDPM step  (mmd/models/diffusion_models/guides.py:152-253, sample_functions.py:41-107,
 * diffusion_model_base.py:126-160, MPB/planners/costs/cost_functions.py, TR/environments/grid_map_sdf.py)
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  /* LimitsNormalizer: mins and (maxs - mins), computed in fp32 by the caller (normalization.py:157-168) */
  float norm_min[MMDK_STATE_DIM];
  float norm_range[MMDK_STATE_DIM];
  /* GridMapSDF (grid_map_sdf.py:84-114): packed (sdf, dsdf/dx, dsdf/dy, 0) per node, [nx, ny] row-major; NULL = no
   * object field */
  const float* grid_dev;
  int nx, ny;
  float grid_lo[2];       /* limits[0] */
  float grid_map_dim[2];  /* |limits[1] - limits[0]| */
  /* hinge margin = link margin (1.1 r) + cutoff margin (robot_planar_disk.py:68, tasks.py:53-58) */
  float margin;
  /* workspace boundaries, already scaled by 1.08 (tasks.py:82-84) */
  float ws_min[2], ws_max[2];
  /* gradient weights after clipping (mmd_params.py:40-43) and clip norm (guides.py:154) */
  float w_collision, w_border, w_smooth;
  float max_grad_norm;
  /* GP prior (MPB/planners/costs/factors/gp_factor.py:34-50): dt (Phi) and the three distinct entries of Q^-1,
   * (12 dt^-3, -6 dt^-2, 4 dt^-1) / sigma_gp^2, evaluated by the caller in double and rounded to fp32 as torch does */
  float dt, gp_q11, gp_q12, gp_q22;
  /* 1 / sigma_coll^2 (field_factor.py:22) */
  float coll_inv_sigma2;
} mmdk_guide_env;

/* Per-group (robot/tile) inputs.  Extra costs = CostConstraint objects (cost_functions.py:275-326), each
 * differentiated, clipped and weighted separately (mpd.py:331-342,409-412).  Vertex constraints are bucketed by
 * waypoint: object o of group g covers entries [bucket_ptr[o*(H+1)+h], bucket_ptr[o*(H+1)+h+1]) of `cons_dev`
 * for waypoint h; an entry is (qx, qy, radius, 0). */
typedef struct {
  int n_groups;
  int K;                        /* samples per group */
  const float* hard_vals_dev;   /* [n_groups, MMDK_MAX_HARD_ROWS, D] normalised states */
  const int* hard_rows_dev;     /* [n_groups, MMDK_MAX_HARD_ROWS] waypoint index or -1 (sample_functions.py:8-14) */
  const int* obj_ptr_dev;       /* [n_groups + 1] range of cost objects per group, or NULL = no extra costs */
  const float* obj_weight_dev;  /* [n_obj] */
  const int* bucket_ptr_dev;    /* [n_obj, H + 1] */
  const float* cons_dev;        /* [n_entries, 4] */
  /* lock-step peer exchange: unnormalised positions [n_peers, H, 2] of every robot's representative
   * sample; group g skips row peer_self_dev[g].  One implicit soft CostConstraint per group with ranges (h, h+1).
   * NULL = off. */
  const float* peers_dev;
  const int* peer_self_dev;     /* [n_groups] */
  int n_peers;
  float peer_radius, peer_weight;
  /* Optional spatial hash of the peer table (mmdk_build_peer_hash): per waypoint a grid x grid uniform grid with
   * cell >= peer_radius over [peer_grid_lo, peer_grid_lo + grid / peer_grid_inv_cell)^2 (clamped at the border).
   * cell_start [H, grid*grid + 1] uint16, sorted [H, n_peers, 4] fp32 (x, y, peer index bits, 0).  With it the step
   * kernel tests only the 3x3 cells around a waypoint instead of all n_peers rows (same exact in-radius test, so the
   * same peers contribute).  NULL = brute-force scan of peers_dev. */
  const uint16_t* peer_cell_start_dev;
  const float* peer_sorted_dev;
  int peer_grid;
  float peer_grid_lo, peer_grid_inv_cell;
  /* Double-buffered peer table (mmdk_peer_exchange): when not NULL, peers_dev is [2][n_peers, H, 2] and readers use half
   * (*peer_seq_dev & 1); NULL = peers_dev is a single [n_peers, H, 2] table. */
  const uint32_t* peer_seq_dev;
} mmdk_groups;

/* One evaluation of GuideManagerTrajectoriesWithVelocity.forward (guides.py:180-226): grad_dev [B,H,D] =
 * -(sum_k w_k clip(dcost_k/dx_unnormalised)).  Optional taps for parity tests: per-cost raw gradients
 * raw_dev [n_costs_max, B, H, D] (order: objects, border, GP, extra objects..., peers), NULL to skip. */
int mmdk_guide_grad(const mmdk_guide_env* env, const mmdk_groups* groups, int H, const float* x_dev,
                    float* grad_dev, float* raw_dev, int n_costs_max, void* stream);

/* Scalars of one reverse step, gathered on the host from the schedule buffers (`extract`, sample_functions.py:34-37). */
typedef struct {
  int do_posterior;      /* 1: x <- q_posterior mean of (x, eps) (diffusion_model_base.py:126-160) */
  float sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod;
  float posterior_mean_coef1, posterior_mean_coef2;
  int clip_denoised;     /* clamp x0 to [-1, 1] (:155) */
  int predict_epsilon;
  int n_guide_steps;     /* 0 = unguided (t >= t_start_guide) */
  int add_noise;         /* 0 when t == 0 (noise[t == 0] = 0, sample_functions.py:76) or noise_dev is NULL */
  float model_std;       /* exp(0.5 * posterior_log_variance_clipped[t]) */
  float noise_std;       /* noise_std_extra_schedule_fn(t), 0.5 in MPD (mpd.py:303) */
  int final_hard_conds;  /* 1: finish with apply_hard_conditioning as p_sample_loop does (diffusion_model_base.py:202) */
} mmdk_step_scalars;

/* ddpm_sample_fn minus the UNet (sample_functions.py:41-86) + the trailing apply_hard_conditioning of
 * p_sample_loop (diffusion_model_base.py:202): posterior mean -> n_guide_steps x (x += guide(x); hard conds)
 * -> x += noise_scale * noise -> hard conds.  x_dev is updated in place; if chain_dev != NULL the new x is also
 * written there (chain slot of this step).  eps_dev/noise_dev may be NULL when unused. */
int mmdk_ddpm_step(const mmdk_guide_env* env, const mmdk_groups* groups, const mmdk_step_scalars* sc, int H,
                   float* x_dev, const float* eps_dev, const float* noise_dev, float* chain_dev, void* stream);

/* Lock-step exchange over peer memory -- the all-gather of the representative paths fused into the step kernel:
 * at the END of a reverse step every group's representative sample (the x the next step starts from, unnormalised like
 * mmdk_publish_peers) is stored straight into the peer table of EVERY rank of the fleet (NVLink / NVSwitch stores through
 * CUDA-IPC mappings), and the last group of a rank releases the rank's sequence number into every rank's flag array;
 * mmdk_wait_peers is the matching acquire.  Tables are double buffered by sequence parity: [2][n_rows][H][2] fp32.
 *   tables_dev[r], flags_dev[r]: rank r's table / flag array ([MMDK_MAX_RANKS] uint32) as mapped on THIS device (own
 *   allocation for r == rank, mmdk_p2p_open of rank r's exported handle otherwise; flags unused when world == 1).
 *   state_dev: local uint32[4] = {publish sequence, consume sequence, group counter, error flag (a peer never arrived)}.
 * Readers (step kernel brute-force path, mmdk_build_peer_hash) select the half with mmdk_groups.peer_seq_dev =
 * state_dev + 1 (world > 1: advanced by mmdk_wait_peers) or state_dev + 0 (world == 1: stream order is enough). */
typedef struct {
  int world, rank;
  int row_offset;   /* table row of this call's first group */
  int n_rows;       /* rows of one table half = robots of the whole fleet */
  int rep_index;    /* representative sample of every group */
  float* tables_dev[MMDK_MAX_RANKS];
  uint32_t* flags_dev[MMDK_MAX_RANKS];
  uint32_t* state_dev;
} mmdk_peer_exchange;

/* mmdk_ddpm_step followed by the publication (exchange == NULL: plain mmdk_ddpm_step).  With scalars that leave x unchanged
 * (do_posterior = n_guide_steps = add_noise = final_hard_conds = 0) it is a pure publication of the current x. */
int mmdk_ddpm_step_publish(const mmdk_guide_env* env, const mmdk_groups* groups, const mmdk_step_scalars* sc, int H,
                           float* x_dev, const float* eps_dev, const float* noise_dev, float* chain_dev,
                           const mmdk_peer_exchange* exchange, void* stream);
/* Blocks the STREAM (not the host) until every rank has published its next sequence number to this rank (world > 1). */
int mmdk_wait_peers(const mmdk_peer_exchange* exchange, void* stream);

/* Peer-memory plumbing: zero-initialised cudaMalloc allocations and their CUDA-IPC handles (64 bytes). */
int mmdk_p2p_alloc(size_t bytes, void** out_dev);
int mmdk_p2p_free(void* dev);
int mmdk_p2p_export(void* dev, unsigned char handle[64]);
int mmdk_p2p_open(const unsigned char handle[64], void** out_dev);
int mmdk_p2p_close(void* dev);

/* The whole reverse loop of p_sample_loop (diffusion_model_base.py:163-211) in ONE call: for every step i
 *   [lockstep, guided steps only: mmdk_publish_peers -> mmdk_build_peer_hash]  ->  mmdk_unet_forward(t_index[i])  ->
 *   mmdk_ddpm_step(scalars[i], noise frame i, chain frame i)
 * issued natively on `stream`; with use_graph != 0 the sequence is captured once into a CUDA graph (keyed on every pointer
 * and scalar it bakes in, plus process-unique ids of the UNet handle and of its executor state for this batch size: a
 * handle re-created after a weight reload, or a state evicted from the bounded per-batch-size cache and rebuilt, never
 * matches an old graph) and replayed, so a chain costs one graph launch.  noise_dev [n_steps, B, H, D] step-major (NULL:
 * no noise added), chain_out_dev [n_steps, B, H, D] (NULL: keep only x_dev), eps_dev [B, H, D] scratch.  Lock-step with
 * the fleet sharded over several processes needs an exchange between publication and hash build: that path stays in the
 * host loop (mmd_b200/sampler.py); peers_local_dev = the rows of groups->peers_dev that belong to this call's groups. */
typedef struct {
  int n_steps;
  const int* t_index;                 /* [n_steps] timestep fed to the UNet (max(i, 0), sample_functions.py:51-54) */
  const mmdk_step_scalars* scalars;   /* [n_steps] */
  int lockstep;
  int rep_index;
  float* peers_local_dev;
  /* not NULL: lock-step publication fused into the step kernels (mmdk_ddpm_step_publish / mmdk_wait_peers) -- also across
   * processes, so a sharded fleet's chain is one captured graph too; peers_local_dev is ignored */
  const mmdk_peer_exchange* exchange;
} mmdk_chain_desc;

int mmdk_run_chain(const mmdk_unet* net, int unet_mode, const mmdk_guide_env* env, const mmdk_groups* groups,
                   const mmdk_chain_desc* chain, int H, float* x_dev, float* eps_dev, const float* noise_dev,
                   float* chain_out_dev, int use_graph, void* stream);

/* The multi-tile reverse loop of DiffusionsEnsemble.p_sample_loop (mmd/models/diffusion_models/diffusion_ensemble.py:78-106)
 * in ONE call: for every step i, for every tile m in order
 *   mmdk_unet_forward(tile m, t_index[i]) -> mmdk_ddpm_step(tile m, scalars[i], noise frame i; the step applies the tile's
 *   hard conditions itself) -> every cross condition (apply_cross_conditioning, sample_functions.py:17-31)
 * and, after the last tile of the step, the chain frame i of every tile (the state AFTER all stitches, :103-105).  The reference
 * records x[m] itself (no clone) and stitches in place, so the stitch after tile m's step also rewrites the stitched waypoint of the
 * previously recorded frame of every tile not stepped yet in that reverse step; reproduced (chain_init_dev = the frame before
 * step 0).
 * A tile's batch is n_groups planner calls of K samples (mmdk_groups), so R multi-tile planner calls run as one chain; a cross
 * condition applies to the batch rows [row_lo, row_hi) (planner calls that share the tile transforms).  use_graph != 0:
 * captured once into a CUDA graph keyed on every pointer / scalar, replayed afterwards. */
#define MMDK_MAX_TILES 8
typedef struct {
  const mmdk_unet* net;
  int unet_mode;
  const mmdk_guide_env* env;
  const mmdk_groups* groups;
  const mmdk_step_scalars* scalars;   /* [n_steps] (tile models may carry different schedules) */
  float* x_dev;                       /* [B, H, D] in/out */
  float* eps_dev;                     /* [B, H, D] scratch */
  const float* noise_dev;             /* [n_steps, B, H, D] or NULL */
  float* chain_out_dev;               /* [n_steps, B, H, D] or NULL */
  float* chain_init_dev;              /* [B, H, D] or NULL: the frame recorded BEFORE the first step (it aliases x in the reference, see below) */
} mmdk_ensemble_tile;

typedef struct {
  int m1, m2;        /* tiles: x[m1][:, ind1, :] = min(x[m2][:, ind2, :] + rel, bnd); x[m2][:, ind2, :] = max(x[m1][:, ind1, :] - rel, -bnd) */
  int ind1, ind2;
  int row_lo, row_hi;
  float rel[MMDK_STATE_DIM], bnd[MMDK_STATE_DIM];
} mmdk_cross_cond;

typedef struct {
  int n_tiles;
  const mmdk_ensemble_tile* tiles;
  int n_steps;
  const int* t_index;   /* [n_steps] */
  int n_cross;
  const mmdk_cross_cond* cross;
} mmdk_ensemble_desc;

int mmdk_run_chain_ensemble(const mmdk_ensemble_desc* ensemble, int H, int use_graph, void* stream);

/* Publishes the representative sample of every group for the lock-step exchange: peers_out_dev[g] [H,2] =
 * unnormalise(x[g*K + rep_index])[:, :2] with the normaliser's clip rule applied to that group. */
int mmdk_publish_peers(const mmdk_guide_env* env, int n_groups, int K, int H, int rep_index, const float* x_dev,
                       float* peers_out_dev, void* stream);

/* Builds the spatial hash of a lock-step peer table peers_dev [n_peers, H, 2] (see mmdk_groups): replaces the
 * reference's dense (n_constraints, B, H, 2) distance tensor (cost_functions.py:304-324) by an O(neighbours) lookup.
 * grid <= MMDK_PEER_GRID_MAX; the caller chooses grid / lo / inv_cell with 1 / inv_cell >= peer_radius.  seq_dev != NULL:
 * peers_dev is the double-buffered table of a mmdk_peer_exchange and half (*seq_dev & 1) is hashed. */
int mmdk_build_peer_hash(const float* peers_dev, const uint32_t* seq_dev, int n_peers, int H, int grid, float grid_lo,
                         float grid_inv_cell, uint16_t* cell_start_dev, float* sorted_dev, void* stream);

/* apply_cross_conditioning (sample_functions.py:17-31) for one (m1, ind1) <- (m2, ind2) stitch of an ensemble:
 * x1[:, ind1, :] = min(x2[:, ind2, :] + rel, bnd); x2[:, ind2, :] = max(x1[:, ind1, :] - rel, -bnd). */
int mmdk_cross_condition(float* x1_dev, float* x2_dev, int B, int H, int ind1, int ind2, const float rel[MMDK_STATE_DIM],
                         const float bnd[MMDK_STATE_DIM], void* stream);

/* q_sample (diffusion_model_base.py:425-433): out = a * x_start + b * noise. */
int mmdk_q_sample(const float* x_start_dev, const float* noise_dev, float a, float b, int64_t n, float* out_dev,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Integer outputs (bit-exact targets)
 * ------------------------------------------------------------------------------------------------ */
/* GridMapSDF cell indices (grid_map_sdf.py:93-97): idx_dev [n, 2] int32 for points_dev [n, 2]. */
int mmdk_cell_index(const mmdk_guide_env* env, const float* points_dev, int64_t n, int32_t* idx_dev, void* stream);

/* RobotPlanarDisk.check_rr_collisions (TR/robots/robot_planar_disk.py:173-203): pos_dev [n_batch, R, 2] ->
 * coll_dev [n_batch, R, R] uint8 (1 = ||pi - pj|| < margin, diagonal 0); mid_dev [n_batch, R, R, 2] midpoints
 * (NaN where no collision) or NULL. */
int mmdk_check_rr_collisions(const float* pos_dev, int64_t n_batch, int R, float margin, uint8_t* coll_dev,
                             float* mid_dev, void* stream);

/* CBS.get_conflicts (mmd/planners/multi_agent/cbs.py:166-246) for n_cand candidate joint states at once -- the
 * 'least_collisions' strategy re-runs it once per free sample (cbs.py:446-458).  base_paths_dev [R, T, 2]: the (globally
 * padded, multi_agent_utils.py:120-143) best path positions of every robot; cand_paths_dev [n_cand, T, 2]: candidate paths
 * of robot `agent_id` (NULL with n_cand = 1: the base state itself).  Paths are densified on the fly exactly like
 * densify_trajs (trajectory_utils.py:54-69; densify = 1: identity, 2 when edge conflicts are requested), T_dense =
 * (T - 1) * densify + 1.  Outputs: coll_dev [n_cand, T_dense, R, R] uint8 as check_rr_collisions returns it (margin =
 * 2.1 r), count_dev [n_cand] int32 = number of (t, a, b) rows torch.nonzero would list (NULL to skip), dense_dev
 * [n_cand, R, T_dense, 2] the densified positions (NULL to skip).  Bit-exact with the reference. */
int mmdk_get_conflicts(const float* base_paths_dev, const float* cand_paths_dev, int n_cand, int agent_id, int R, int T,
                       int densify, float margin, uint8_t* coll_dev, int32_t* count_dev, float* dense_dev, void* stream);

/* smooth_trajs (mmd/common/trajectory_utils.py:31-38: scipy.signal.savgol_filter(window 10, polyorder 2, mode 'interp')
 * along the horizon of trajs_dev [B, H, D]) as one linear operator filter_dev [H, H] float64 built by the host
 * (mmd_b200/smoothing.py restates scipy's coefficients and its polynomial edge fit); double accumulation, fp32 out. */
int mmdk_smooth_trajs(const double* filter_dev, const float* trajs_dev, int B, int H, int D, float* out_dev, void* stream);

/* PlanningTask.get_trajs_collision_and_free + cost_smoothness/path_length (TR/tasks/tasks.py:236-311,
 * TR/trajectory/metrics.py:7-39, mpd.py:356-382) on UNNORMALISED trajectories trajs_dev [B, H, D]:
 * free_dev [B] uint8 (no interpolated waypoint in collision AND inside joint limits), cost_dev [B] =
 * path_length + smoothness, waypoint_coll_dev [B, (H-1)*n_interp] uint8 or NULL. */
int mmdk_classify_trajs(const mmdk_guide_env* env, const float* trajs_dev, int B, int H, int n_interp, float radius,
                        const float q_min[2], const float q_max[2], uint8_t* free_dev, float* cost_dev,
                        uint8_t* waypoint_coll_dev, void* stream);

/* Elementwise unnormalise of a whole chain with the per-call global clip rule disabled/enabled by `clip`
 * (mpd.py:354: dataset.unnormalize_trajectories on the [T+2, B, H, D] chain). */
int mmdk_unnormalize(const mmdk_guide_env* env, const float* x_dev, int64_t n_rows, int clip, float* out_dev,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDK_H_ */
