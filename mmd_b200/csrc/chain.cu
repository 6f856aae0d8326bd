// Whole reverse chain behind ONE C call: the per-timestep loop of p_sample_loop (diffusion_model_base.py:163-211 ->
// sample_functions.py:41-86) with every step's launches (lock-step peer publication + hash, TemporalUnet forward, fused
// posterior/guide/noise step) issued natively, optionally captured once into a CUDA graph and replayed.
//
// Why: at B = R*K = 4096 a reverse step is ~1.4 ms of GPU work and the Python loop's ~6 ctypes launches per step are hidden;
// at the planner's B = K = 64..128 (one MPD.__call__) or with the fleet sharded 8 ways (512 trajectories per GPU) a step
// is ~0.2 ms and the host loop is the bottleneck.  One graph launch per chain removes it.
#include <map>
#include <vector>

#include "common.cuh"

struct mmdk_unet;
extern "C" uint64_t mmdk_internal_unet_uid(const mmdk_unet* net);                      // api.cu
extern "C" uint64_t mmdk_internal_state_uid(const mmdk_unet* net, int mode, int B);   // api.cu

namespace mmdk {

namespace {

struct GraphKey {
  std::vector<uint64_t> words;
  bool operator<(const GraphKey& o) const { return words < o.words; }
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  uint64_t last_use = 0;
};

struct ChainCache {
  std::map<GraphKey, GraphEntry> graphs;
  cudaStream_t side = nullptr;       // capture / replay stream (the caller's stream may be the legacy default stream,
  cudaEvent_t ev_in = nullptr;       // which cannot be captured); ordered with the caller's stream through events
  cudaEvent_t ev_out = nullptr;
  uint64_t tick = 0;
};

ChainCache& cache_for_device(int dev) {
  static std::map<int, ChainCache> caches;
  return caches[dev];
}

void push_bytes(GraphKey& k, const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; i += 8) {
    uint64_t w = 0;
    for (size_t j = 0; j < 8 && i + j < n; ++j) w |= (uint64_t)b[i + j] << (8 * j);
    k.words.push_back(w);
  }
}

}  // namespace

// Captures `issue` once per key into a CUDA graph (after `warm` ran everything that allocates / configures outside the
// capture) and replays it; ordered after the caller's pending work on `s` and before its later work.
template <class Warm, class Issue>
static int run_graph(const GraphKey& key, Warm&& warm, Issue&& issue, cudaStream_t s) {
  int dev = 0;
  MMDK_CUDA(cudaGetDevice(&dev));
  ChainCache& cc = cache_for_device(dev);
  if (!cc.side) {
    MMDK_CUDA(cudaStreamCreateWithFlags(&cc.side, cudaStreamNonBlocking));
    MMDK_CUDA(cudaEventCreateWithFlags(&cc.ev_in, cudaEventDisableTiming));
    MMDK_CUDA(cudaEventCreateWithFlags(&cc.ev_out, cudaEventDisableTiming));
  }
  // everything below runs on the side stream, ordered after the caller's pending work on `s`
  MMDK_CUDA(cudaEventRecord(cc.ev_in, s));
  MMDK_CUDA(cudaStreamWaitEvent(cc.side, cc.ev_in, 0));
  auto it = cc.graphs.find(key);
  if (it == cc.graphs.end()) {
    int rc = warm(cc.side);
    if (rc != MMDK_OK) return rc;
    MMDK_CUDA(cudaStreamSynchronize(cc.side));
    MMDK_CUDA(cudaStreamBeginCapture(cc.side, cudaStreamCaptureModeThreadLocal));
    rc = issue(cc.side);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(cc.side, &graph);
    if (rc != MMDK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    MMDK_CUDA(ce);
    GraphEntry ge;
    ce = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    MMDK_CUDA(ce);
    if (cc.graphs.size() >= 8) {   // bounded: evict the least recently used graph
      auto victim = cc.graphs.begin();
      for (auto j = cc.graphs.begin(); j != cc.graphs.end(); ++j) if (j->second.last_use < victim->second.last_use) victim = j;
      cudaGraphExecDestroy(victim->second.exec);
      cc.graphs.erase(victim);
    }
    it = cc.graphs.emplace(key, ge).first;
  }
  it->second.last_use = ++cc.tick;
  MMDK_CUDA(cudaGraphLaunch(it->second.exec, cc.side));
  MMDK_CUDA(cudaEventRecord(cc.ev_out, cc.side));
  MMDK_CUDA(cudaStreamWaitEvent(s, cc.ev_out, 0));
  return MMDK_OK;
}

static int issue_chain(const mmdk_unet* net, int unet_mode, const mmdk_guide_env* env, const mmdk_groups* groups,
                       const mmdk_chain_desc* ch, int H, float* x, float* eps, const float* noise, float* chain_out,
                       cudaStream_t s) {
  const size_t B = (size_t)groups->n_groups * groups->K;
  const size_t frame = B * H * MMDK_STATE_DIM;
  const mmdk_peer_exchange* ex = ch->exchange;
  auto lock_guided = [&](int i) {
    return i >= 0 && i < ch->n_steps && ch->lockstep && ch->scalars[i].n_guide_steps > 0 && groups->n_peers > 1 && groups->peers_dev;
  };
  bool published = false;   // the previous step kernel already published the current x
  for (int i = 0; i < ch->n_steps; ++i) {
    const mmdk_step_scalars& sc = ch->scalars[i];
    if (lock_guided(i)) {
      int rc;
      if (ex) {
        if (!published) {   // first guided step of the chain: a pure publication of the current x (x <- x)
          mmdk_step_scalars idle = sc;
          idle.do_posterior = 0; idle.n_guide_steps = 0; idle.add_noise = 0; idle.final_hard_conds = 0;
          rc = mmdk_ddpm_step_publish(env, groups, &idle, H, x, nullptr, nullptr, nullptr, ex, s);
          if (rc != MMDK_OK) return rc;
        }
        rc = mmdk_wait_peers(ex, s);
        if (rc != MMDK_OK) return rc;
      } else {
        // every group publishes its representative sample into its row of the peer table (single rank: the table IS local)
        rc = mmdk_publish_peers(env, groups->n_groups, groups->K, H, ch->rep_index, x, ch->peers_local_dev, s);
        if (rc != MMDK_OK) return rc;
      }
      if (groups->peer_cell_start_dev) {
        rc = mmdk_build_peer_hash(groups->peers_dev, groups->peer_seq_dev, groups->n_peers, H, groups->peer_grid,
                                  groups->peer_grid_lo, groups->peer_grid_inv_cell,
                                  const_cast<uint16_t*>(groups->peer_cell_start_dev), const_cast<float*>(groups->peer_sorted_dev), s);
        if (rc != MMDK_OK) return rc;
      }
    }
    int rc = mmdk_unet_forward(net, unet_mode, x, (int)B, ch->t_index[i], eps, s);
    if (rc != MMDK_OK) return rc;
    published = ex && lock_guided(i + 1);   // this step's kernel publishes the x the next (guided) step starts from
    rc = mmdk_ddpm_step_publish(env, groups, &sc, H, x, eps, noise ? noise + (size_t)i * frame : nullptr,
                                chain_out ? chain_out + (size_t)i * frame : nullptr, published ? ex : nullptr, s);
    if (rc != MMDK_OK) return rc;
  }
  return MMDK_OK;
}

}  // namespace mmdk

using namespace mmdk;

extern "C" {

int mmdk_run_chain(const mmdk_unet* net, int unet_mode, const mmdk_guide_env* env, const mmdk_groups* groups,
                   const mmdk_chain_desc* chain, int H, float* x_dev, float* eps_dev, const float* noise_dev,
                   float* chain_out_dev, int use_graph, void* stream) {
  if (!net || !env || !groups || !chain || !x_dev || !eps_dev) return fail(MMDK_EINVAL, "null argument");
  if (chain->n_steps < 1 || !chain->scalars || !chain->t_index) return fail(MMDK_EINVAL, "empty chain description");
  if (chain->lockstep && groups->peers_dev && !chain->peers_local_dev && !chain->exchange)
    return fail(MMDK_EINVAL, "lockstep needs peers_local_dev or an exchange");
  if (chain->exchange && groups->peers_dev && !groups->peer_seq_dev)
    return fail(MMDK_EINVAL, "an exchange publishes into a double-buffered table: set groups->peer_seq_dev");
  cudaStream_t s = (cudaStream_t)stream;
  if (!use_graph) return issue_chain(net, unet_mode, env, groups, chain, H, x_dev, eps_dev, noise_dev, chain_out_dev, s);

  // the graph bakes in every pointer and scalar: key on all of them
  // the executor state of this batch size must exist before the key is built: its id is part of the key (a state that was
  // evicted and rebuilt, or a handle re-created at the same address after a weight reload, must never match an old graph)
  const int B_all = groups->n_groups * groups->K;
  if (mmdk_internal_state_uid(net, unet_mode, B_all) == 0) {
    int rc = mmdk_unet_forward(net, unet_mode, x_dev, B_all, chain->t_index[0], eps_dev, s);   // eps is scratch: harmless
    if (rc != MMDK_OK) return rc;
  }
  GraphKey key;
  key.words.push_back(1);   // kind: single-model chain
  key.words.push_back(mmdk_internal_unet_uid(net));
  key.words.push_back(mmdk_internal_state_uid(net, unet_mode, B_all));
  key.words.push_back((uint64_t)(uintptr_t)net);
  key.words.push_back((uint64_t)unet_mode | ((uint64_t)H << 8) | ((uint64_t)chain->n_steps << 24) | ((uint64_t)chain->lockstep << 48) |
                      ((uint64_t)chain->rep_index << 49));
  key.words.push_back((uint64_t)(uintptr_t)x_dev);
  key.words.push_back((uint64_t)(uintptr_t)eps_dev);
  key.words.push_back((uint64_t)(uintptr_t)noise_dev);
  key.words.push_back((uint64_t)(uintptr_t)chain_out_dev);
  key.words.push_back((uint64_t)(uintptr_t)chain->peers_local_dev);
  if (chain->exchange) push_bytes(key, chain->exchange, sizeof(*chain->exchange));
  else key.words.push_back(0);
  push_bytes(key, env, sizeof(*env));
  push_bytes(key, groups, sizeof(*groups));
  push_bytes(key, chain->scalars, sizeof(mmdk_step_scalars) * chain->n_steps);
  push_bytes(key, chain->t_index, sizeof(int) * chain->n_steps);
  auto warm = [&](cudaStream_t side) -> int {
    // everything that allocates or configures (executor state per batch size, function attributes) OUTSIDE the capture
    const int B = groups->n_groups * groups->K;
    int rc = mmdk_unet_forward(net, unet_mode, x_dev, B, chain->t_index[0], eps_dev, side);
    if (rc != MMDK_OK) return rc;
    mmdk_step_scalars dry = chain->scalars[0];
    dry.n_guide_steps = 0; dry.do_posterior = 0; dry.add_noise = 0; dry.final_hard_conds = 0;   // x <- x: configures the launch only
    return mmdk_ddpm_step(env, groups, &dry, H, x_dev, nullptr, nullptr, nullptr, side);
  };
  auto issue = [&](cudaStream_t side) -> int {
    return issue_chain(net, unet_mode, env, groups, chain, H, x_dev, eps_dev, noise_dev, chain_out_dev, side);
  };
  return run_graph(key, warm, issue, s);
}

int mmdk_run_chain_ensemble(const mmdk_ensemble_desc* ens, int H, int use_graph, void* stream) {
  if (!ens || ens->n_tiles < 1 || ens->n_tiles > MMDK_MAX_TILES || !ens->tiles) return fail(MMDK_EINVAL, "bad ensemble description");
  if (ens->n_steps < 1 || !ens->t_index) return fail(MMDK_EINVAL, "empty chain description");
  if (ens->n_cross > 0 && !ens->cross) return fail(MMDK_EINVAL, "null cross-condition list");
  for (int m = 0; m < ens->n_tiles; ++m) {
    const mmdk_ensemble_tile& t = ens->tiles[m];
    if (!t.net || !t.env || !t.groups || !t.scalars || !t.x_dev || !t.eps_dev) return fail(MMDK_EINVAL, "null argument in a tile");
    if (t.groups->peers_dev) return fail(MMDK_EINVAL, "lock-step peers are not part of the ensemble loop");
  }
  for (int c = 0; c < ens->n_cross; ++c) {
    const mmdk_cross_cond& cc = ens->cross[c];
    if (cc.m1 < 0 || cc.m1 >= ens->n_tiles || cc.m2 < 0 || cc.m2 >= ens->n_tiles || cc.row_lo < 0 || cc.row_hi < cc.row_lo)
      return fail(MMDK_EINVAL, "bad cross condition");
  }
  cudaStream_t s = (cudaStream_t)stream;
  auto issue = [&](cudaStream_t q) -> int {
    // m_done: the tile that has just been stepped in reverse step `step` (-1: none).  The reference appends x[m] itself to its
    // chain lists (diffusion_ensemble.py:77,103-105: no clone) and apply_cross_conditioning writes in place, so the stitch that
    // follows tile m's step also rewrites the stitched waypoint of the frame already recorded for every tile that has NOT been
    // stepped yet in this reverse step (its x is still the recorded tensor).  Reproduced here: the row goes into that frame too.
    auto cross_all = [&](int m_done, int step) -> int {
      for (int c = 0; c < ens->n_cross; ++c) {
        const mmdk_cross_cond& cc = ens->cross[c];
        const size_t o = (size_t)cc.row_lo * H * MMDK_STATE_DIM;
        int rc = mmdk_cross_condition(ens->tiles[cc.m1].x_dev + o, ens->tiles[cc.m2].x_dev + o, cc.row_hi - cc.row_lo, H, cc.ind1,
                                      cc.ind2, cc.rel, cc.bnd, q);
        if (rc != MMDK_OK) return rc;
        for (int side = 0; side < 2; ++side) {
          const int tm = side ? cc.m2 : cc.m1;
          int ind = side ? cc.ind2 : cc.ind1;
          if (tm <= m_done || cc.row_hi <= cc.row_lo) continue;
          const mmdk_ensemble_tile& t = ens->tiles[tm];
          const size_t frame = (size_t)t.groups->n_groups * t.groups->K * H * MMDK_STATE_DIM;
          float* prev = (step == 0) ? t.chain_init_dev : (t.chain_out_dev ? t.chain_out_dev + (size_t)(step - 1) * frame : nullptr);
          if (!prev) continue;
          if (ind < 0) ind += H;
          const size_t off = o + (size_t)ind * MMDK_STATE_DIM, pitch = (size_t)H * MMDK_STATE_DIM * sizeof(float);
          MMDK_CUDA(cudaMemcpy2DAsync(prev + off, pitch, t.x_dev + off, pitch, MMDK_STATE_DIM * sizeof(float),
                                      (size_t)(cc.row_hi - cc.row_lo), cudaMemcpyDeviceToDevice, q));
        }
      }
      return MMDK_OK;
    };
    for (int i = 0; i < ens->n_steps; ++i) {
      for (int m = 0; m < ens->n_tiles; ++m) {
        const mmdk_ensemble_tile& t = ens->tiles[m];
        const size_t B = (size_t)t.groups->n_groups * t.groups->K;
        const size_t frame = B * H * MMDK_STATE_DIM;
        int rc = mmdk_unet_forward(t.net, t.unet_mode, t.x_dev, (int)B, ens->t_index[i], t.eps_dev, q);
        if (rc != MMDK_OK) return rc;
        rc = mmdk_ddpm_step(t.env, t.groups, &t.scalars[i], H, t.x_dev, t.eps_dev, t.noise_dev ? t.noise_dev + (size_t)i * frame : nullptr,
                            nullptr, q);
        if (rc != MMDK_OK) return rc;
        rc = cross_all(m, i);   // diffusion_ensemble.py:99-101: after EVERY tile's step
        if (rc != MMDK_OK) return rc;
      }
      for (int m = 0; m < ens->n_tiles; ++m) {   // the chain frame holds the step's state after all stitches (:103-105)
        const mmdk_ensemble_tile& t = ens->tiles[m];
        if (!t.chain_out_dev) continue;
        const size_t frame = (size_t)t.groups->n_groups * t.groups->K * H * MMDK_STATE_DIM;
        MMDK_CUDA(cudaMemcpyAsync(t.chain_out_dev + (size_t)i * frame, t.x_dev, frame * sizeof(float), cudaMemcpyDeviceToDevice, q));
      }
    }
    return MMDK_OK;
  };
  if (!use_graph) return issue(s);
  for (int m = 0; m < ens->n_tiles; ++m) {   // executor states first: their ids are part of the key (see mmdk_run_chain)
    const mmdk_ensemble_tile& t = ens->tiles[m];
    const int B = t.groups->n_groups * t.groups->K;
    if (mmdk_internal_state_uid(t.net, t.unet_mode, B) == 0) {
      int rc = mmdk_unet_forward(t.net, t.unet_mode, t.x_dev, B, ens->t_index[0], t.eps_dev, s);
      if (rc != MMDK_OK) return rc;
    }
  }
  GraphKey key;
  key.words.push_back(2);   // kind: ensemble chain
  key.words.push_back(((uint64_t)H << 8) | ((uint64_t)ens->n_steps << 24) | ((uint64_t)ens->n_tiles << 48) | ((uint64_t)ens->n_cross << 52));
  push_bytes(key, ens->t_index, sizeof(int) * ens->n_steps);
  for (int m = 0; m < ens->n_tiles; ++m) {
    const mmdk_ensemble_tile& t = ens->tiles[m];
    key.words.push_back((uint64_t)(uintptr_t)t.net);
    key.words.push_back(mmdk_internal_unet_uid(t.net));
    key.words.push_back(mmdk_internal_state_uid(t.net, t.unet_mode, t.groups->n_groups * t.groups->K));
    key.words.push_back((uint64_t)t.unet_mode);
    key.words.push_back((uint64_t)(uintptr_t)t.x_dev);
    key.words.push_back((uint64_t)(uintptr_t)t.eps_dev);
    key.words.push_back((uint64_t)(uintptr_t)t.noise_dev);
    key.words.push_back((uint64_t)(uintptr_t)t.chain_out_dev);
    key.words.push_back((uint64_t)(uintptr_t)t.chain_init_dev);
    push_bytes(key, t.env, sizeof(*t.env));
    push_bytes(key, t.groups, sizeof(*t.groups));
    push_bytes(key, t.scalars, sizeof(mmdk_step_scalars) * ens->n_steps);
  }
  if (ens->n_cross) push_bytes(key, ens->cross, sizeof(mmdk_cross_cond) * ens->n_cross);
  auto warm = [&](cudaStream_t side) -> int {
    for (int m = 0; m < ens->n_tiles; ++m) {
      const mmdk_ensemble_tile& t = ens->tiles[m];
      const int B = t.groups->n_groups * t.groups->K;
      int rc = mmdk_unet_forward(t.net, t.unet_mode, t.x_dev, B, ens->t_index[0], t.eps_dev, side);
      if (rc != MMDK_OK) return rc;
      mmdk_step_scalars dry = t.scalars[0];
      dry.n_guide_steps = 0; dry.do_posterior = 0; dry.add_noise = 0; dry.final_hard_conds = 0;
      rc = mmdk_ddpm_step(t.env, t.groups, &dry, H, t.x_dev, nullptr, nullptr, nullptr, side);
      if (rc != MMDK_OK) return rc;
    }
    return MMDK_OK;
  };
  return run_graph(key, warm, issue, s);
}

}  // extern "C"
