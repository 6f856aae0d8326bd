"""Batched (robots x samples) guided-diffusion sampler: the B200-native execution order.

The reference plans robots one planner call at a time (cbs.py:316-324).  Here all R robots x K samples denoise as ONE
batch of R "groups" (include/mmdk.h) -- two kernel launches per reverse step for the whole batch -- in one of two modes:

  independent  every robot has a fixed constraint set known up front (plain CBS/XCBS root, cbs.py:398; or any later
               replanning call): bit-for-bit the same arithmetic as R separate reference calls with the same noise.
  lockstep     (north star) before each reverse step every robot publishes the unnormalised positions of
               one representative sample; the [R, H, 2] table is all-gathered (one NCCL all-gather per timestep when the
               robots are sharded over GPUs) and acts on every other robot as one soft CostConstraint with ranges
               (h, h+1), radius 2.4 r, weight 2e-2 (mmd_params.py:43,52; cbs.py:468-508).  Oracle:
               oracle.port.lockstep_sample, a composition of the reference's own per-timestep function.
"""
import ctypes as C
import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from .diffusion import GaussianDiffusionModel, _run_step, lower_for_step
from .guides import PeerHash


def shard_robots(n_robots_total: int, world_size: int, rank: int):
    """Contiguous robot block of `rank` (flatten (R, K), split contiguous robot blocks across GPUs)."""
    per = (n_robots_total + world_size - 1) // world_size
    lo = min(rank * per, n_robots_total)
    return lo, min(lo + per, n_robots_total)


def gather_peers(peers_local: torch.Tensor, n_robots_total: int, group=None) -> torch.Tensor:
    """All-gather of the representative paths: [R_local, H, 2] per rank -> [R_total, H, 2] in robot order.  One
    collective per guided timestep (16 KiB at R=32): latency-bound; NCCL over NVLink on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = peers_local.shape[0]
    if per * world == n_robots_total:
        out = torch.empty(n_robots_total, *peers_local.shape[1:], dtype=peers_local.dtype, device=peers_local.device)
        dist.all_gather_into_tensor(out, peers_local.contiguous(), group=group)
        return out
    # ragged last shard: pad to the common block size
    blk = (n_robots_total + world - 1) // world
    pad = torch.zeros(blk, *peers_local.shape[1:], dtype=peers_local.dtype, device=peers_local.device)
    pad[:per] = peers_local
    out = torch.empty(blk * world, *peers_local.shape[1:], dtype=peers_local.dtype, device=peers_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_robots_total]


class MultiRobotSampler:
    def __init__(self, model: GaussianDiffusionModel, guide, n_guide_steps=20, t_start_guide=None, noise_std=0.5,
                 n_diffusion_steps_without_noise=1, peer_radius=2.4 * 0.05, peer_weight=2e-2, rep_index=0,
                 process_group=None, use_peer_hash="auto", use_graph=True, exchange="p2p"):
        self.model, self.guide = model, guide
        self.n_guide_steps = n_guide_steps
        T = model.n_diffusion_steps
        self.t_start_guide = math.ceil(0.5 * T) if t_start_guide is None else t_start_guide  # mpd.py:267
        self.noise_std = noise_std
        self.n_extra = n_diffusion_steps_without_noise
        self.peer_radius, self.peer_weight, self.rep_index = peer_radius, peer_weight, rep_index
        self.pg = process_group
        # peer term: spatial hash (O(neighbours)) or brute-force scan of the staged table; "auto" = hash for fleets > 64 robots
        # (measured on B200: at 32 robots the scan executes fewer instructions than the 3x3-cell walk)
        self.use_peer_hash = use_peer_hash
        self.use_graph = use_graph          # capture the step loop once, replay (mmdk_run_chain)
        # lock-step exchange: "p2p" = publication fused into the step kernel over peer memory (mmd_b200/exchange.py; also
        # across processes -> one graph per rank); "nccl" = host loop with one all-gather per guided step (fallback, A/B);
        # None = separate publish kernel, single process only
        self.exchange = exchange
        self._ws = {}

    # -- persistent buffers: stable device pointers let mmdk_run_chain replay its captured CUDA graph -------------------
    def _workspace(self, key, B, H, D, n_steps, R, R_total, robot_offset, mode, return_chain, dev):
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            ws = dict(x=torch.empty(B, H, D, device=dev), eps=torch.empty(B, H, D, device=dev), noise=None,
                      chain=torch.empty(n_steps + 1, B, H, D, device=dev) if return_chain else None,
                      peers=None, peer_self=None, peer_hash=None, lowered=None, keep=None, hc_sig=None, ex=None)
            if mode == "lockstep":
                distributed = R_total != R
                use_ex = self.exchange == "p2p" and R_total > 1
                if use_ex:
                    from .exchange import PeerExchange
                    try:
                        ws["ex"] = PeerExchange(R_total, H, robot_offset, self.rep_index, dev, group=self.pg,
                                                distributed=distributed)
                    except _lib.MMDKError as e:   # CUDA IPC unavailable (e.g. no peer access): NCCL all-gather loop instead
                        if not distributed:
                            raise
                        import warnings
                        warnings.warn(f"peer-memory exchange unavailable ({e}); using the NCCL all-gather loop")
                if ws["ex"] is not None:
                    ws["peers"] = ws["ex"].table[0]
                else:
                    ws["peers"] = torch.zeros(R_total, H, 2, device=dev)
                ws["peer_self"] = torch.arange(robot_offset, robot_offset + R, dtype=torch.int32, device=dev)
                hash_on = (R_total > 64) if self.use_peer_hash == "auto" else bool(self.use_peer_hash)
                if hash_on and R_total > 1:
                    ws["peer_hash"] = PeerHash(R_total, H, self.peer_radius, dev)
            self._ws[key] = ws
        return ws

    @torch.no_grad()
    def sample(self, hard_conds_l: Sequence[dict], n_samples: int, noise: Optional[torch.Tensor] = None,
               mode="lockstep", constraints_l=None, return_chain=False, robot_offset=0, n_robots_total=None,
               x_init=None, n_diffusion_steps=None, generator=None):
        """hard_conds_l: per LOCAL robot {row: state[D]} (normalised).  noise: [R_local, steps+1, K, H, D] (robot-major,
        the oracle's layout), or step-major [steps+1, R_local*K, H, D] (no per-step gather), or None (drawn with
        torch.randn on the device).  Returns final [R_local, K, H, D] (and the chain [R_local, steps+1, K, H, D] if
        return_chain)."""
        lib = _lib.lib()
        model, guide = self.model, self.guide
        dev = model.betas.device
        R, K = len(hard_conds_l), n_samples
        H, D = model.model.n_support_points, model.state_dim
        T = model.n_diffusion_steps if n_diffusion_steps is None else n_diffusion_steps
        n_steps = T + self.n_extra
        B = R * K
        distributed = self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                              and n_robots_total is not None and n_robots_total != R)
        R_total = n_robots_total if n_robots_total is not None else R
        if mode not in ("lockstep", "independent"):
            raise ValueError("mode must be 'lockstep' or 'independent'")
        has_constraints = constraints_l is not None and any(len(c[0]) for c in constraints_l)
        key = (R, K, H, T, n_steps, mode, R_total, robot_offset, bool(return_chain))
        ws = self._workspace(key, B, H, D, n_steps, R, R_total, robot_offset, mode, return_chain, dev)
        x, eps, chain = ws["x"], ws["eps"], ws["chain"]

        # ---- inputs into the persistent buffers ---------------------------------------------------------------------
        step_major = noise is not None and noise.dim() == 4
        if noise is not None and not step_major:          # robot-major (oracle layout) -> step-major frames
            noise = noise.to(dev).to(torch.float32).permute(1, 0, 2, 3, 4).reshape(n_steps + 1, B, H, D).contiguous()
        if x_init is not None:
            x.copy_(x_init.reshape(B, H, D))
        elif noise is not None:
            x.copy_(noise[0])
        else:
            x.normal_(generator=generator)
        if noise is not None:
            nz_all = noise[1:]
            nz_all = nz_all if nz_all.is_contiguous() else nz_all.contiguous()
        else:
            if ws["noise"] is None:
                ws["noise"] = torch.empty(n_steps, B, H, D, device=dev)
            ws["noise"].normal_(generator=generator)      # one draw for the whole chain (the reference: one randn_like per step)
            nz_all = ws["noise"]
        # hard conditions on x_T (diffusion_model_base.py:195), vectorised over the robots
        rows = sorted({int(r) for hc in hard_conds_l for r in hc})
        if rows:
            vals = torch.stack([torch.stack([torch.as_tensor(hc[r], dtype=torch.float32).reshape(-1)[:D].to(dev)
                                             if r in hc else torch.full((D,), float("nan"), device=dev) for r in rows])
                                for hc in hard_conds_l])                                  # [R, n_rows, D]
            xv = x.view(R, K, H, D)
            for j, r in enumerate(rows):
                v = vals[:, j][:, None, :]
                xv[:, :, r, :] = torch.where(torch.isnan(v), xv[:, :, r, :], v)
        if return_chain:
            chain[0].copy_(x)

        # ---- lowering (cached while the structure is unchanged; hard-condition VALUES are refreshed in place) ---------
        if constraints_l is None:
            constraints_l = [([], [])] * R
        hc_sig = tuple(tuple(sorted(int(r) for r in hc)) for hc in hard_conds_l)
        if ws["lowered"] is None or has_constraints or ws["hc_sig"] != hc_sig:
            ws["lowered"] = lower_for_step(guide, R, K, H, dev, list(hard_conds_l), list(constraints_l), ws["peers"],
                                           ws["peer_self"], self.peer_radius, self.peer_weight, ws["peer_hash"],
                                           ws["ex"].seq_ptr if ws["ex"] is not None else None)
            ws["hc_sig"] = hc_sig
            ws["keep"] = None
        else:
            _refresh_hard_vals(hard_conds_l, ws["lowered"][2].rows_d, ws["lowered"][2].vals_d, H)
        env, grp, keep = ws["lowered"]

        steps = []
        for i in reversed(range(-self.n_extra, T)):
            guided = guide is not None and i < self.t_start_guide
            steps.append((max(i, 0), model.step_scalars(i, self.n_guide_steps if guided else 0, self.noise_std, True), guided))

        lock = mode == "lockstep" and R_total > 1
        if not (distributed and lock) or ws["ex"] is not None:
            # one native call for the whole chain (graph replay when the pointers are the persistent ones); with the
            # peer-memory exchange this also covers a fleet sharded over processes
            from .diffusion import run_chain_native
            graph_ok = self.use_graph and not has_constraints
            peers_local = ws["peers"][robot_offset:robot_offset + R] if (lock and ws["ex"] is None) else None
            ws["keep"] = run_chain_native(model, ws["lowered"], [(t, sc) for t, sc, _ in steps], x, eps, nz_all,
                                          chain[1:] if return_chain else None, lockstep=lock, rep_index=self.rep_index,
                                          peers_local=peers_local, use_graph=graph_ok, keep=ws["keep"],
                                          exchange=ws["ex"] if lock else None)
        else:
            # fleet sharded over processes: the representative paths are exchanged between publication and hash build
            peers = ws["peers"]
            peers_local = torch.zeros(R, H, 2, device=dev)
            for k, (t, sc, guided) in enumerate(steps):
                if guided:
                    _lib.check(lib.mmdk_publish_peers(C.byref(env), R, K, H, self.rep_index, _lib.ptr(x),
                                                      _lib.ptr(peers_local), _lib.stream_ptr()))
                    peers.copy_(gather_peers(peers_local, R_total, self.pg))
                    if ws["peer_hash"] is not None:
                        ws["peer_hash"].build(peers)
                model.model.forward_t(x, t, precision=model.unet_precision, out=eps)
                _lib.check(lib.mmdk_ddpm_step(C.byref(env), C.byref(grp), C.byref(sc), H, _lib.ptr(x), _lib.ptr(eps),
                                              _lib.ptr(nz_all[k]), _lib.ptr(chain[k + 1]) if return_chain else None,
                                              _lib.stream_ptr()))
        out = x.reshape(R, K, H, D).clone()
        if return_chain:
            return out, chain.reshape(n_steps + 1, R, K, H, D).transpose(0, 1).clone()
        return out


def _refresh_hard_vals(hard_conds_l, rows_d, vals_d, H):
    """New hard-condition VALUES into the cached lowering (same rows): one stacked copy, no per-robot host sync."""
    n_g, n_r, D = vals_d.shape
    cols = []
    for hc in hard_conds_l:
        vs = [torch.as_tensor(v, dtype=torch.float32).reshape(-1, D)[0] for v in hc.values()]
        vs += [torch.zeros(D, dtype=torch.float32, device=vs[0].device if vs else None)] * (n_r - len(vs))
        cols.append(torch.stack([v.to(vals_d.device, non_blocking=True) for v in vs]))
    vals_d.copy_(torch.stack(cols))
