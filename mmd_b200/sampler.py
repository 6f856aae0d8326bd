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
                 process_group=None, use_peer_hash=True):
        self.model, self.guide = model, guide
        self.n_guide_steps = n_guide_steps
        T = model.n_diffusion_steps
        self.t_start_guide = math.ceil(0.5 * T) if t_start_guide is None else t_start_guide  # mpd.py:267
        self.noise_std = noise_std
        self.n_extra = n_diffusion_steps_without_noise
        self.peer_radius, self.peer_weight, self.rep_index = peer_radius, peer_weight, rep_index
        self.pg = process_group
        self.use_peer_hash = use_peer_hash  # False: brute-force scan of the peer table (kept for the equality test)

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

        if x_init is not None:
            x = x_init.reshape(B, H, D).to(torch.float32).clone().contiguous()
        elif noise is not None:
            step_major = noise.dim() == 4
            x = (noise[0] if step_major else noise[:, 0].reshape(B, H, D)).to(dev).clone().contiguous()
        else:
            x = torch.randn(B, H, D, device=dev, generator=generator)
        # hard conditions on x_T (diffusion_model_base.py:195)
        for r, hc in enumerate(hard_conds_l):
            for row, v in hc.items():
                x[r * K:(r + 1) * K, row, :] = torch.as_tensor(v, dtype=torch.float32).to(dev)
        chain = torch.empty(n_steps + 1, B, H, D, device=dev) if return_chain else None
        if return_chain:
            chain[0].copy_(x)

        peers = peer_self = peers_local = peer_hash = None
        if mode == "lockstep":
            peers = torch.zeros(R_total, H, 2, device=dev)
            peers_local = peers[robot_offset:robot_offset + R] if not distributed else torch.zeros(R, H, 2, device=dev)
            peer_self = torch.arange(robot_offset, robot_offset + R, dtype=torch.int32, device=dev)
            if self.use_peer_hash and R_total > 1:
                peer_hash = PeerHash(R_total, H, self.peer_radius, dev)
        elif mode != "independent":
            raise ValueError("mode must be 'lockstep' or 'independent'")
        if constraints_l is None:
            constraints_l = [([], [])] * R
        env, grp, keep = lower_for_step(guide, R, K, H, dev, list(hard_conds_l), list(constraints_l), peers, peer_self,
                                        self.peer_radius, self.peer_weight, peer_hash)
        eps = torch.empty_like(x)
        k = 1
        for i in reversed(range(-self.n_extra, T)):
            guided = guide is not None and i < self.t_start_guide
            if mode == "lockstep" and guided and R_total > 1:
                _lib.check(lib.mmdk_publish_peers(C.byref(env), R, K, H, self.rep_index, _lib.ptr(x),
                                                  _lib.ptr(peers_local), _lib.stream_ptr()))
                if distributed:
                    peers.copy_(gather_peers(peers_local, R_total, self.pg))
                if peer_hash is not None:
                    peer_hash.build(peers)
            t = max(i, 0)
            model.model.forward_t(x, t, precision=model.unet_precision, out=eps)
            sc = model.step_scalars(i, self.n_guide_steps if guided else 0, self.noise_std, True)
            if noise is not None:
                nz = noise[k] if noise.dim() == 4 else noise[:, k].reshape(B, H, D).to(dev).contiguous()
            else:
                nz = torch.randn(B, H, D, device=dev, generator=generator)
            _lib.check(lib.mmdk_ddpm_step(C.byref(env), C.byref(grp), C.byref(sc), H, _lib.ptr(x), _lib.ptr(eps),
                                          _lib.ptr(nz), _lib.ptr(chain[k]) if return_chain else None,
                                          _lib.stream_ptr()))
            k += 1
        del keep
        out = x.reshape(R, K, H, D)
        if return_chain:
            return out, chain.reshape(n_steps + 1, R, K, H, D).transpose(0, 1)
        return out
