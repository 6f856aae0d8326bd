"""GaussianDiffusionModel / ddpm_sample_fn / guide_gradient_steps with the reference's call surface, run by libmmdk.

Reference: mmd/models/diffusion_models/diffusion_model_base.py:48-433, sample_functions.py:8-107, helpers.py:29-49.
Host code stays Python/PyTorch (schedule buffers, the step loop, noise draws in the reference's order); every
per-timestep tensor op is one of two hand-written kernels: the TemporalUnet forward and the fused
posterior -> 20 x guide -> noise -> hard-conditioning step.
"""
import ctypes as C
import math
from copy import copy

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .guides import GuideManagerTrajectoriesWithVelocity


# ---------------------------------------------------------------------------------------------------------------------
# schedules (helpers.py:29-49) -- computed once on the host with the reference's expressions
# ---------------------------------------------------------------------------------------------------------------------
def cosine_beta_schedule(n_diffusion_steps, s=0.008, a_min=0, a_max=0.999, dtype=torch.float32):
    steps = n_diffusion_steps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.tensor(np.clip(betas, a_min=a_min, a_max=a_max), dtype=dtype)


def exponential_beta_schedule(n_diffusion_steps, beta_start=1e-4, beta_end=1.0):
    x = torch.linspace(0, n_diffusion_steps, n_diffusion_steps)
    beta_start, beta_end = torch.tensor(beta_start), torch.tensor(beta_end)
    a = 1 / n_diffusion_steps * torch.log(beta_end / beta_start)
    return beta_start * torch.exp(a * x)


def make_timesteps(batch_size, i, device):  # diffusion_model_base.py:27-29
    return torch.full((batch_size,), i, device=device, dtype=torch.long)


# ---------------------------------------------------------------------------------------------------------------------
# sample functions (sample_functions.py)
# ---------------------------------------------------------------------------------------------------------------------
def apply_hard_conditioning(x, conditions):  # sample_functions.py:8-14 (host-API helper; the fused step does this itself)
    for t, val in conditions.items():
        x[:, t, :] = val.clone()
    return x


def apply_cross_conditioning(x, conditions, transforms):  # sample_functions.py:17-31
    lib = _lib.lib()
    for (m1, m2), (ind1, ind2) in conditions.items():
        rel = (transforms[m2] - transforms[m1]).detach().to(torch.float32).cpu()
        D = x[m1].shape[2]
        if D > rel.shape[0]:
            rel = torch.cat([rel, torch.zeros(D - rel.shape[0])])
        bnd = rel / torch.norm(rel, keepdim=True)
        bnd[bnd == 0] = 1e6
        B, H, _ = x[m1].shape
        _lib.check(lib.mmdk_cross_condition(_lib.ptr(x[m1]), _lib.ptr(x[m2]), B, H, int(ind1), int(ind2),
                                            (C.c_float * 4)(*rel.tolist()), (C.c_float * 4)(*bnd.tolist()),
                                            _lib.stream_ptr()))
    return x


def extract(a, t, x_shape):  # sample_functions.py:34-37
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def _t_int(t):
    return int(t[0]) if torch.is_tensor(t) else int(t)


def _hard_rows(hard_conds):
    """{row: [B, D] or [D]} -> {row: [D]} (run_inference repeats one state over the batch, :327-329)."""
    out = {}
    for k, v in (hard_conds or {}).items():
        if torch.is_tensor(v) and v.dim() == 2:
            # a planner call shares ONE state per conditioned waypoint (the kernel's group layout); per-sample conditions
            # would be silently dropped otherwise
            if v.shape[0] > 1 and not bool((v == v[0:1]).all()):
                raise ValueError("per-sample hard conditions are not supported: all rows of a [B, D] condition must be equal")
            v = v[0]
        out[k] = v
    return out


@torch.no_grad()
def ddpm_sample_fn(model, x, hard_conds, context, t, guide=None, n_guide_steps=1, scale_grad_by_std=False,
                   t_start_guide=torch.inf, noise_std_extra_schedule_fn=None, debug=False, noise=None, **kwargs):
    """sample_functions.py:41-86.  Returns (x_new, None); `x` is left untouched like the reference.
    `noise` (optional) replaces the torch.randn_like draw (parity tests hand the oracle's noise in)."""
    if context is not None:
        raise NotImplementedError("context models are not on the sampling path (mpd.py:210)")
    if scale_grad_by_std:
        raise NotImplementedError("scale_grad_by_std=True is never used by the planners (sample_functions.py:45)")
    t_single = _t_int(t)
    if noise is None:
        noise = torch.randn_like(x)
    noise_std = 1.0 if noise_std_extra_schedule_fn is None else float(noise_std_extra_schedule_fn(t_single))
    guided = guide is not None and t_single < t_start_guide
    x_new = x.clone().contiguous()
    model._fused_step(x_new, _hard_rows(hard_conds), t_single, guide if guided else None, n_guide_steps, noise,
                      noise_std, chain_slot=None, final_hard_cond=False)
    return x_new, None


@torch.no_grad()
def guide_gradient_steps(x, hard_conds=None, guide=None, n_guide_steps=1, scale_grad_by_std=False, model_var=None,
                         debug=False, **kwargs):
    """sample_functions.py:89-107: n x (x += guide(x); apply_hard_conditioning), fused in one launch."""
    if scale_grad_by_std:
        raise NotImplementedError("scale_grad_by_std=True is never used by the planners")
    if not isinstance(guide, GuideManagerTrajectoriesWithVelocity):
        raise TypeError("guide must be a mmd_b200 GuideManagerTrajectoriesWithVelocity")
    x_new = x.clone().contiguous()
    _run_step(guide, x_new, [_hard_rows(hard_conds)], x_new.shape[0], None, None, None,
              _lib.StepScalars(do_posterior=0, n_guide_steps=int(n_guide_steps), add_noise=0, clip_denoised=1,
                               predict_epsilon=1, final_hard_conds=0), [guide._own_constraints()])
    return x_new


def _run_step(guide, x, hard_conds_per_group, K, eps, noise, chain_slot, sc, constraints_per_group=None, peers=None,
              peer_self=None, peer_radius=0.0, peer_weight=0.0, lowered=None):
    lib = _lib.lib()
    B, H, D = x.shape
    n_groups = B // K
    if lowered is None:
        lowered = lower_for_step(guide, n_groups, K, H, x.device, hard_conds_per_group, constraints_per_group, peers,
                                 peer_self, peer_radius, peer_weight)
    env, grp, _keep = lowered
    _lib.check(lib.mmdk_ddpm_step(C.byref(env), C.byref(grp), C.byref(sc), H, _lib.ptr(x), _lib.ptr(eps),
                                  _lib.ptr(noise), _lib.ptr(chain_slot), _lib.stream_ptr()))


def lower_for_step(guide, n_groups, K, H, device, hard_conds_per_group, constraints_per_group=None, peers=None,
                   peer_self=None, peer_radius=0.0, peer_weight=0.0, peer_hash=None, peer_seq_ptr=None):
    """(mmdk_guide_env, mmdk_groups, keepalive) for a batch of `n_groups` planner calls of K samples."""
    if guide is not None:
        env, keep = guide.lower_env(device)
        helper = guide
    else:
        env, keep = _lib.GuideEnv(), []
        helper = GuideManagerTrajectoriesWithVelocity.__new__(GuideManagerTrajectoriesWithVelocity)
    if constraints_per_group is None:
        constraints_per_group = [([], [])] * n_groups
    grp, keep2 = GuideManagerTrajectoriesWithVelocity.lower_groups(
        helper, n_groups, K, H, device, constraints_per_group, hard_conds_per_group, peers, peer_self, peer_radius,
        peer_weight, peer_hash, peer_seq_ptr)
    allk = type(keep2)(list(keep) + list(keep2))
    allk.rows_d, allk.vals_d = keep2.rows_d, keep2.vals_d
    return env, grp, allk


def run_chain_native(model, lowered, steps, x, eps, noise_steps=None, chain_steps=None, lockstep=False, rep_index=0,
                     peers_local=None, use_graph=False, keep=None, exchange=None):
    """One mmdk_run_chain call for a list of reverse steps.  steps: [(t_index, StepScalars)]; noise_steps / chain_steps:
    contiguous [n_steps, B, H, D] tensors (or None).  Returns the ctypes objects that must stay alive with a captured graph."""
    lib = _lib.lib()
    env, grp, _k = lowered
    n = len(steps)
    if keep is None:
        t_arr = (C.c_int * n)(*[int(t) for t, _ in steps])
        sc_arr = (_lib.StepScalars * n)()
        for i, (_, sc) in enumerate(steps):
            sc_arr[i] = sc
        desc = _lib.ChainDesc()
        desc.n_steps, desc.t_index, desc.scalars = n, t_arr, sc_arr
        keep = (desc, t_arr, sc_arr)
    desc = keep[0]
    desc.lockstep, desc.rep_index = int(bool(lockstep)), int(rep_index)
    desc.peers_local_dev = peers_local.data_ptr() if peers_local is not None else None
    desc.exchange = C.pointer(exchange.struct) if exchange is not None else None
    B, H, D = x.shape
    if noise_steps is not None:
        assert noise_steps.is_contiguous() and noise_steps.shape[0] >= n and tuple(noise_steps.shape[1:]) == (B, H, D)
    if chain_steps is not None:
        assert chain_steps.is_contiguous() and chain_steps.shape[0] >= n and tuple(chain_steps.shape[1:]) == (B, H, D)
    unet = model.model
    unet.ensure_time_table(max(t for t, _ in steps) + 1)
    def issue():
        mode = unet.native_mode(model.unet_precision)
        _lib.check(lib.mmdk_run_chain(unet.native(), mode, C.byref(env), C.byref(grp), C.byref(desc), H, _lib.ptr(x), _lib.ptr(eps),
                                      _lib.ptr(noise_steps), _lib.ptr(chain_steps), int(bool(use_graph)), _lib.stream_ptr()))
        return mode
    while True:
        mode = unet.native_mode(model.unet_precision)
        try:
            issue()
            break
        except ValueError:
            # unet_precision='auto' on a shape a tensor-core builder rejects: the next native executor instead (a builder
            # fails before the first launch of the chain, so x is untouched)
            if (model.unet_precision or unet.unet_precision) != "auto" or mode == _lib.UNET_FP32:
                raise
            unet._note_rejection(mode)
    return keep


def cross_cond_entries(cross_conds, transforms, D, row_lo, row_hi):
    """[_lib.CrossCond] for the batch rows [row_lo, row_hi) from the reference's cross-condition dict and tile transforms
    (sample_functions.py:17-31: rel = transforms[m2] - transforms[m1] padded to D, bnd = rel / |rel| with zeros -> 1e6)."""
    out = []
    for (m1, m2), (ind1, ind2) in cross_conds.items():
        rel = (transforms[m2] - transforms[m1]).detach().to(torch.float32).cpu()
        if D > rel.shape[0]:
            rel = torch.cat([rel, torch.zeros(D - rel.shape[0])])
        bnd = rel / torch.norm(rel, keepdim=True)
        bnd[bnd == 0] = 1e6
        cc = _lib.CrossCond()
        cc.m1, cc.m2, cc.ind1, cc.ind2, cc.row_lo, cc.row_hi = int(m1), int(m2), int(ind1), int(ind2), int(row_lo), int(row_hi)
        for d in range(4):
            cc.rel[d], cc.bnd[d] = float(rel[d]), float(bnd[d])
        out.append(cc)
    return out


def run_ensemble_chain_native(models, lowered, step_list, x, eps, noise_steps, chain_steps, cross_entries, use_graph=False,
                              chain_init=None):
    """One mmdk_run_chain_ensemble call: the multi-tile reverse loop of DiffusionsEnsemble.p_sample_loop
    (diffusion_ensemble.py:78-106).  models / lowered / x / eps / noise_steps / chain_steps: dicts keyed by tile (tiles must be
    0..n-1 in order); step_list[m] = [(t_index, StepScalars)] per tile; noise_steps[m] / chain_steps[m]: contiguous
    [n_steps, B, H, D] (or None); cross_entries: [_lib.CrossCond]."""
    lib = _lib.lib()
    tiles = list(models.keys())
    assert tiles == list(range(len(tiles))), "ensemble tiles must be keyed 0..n-1"
    n = len(step_list[tiles[0]])
    t_arr = (C.c_int * n)(*[int(t) for t, _ in step_list[tiles[0]]])
    tile_arr = (_lib.EnsembleTile * len(tiles))()
    keep = [t_arr, tile_arr]
    H = x[tiles[0]].shape[1]
    for m in tiles:
        assert [int(t) for t, _ in step_list[m]] == list(t_arr), "tile models must share the timestep sequence"
        B, Hm, D = x[m].shape
        sc_arr = (_lib.StepScalars * n)()
        for i, (_, sc) in enumerate(step_list[m]):
            sc_arr[i] = sc
        unet = models[m].model
        unet.ensure_time_table(max(t_arr) + 1)
        env, grp, k = lowered[m]
        keep += [sc_arr, k]
        for nm, buf in (("noise", noise_steps), ("chain", chain_steps)):
            if buf is not None and buf.get(m) is not None:
                assert buf[m].is_contiguous() and buf[m].shape[0] >= n and tuple(buf[m].shape[1:]) == (B, Hm, D), nm
        te = tile_arr[m]
        te.net, te.unet_mode = unet.native(), unet.native_mode(models[m].unet_precision)
        te.env, te.groups, te.scalars = C.pointer(env), C.pointer(grp), sc_arr
        te.x_dev, te.eps_dev = x[m].data_ptr(), eps[m].data_ptr()
        te.noise_dev = noise_steps[m].data_ptr() if noise_steps is not None and noise_steps.get(m) is not None else None
        te.chain_out_dev = chain_steps[m].data_ptr() if chain_steps is not None and chain_steps.get(m) is not None else None
        # the frame recorded before the first step: in the reference it IS x[m] until tile m is stepped (aliasing, mmdk.h)
        te.chain_init_dev = chain_init[m].data_ptr() if chain_init is not None and chain_init.get(m) is not None else None
    cross_arr = (_lib.CrossCond * max(1, len(cross_entries)))(*cross_entries)
    desc = _lib.EnsembleDesc()
    desc.n_tiles, desc.tiles, desc.n_steps, desc.t_index = len(tiles), tile_arr, n, t_arr
    desc.n_cross, desc.cross = len(cross_entries), cross_arr
    keep += [cross_arr, desc]
    _lib.check(lib.mmdk_run_chain_ensemble(C.byref(desc), H, int(bool(use_graph)), _lib.stream_ptr()))
    return keep


class GaussianDiffusionModel(nn.Module):
    def __init__(self, model=None, variance_schedule='exponential', n_diffusion_steps=100, clip_denoised=True,
                 predict_epsilon=False, loss_type='l2', context_model=None, **kwargs):
        super().__init__()
        self.model = model
        self.context_model = context_model
        self.n_diffusion_steps = n_diffusion_steps
        self.state_dim = self.model.state_dim
        if variance_schedule == 'cosine':
            betas = cosine_beta_schedule(n_diffusion_steps, s=0.008, a_min=0, a_max=0.999)
        elif variance_schedule == 'exponential':
            betas = exponential_beta_schedule(n_diffusion_steps, beta_start=1e-4, beta_end=1.0)
        else:
            raise NotImplementedError
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = torch.cat([torch.ones(1), alphas_cumprod[:-1]])
        self.clip_denoised = clip_denoised
        self.predict_epsilon = predict_epsilon
        # diffusion_model_base.py:76-105 (same names: part of the checkpoint contract)
        self.register_buffer('betas', betas)
        self.register_buffer('alphas_cumprod', alphas_cumprod)
        self.register_buffer('alphas_cumprod_prev', alphas_cumprod_prev)
        self.register_buffer('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        self.register_buffer('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        self.register_buffer('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer('posterior_variance', posterior_variance)
        self.register_buffer('posterior_log_variance_clipped', torch.log(torch.clamp(posterior_variance, min=1e-20)))
        self.register_buffer('posterior_mean_coef1', betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        self.register_buffer('posterior_mean_coef2', (1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod))
        self.model.ensure_time_table(n_diffusion_steps)
        self._host = None
        self.unet_precision = kwargs.get('unet_precision', None)

    # -- host copies of the schedule (one D2H at first use; `extract` then never touches the device) --------------
    def _sched(self):
        if self._host is None:
            names = ['sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod', 'posterior_mean_coef1',
                     'posterior_mean_coef2', 'posterior_log_variance_clipped', 'sqrt_alphas_cumprod',
                     'sqrt_one_minus_alphas_cumprod']
            h = {n: getattr(self, n).detach().cpu() for n in names}
            h['model_std'] = torch.exp(0.5 * h['posterior_log_variance_clipped'])  # sample_functions.py:60
            self._host = {k: v.tolist() for k, v in h.items()}
        return self._host

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._host = None
        return out

    def step_scalars(self, t_single, n_guide_steps, noise_std, has_noise=True):
        """mmdk_step_scalars of reverse step `t_single` (negative = extra noise-free steps at t=0, :51-54)."""
        s = self._sched()
        t = max(t_single, 0)
        sc = _lib.StepScalars()
        sc.do_posterior = 1
        sc.sqrt_recip_alphas_cumprod = s['sqrt_recip_alphas_cumprod'][t]
        sc.sqrt_recipm1_alphas_cumprod = s['sqrt_recipm1_alphas_cumprod'][t]
        sc.posterior_mean_coef1 = s['posterior_mean_coef1'][t]
        sc.posterior_mean_coef2 = s['posterior_mean_coef2'][t]
        sc.clip_denoised = int(bool(self.clip_denoised))
        sc.predict_epsilon = int(bool(self.predict_epsilon))
        sc.n_guide_steps = int(n_guide_steps)
        sc.add_noise = int(has_noise and t != 0)  # noise[t == 0] = 0 (sample_functions.py:76)
        sc.model_std = s['model_std'][t]
        sc.noise_std = float(noise_std)
        sc.final_hard_conds = 1
        return sc

    # -- one fused reverse step --------------------------------------------------------------------------------------
    def _fused_step(self, x, hard_rows, t_single, guide, n_guide_steps, noise, noise_std, chain_slot,
                    final_hard_cond=True, lowered=None, eps_buf=None, K=None, constraints_per_group=None):
        if not self.clip_denoised:
            raise RuntimeError("clip_denoised=False is rejected by the reference too (diffusion_model_base.py:157)")
        t = max(t_single, 0)
        eps = self.model.forward_t(x, t, precision=self.unet_precision, out=eps_buf)
        sc = self.step_scalars(t_single, n_guide_steps if guide is not None else 0, noise_std, noise is not None)
        K = K or x.shape[0]
        if constraints_per_group is None and guide is not None:
            constraints_per_group = [guide._own_constraints()] * (x.shape[0] // K)
        hcs = hard_rows if isinstance(hard_rows, list) else [hard_rows] * (x.shape[0] // K)
        sc.final_hard_conds = int(final_hard_cond)  # the trailing apply_hard_conditioning belongs to p_sample_loop
        _run_step(guide, x, hcs, K, eps, noise.contiguous() if noise is not None else None, chain_slot, sc,
                  constraints_per_group, lowered=lowered)
        return x

    # ------------------------------------------ sampling ------------------------------------------#
    @torch.no_grad()
    def p_sample_loop(self, shape, hard_conds, n_diffusion_steps: int, context=None, return_chain=False,
                      sample_fn=ddpm_sample_fn, n_diffusion_steps_without_noise=0, warm_start_path_b=None, noise=None,
                      guide=None, n_guide_steps=1, t_start_guide=torch.inf, noise_std_extra_schedule_fn=None,
                      scale_grad_by_std=False, **sample_kwargs):
        """diffusion_model_base.py:163-211.  `noise` (optional, [steps+1, B, H, D]) replaces the generator draws:
        noise[0] = x_T, noise[1+k] = k-th randn_like."""
        if context is not None:
            raise NotImplementedError("context models are not on the sampling path")
        if sample_fn is not ddpm_sample_fn:
            raise NotImplementedError("only ddpm_sample_fn is lowered (mpd.py:421)")
        device = self.betas.device
        B, H, D = shape
        n_steps = n_diffusion_steps + n_diffusion_steps_without_noise
        if warm_start_path_b is not None:
            x = warm_start_path_b.to(torch.float32).clone().contiguous()
        elif noise is not None:
            x = noise[0].to(device).clone().contiguous()
        else:
            x = torch.randn(shape, device=device)
        hard_rows = _hard_rows(hard_conds)
        x = apply_hard_conditioning(x, hard_rows)
        chain = torch.empty(n_steps + 1, B, H, D, device=device) if return_chain else None
        if return_chain:
            chain[0].copy_(x)
        lowered = lower_for_step(guide, 1, B, H, device, [hard_rows],
                                 [guide._own_constraints()] if guide is not None else None)
        eps_buf = torch.empty_like(x)
        # noise frames in the reference's draw order (one randn_like per step, sample_functions.py:74), then the whole loop
        # as ONE native call (mmdk_run_chain)
        if noise is not None:
            nz_all = noise[1:1 + n_steps].to(device).to(torch.float32).contiguous()
        else:
            nz_all = torch.stack([torch.randn_like(x) for _ in range(n_steps)])
        steps = []
        for i in reversed(range(-n_diffusion_steps_without_noise, n_diffusion_steps)):
            noise_std = 1.0 if noise_std_extra_schedule_fn is None else float(noise_std_extra_schedule_fn(i))
            guided = guide is not None and i < t_start_guide
            sc = self.step_scalars(i, n_guide_steps if guided else 0, noise_std, True)
            steps.append((max(i, 0), sc))
        if not self.clip_denoised:
            raise RuntimeError("clip_denoised=False is rejected by the reference too (diffusion_model_base.py:157)")
        run_chain_native(self, lowered, steps, x, eps_buf, nz_all, chain[1:] if return_chain else None)
        if return_chain:
            return x, chain.transpose(0, 1)  # [B, steps+1, H, D] like torch.stack(chain, dim=1)
        return x

    @torch.no_grad()
    def conditional_sample(self, hard_conds, n_diffusion_steps: int, horizon=None, batch_size=1, ddim=False,
                           warm_start_path_b=None, **sample_kwargs):  # :293-307
        if ddim:
            raise NotImplementedError("ddim sampling is not used by the planners (mpd.py:423 commented out)")
        shape = (batch_size, horizon or self.model.n_support_points, self.state_dim)
        return self.p_sample_loop(shape, hard_conds, n_diffusion_steps=n_diffusion_steps,
                                  warm_start_path_b=warm_start_path_b, **sample_kwargs)

    def forward(self, cond, *args, **kwargs):  # :309-311
        raise NotImplementedError

    @torch.no_grad()
    def warmup(self, horizon=64, device='cuda'):  # :313-318
        x = torch.randn((2, horizon, self.state_dim), device=device)
        self.model.forward_t(x, 1, precision=self.unet_precision)

    @torch.no_grad()
    def run_inference(self, context=None, hard_conds=None, n_samples=1, return_chain=False, **diffusion_kwargs):
        """diffusion_model_base.py:321-351 -> [T+2, B, H, D] chain (normalised) or the last frame."""
        hard_conds = copy(hard_conds)
        for k, v in hard_conds.items():
            hard_conds[k] = v.reshape(1, -1).repeat(n_samples, 1)
        samples, chain = self.conditional_sample(hard_conds, n_diffusion_steps=self.n_diffusion_steps, context=context,
                                                 batch_size=n_samples, return_chain=True, **diffusion_kwargs)
        chain = chain.transpose(0, 1)  # 'b diffsteps h d -> diffsteps b h d'
        return chain if return_chain else chain[-1]

    @torch.no_grad()
    def run_local_inference(self, seed_trajectory_b, n_noising_steps, n_denoising_steps, context=None, hard_conds=None,
                            n_samples=1, return_chain=False, noise=None, **diffusion_kwargs):
        """diffusion_model_base.py:353-421: q_sample(seed, t=n_noising) then n_denoising reverse steps (warm start)."""
        hard_conds = copy(hard_conds)
        for k, v in hard_conds.items():
            hard_conds[k] = v.reshape(1, -1).repeat(n_samples, 1)
        if n_noising_steps is None:
            seed_noised = None
        else:
            B = seed_trajectory_b.shape[0]
            t = make_timesteps(B, n_noising_steps, seed_trajectory_b.device)
            seed_noised = self.q_sample(seed_trajectory_b, t, noise=None if noise is None else noise[0])
        samples, chain = self.conditional_sample(hard_conds, n_diffusion_steps=n_denoising_steps, context=context,
                                                 batch_size=n_samples, return_chain=True,
                                                 warm_start_path_b=seed_noised, noise=noise, **diffusion_kwargs)
        chain = chain.transpose(0, 1)
        return chain if return_chain else chain[-1]

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):  # :425-433
        lib = _lib.lib()
        if noise is None:
            noise = torch.randn_like(x_start)
        s = self._sched()
        ti = _t_int(t)
        xs = x_start.to(torch.float32).contiguous()
        nz = noise.to(xs.device).to(torch.float32).contiguous()
        out = torch.empty_like(xs)
        _lib.check(lib.mmdk_q_sample(_lib.ptr(xs), _lib.ptr(nz), s['sqrt_alphas_cumprod'][ti],
                                     s['sqrt_one_minus_alphas_cumprod'][ti], xs.numel(), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def loss(self, *a, **k):
        raise NotImplementedError("training is outside the sampling hot path (SURVEY row 18)")
