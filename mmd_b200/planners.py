"""Planner objects of the reference's boundary: MPD (single agent) and DiffusionsEnsemble (multi-tile).

Reference: mmd/planners/single_agent/mpd.py:58-520, mmd/planners/single_agent/common.py:26-46,
mmd/models/diffusion_models/diffusion_ensemble.py:25-312.  Host logic only; every tensor op is a libmmdk kernel
(UNet, fused DDPM/guide step, classification, unnormalise) except the final argmin over K costs and the Savitzky-Golay
smoothing now runs on the device too (smoothing.py).
"""
import math
import time
from copy import deepcopy
from math import ceil
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import envs as _envs
from .costs import CostCollision, CostComposite, CostConstraint, CostGPTrajectory
from .datasets import TrajectoryDataset
from .diffusion import (GaussianDiffusionModel, apply_cross_conditioning, apply_hard_conditioning, cross_cond_entries,
                        ddpm_sample_fn, lower_for_step, run_ensemble_chain_native,
                        guide_gradient_steps, lower_for_step, _hard_rows)
from .guides import GuideManagerTrajectoriesWithVelocity
from .tasks import PlanningTask, RobotPlanarDisk
from .unet import TemporalUnet, UNET_DIM_MULTS


class PlannerOutput:  # common.py:26-46
    def __init__(self):
        self.trajs_iters = None
        self.trajs_final = None
        self.trajs_final_coll = None
        self.trajs_final_coll_idxs = None
        self.trajs_final_free = None
        self.trajs_final_free_idxs = None
        self.success_free_trajs = None
        self.fraction_free_trajs = None
        self.collision_intensity_trajs = None
        self.idx_best_traj = None
        self.traj_final_free_best = None
        self.cost_best_free_traj = None
        self.cost_smoothness = None
        self.cost_path_length = None
        self.cost_all = None
        self.variance_waypoint_trajs_final_free = None
        self.t_total = None
        self.constraints_l = None


from .smoothing import smooth_trajs  # noqa: E402  (trajectory_utils.py:31-38 on the device: mmdk_smooth_trajs)


# TR/trajectory/metrics.py:7-39 -- planner post-processing statistics over the collision-free trajectories [n, H, D]
def compute_path_length(trajs):
    return torch.linalg.norm(torch.diff(trajs[..., :2], dim=-2), dim=-1).sum(-1)


def compute_smoothness(trajs):
    return torch.linalg.norm(torch.diff(trajs[..., 2:4], dim=-2), dim=-1).sum(-1)


def compute_variance_waypoints(trajs):
    pos = trajs[..., :2].permute(1, 0, 2)                       # [H, n, 2]
    d = torch.triu(torch.cdist(pos, pos, p=2), diagonal=1).flatten(1)   # metrics.py:24-27 keeps the zeros of the lower triangle
    return torch.var(d, dim=1).sum()


def _collision_intensity(task, trajs):
    """tasks.py compute_collision_intensity_trajs: fraction of interpolated waypoints in collision."""
    _, _, _, _, wp = task.get_trajs_collision_and_free(trajs, return_indices=True)
    return wp.float().mean()


class MPD:
    """MPD with the reference call surface (mpd.py:64-88, 306-405).  The reference constructor needs downloaded
    checkpoints/datasets (`trained_models_dir/model_id/{args.yaml,checkpoints}`), which do not ship with it; when
    `model` (a GaussianDiffusionModel) is passed it is used instead, and `model_id` selects the environment
    ('EnvHighways2D-RobotPlanarDisk')."""

    def __init__(self, model_id: str, planner_alg: str = 'mmd', start_state_pos=None, goal_state_pos=None,
                 use_guide_on_extra_objects_only: bool = False, start_guide_steps_fraction: float = 0.5,
                 n_guide_steps: int = 20, n_diffusion_steps_without_noise: int = 1,
                 weight_grad_cost_collision: float = 2e-2, weight_grad_cost_smoothness: float = 8e-2,
                 weight_grad_cost_constraints: float = 2e-1, weight_grad_cost_soft_constraints: float = 2e-2,
                 factor_num_interpolated_points_for_collision: float = 1.5, trajectory_duration: float = 5.0,
                 device: str = 'cuda', debug: bool = False, seed: int = 18, results_dir: str = 'logs',
                 trained_models_dir: str = None, n_samples: int = 64, n_local_inference_noising_steps: int = 3,
                 n_local_inference_denoising_steps: int = 3, model: GaussianDiffusionModel = None,
                 normalizer_limits=((-1., -1., -2., -2.), (1., 1., 2., 2.)), **kwargs):
        self.constraints = []
        self.weight_grad_cost_constraints = weight_grad_cost_constraints
        self.weight_grad_cost_soft_constraints = weight_grad_cost_soft_constraints
        torch.manual_seed(seed)  # fix_random_seed(seed) (mpd.py:96): noise draws come from the global generator
        dev = torch.device(device)
        tensor_args = {'device': dev, 'dtype': torch.float32}
        if planner_alg not in ('mmd', 'diffusion_prior_then_guide', 'diffusion_prior'):
            raise NotImplementedError
        self.run_prior_only = planner_alg == 'diffusion_prior'
        self.run_prior_then_guidance = planner_alg == 'diffusion_prior_then_guide'
        if use_guide_on_extra_objects_only:
            raise NotImplementedError("every reference env keeps extra objects empty: nothing to guide on")
        env_name = model_id.split('-')[0]
        env = _envs.get_env(env_name + 'ExtraObjects', tensor_args=tensor_args)  # mpd.py:126 use_extra_objects=True
        robot = RobotPlanarDisk(tensor_args=tensor_args)
        task = PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.05,
                            tensor_args=tensor_args)
        dataset = TrajectoryDataset(env, robot, task, *normalizer_limits, tensor_args=tensor_args)
        n_support_points = dataset.n_support_points
        dt = trajectory_duration / n_support_points
        robot.dt = dt
        if model is None:
            model = self._load_reference_checkpoint(trained_models_dir, model_id, dataset, dev)
        model = model.to(dev)
        model.eval()
        model.warmup(horizon=n_support_points, device=dev)
        if start_state_pos is None or goal_state_pos is None:
            raise ValueError("start_state_pos and goal_state_pos are required")
        start_state_pos = torch.as_tensor(start_state_pos, dtype=torch.float32, device=dev)
        goal_state_pos = torch.as_tensor(goal_state_pos, dtype=torch.float32, device=dev)
        hard_conds = dataset.get_hard_conditions(torch.vstack((start_state_pos, goal_state_pos)), normalize=True)

        cost_l, w_l = [], []
        for field in task.get_collision_fields():  # mpd.py:221-235
            cost_l.append(CostCollision(robot, n_support_points, field=field, sigma_coll=1.0, tensor_args=tensor_args))
            w_l.append(weight_grad_cost_collision)
        cost_l.append(CostGPTrajectory(robot, n_support_points, dt, sigma_gp=1.0, tensor_args=tensor_args))
        w_l.append(weight_grad_cost_smoothness)
        composite = CostComposite(robot, n_support_points, cost_l, weights_cost_l=w_l, tensor_args=tensor_args)
        guide = GuideManagerTrajectoriesWithVelocity(
            dataset, composite, clip_grad=True, interpolate_trajectories_for_collision=True,
            num_interpolated_points=ceil(n_support_points * factor_num_interpolated_points_for_collision),
            tensor_args=tensor_args)

        self.start_state_pos, self.goal_state_pos = start_state_pos.clone(), goal_state_pos.clone()
        self.robot, self.task, self.dataset, self.context = robot, task, dataset, None
        self.n_diffusion_steps_without_noise = n_diffusion_steps_without_noise
        self.hard_conds, self.model, self.guide = hard_conds, model, guide
        self.n_support_points = n_support_points
        self.t_start_guide = ceil(start_guide_steps_fraction * model.n_diffusion_steps)
        self.n_guide_steps = n_guide_steps
        self.tensor_args = tensor_args
        self.num_samples = n_samples
        self.n_local_inference_noising_steps = n_local_inference_noising_steps
        self.n_local_inference_denoising_steps = n_local_inference_denoising_steps
        self.results_dir = results_dir
        self.recent_call_data = PlannerOutput()
        self.sample_fn_kwargs = dict(
            guide=None if self.run_prior_then_guidance or self.run_prior_only else self.guide,
            n_guide_steps=self.n_guide_steps, t_start_guide=self.t_start_guide,
            noise_std_extra_schedule_fn=lambda x: 0.5)

    @staticmethod
    def _load_reference_checkpoint(trained_models_dir, model_id, dataset, dev):  # mpd.py:116-171
        import os
        import yaml
        model_dir = os.path.join(trained_models_dir or '', model_id)
        args_path = os.path.join(model_dir, 'args.yaml')
        if not os.path.exists(args_path):
            raise FileNotFoundError(f"{args_path} not found: the reference's checkpoints are downloaded separately "
                                    f"(README.md:82-93); pass model=GaussianDiffusionModel(...) instead")
        args = yaml.safe_load(open(args_path))
        unet = TemporalUnet(state_dim=dataset.state_dim, n_support_points=dataset.n_support_points,
                            unet_input_dim=args['unet_input_dim'], dim_mults=UNET_DIM_MULTS[args['unet_dim_mults_option']])
        model = GaussianDiffusionModel(model=unet, variance_schedule=args['variance_schedule'],
                                       n_diffusion_steps=args['n_diffusion_steps'], predict_epsilon=args['predict_epsilon'])
        ck = 'ema_model_current_state_dict.pth' if args.get('use_ema') else 'model_current_state_dict.pth'
        model.load_state_dict(torch.load(os.path.join(model_dir, 'checkpoints', ck), map_location=dev))
        return model

    # -- mpd.py:306-405 -------------------------------------------------------------------------------------------------
    def __call__(self, start_state_pos, goal_state_pos, constraints_l=None, experience=None, *args, **kwargs):
        if not torch.allclose(torch.as_tensor(start_state_pos).to(self.start_state_pos), self.start_state_pos):
            raise ValueError("The start state is different from the one stored in the planner.")
        if not torch.allclose(torch.as_tensor(goal_state_pos).to(self.goal_state_pos), self.goal_state_pos):
            raise ValueError("The goal state is different from the one stored in the planner.")
        cost_constraints_l = []
        for c in (constraints_l or []):
            cost_constraints_l.append(CostConstraint(self.robot, self.n_support_points, q_l=c.get_q_l(),
                                                     traj_range_l=c.get_t_range_l(), radius_l=c.radius_l,
                                                     is_soft=c.is_soft, tensor_args=self.tensor_args))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if experience is None:
            chain, _, _ = self.run_constrained_inference(cost_constraints_l)
        else:
            chain, _, _ = self.run_constrained_local_inference(cost_constraints_l, experience)
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t0

        return self._finish(chain, constraints_l, t_total)

    def _finish(self, chain, constraints_l, t_total):
        """mpd.py:352-405: unnormalise the chain, classify, pick the best free trajectory, smooth."""
        trajs_iters = self.dataset.unnormalize_trajectories(chain)
        trajs_final = trajs_iters[-1]
        # ONE classification launch serves the free / colliding split, the per-trajectory cost (path length + smoothness do
        # not depend on the other trajectories) and the collision intensity (tasks.py:236-311, mpd.py:356-382)
        free_mask, cost_every, wp = self.task.classify(trajs_final, return_waypoints=True)
        free_b = free_mask.bool()
        free_idxs, coll_idxs = torch.argwhere(free_b), torch.argwhere(~free_b)
        free = trajs_final[free_idxs.squeeze(-1)] if free_idxs.numel() else None
        coll = trajs_final[coll_idxs.squeeze(-1)] if coll_idxs.numel() else None
        out = PlannerOutput()
        n_all = trajs_final.shape[0]
        n_free = 0 if free is None else free.shape[0]
        out.success_free_trajs = 1 if n_free > 0 else 0              # tasks.py compute_success_free_trajs
        out.fraction_free_trajs = n_free / n_all
        out.collision_intensity_trajs = float(wp.float().mean())     # tasks.py compute_collision_intensity_trajs
        if free is not None:
            cost_all = cost_every[free_idxs.squeeze(-1)]
            idx_best_free = int(torch.argmin(cost_all))
            out.idx_best_traj = free_idxs[idx_best_free]
            out.idx_best_free_traj = idx_best_free
            out.cost_best_free_traj = float(cost_all[idx_best_free])
            out.cost_all = cost_all
            out.cost_smoothness = compute_smoothness(free)
            out.cost_path_length = compute_path_length(free)
            out.variance_waypoint_trajs_final_free = compute_variance_waypoints(free) if n_free > 1 else torch.zeros((), device=free.device)
            out.traj_final_free_best = free[idx_best_free]
        out.trajs_iters, out.trajs_final = trajs_iters, trajs_final
        out.trajs_final_coll, out.trajs_final_coll_idxs = coll, coll_idxs
        out.trajs_final_free, out.trajs_final_free_idxs = free, free_idxs
        out.t_total = t_total
        out.constraints_l = constraints_l
        if out.trajs_final is not None:
            out.trajs_final = smooth_trajs(out.trajs_final)
        self.recent_call_data = out
        return out

    def _weights(self, cost_constraints_l):
        return [self.weight_grad_cost_soft_constraints if c.is_soft else self.weight_grad_cost_constraints
                for c in cost_constraints_l]

    def run_constrained_inference(self, cost_constraints_l: List[CostConstraint], noise=None):  # mpd.py:407-458
        """Returns (chain [steps + 2, B, H, D] normalised, t_model_sampling, t_post_diffusion_guide) like the reference."""
        self.guide.add_extra_costs(cost_constraints_l, self._weights(cost_constraints_l))
        t_post = 0.0
        t0 = time.perf_counter()
        try:
            chain = self.model.run_inference(self.context, self.hard_conds, n_samples=self.num_samples,
                                             horizon=self.n_support_points, return_chain=True, sample_fn=ddpm_sample_fn,
                                             **self.sample_fn_kwargs,
                                             n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise,
                                             noise=noise)
            t_sampling = time.perf_counter() - t0
            if self.run_prior_then_guidance:
                t1 = time.perf_counter()
                chain = self._post_guide(chain)
                t_post = time.perf_counter() - t1
        finally:
            self.guide.reset_extra_costs()
        return chain, t_sampling, t_post

    def _post_guide(self, chain):
        """mpd.py:432-452: (t_start_guide + n_without_noise) * n_guide_steps extra guide steps on the final sample."""
        n_post = (self.t_start_guide + self.n_diffusion_steps_without_noise) * self.n_guide_steps
        trajs, post = chain[-1], []
        for _ in range(n_post):
            trajs = guide_gradient_steps(trajs, hard_conds=self.hard_conds, guide=self.guide, n_guide_steps=1)
            post.append(trajs)
        return torch.cat((chain, torch.stack(post, 0)))

    def run_constrained_local_inference(self, cost_constraints_l, experience, noise=None):  # mpd.py:460-520
        self.guide.add_extra_costs(cost_constraints_l, self._weights(cost_constraints_l))
        t_post = 0.0
        t0 = time.perf_counter()
        try:
            chain = self.model.run_local_inference(
                experience.path_b, self.n_local_inference_noising_steps, self.n_local_inference_denoising_steps,
                self.context, self.hard_conds, n_samples=self.num_samples, horizon=self.n_support_points,
                return_chain=True, sample_fn=ddpm_sample_fn, **self.sample_fn_kwargs,
                n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise, noise=noise)
            t_sampling = time.perf_counter() - t0
            if self.run_prior_then_guidance:   # mpd.py:494-514: the local variant runs the same post-guide loop
                t1 = time.perf_counter()
                chain = self._post_guide(chain)
                t_post = time.perf_counter() - t1
        finally:
            self.guide.reset_extra_costs()
        return chain, t_sampling, t_post


def plan_batch(planners: List["MPD"], constraints_l_l: Optional[List] = None, rng: str = "sequential",
               starts_goals: Optional[torch.Tensor] = None) -> List[PlannerOutput]:
    """Serves R planner calls (cbs.py:316-324 root loop; :390-430 the two children of an expansion) as ONE batched chain:
    planner r is group r of MultiRobotSampler(mode="independent"), i.e. its own hard conditions, constraint set and clip
    decision -- arithmetically the same as calling the planners one after the other, and the noise is drawn in the same order
    (per planner: x_T, then one frame per reverse step), so the results are bit-identical to sequential calls from the same
    torch RNG state.  The planners must share the diffusion model, the environment and the guide weights.
    MPDEnsemble planners (multi-tile, BASELINE config 5) are served the same way by _plan_batch_ensemble.
    rng="batched" draws all noise in one call per tile instead (faster on the host, not the sequential calls' draw order).
    starts_goals (optional, [R, 2, 2] on the device): the (start, goal) positions the caller plans for, checked against the
    planners' stored states exactly as every planner's __call__ checks its arguments (mpd.py:308-315)."""
    from .sampler import MultiRobotSampler
    p0 = planners[0]
    R = len(planners)
    if constraints_l_l is None:
        constraints_l_l = [None] * R
    if starts_goals is not None:
        stored = torch.stack([torch.stack((p.start_state_pos, p.goal_state_pos)) for p in planners])
        if not torch.allclose(starts_goals.to(stored), stored):
            raise ValueError("A start or goal state is different from the one stored in its planner.")
    if isinstance(p0, MPDEnsemble):
        return _plan_batch_ensemble(planners, constraints_l_l, rng=rng)
    for p in planners:
        if p.model is not p0.model or p.num_samples != p0.num_samples or p.run_prior_only or p.run_prior_then_guidance:
            raise ValueError("plan_batch needs planners that share one model / n_samples and run the 'mmd' algorithm")
    K, H, D = p0.num_samples, p0.n_support_points, p0.model.state_dim
    dev = p0.tensor_args['device']
    n_steps = p0.model.n_diffusion_steps + p0.n_diffusion_steps_without_noise
    noise = torch.empty(R, n_steps + 1, K, H, D, device=dev)
    if rng == "batched":
        noise.normal_()
    else:
        for r in range(R):   # the draw order of R sequential run_inference calls (diffusion.py run_inference): x_T, then one
            for k in range(n_steps + 1):   # frame per step -- drawn straight into the batch buffer (same Philox offsets, no copies)
                torch.randn((K, H, D), out=noise[r, k])
    cons, cons_objs = [], []
    for p, cl in zip(planners, constraints_l_l):
        ccs = [CostConstraint(p.robot, H, q_l=c.get_q_l(), traj_range_l=c.get_t_range_l(), radius_l=c.radius_l,
                              is_soft=c.is_soft, tensor_args=p.tensor_args) for c in (cl or [])]
        cons.append((ccs, p._weights(ccs)))
        cons_objs.append(cl)
    key = (id(p0.model), K)
    smp = _batch_samplers.get(key)
    if smp is None:
        smp = _batch_samplers[key] = MultiRobotSampler(p0.model, p0.guide, n_guide_steps=p0.n_guide_steps,
                                                       t_start_guide=p0.t_start_guide, noise_std=0.5,
                                                       n_diffusion_steps_without_noise=p0.n_diffusion_steps_without_noise)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev = _chain_events()
    ev[0].record()
    _, chains = smp.sample([_hard_rows(p.hard_conds) for p in planners], K, noise=noise, mode="independent",
                           constraints_l=cons, return_chain=True)
    ev[1].record()
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    outs = []
    for r, p in enumerate(planners):
        out = p._finish(chains[r], cons_objs[r], t_total)
        p.recent_call_data = out
        outs.append(out)
    return outs


_batch_samplers = {}
_chain_ev = []


def _chain_events():
    if not _chain_ev:
        _chain_ev.extend([torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)])
    return _chain_ev


def last_batch_chain_ms():
    """Device time (CUDA events on the launching stream) of the sampling chain of the last plan_batch call."""
    torch.cuda.synchronize()
    return _chain_ev[0].elapsed_time(_chain_ev[1])


class DiffusionsEnsemble(nn.Module):
    """diffusion_ensemble.py:25-312: one GaussianDiffusionModel per tile, stepped tile after tile inside every reverse
    step and stitched by apply_cross_conditioning."""

    def __init__(self, models: Dict[int, GaussianDiffusionModel], transforms: Dict[int, torch.Tensor],
                 context_model=None, **kwargs):
        super().__init__()
        self.models = models
        assert len(set(m.n_diffusion_steps for m in models.values())) == 1
        self.n_diffusion_steps = models[0].n_diffusion_steps
        assert len(set(m.predict_epsilon for m in models.values())) == 1
        self.transforms = transforms
        self.context_model = context_model

    @torch.no_grad()
    def p_sample_loop(self, shape, hard_conds, cross_conds, n_diffusion_steps=None, contexts=None, return_chain=False,
                      sample_fn=ddpm_sample_fn, n_diffusion_steps_without_noise=0, warm_start_path_b=None, noise=None,
                      **sample_kwargs):
        """diffusion_ensemble.py:56-106.  noise (optional): {m: [steps + 1, B, H, D]}."""
        dev = self.models[0].betas.device
        Hh = shape[1]
        x = {}
        for m in self.models:
            if warm_start_path_b is not None:   # diffusion_ensemble.py:69-72: tile m's slice of the seed, moved into its frame
                x[m] = warm_start_path_b[:, m * Hh:(m + 1) * Hh, :].to(dev).clone().contiguous()
                x[m][:, :, :2] -= self.transforms[m].to(dev)
            else:
                x[m] = (noise[m][0].to(dev).clone() if noise is not None else torch.randn(shape, device=dev)).contiguous()
            x[m] = apply_hard_conditioning(x[m], _hard_rows(hard_conds.setdefault(m, {})))
        x = apply_cross_conditioning(x, cross_conds, self.transforms)
        # the whole multi-tile loop (:78-106) as ONE native call: per step, per tile, UNet forward -> fused posterior / guide /
        # noise step (+ the tile's hard conditions) -> cross conditions; chain frames after the last tile of each step
        if sample_fn is not ddpm_sample_fn:
            raise NotImplementedError("only ddpm_sample_fn is lowered (mpd_ensemble.py:527)")
        steps_i = list(reversed(range(-n_diffusion_steps_without_noise, n_diffusion_steps)))
        n_steps, (B, _, D) = len(steps_i), shape
        tiles = list(self.models.keys())
        nz = {m: (noise[m][1:1 + n_steps].to(dev).to(torch.float32).contiguous() if noise is not None
                  else torch.empty(n_steps, B, Hh, D, device=dev)) for m in tiles}
        if noise is None:   # the reference's draw order: per step, per tile, one randn_like (:89-95)
            for k in range(n_steps):
                for m in tiles:
                    nz[m][k] = torch.randn_like(x[m])
        lowered, step_list, eps = {}, {}, {}
        for m in tiles:
            kw = dict(sample_kwargs['sample_kwargs'][m])
            if kw.get('scale_grad_by_std'):
                raise NotImplementedError("scale_grad_by_std=True is never used by the planners (sample_functions.py:45)")
            if not self.models[m].clip_denoised:
                raise RuntimeError("clip_denoised=False is rejected by the reference too (diffusion_model_base.py:157)")
            guide, fn = kw.get('guide'), kw.get('noise_std_extra_schedule_fn')
            n_g, t_sg = kw.get('n_guide_steps', 1), kw.get('t_start_guide', torch.inf)
            hard_rows = _hard_rows(hard_conds[m])
            lowered[m] = lower_for_step(guide, 1, B, Hh, dev, [hard_rows], [guide._own_constraints()] if guide is not None else None)
            step_list[m] = [(max(i, 0), self.models[m].step_scalars(i, n_g if (guide is not None and i < t_sg) else 0,
                                                                     1.0 if fn is None else float(fn(i)), True)) for i in steps_i]
            eps[m] = torch.empty_like(x[m])
        chain = {m: torch.empty(n_steps + 1, B, Hh, D, device=dev) for m in tiles} if return_chain else None
        if return_chain:
            for m in tiles:
                chain[m][0].copy_(x[m])
        run_ensemble_chain_native(self.models, lowered, step_list, x, eps, nz, {m: chain[m][1:] for m in tiles} if return_chain else None,
                                  cross_cond_entries(cross_conds, self.transforms, D, 0, B),
                                  chain_init={m: chain[m][0] for m in tiles} if return_chain else None)
        if return_chain:
            return x, {m: chain[m].transpose(0, 1) for m in tiles}   # [B, steps + 1, H, D] like torch.stack(chain, dim=1)
        return x

    @torch.no_grad()
    def run_inference(self, contexts=None, hard_conds: Dict[int, dict] = None, cross_conds=None, n_samples: int = 1,
                      return_chain: bool = False, **diffusion_kwargs):  # diffusion_ensemble.py:224-268
        hard_conds = deepcopy(hard_conds)
        for m, c in hard_conds.items():
            for k, v in c.items():
                hard_conds[m][k] = v.reshape(1, -1).repeat(n_samples, 1)
        H, D = self.models[0].model.n_support_points, self.models[0].state_dim
        out = self.p_sample_loop((n_samples, H, D), hard_conds, deepcopy(cross_conds),
                                 n_diffusion_steps=self.n_diffusion_steps, return_chain=True, **diffusion_kwargs)
        _, chains = out
        chains = {m: c.transpose(0, 1) for m, c in chains.items()}
        return chains if return_chain else {m: c[-1] for m, c in chains.items()}


    @torch.no_grad()
    def run_local_inference(self, seed_trajectory_b, n_noising_steps, n_denoising_steps, contexts=None,
                            hard_conds: Dict[int, dict] = None, cross_conds=None, n_samples: int = 1,
                            return_chain: bool = False, q_noise=None, **diffusion_kwargs):  # diffusion_ensemble.py:270-312
        """seed_trajectory_b [B, n_tiles * H, D]: the whole multi-tile path; noised ONCE by tile 0's schedule (the
        reference calls models[0].q_sample on the concatenated path), then every tile denoises its slice for
        n_denoising_steps steps.  q_noise (optional): the q_sample draw, for parity tests."""
        hard_conds = deepcopy(hard_conds)
        for m, c in hard_conds.items():
            for k, v in c.items():
                hard_conds[m][k] = v.reshape(1, -1).repeat(n_samples, 1)
        diffusion_kwargs.pop('horizon', None)
        dev = self.models[0].betas.device
        seed = seed_trajectory_b.to(dev).to(torch.float32).contiguous()
        if n_noising_steps is None:
            noised = None
        else:
            t = torch.full((seed.shape[0],), int(n_noising_steps), dtype=torch.long)
            noised = self.models[0].q_sample(seed, t, noise=q_noise)
        H, D = self.models[0].model.n_support_points, self.models[0].state_dim
        _, chains = self.p_sample_loop((n_samples, H, D), hard_conds, deepcopy(cross_conds), n_diffusion_steps=n_denoising_steps,
                                       return_chain=True, warm_start_path_b=noised, **diffusion_kwargs)
        chains = {m: c.transpose(0, 1) for m, c in chains.items()}
        return chains if return_chain else {m: c[-1] for m, c in chains.items()}


class PlanningTaskEnsemble:
    """TR/tasks/tasks_ensemble.py:9-360, the parts the planner uses: tile frames (transform_q / inverse_transform_q),
    waypoint-index -> tile (infer_task_id_from_q_idx), per-tile classification and the combination of the tile results."""

    def __init__(self, tasks: Dict[int, PlanningTask], transforms: Dict[int, torch.Tensor], horizon=64, tensor_args=None):
        self.tasks, self.transforms, self.horizon = tasks, transforms, horizon
        self.robot = tasks[0].robot
        self.tensor_args = tensor_args or tasks[0].tensor_args

    def _pad(self, task_id, q):
        tr = self.transforms[task_id].to(q.device)
        if q.shape[-1] > tr.shape[0]:
            tr = torch.cat([tr, torch.zeros(q.shape[-1] - tr.shape[0], device=q.device)])
        return tr

    def transform_q(self, task_id, q):
        return q + self._pad(task_id, q)

    def inverse_transform_q(self, task_id, q):
        return q - self._pad(task_id, q)

    def infer_task_id_from_q_idx(self, q_idx):
        task_id = int(q_idx // self.horizon)
        return task_id, self.tasks[task_id]

    def get_traj_unnormalized(self, task_id, datasets, traj_normalized):
        trajs_iters = datasets[task_id].unnormalize_trajectories(traj_normalized)
        trajs_final = trajs_iters[-1]
        coll, coll_idxs, free, free_idxs, _ = self.tasks[task_id].get_trajs_collision_and_free(trajs_final, return_indices=True)
        return trajs_iters, trajs_final, coll, coll_idxs, free, free_idxs

    def combine_trajs(self, results_ensemble: Dict[int, dict]):
        """tasks_ensemble.py:162-238: tile chains moved to the global frame and concatenated along the horizon; a sample is
        free iff it is free in every tile; costs are recomputed on the combined free trajectories."""
        chains = [self.transform_q(m, r['trajs_iters']) for m, r in results_ensemble.items()]
        res = {'trajs_iters': torch.cat(chains, dim=-2)}
        trajs_final = res['trajs_iters'][-1]
        B = trajs_final.shape[0]
        coll_mask = torch.zeros(B, dtype=torch.bool, device=trajs_final.device)
        for r in results_ensemble.values():
            if r['trajs_final_coll_idxs'] is not None and r['trajs_final_coll_idxs'].numel():
                coll_mask[r['trajs_final_coll_idxs'].reshape(-1)] = True
        coll_ids, free_ids = torch.nonzero(coll_mask).reshape(-1), torch.nonzero(~coll_mask).reshape(-1)
        res['trajs_final_coll_idxs'], res['trajs_final_free_idxs'] = coll_ids, free_ids
        res['trajs_final_coll'] = trajs_final[coll_ids] if coll_ids.numel() else None
        res['trajs_final_free'] = trajs_final[free_ids] if free_ids.numel() else None
        res['success_free_trajs'] = 1 if free_ids.numel() else 0
        res['fraction_free_trajs'] = free_ids.numel() / B
        res['collision_intensity_trajs'] = 1 - res['fraction_free_trajs']
        if free_ids.numel():
            free = res['trajs_final_free']
            res['cost_smoothness_trajs_final_free'] = compute_smoothness(free)
            res['cost_path_length_trajs_final_free'] = compute_path_length(free)
            res['cost_all_trajs_final_free'] = res['cost_smoothness_trajs_final_free'] + res['cost_path_length_trajs_final_free']
            res['variance_waypoint_trajs_final_free'] = compute_variance_waypoints(free)
            best = torch.argmin(res['cost_all_trajs_final_free'], dim=-1)
            res['idx_best_traj'] = free_ids[best]
            res['traj_final_free_best'] = free[best]
            res['cost_best_free_traj'] = torch.min(res['cost_all_trajs_final_free'], dim=-1)[0]
        res['t_total'] = next(iter(results_ensemble.values()))['t_total']
        return res


class MPDEnsemble:
    """MPDEnsemble with the reference call surface (mpd_ensemble.py:65-640): one diffusion model + guide per tile, hard
    conditions on the first / last tile, cross conditions stitching tile i's last waypoint to tile i + 1's first, constraints
    split per tile (waypoint ranges shifted by task_id * H, positions by the tile transform).

    As for MPD, the reference constructor needs downloaded checkpoints; `models` ({tile: GaussianDiffusionModel}) may be
    passed instead and `model_ids[j]` then only names tile j's environment ('EnvEmpty2D-RobotPlanarDisk')."""

    def __init__(self, model_ids: tuple, transforms: Dict[int, torch.Tensor], planner_alg: str = 'mmd',
                 start_state_pos=None, goal_state_pos=None, use_guide_on_extra_objects_only: bool = False,
                 start_guide_steps_fraction: float = 0.5, n_guide_steps: int = 20, n_diffusion_steps_without_noise: int = 1,
                 weight_grad_cost_collision: float = 2e-2, weight_grad_cost_smoothness: float = 8e-2,
                 weight_grad_cost_constraints: float = 2e-1, weight_grad_cost_soft_constraints: float = 2e-2,
                 factor_num_interpolated_points_for_collision: float = 1.5, trajectory_duration: float = 5.0,
                 device: str = 'cuda', debug: bool = False, seed: int = 18, results_dir: str = 'logs',
                 trained_models_dir: str = None, n_samples: int = 64, n_local_inference_noising_steps: int = 3,
                 n_local_inference_denoising_steps: int = 3, models: Dict[int, GaussianDiffusionModel] = None,
                 normalizer_limits=((-1., -1., -2., -2.), (1., 1., 2., 2.)), **kwargs):
        self.constraints = []
        self.weight_grad_cost_constraints = weight_grad_cost_constraints
        self.weight_grad_cost_soft_constraints = weight_grad_cost_soft_constraints
        torch.manual_seed(seed)
        dev = torch.device(device)
        tensor_args = {'device': dev, 'dtype': torch.float32}
        if planner_alg not in ('mmd', 'diffusion_prior_then_guide', 'diffusion_prior'):
            raise NotImplementedError
        self.run_prior_only = planner_alg == 'diffusion_prior'
        self.run_prior_then_guidance = planner_alg == 'diffusion_prior_then_guide'
        if use_guide_on_extra_objects_only:
            raise NotImplementedError("every reference env keeps extra objects empty: nothing to guide on")
        self.models, self.guides, tasks, datasets, sample_kwargs = {}, {}, {}, [], []
        for j, model_id in enumerate(model_ids):
            env = _envs.get_env(model_id.split('-')[0] + 'ExtraObjects', tensor_args=tensor_args)
            robot = RobotPlanarDisk(tensor_args=tensor_args)
            task = PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.01,   # mpd_ensemble.py:139
                                tensor_args=tensor_args)
            dataset = TrajectoryDataset(env, robot, task, *normalizer_limits, tensor_args=tensor_args)
            datasets.append(dataset)
            n_support_points = dataset.n_support_points
            dt = trajectory_duration / n_support_points
            robot.dt = dt
            if j == 0:
                self.robot = robot
            model = models[j] if models is not None else MPD._load_reference_checkpoint(trained_models_dir, model_id, dataset, dev)
            model = model.to(dev)
            model.eval()
            model.warmup(horizon=n_support_points, device=dev)
            self.models[j], tasks[j] = model, task
            cost_l, w_l = [], []
            for field in task.get_collision_fields():
                cost_l.append(CostCollision(robot, n_support_points, field=field, sigma_coll=1.0, tensor_args=tensor_args))
                w_l.append(weight_grad_cost_collision)
            cost_l.append(CostGPTrajectory(robot, n_support_points, dt, sigma_gp=1.0, tensor_args=tensor_args))
            w_l.append(weight_grad_cost_smoothness)
            composite = CostComposite(robot, n_support_points, cost_l, weights_cost_l=w_l, tensor_args=tensor_args)
            guide = GuideManagerTrajectoriesWithVelocity(
                dataset, composite, clip_grad=True, interpolate_trajectories_for_collision=True,
                num_interpolated_points=ceil(n_support_points * factor_num_interpolated_points_for_collision),
                tensor_args=tensor_args)
            self.guides[j] = guide
            t_start_guide = ceil(start_guide_steps_fraction * model.n_diffusion_steps)
            sample_kwargs.append(dict(guide=None if self.run_prior_then_guidance or self.run_prior_only else guide,
                                      n_guide_steps=n_guide_steps, t_start_guide=t_start_guide,
                                      noise_std_extra_schedule_fn=lambda x: 0.5))
        transforms = {k: torch.as_tensor(v, dtype=torch.float32, device=dev) for k, v in transforms.items()}
        self.task = PlanningTaskEnsemble(tasks, transforms, horizon=n_support_points, tensor_args=tensor_args)
        if start_state_pos is None or goal_state_pos is None:
            raise ValueError("start_state_pos and goal_state_pos are required")
        start_state_pos = torch.as_tensor(start_state_pos, dtype=torch.float32, device=dev)
        goal_state_pos = torch.as_tensor(goal_state_pos, dtype=torch.float32, device=dev)
        last = len(model_ids) - 1
        start_local = self.task.inverse_transform_q(0, start_state_pos)       # global frame -> tile frames
        goal_local = self.task.inverse_transform_q(last, goal_state_pos)
        hard_conds = {0: datasets[0].get_single_pt_hard_conditions(start_local, 0, True)}
        goal_hc = datasets[last].get_single_pt_hard_conditions(goal_local, -1, True)
        if last in hard_conds:
            hard_conds[last].update(goal_hc)
        else:
            hard_conds[last] = goal_hc
        self.transforms = transforms
        self.model = DiffusionsEnsemble(self.models, transforms)
        self.cross_conds = {(i, i + 1): (n_support_points - 1, 0) for i in range(last)}
        self.start_state_pos, self.goal_state_pos = start_state_pos.clone(), goal_state_pos.clone()
        self.sample_kwargs, self.contexts = sample_kwargs, None
        self.n_diffusion_steps_without_noise = n_diffusion_steps_without_noise
        self.hard_conds = hard_conds
        self.n_support_points, self.t_start_guide, self.n_guide_steps = n_support_points, t_start_guide, n_guide_steps
        self.tensor_args = tensor_args
        self.num_samples = n_samples
        self.n_local_inference_noising_steps = n_local_inference_noising_steps
        self.n_local_inference_denoising_steps = n_local_inference_denoising_steps
        self.datasets = datasets
        self.results_dir = results_dir
        self.recent_call_data = PlannerOutput()

    # -- mpd_ensemble.py:335-429 ---------------------------------------------------------------------------------------
    def __call__(self, start_state_pos, goal_state_pos, constraints_l=None, experience=None, *args, **kwargs):
        if not torch.allclose(torch.as_tensor(start_state_pos).to(self.start_state_pos), self.start_state_pos):
            raise ValueError("The start state is different from the one stored in the planner.")
        if not torch.allclose(torch.as_tensor(goal_state_pos).to(self.goal_state_pos), self.goal_state_pos):
            raise ValueError("The goal state is different from the one stored in the planner.")
        cost_constraints_l = [CostConstraint(self.robot, self.n_support_points, q_l=c.get_q_l(), traj_range_l=c.get_t_range_l(),
                                             radius_l=c.radius_l, is_soft=c.is_soft, tensor_args=self.tensor_args)
                              for c in (constraints_l or [])]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if experience is None:
            chains, _, _ = self.run_constrained_inference(cost_constraints_l)
        else:
            chains, _, _ = self.run_constrained_local_inference(cost_constraints_l, experience)
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t0
        return self._finish(chains, constraints_l, t_total)

    def _finish(self, chains, constraints_l, t_total):
        """mpd_ensemble.py:352-429: per tile unnormalise + classify, stitch the tiles in the global frame, statistics, smoothing."""
        per_tile = {}
        for m in self.models:
            iters, final, coll, coll_idxs, free, free_idxs = self.task.get_traj_unnormalized(m, self.datasets, chains[m])
            per_tile[m] = dict(trajs_iters=iters, trajs_final=final, trajs_final_coll=coll, trajs_final_coll_idxs=coll_idxs,
                               trajs_final_free=free, trajs_final_free_idxs=free_idxs, t_total=t_total)
        res = self.task.combine_trajs(per_tile)
        out = PlannerOutput()
        out.trajs_iters, out.trajs_final = res['trajs_iters'], res['trajs_iters'][-1]
        out.trajs_final_coll, out.trajs_final_coll_idxs = res['trajs_final_coll'], res['trajs_final_coll_idxs']
        out.trajs_final_free, out.trajs_final_free_idxs = res['trajs_final_free'], res['trajs_final_free_idxs']
        out.success_free_trajs, out.fraction_free_trajs = res['success_free_trajs'], res['fraction_free_trajs']
        out.collision_intensity_trajs = res['collision_intensity_trajs']
        if out.success_free_trajs:
            out.idx_best_traj, out.traj_final_free_best = res['idx_best_traj'], res['traj_final_free_best']
            out.cost_best_free_traj = res['cost_best_free_traj']
            out.cost_smoothness, out.cost_path_length = res['cost_smoothness_trajs_final_free'], res['cost_path_length_trajs_final_free']
            out.cost_all = res['cost_all_trajs_final_free']
            out.variance_waypoint_trajs_final_free = res['variance_waypoint_trajs_final_free']
        out.t_total, out.constraints_l = t_total, constraints_l
        if out.trajs_final is not None:
            out.trajs_final = smooth_trajs(out.trajs_final)
        self.recent_call_data = out
        return out

    def split_cost_constraints_to_tasks(self, cost_constraints_l: List[CostConstraint]):  # mpd_ensemble.py:431-507
        """{tile: [CostConstraint]}: every (q, range, radius) entry goes to the tile its FIRST waypoint index falls in
        (ranges are not broken at tile borders, as in the reference); one hard and one soft object per tile."""
        buckets = {}
        for c in cost_constraints_l:
            for j in range(len(c.qs)):
                task_id, _ = self.task.infer_task_id_from_q_idx(int(c.traj_ranges[j][0]))
                buckets.setdefault((task_id, bool(c.is_soft)), []).append((c.qs[j], c.traj_ranges[j], c.radii[j]))
        out = {}
        for soft in (False, True):   # hard objects first, then soft (reference order)
            for (task_id, is_soft), items in buckets.items():
                if is_soft != soft:
                    continue
                q_l, r_l, rad_l = zip(*items)
                cc = CostConstraint(self.robot, self.n_support_points, q_l=list(q_l),
                                    traj_range_l=[(int(r[0]), int(r[1])) for r in r_l], radius_l=[float(r) for r in rad_l],
                                    is_soft=soft, tensor_args=self.tensor_args)
                out.setdefault(task_id, []).append(cc)
        return out

    def _install_constraints(self, cost_constraints_l):
        split = self.split_cost_constraints_to_tasks(cost_constraints_l)
        for task_id, ccs in split.items():
            for cc in ccs:
                cc.traj_ranges = cc.traj_ranges - task_id * self.n_support_points      # mpd_ensemble.py:516-517
                cc.qs = cc.qs - self.transforms[task_id].to(cc.qs.device)
                self.guides[task_id].add_extra_costs([cc], [self.weight_grad_cost_soft_constraints if cc.is_soft
                                                            else self.weight_grad_cost_constraints])
        return split

    def _post_guide(self, chains):
        n_post = (self.t_start_guide + self.n_diffusion_steps_without_noise) * self.n_guide_steps
        for task_id in self.guides:
            trajs, post = chains[task_id][-1], []
            for _ in range(n_post):
                trajs = guide_gradient_steps(trajs, hard_conds=self.hard_conds.get(task_id, {}), guide=self.guides[task_id],
                                             n_guide_steps=1)
                post.append(trajs)
            chains[task_id] = torch.cat((chains[task_id], torch.stack(post, 0)))
        return chains

    def run_constrained_inference(self, cost_constraints_l, noise=None):  # mpd_ensemble.py:509-571
        split = self._install_constraints(cost_constraints_l)
        t_post = 0.0
        t0 = time.perf_counter()
        try:
            chains = self.model.run_inference(self.contexts, self.hard_conds, cross_conds=self.cross_conds,
                                              n_samples=self.num_samples, return_chain=True, sample_fn=ddpm_sample_fn,
                                              sample_kwargs=self.sample_kwargs,
                                              n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise, noise=noise)
            t_sampling = time.perf_counter() - t0
            if self.run_prior_then_guidance:
                t1 = time.perf_counter()
                chains = self._post_guide(chains)
                t_post = time.perf_counter() - t1
        finally:
            for task_id in split:
                self.guides[task_id].reset_extra_costs()
        return chains, t_sampling, t_post

    def run_constrained_local_inference(self, cost_constraints_l, experience, noise=None, q_noise=None):  # :573-640
        split = self._install_constraints(cost_constraints_l)
        t_post = 0.0
        t0 = time.perf_counter()
        try:
            chains = self.model.run_local_inference(
                experience.path_b, self.n_local_inference_noising_steps, self.n_local_inference_denoising_steps, self.contexts,
                self.hard_conds, cross_conds=self.cross_conds, n_samples=self.num_samples, horizon=self.n_support_points,
                return_chain=True, sample_fn=ddpm_sample_fn, sample_kwargs=self.sample_kwargs,
                n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise, noise=noise, q_noise=q_noise)
            t_sampling = time.perf_counter() - t0
            if self.run_prior_then_guidance:
                t1 = time.perf_counter()
                chains = self._post_guide(chains)
                t_post = time.perf_counter() - t1
        finally:
            for task_id in split:
                self.guides[task_id].reset_extra_costs()
        return chains, t_sampling, t_post


def _plan_batch_ensemble(planners: List["MPDEnsemble"], constraints_l_l, rng="sequential") -> List[PlannerOutput]:
    """R MPDEnsemble planner calls (one per robot, cbs.py:316-324) as ONE batched multi-tile chain.

    DiffusionsEnsemble.p_sample_loop (diffusion_ensemble.py:56-106) steps tile after tile inside every reverse step; here
    every tile step runs for all R planners at once: planner r is group r of tile m's batch [R * K, H, D] (its own hard
    conditions, split constraints and clip decision), and the cross conditioning runs per set of planners that share the tile
    transforms (robots may traverse the tiles in different orders: inference_multi_agent.py:205-222).  Arithmetic and noise
    draw order (per planner: x_T of every tile, then per step, per tile, one frame) are those of sequential calls, so from the
    same torch RNG state the results are bit-identical to `[p(start, goal, constraints) for p in planners]`."""
    p0 = planners[0]
    R = len(planners)
    tiles = list(p0.models.keys())
    for p in planners:
        if not isinstance(p, MPDEnsemble) or list(p.models.keys()) != tiles or any(p.models[m] is not p0.models[m] for m in tiles):
            raise ValueError("plan_batch needs MPDEnsemble planners that share the per-tile diffusion models")
        if (p.num_samples != p0.num_samples or p.run_prior_only or p.run_prior_then_guidance or p.n_guide_steps != p0.n_guide_steps
                or p.t_start_guide != p0.t_start_guide or p.n_diffusion_steps_without_noise != p0.n_diffusion_steps_without_noise
                or p.cross_conds != p0.cross_conds):
            raise ValueError("plan_batch needs planners with one n_samples / guide schedule running the 'mmd' algorithm")
    K, H = p0.num_samples, p0.n_support_points
    D = p0.models[tiles[0]].state_dim
    dev = p0.tensor_args['device']
    T, n_extra = p0.model.n_diffusion_steps, p0.n_diffusion_steps_without_noise
    n_steps = T + n_extra
    B = R * K
    noise = {m: torch.empty(n_steps + 1, B, H, D, device=dev) for m in tiles}
    if rng == "sequential":
        for r in range(R):   # the draw order of R sequential DiffusionsEnsemble.p_sample_loop calls
            sl = slice(r * K, (r + 1) * K)
            for m in tiles:
                torch.randn((K, H, D), out=noise[m][0, sl])
            for k in range(1, n_steps + 1):
                for m in tiles:
                    torch.randn((K, H, D), out=noise[m][k, sl])
    elif rng == "batched":
        for m in tiles:
            noise[m].normal_()
    else:
        raise ValueError("rng must be 'sequential' or 'batched'")
    # planners that share the tile transforms form one cross-conditioning call; keep the planners' order inside a set
    sets = {}
    for r, p in enumerate(planners):
        sig = tuple(tuple(float(v) for v in p.transforms[m].reshape(-1).tolist()) for m in tiles)
        sets.setdefault(sig, []).append(r)
    order = [r for rs in sets.values() for r in rs]                # batch position -> planner index
    pos = {r: i for i, r in enumerate(order)}
    segs, a = [], 0
    for rs in sets.values():
        segs.append((a * K, (a + len(rs)) * K, planners[rs[0]].transforms))
        a += len(rs)
    if order != list(range(R)):
        for m in tiles:
            noise[m] = noise[m].view(n_steps + 1, R, K, H, D)[:, order].reshape(n_steps + 1, B, H, D).contiguous()
    ordered = [planners[r] for r in order]
    installed = []
    try:
        for r in order:
            p, cl = planners[r], constraints_l_l[r]
            ccs = [CostConstraint(p.robot, H, q_l=c.get_q_l(), traj_range_l=c.get_t_range_l(), radius_l=c.radius_l,
                                  is_soft=c.is_soft, tensor_args=p.tensor_args) for c in (cl or [])]
            installed.append((p, p._install_constraints(ccs)))
        cons = {m: [p.guides[m]._own_constraints() for p in ordered] for m in tiles}
        hcs = {m: [_hard_rows(p.hard_conds.get(m, {})) for p in ordered] for m in tiles}
        x, eps, lowered, step_list = {}, {}, {}, {}
        steps_i = list(reversed(range(-n_extra, T)))
        for m in tiles:
            x[m] = noise[m][0].clone()
            for g, hc in enumerate(hcs[m]):
                for row, val in hc.items():
                    x[m][g * K:(g + 1) * K, row, :] = val
            eps[m] = torch.empty_like(x[m])
            kw = p0.sample_kwargs[m]
            guide, fn = kw['guide'], kw.get('noise_std_extra_schedule_fn')
            lowered[m] = lower_for_step(guide, R, K, H, dev, hcs[m], cons[m] if guide is not None else None)
            step_list[m] = [(max(i, 0), p0.models[m].step_scalars(i, kw['n_guide_steps'] if (guide is not None and i < kw['t_start_guide']) else 0,
                                                                  1.0 if fn is None else float(fn(i)), True)) for i in steps_i]
        cross_entries = []
        for lo, hi, tr in segs:
            apply_cross_conditioning({m: x[m][lo:hi] for m in tiles}, p0.cross_conds, tr)
            cross_entries += cross_cond_entries(p0.cross_conds, tr, D, lo, hi)
        chains = {m: torch.empty(n_steps + 1, B, H, D, device=dev) for m in tiles}
        for m in tiles:
            chains[m][0].copy_(x[m])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = _chain_events()
        ev[0].record()
        # the whole multi-tile chain of all planner calls: one native call (mmdk_run_chain_ensemble)
        run_ensemble_chain_native(p0.models, lowered, step_list, x, eps, {m: noise[m][1:] for m in tiles},
                                  {m: chains[m][1:] for m in tiles}, cross_entries, chain_init={m: chains[m][0] for m in tiles})
        ev[1].record()
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t0
    finally:
        for p, split in installed:
            for task_id in split:
                p.guides[task_id].reset_extra_costs()
    outs = [None] * R
    for r, p in enumerate(planners):
        sl = slice(pos[r] * K, (pos[r] + 1) * K)
        outs[r] = p._finish({m: chains[m][:, sl] for m in tiles}, constraints_l_l[r], t_total)
    return outs


class MultiPointConstraint:
    """mmd/common/constraints.py:46-86: a batch of vertex constraints (centre, inclusive-exclusive waypoint range,
    radius), soft or hard; one object becomes one CostConstraint (mpd.py:331-342)."""

    def __init__(self, q_l, t_range_l, radius_l=None, is_soft=False):
        self.q_l = q_l
        self.t_range_l = t_range_l
        self.radius_l = [2.4 * 0.05] * len(q_l) if radius_l is None else radius_l  # mmd_params.py:52
        self.is_soft = is_soft

    def get_q_l(self):
        return self.q_l

    def get_t_range_l(self):
        return self.t_range_l

    def get_radius_l(self):
        return self.radius_l

    def get_is_soft(self):
        return self.is_soft
