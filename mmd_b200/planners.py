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
from .diffusion import (GaussianDiffusionModel, apply_cross_conditioning, apply_hard_conditioning, ddpm_sample_fn,
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
            chain = self.run_constrained_inference(cost_constraints_l)
        else:
            chain = self.run_constrained_local_inference(cost_constraints_l, experience)
        torch.cuda.synchronize()
        t_total = time.perf_counter() - t0

        trajs_iters = self.dataset.unnormalize_trajectories(chain)
        trajs_final = trajs_iters[-1]
        coll, coll_idxs, free, free_idxs, _ = self.task.get_trajs_collision_and_free(trajs_final, return_indices=True)
        out = PlannerOutput()
        if free is not None:
            _, cost_all, _ = self.task.classify(free)  # path length + smoothness (TR/trajectory/metrics.py)
            idx_best_free = int(torch.argmin(cost_all))
            out.idx_best_traj = free_idxs[idx_best_free]
            out.idx_best_free_traj = idx_best_free
            out.cost_best_free_traj = float(cost_all[idx_best_free])
            out.cost_all = cost_all
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
        self.guide.add_extra_costs(cost_constraints_l, self._weights(cost_constraints_l))
        try:
            chain = self.model.run_inference(self.context, self.hard_conds, n_samples=self.num_samples,
                                             horizon=self.n_support_points, return_chain=True, sample_fn=ddpm_sample_fn,
                                             **self.sample_fn_kwargs,
                                             n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise,
                                             noise=noise)
            if self.run_prior_then_guidance:
                n_post = (self.t_start_guide + self.n_diffusion_steps_without_noise) * self.n_guide_steps
                trajs, post = chain[-1], []
                for _ in range(n_post):
                    trajs = guide_gradient_steps(trajs, hard_conds=self.hard_conds, guide=self.guide, n_guide_steps=1)
                    post.append(trajs)
                chain = torch.cat((chain, torch.stack(post, 0)))
        finally:
            self.guide.reset_extra_costs()
        return chain

    def run_constrained_local_inference(self, cost_constraints_l, experience, noise=None):  # mpd.py:460-520
        self.guide.add_extra_costs(cost_constraints_l, self._weights(cost_constraints_l))
        try:
            chain = self.model.run_local_inference(
                experience.path_b, self.n_local_inference_noising_steps, self.n_local_inference_denoising_steps,
                self.context, self.hard_conds, n_samples=self.num_samples, horizon=self.n_support_points,
                return_chain=True, sample_fn=ddpm_sample_fn, **self.sample_fn_kwargs,
                n_diffusion_steps_without_noise=self.n_diffusion_steps_without_noise, noise=noise)
        finally:
            self.guide.reset_extra_costs()
        return chain


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
        if warm_start_path_b is not None:
            raise NotImplementedError("ensemble warm start (run_local_inference) is not lowered yet")
        dev = self.models[0].betas.device
        x = {}
        for m in self.models:
            x[m] = (noise[m][0].to(dev).clone() if noise is not None else torch.randn(shape, device=dev)).contiguous()
            x[m] = apply_hard_conditioning(x[m], _hard_rows(hard_conds.setdefault(m, {})))
        x = apply_cross_conditioning(x, cross_conds, self.transforms)
        chains = {m: [x[m].clone()] for m in self.models} if return_chain else None
        k = 1
        for i in reversed(range(-n_diffusion_steps_without_noise, n_diffusion_steps)):
            t = torch.full((shape[0],), i, dtype=torch.long)
            for m in self.models:
                kw = dict(sample_kwargs['sample_kwargs'][m])
                nz = noise[m][k].to(dev) if noise is not None else None
                x[m], _ = sample_fn(self.models[m], x[m], hard_conds[m], None, t, noise=nz, **kw)
                x[m] = apply_hard_conditioning(x[m], _hard_rows(hard_conds[m]))
                x = apply_cross_conditioning(x, cross_conds, self.transforms)
            if return_chain:
                for m in self.models:
                    chains[m].append(x[m].clone())
            k += 1
        if return_chain:
            return x, {m: torch.stack(v, dim=1) for m, v in chains.items()}
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
