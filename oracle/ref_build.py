"""Assemble the REAL reference objects (build container only) exactly as mpd.py:116-304 does, minus the parts that
need downloaded checkpoints/datasets.  Test infrastructure only: used by oracle/gen_golden.py and
tests/test_oracle_vs_reference.py to pin oracle/port.py against the reference itself.
"""
import contextlib
import io
from math import ceil

import torch

from . import ref_shim


class _DatasetStub:
    """Stands in for TrajectoryDataset around a LimitsNormalizer (mmd/datasets/trajectories.py:200-214)."""

    def __init__(self, normalizer):
        self.normalizer = normalizer

    def unnormalize_trajectories(self, x):
        return self.normalizer.unnormalize(x)

    def normalize_trajectories(self, x):
        return self.normalizer.normalize(x)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def build_reference(env_name="EnvHighways2D", n_diffusion_steps=25, unet_params=None, dim_mults=(1, 2, 4),
                    mins=(-1.0, -1.0, -2.0, -2.0), maxs=(1.0, 1.0, 2.0, 2.0), cutoff_margin=0.05,
                    self_attention=False, with_model=True):
    ref_shim.install()
    import torch_robotics.environments as E
    from torch_robotics.robots import RobotPlanarDisk
    from torch_robotics.tasks.tasks import PlanningTask
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite, CostGPTrajectory
    from mmd.models.diffusion_models.guides import GuideManagerTrajectoriesWithVelocity
    from mmd.models.diffusion_models.temporal_unet import TemporalUnet
    from mmd.models.diffusion_models.diffusion_model_base import GaussianDiffusionModel
    from mmd.datasets.normalization import LimitsNormalizer

    ta = {"device": torch.device("cpu"), "dtype": torch.float32}
    with quiet():
        env = getattr(E, env_name + "ExtraObjects")(tensor_args=ta)  # mpd.py:126 use_extra_objects=True
        robot = RobotPlanarDisk(tensor_args=ta)
        task = PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=cutoff_margin, tensor_args=ta)
    n_support_points = 64
    dt = 5.0 / n_support_points  # mpd.py:142, mmd_params.py:45
    robot.dt = dt
    normalizer = LimitsNormalizer(torch.stack((torch.tensor(mins), torch.tensor(maxs))))
    normalizer.mins, normalizer.maxs = torch.tensor(mins), torch.tensor(maxs)
    dataset = _DatasetStub(normalizer)

    costs, weights = [], []
    for field in task.get_collision_fields():  # [None, objects, ws-boundaries]  (mpd.py:221-235)
        costs.append(CostCollision(robot, n_support_points, field=field, sigma_coll=1.0, tensor_args=ta))
        weights.append(2e-2)
    costs.append(CostGPTrajectory(robot, n_support_points, dt, sigma_gp=1.0, tensor_args=ta))
    weights.append(8e-2)
    composite = CostComposite(robot, n_support_points, costs, weights_cost_l=weights, tensor_args=ta)
    guide = GuideManagerTrajectoriesWithVelocity(
        dataset, composite, clip_grad=True, interpolate_trajectories_for_collision=True,
        num_interpolated_points=ceil(n_support_points * 1.5), tensor_args=ta)

    model = None
    if with_model:
        with quiet():
            unet = TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=dim_mults,
                                self_attention=self_attention)
            model = GaussianDiffusionModel(model=unet, variance_schedule="exponential",
                                           n_diffusion_steps=n_diffusion_steps, predict_epsilon=True)
        if unet_params is not None:
            unet.load_state_dict(unet_params, strict=True)
        model.eval()
    return dict(env=env, robot=robot, task=task, dataset=dataset, guide=guide, model=model, tensor_args=ta)


def make_cost_constraint(ref, qs, traj_ranges, radii, is_soft=True):
    from mp_baselines.planners.costs.cost_functions import CostConstraint
    return CostConstraint(ref["robot"], 64, q_l=[q for q in qs], traj_range_l=[tuple(r) for r in traj_ranges.tolist()],
                          radius_l=[float(r) for r in radii], is_soft=is_soft, tensor_args=ref["tensor_args"])


@contextlib.contextmanager
def scripted_noise(noise_list):
    """Feed torch.randn / torch.randn_like from a pre-drawn list, in call order (SURVEY hard part d)."""
    it = iter(noise_list)
    orig_randn, orig_randn_like = torch.randn, torch.randn_like

    def _randn(*a, **k):
        return next(it).clone()

    def _randn_like(x, **k):
        return next(it).clone()

    torch.randn, torch.randn_like = _randn, _randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_randn_like


def reference_run_inference(ref, hard_conds, n_samples, noise, n_guide_steps=20, t_start_guide=None,
                            extra_costs=None, extra_weights=None, n_extra=1):
    """model.run_inference exactly as MPD.run_constrained_inference calls it (mpd.py:407-424)."""
    from mmd.models.diffusion_models.sample_functions import ddpm_sample_fn
    model, guide = ref["model"], ref["guide"]
    if t_start_guide is None:
        t_start_guide = ceil(0.5 * model.n_diffusion_steps)
    if extra_costs:
        guide.add_extra_costs(extra_costs, extra_weights)
    try:
        with scripted_noise([noise[i] for i in range(noise.shape[0])]), quiet():
            chain = model.run_inference(None, hard_conds, n_samples=n_samples, horizon=64, return_chain=True,
                                        sample_fn=ddpm_sample_fn, guide=guide, n_guide_steps=n_guide_steps,
                                        t_start_guide=t_start_guide, noise_std_extra_schedule_fn=lambda x: 0.5,
                                        n_diffusion_steps_without_noise=n_extra)
    finally:
        guide.reset_extra_costs()
    return chain
