"""Shared builders for the parity tests: the same synthetic problem on the oracle (CPU) and on mmd_b200 (CUDA)."""
import math

import torch

from oracle import port

LIMITS = port.DEFAULT_NORMALIZER_LIMITS


def build_oracle(env_name="EnvHighways2D", T=25, seed=0, out_scale=1.0, dim_mults=(1, 2, 4), cutoff_margin=0.05,
                 w_smooth=8e-2):
    P = port.make_unet_params(seed=seed, out_scale=out_scale, dim_mults=dim_mults)
    sdf, grad = port.build_sdf_grid(env_name)
    norm = port.LimitsNormalizer(*LIMITS)
    guide = port.GuideSpec(port.GridSDF(sdf, grad), norm, cutoff_margin=cutoff_margin, w_smooth=w_smooth)
    model = port.DiffusionModel(P, T)
    return dict(P=P, guide=guide, model=model, norm=norm, sdf=sdf, grad=grad)


def build_product(dev, env_name="EnvHighways2D", T=25, P=None, dim_mults=(1, 2, 4), cutoff_margin=0.05,
                  precision="fp32", w_smooth=8e-2):
    import mmd_b200 as M
    ta = {"device": dev, "dtype": torch.float32}
    env = M.envs.get_env(env_name + "ExtraObjects", tensor_args=ta)
    robot = M.RobotPlanarDisk(tensor_args=ta)
    task = M.PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=cutoff_margin, tensor_args=ta)
    dataset = M.TrajectoryDataset(env, robot, task, *LIMITS, tensor_args=ta)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=dim_mults,
                          unet_precision=precision)
    if P is not None:
        unet.load_state_dict(P, strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T,
                                     predict_epsilon=True).to(dev)
    dt = 5.0 / 64
    costs, weights = [], []
    for field in task.get_collision_fields():  # mpd.py:221-235
        costs.append(M.CostCollision(robot, 64, field=field, sigma_coll=1.0, tensor_args=ta))
        weights.append(2e-2)
    costs.append(M.CostGPTrajectory(robot, 64, dt, sigma_gp=1.0, tensor_args=ta))
    weights.append(w_smooth)
    comp = M.CostComposite(robot, 64, costs, weights_cost_l=weights, tensor_args=ta)
    guide = M.GuideManagerTrajectoriesWithVelocity(dataset, comp, clip_grad=True,
                                                   interpolate_trajectories_for_collision=True,
                                                   num_interpolated_points=math.ceil(64 * 1.5), tensor_args=ta)
    return dict(model=model, unet=unet, guide=guide, task=task, robot=robot, dataset=dataset, env=env, ta=ta)


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a, b):
    return float((a.detach().cpu() - b.detach().cpu()).abs().max())


def random_constraints(n, seed=0, radius=0.12):
    g = torch.Generator().manual_seed(seed)
    qs = torch.rand(n, 2, generator=g) * 2 - 1
    hh = torch.randint(0, 64, (n,), generator=g).float()
    return qs, torch.stack((hh, hh + 1), -1), torch.full((n,), radius)
