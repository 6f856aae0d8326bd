"""LimitsNormalizer + the slice of TrajectoryDataset the planners use (mmd/datasets/normalization.py:150-168,
mmd/datasets/trajectories.py:200-250).  No dataset files ship with the reference: the normaliser
limits are given explicitly (or read from a checkpoint directory when one exists)."""
import ctypes as C

import torch

from . import _lib


class LimitsNormalizer:
    def __init__(self, mins, maxs):
        self.mins = torch.as_tensor(mins, dtype=torch.float32)
        self.maxs = torch.as_tensor(maxs, dtype=torch.float32)

    def normalize(self, x):  # host-side, tiny tensors (start/goal states)
        mins, maxs = self.mins.to(x.device), self.maxs.to(x.device)
        x = (x - mins) / (maxs - mins)
        return 2 * x - 1

    def _env(self):
        env = _lib.GuideEnv()
        rng = self.maxs - self.mins
        for d in range(4):
            env.norm_min[d] = float(self.mins[d])
            env.norm_range[d] = float(rng[d])
        return env

    def unnormalize(self, x, eps=1e-4):
        """normalization.py:157-168 on the device: clip the WHOLE tensor iff any element leaves [-1-eps, 1+eps]."""
        lib = _lib.lib()
        x = x.contiguous()
        assert x.shape[-1] == 4 and x.dtype == torch.float32
        clip = bool((x.max() > 1 + eps) | (x.min() < -1 - eps))
        out = torch.empty_like(x)
        env = self._env()
        _lib.check(lib.mmdk_unnormalize(C.byref(env), _lib.ptr(x), x.numel() // 4, int(clip), _lib.ptr(out), _lib.stream_ptr()))
        return out


class TrajectoryDataset:
    """Minimal stand-in: normaliser + hard conditions + the env/robot/task triple (trajectories.py:37-101,200-250)."""

    def __init__(self, env, robot, task, mins=(-1., -1., -2., -2.), maxs=(1., 1., 2., 2.), n_support_points=64,
                 tensor_args=None):
        self.env, self.robot, self.task = env, robot, task
        self.normalizer = LimitsNormalizer(mins, maxs)
        self.n_support_points = n_support_points
        self.state_dim = 4
        self.include_velocity = True
        self.field_key_traj = 'traj'
        self.tensor_args = tensor_args
        self.threshold_start_goal_pos = 1.0

    def unnormalize_trajectories(self, x):
        return self.normalizer.unnormalize(x)

    def normalize_trajectories(self, x):
        return self.normalizer.normalize(x)

    def get_hard_conditions(self, traj, horizon=None, normalize=False):  # trajectories.py:216-239
        start = torch.cat((traj[0][..., :2], torch.zeros_like(traj[0][..., :2])), dim=-1)
        goal = torch.cat((traj[-1][..., :2], torch.zeros_like(traj[-1][..., :2])), dim=-1)
        if normalize:
            start, goal = self.normalizer.normalize(start), self.normalizer.normalize(goal)
        horizon = horizon or self.n_support_points
        return {0: start, horizon - 1: goal}

    def get_single_pt_hard_conditions(self, state_position, idx, normalize=False):  # trajectories.py:241-250
        state = torch.cat((state_position[..., :2], torch.zeros_like(state_position[..., :2])), dim=-1)
        if normalize:
            state = self.normalizer.normalize(state)
        return {idx: state}
