"""Cost descriptors with the reference constructors (MPB/planners/costs/cost_functions.py:18-326,505-542).

They carry parameters only; GuideManagerTrajectoriesWithVelocity lowers them (or the reference's own objects, which
expose the same attributes) into the fused CUDA guide.  No `eval`: gradients are closed form on the device.
"""
import torch


class Cost:
    def __init__(self, robot, n_support_points, tensor_args=None, **kwargs):
        self.robot = robot
        self.n_dof = getattr(robot, "q_dim", 2)
        self.dim = 2 * self.n_dof
        self.n_support_points = n_support_points
        self.tensor_args = tensor_args


class CostCollision(Cost):  # cost_functions.py:150-193
    def __init__(self, robot, n_support_points, field=None, sigma_coll=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.field = field
        self.sigma_coll = sigma_coll


class CostGPTrajectory(Cost):  # cost_functions.py:505-542
    def __init__(self, robot, n_support_points, dt, sigma_gp=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.dt = dt
        self.sigma_gp = sigma_gp


class CostComposite(Cost):  # cost_functions.py:57-106
    def __init__(self, robot, n_support_points, cost_list, weights_cost_l=None, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        self.cost_l = cost_list
        self.weight_cost_l = weights_cost_l if weights_cost_l is not None else [1.0] * len(cost_list)


class CostConstraint(Cost):  # cost_functions.py:275-326
    def __init__(self, robot, n_support_points, q_l, traj_range_l, radius_l, is_soft=False, **kwargs):
        super().__init__(robot, n_support_points, **kwargs)
        dev = (self.tensor_args or {}).get("device", None)
        # tensors pass straight through (device-side constraint construction, conflicts.soft_constraints_from_paths)
        if torch.is_tensor(q_l):
            self.qs = q_l.to(torch.float32).reshape(-1, q_l.shape[-1]).to(dev)
        else:
            self.qs = torch.stack([torch.as_tensor(q, dtype=torch.float32) for q in q_l], dim=0).to(dev)
        self.traj_ranges = torch.as_tensor(traj_range_l, dtype=torch.float32).reshape(-1, 2).to(dev)
        self.radii = torch.as_tensor(radius_l, dtype=torch.float32).reshape(-1).to(dev)
        self.is_soft = is_soft
