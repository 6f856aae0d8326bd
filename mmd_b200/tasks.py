"""RobotPlanarDisk / PlanningTask with the attributes the sampling path reads, integer outputs computed by libmmdk.

Reference: TR/robots/robot_planar_disk.py:59-203, TR/robots/robot_base.py:149-161, TR/tasks/tasks.py:22-311,
TR/torch_planning_objectives/fields/distance_fields.py:290-367.
"""
import ctypes as C

import torch

from . import _lib
from ._gridpack import packed_grid

ROBOT_PLANAR_DISK_RADIUS = 0.05  # mmd/config/mmd_params.py:30


class RobotPlanarDisk:
    def __init__(self, name='RobotPlanarDisk', radius=ROBOT_PLANAR_DISK_RADIUS, q_limits=((-1., -1.), (1., 1.)),
                 tensor_args=None, **kwargs):
        self.name = name
        self.radius = radius
        self.tensor_args = tensor_args or {"device": torch.device("cpu"), "dtype": torch.float32}
        self.q_limits = torch.tensor(q_limits, dtype=torch.float32)
        self.q_min, self.q_max = self.q_limits[0], self.q_limits[1]
        self.q_dim = 2
        self.dt = None
        # robot_planar_disk.py:68 link margin 1.1 r as an fp32 tensor
        self.link_margins_for_object_collision_checking = [radius * 1.1]
        self.link_margins_for_object_collision_checking_tensor = torch.tensor([radius * 1.1], dtype=torch.float32)
        self.df_collision_self = None  # robot_base.py:87-88

    def get_position(self, x):  # robot_base.py:149-153
        return x[..., :self.q_dim]

    def get_velocity(self, x):  # robot_base.py:155-161
        return x[..., self.q_dim:2 * self.q_dim]

    def check_rr_collisions(self, robot_q):
        """(..., n_robots, 2) -> (collisions bool (..., R, R), midpoints (..., R, R, 2) NaN where free)
        (robot_planar_disk.py:173-203)."""
        lib = _lib.lib()
        q = robot_q.to(torch.float32).contiguous()
        R = q.shape[-2]
        lead = q.shape[:-2]
        n = int(torch.tensor(lead).prod()) if len(lead) else 1
        coll = torch.empty(*lead, R, R, dtype=torch.uint8, device=q.device)
        mid = torch.empty(*lead, R, R, 2, dtype=torch.float32, device=q.device)
        margin = 2.1 * self.radius
        _lib.check(lib.mmdk_check_rr_collisions(_lib.ptr(q), n, R, margin, _lib.ptr(coll), _lib.ptr(mid), _lib.stream_ptr()))
        return coll.bool(), mid


class _CollisionField:
    """Parameters of a CollisionObjectDistanceField / CollisionWorkspaceBoundariesDistanceField
    (distance_fields.py:290-367)."""

    def __init__(self, collision_margins, cutoff_margin, df_obj_list_fn=None, ws_min=None, ws_max=None):
        self.collision_margins = collision_margins
        self.cutoff_margin = cutoff_margin
        self.df_obj_list_fn = df_obj_list_fn
        self.ws_min, self.ws_max = ws_min, ws_max


class PlanningTask:
    def __init__(self, env=None, robot=None, ws_limits=None, obstacle_cutoff_margin=0.01, tensor_args=None, **kwargs):
        self.env, self.robot = env, robot
        self.tensor_args = tensor_args or {"device": torch.device("cpu"), "dtype": torch.float32}
        self.ws_limits = env.limits if ws_limits is None else torch.as_tensor(ws_limits, dtype=torch.float32)
        self.ws_min, self.ws_max = self.ws_limits[0], self.ws_limits[1]
        self.obstacle_cutoff_margin = obstacle_cutoff_margin
        margins = robot.link_margins_for_object_collision_checking_tensor
        self.df_collision_self = robot.df_collision_self
        self.df_collision_objects = _CollisionField(margins, obstacle_cutoff_margin, df_obj_list_fn=env.get_df_obj_list)
        # tasks.py:82-84: boundaries scaled by 1.08 (fp32 tensor * python float)
        self.df_collision_ws_boundaries = _CollisionField(margins, obstacle_cutoff_margin, ws_min=self.ws_min * 1.08,
                                                          ws_max=self.ws_max * 1.08)
        self._collision_fields = [self.df_collision_self, self.df_collision_objects, self.df_collision_ws_boundaries]
        self._packed = {}

    def get_collision_fields(self):
        return self._collision_fields

    # -- integer outputs -------------------------------------------------------------------------------------------
    def _env_struct(self, device):
        env = _lib.GuideEnv()
        keep = None
        grid = self.env.grid_map_sdf_obj_fixed
        if grid is not None:
            keep = packed_grid(grid, device)   # shared with the guides (mmd_b200/_gridpack.py)
            env.grid_dev = keep.data_ptr()
            env.nx, env.ny = keep.shape[0], keep.shape[1]
            md = torch.abs(self.env.limits[1] - self.env.limits[0])
            env.grid_lo[0], env.grid_lo[1] = float(self.env.limits[0][0]), float(self.env.limits[0][1])
            env.grid_map_dim[0], env.grid_map_dim[1] = float(md[0]), float(md[1])
        else:
            env.grid_dev = None
        b = self.df_collision_ws_boundaries
        env.ws_min[0], env.ws_min[1] = float(b.ws_min[0]), float(b.ws_min[1])
        env.ws_max[0], env.ws_max[1] = float(b.ws_max[0]), float(b.ws_max[1])
        return env, keep

    def classify(self, trajs, num_interpolation=5, return_waypoints=False):
        """free mask [B] (uint8), cost [B] = path length + smoothness, optional waypoint collisions."""
        lib = _lib.lib()
        t = trajs.to(torch.float32).contiguous()
        B, H, _ = t.shape
        env, keep = self._env_struct(t.device)
        free = torch.empty(B, dtype=torch.uint8, device=t.device)
        cost = torch.empty(B, dtype=torch.float32, device=t.device)
        wp = torch.empty(B, (H - 1) * num_interpolation, dtype=torch.uint8, device=t.device) if return_waypoints else None
        qmin = (C.c_float * 2)(float(self.robot.q_min[0]), float(self.robot.q_min[1]))
        qmax = (C.c_float * 2)(float(self.robot.q_max[0]), float(self.robot.q_max[1]))
        _lib.check(lib.mmdk_classify_trajs(C.byref(env), _lib.ptr(t), B, H, num_interpolation, float(self.robot.radius),
                                           qmin, qmax, _lib.ptr(free), _lib.ptr(cost), _lib.ptr(wp), _lib.stream_ptr()))
        del keep
        return free, cost, wp

    def get_trajs_collision_and_free(self, trajs, return_indices=False, num_interpolation=5):
        """tasks.py:236-311 for a [B, H, D] batch: (trajs_coll, idxs_coll, trajs_free, idxs_free, waypoint collisions)."""
        assert trajs.ndim == 3
        free, _, wp = self.classify(trajs, num_interpolation, return_waypoints=True)
        free_b = free.bool()
        idx_free = torch.argwhere(free_b)
        idx_coll = torch.argwhere(~free_b)
        trajs_free = trajs[idx_free.squeeze(-1)] if idx_free.numel() else None
        trajs_coll = trajs[idx_coll.squeeze(-1)] if idx_coll.numel() else None
        if return_indices:
            return trajs_coll, idx_coll, trajs_free, idx_free, wp.bool()
        return trajs_coll, trajs_free
