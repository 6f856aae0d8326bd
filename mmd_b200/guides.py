"""GuideManagerTrajectoriesWithVelocity with the reference's call surface (guides.py:152-253), lowered to libmmdk.

The cost objects are introspected by duck typing: the reference's own CostComposite / CostCollision /
CostGPTrajectory / CostConstraint instances work, and so do the light descriptors in mmd_b200.costs.  The gradient is
closed form, evaluated by the fused CUDA kernel -- there is no autograd and no torch fallback.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._gridpack import packed_grid


def _f(v):
    if torch.is_tensor(v):
        return float(v.reshape(-1)[0])
    return float(v)


def _normalizer_limits(dataset):
    """mins/maxs [D] of the trajectory LimitsNormalizer (mmd/datasets/normalization.py:150-168)."""
    nz = getattr(dataset, "normalizer", dataset)
    if hasattr(nz, "normalizers"):  # DatasetNormalizer
        key = getattr(dataset, "field_key_traj", "traj")
        nz = nz.normalizers[key]
    return torch.as_tensor(nz.mins, dtype=torch.float32).cpu(), torch.as_tensor(nz.maxs, dtype=torch.float32).cpu()


class KeepList(list):
    """Device tensors a lowered mmdk_groups points into (kept alive by the caller), with named handles on the ones that may
    be refreshed in place."""
    rows_d = None
    vals_d = None


class ConstraintSet:
    """Device arrays of the CostConstraint objects of ONE group, bucketed by waypoint (include/mmdk.h mmdk_groups)."""

    def __init__(self, costs: Sequence, weights: Sequence[float], H: int, device):
        base = 0
        bps, entries = [], []
        hh = torch.arange(H, device=device, dtype=torch.float32)
        for c in costs:
            qs = torch.as_tensor(c.qs, dtype=torch.float32).reshape(-1, 2).to(device)
            rng = torch.as_tensor(c.traj_ranges, dtype=torch.float32).reshape(-1, 2).to(device)
            rad = torch.as_tensor(c.radii, dtype=torch.float32).reshape(-1).to(device)
            mask = (hh[:, None] >= rng[None, :, 0]) & (hh[:, None] < rng[None, :, 1])  # [H, n]  cost_functions.py:304-305
            idx = mask.nonzero()  # sorted by (h, constraint index)
            local = torch.zeros(H + 1, dtype=torch.int64)
            local[1:] = torch.cumsum(mask.sum(1), 0).cpu()
            bps.append((local + base).to(torch.int32))
            base += idx.shape[0]
            e = torch.zeros(idx.shape[0], 4, device=device)
            e[:, :2] = qs[idx[:, 1]]
            e[:, 2] = rad[idx[:, 1]]
            entries.append(e)
        self.n_obj = len(costs)
        self.bucket_ptr = torch.stack(bps, 0) if bps else torch.zeros(0, H + 1, dtype=torch.int32)
        self.cons = torch.cat(entries, 0) if entries else torch.zeros(0, 4, device=device)
        self.weights = torch.tensor([float(w) for w in weights], dtype=torch.float32)


class GuideManagerTrajectoriesWithVelocity(nn.Module):
    def __init__(self, dataset, cost, clip_grad=False, clip_grad_rule='norm', max_grad_norm=1., max_grad_value=0.1,
                 interpolate_trajectories_for_collision=False, num_interpolated_points_for_collision=128,
                 start_state_pos=None, goal_state_pos=None, num_steps=100, robot=None, n_samples=1, tensor_args=None,
                 **kwargs):
        super().__init__()
        self.cost = cost
        self.dataset = dataset
        # guides.py:189-191: the interpolated trajectory is computed and never consumed -> nothing to do
        self.interpolate_trajectories_for_collision = interpolate_trajectories_for_collision
        self.num_interpolated_points_for_collision = num_interpolated_points_for_collision
        self.clip_grad = clip_grad
        self.clip_grad_rule = clip_grad_rule
        self.max_grad_norm = max_grad_norm
        self.max_grad_value = max_grad_value
        if clip_grad and clip_grad_rule != 'norm':
            raise NotImplementedError("only clip_grad_rule='norm' is lowered (mpd.py:257-265 uses it)")
        self.extra_cost_l = []
        self.extra_costs_grad_weight_l = []
        self.tensor_args = tensor_args
        self._env_cache = None
        self._grid_cache = {}

    # -- reference API ---------------------------------------------------------------------------------------------
    def add_extra_costs(self, extra_costs, extra_costs_grad_weights):  # guides.py:228-230
        self.extra_cost_l.extend(extra_costs)
        self.extra_costs_grad_weight_l.extend(extra_costs_grad_weights)

    def reset_extra_costs(self):  # guides.py:232-234
        self.extra_cost_l = []
        self.extra_costs_grad_weight_l = []

    @torch.no_grad()
    def forward(self, x_normalized, hard_conds=None, return_raw=False):
        """grad [B,H,D] of one guide evaluation; the whole batch is one planner call (one clip decision)."""
        lib = _lib.lib()
        x = x_normalized.contiguous()
        B, H, D = x.shape
        env, keep = self.lower_env(x.device)
        grp, keep2 = self.lower_groups(1, B, H, x.device, [self._own_constraints()], None)
        grad = torch.empty_like(x)
        n_costs = 3 + len(self.extra_cost_l)
        raw = torch.zeros(n_costs, B, H, D, device=x.device) if return_raw else None
        _lib.check(lib.mmdk_guide_grad(C.byref(env), C.byref(grp), H, _lib.ptr(x), _lib.ptr(grad), _lib.ptr(raw),
                                       n_costs, _lib.stream_ptr()))
        del keep, keep2
        return (grad, raw) if return_raw else grad

    # -- lowering --------------------------------------------------------------------------------------------------
    def _own_constraints(self):
        return list(self.extra_cost_l), list(self.extra_costs_grad_weight_l)

    def _pack_grid(self, grid_obj, device):
        return packed_grid(grid_obj, device)   # process-wide cache (mmd_b200/_gridpack.py)

    def lower_env(self, device):
        """mmdk_guide_env from the CostComposite (mpd.py:215-255)."""
        env = _lib.GuideEnv()
        keep = []
        mins, maxs = _normalizer_limits(self.dataset)
        rng = maxs - mins  # fp32, as normalization.py:168 computes it
        for d in range(4):
            env.norm_min[d] = float(mins[d])
            env.norm_range[d] = float(rng[d])
        env.grid_dev = None
        env.w_collision = env.w_border = env.w_smooth = 0.0
        env.coll_inv_sigma2 = 1.0
        env.margin = 0.0
        env.max_grad_norm = float(self.max_grad_norm) if self.clip_grad else float("inf")
        have_border = have_gp = False
        weights = list(self.cost.weight_cost_l)
        for cost, w in zip(self.cost.cost_l, weights):
            if hasattr(cost, "field"):
                field = cost.field
                if field is None:  # non-tensor cost, skipped by the reference (guides.py:208)
                    continue
                # distance_fields.py:115: fp32 tensor + python float, rounded in fp32
                margin = float(torch.as_tensor(field.collision_margins, dtype=torch.float32).reshape(-1)[0]
                               + float(_f(field.cutoff_margin)))
                inv_s2 = 1.0 / (float(cost.sigma_coll) ** 2)
                if getattr(field, "ws_min", None) is not None:
                    env.ws_min[0], env.ws_min[1] = float(field.ws_min[0]), float(field.ws_min[1])
                    env.ws_max[0], env.ws_max[1] = float(field.ws_max[0]), float(field.ws_max[1])
                    env.w_border = float(w)
                    have_border = True
                else:
                    grid_obj = None
                    for df in field.df_obj_list_fn():
                        if hasattr(df, "sdf_tensor"):
                            grid_obj = df
                        elif not _is_empty_object_field(df):
                            raise NotImplementedError("analytic (non-grid) object fields are not lowered; "
                                                      "all reference envs keep extra objects empty")
                    if grid_obj is not None:
                        packed = self._pack_grid(grid_obj, device)
                        keep.append(packed)
                        env.grid_dev = packed.data_ptr()
                        env.nx, env.ny = packed.shape[0], packed.shape[1]
                        lim = torch.as_tensor(grid_obj.limits, dtype=torch.float32).cpu()
                        md = torch.abs(lim[1] - lim[0])
                        env.grid_lo[0], env.grid_lo[1] = float(lim[0][0]), float(lim[0][1])
                        env.grid_map_dim[0], env.grid_map_dim[1] = float(md[0]), float(md[1])
                        env.w_collision = float(w)
                env.margin = margin
                env.coll_inv_sigma2 = inv_s2
            elif hasattr(cost, "dt") and hasattr(cost, "sigma_gp"):
                dt = float(cost.dt)
                s2 = float(cost.sigma_gp) ** 2
                qc = float(torch.tensor(1.0) / s2)  # gp_factor.py:25 eye / sigma**2 in fp32
                env.dt = dt
                # gp_factor.py:42-50: python-double coefficient times the fp32 Q_c_inv
                env.gp_q11 = float(torch.tensor(qc) * (12. * (dt ** -3.)))
                env.gp_q12 = float(torch.tensor(qc) * (-6. * (dt ** -2.)))
                env.gp_q22 = float(torch.tensor(qc) * (4. * (dt ** -1.)))
                env.w_smooth = float(w)
                have_gp = True
            else:
                raise NotImplementedError(f"cost term {type(cost).__name__} is not lowered")
        if not have_border:
            env.ws_min[0] = env.ws_min[1] = -1e30
            env.ws_max[0] = env.ws_max[1] = 1e30
        if not have_gp:
            env.dt = 1.0
            env.gp_q11 = env.gp_q12 = env.gp_q22 = 0.0
        return env, keep

    def lower_groups(self, n_groups, K, H, device, constraints_per_group, hard_conds_per_group, peers=None,
                     peer_self=None, peer_radius=0.0, peer_weight=0.0, peer_hash=None, peer_seq_ptr=None):
        """mmdk_groups for `n_groups` planner calls of K samples.  constraints_per_group[g] = (costs, weights);
        hard_conds_per_group[g] = {row: tensor[D]} (normalised) or None."""
        grp = _lib.Groups()
        keep = KeepList()
        grp.n_groups, grp.K = n_groups, K
        rows = torch.full((n_groups, _lib.MAX_HARD_ROWS), -1, dtype=torch.int32)
        vals = torch.zeros(n_groups, _lib.MAX_HARD_ROWS, 4, dtype=torch.float32)
        if hard_conds_per_group is not None:
            for g, hc in enumerate(hard_conds_per_group):
                if hc is None:
                    continue
                if len(hc) > _lib.MAX_HARD_ROWS:
                    raise ValueError("too many hard-conditioned waypoints")
                for r, (row, v) in enumerate(hc.items()):
                    v = torch.as_tensor(v, dtype=torch.float32).cpu()
                    rows[g, r] = int(row) % H
                    vals[g, r] = v[0] if v.dim() == 2 else v
        rows_d, vals_d = rows.to(device), vals.to(device)
        keep += [rows_d, vals_d]
        keep.rows_d, keep.vals_d = rows_d, vals_d   # refreshed in place by callers that cache the lowering
        grp.hard_rows_dev, grp.hard_vals_dev = rows_d.data_ptr(), vals_d.data_ptr()
        n_obj_total = sum(len(c[0]) for c in constraints_per_group) if constraints_per_group else 0
        if n_obj_total:
            obj_ptr = [0]
            bps, conss, ws = [], [], []
            ent_base = 0
            for costs, weights in constraints_per_group:
                cs = ConstraintSet(costs, weights, H, device)
                bps.append(cs.bucket_ptr + ent_base)
                conss.append(cs.cons)
                ws.append(cs.weights)
                ent_base += cs.cons.shape[0]
                obj_ptr.append(obj_ptr[-1] + cs.n_obj)
            obj_ptr_d = torch.tensor(obj_ptr, dtype=torch.int32).to(device)
            bp_d = torch.cat(bps, 0).to(device).contiguous()
            cons_d = torch.cat(conss, 0).contiguous()
            if cons_d.shape[0] == 0:
                cons_d = torch.zeros(1, 4, device=device)
            w_d = torch.cat(ws, 0).to(device)
            keep += [obj_ptr_d, bp_d, cons_d, w_d]
            grp.obj_ptr_dev, grp.bucket_ptr_dev = obj_ptr_d.data_ptr(), bp_d.data_ptr()
            grp.cons_dev, grp.obj_weight_dev = cons_d.data_ptr(), w_d.data_ptr()
        else:
            grp.obj_ptr_dev = None
        if peers is not None:
            keep += [peers, peer_self]
            grp.peers_dev, grp.peer_self_dev = peers.data_ptr(), peer_self.data_ptr()
            grp.n_peers = peers.shape[0]
            grp.peer_seq_dev = peer_seq_ptr   # not None: `peers` is half 0 of a double-buffered exchange table
            grp.peer_radius, grp.peer_weight = float(peer_radius), float(peer_weight)
            if peer_hash is not None:
                keep.append(peer_hash)
                grp.peer_cell_start_dev = peer_hash.cell_start.data_ptr()
                grp.peer_sorted_dev = peer_hash.sorted.data_ptr()
                grp.peer_grid, grp.peer_grid_lo = peer_hash.grid, peer_hash.lo
                grp.peer_grid_inv_cell = peer_hash.inv_cell
            else:
                grp.peer_cell_start_dev = None
                grp.peer_sorted_dev = None
        else:
            grp.peers_dev = None
            grp.peer_seq_dev = None
            grp.n_peers = 0
            grp.peer_cell_start_dev = None
            grp.peer_sorted_dev = None
        return grp, keep


class PeerHash:
    """Per-waypoint uniform-grid hash of a lock-step peer table [n_peers, H, 2] (include/mmdk.h mmdk_build_peer_hash).
    Cell size >= radius (a hair above it, so fp32 rounding of the cell coordinate can never push a peer within the
    radius two cells away); the grid covers [-1.2, 1.2]^2 and clamps at its border (clamping is 1-Lipschitz, so the
    3x3 neighbourhood stays sufficient outside the box too)."""
    SPAN, LO, GRID_MAX = 2.4, -1.2, 32

    def __init__(self, n_peers, H, radius, device):
        import math
        self.n_peers, self.H = int(n_peers), int(H)
        self.grid = max(1, min(self.GRID_MAX, int(math.floor(self.SPAN / (float(radius) * 1.001)))))
        self.lo = self.LO
        self.inv_cell = self.grid / self.SPAN
        assert 1.0 / self.inv_cell >= float(radius)
        self.cell_start = torch.zeros(H, self.grid * self.grid + 1, dtype=torch.int16, device=device)
        self.sorted = torch.zeros(H, self.n_peers, 4, dtype=torch.float32, device=device)

    def build(self, peers):
        assert peers.shape == (self.n_peers, self.H, 2) and peers.is_contiguous()
        _lib.check(_lib.lib().mmdk_build_peer_hash(_lib.ptr(peers), None, self.n_peers, self.H, self.grid, self.lo,
                                                   self.inv_cell, _lib.ptr(self.cell_start), _lib.ptr(self.sorted),
                                                   _lib.stream_ptr()))
        return self


def _is_empty_object_field(df):
    """True for an ObjectField whose primitives are all empty sphere fields (sdf == 1 everywhere,
    TR/environments/primitives.py:108-110); every env_*_extra_objects.py of the reference builds exactly that."""
    fields = getattr(df, "fields", None)
    if fields is None:
        return getattr(df, "is_empty", False)
    for f in fields:
        centers = getattr(f, "centers", None)
        if centers is None or len(centers) != 0:
            return False
    return True
