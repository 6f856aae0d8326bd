"""Conflict detection between robot paths on the device, one state or a batch of candidate states.

Reference: mmd/planners/multi_agent/cbs.py:166-246 (CBS.get_conflicts), :446-458 (the 'least_collisions' strategy calls it
once per free sample), mmd/common/conflicts.py (Vertex / Edge / PointConflict), mmd/common/multi_agent_utils.py:120-143
(global_pad_paths), mmd/common/trajectory_utils.py:54-69 (densify_trajs).  The reference runs Python loops over waypoints
(densify) and K host round trips; here one kernel evaluates every candidate joint state (mmdk_get_conflicts) and the host
only reads back the counts and the (t, a, b) rows of the state it keeps.  Integer outputs are bit-exact.
"""
import ctypes as C
from math import ceil, floor
from typing import List, Sequence

import torch

from . import _lib


class Conflict:
    def __init__(self):
        self.time_interval = None


class VertexConflict(Conflict):  # conflicts.py:36-53
    def __init__(self, agent_ids, q_l, t):
        super().__init__()
        self.agent_ids = agent_ids
        self.q_map = {a: q for a, q in zip(agent_ids, q_l)}
        self.t = t

    def get_t_range(self):
        return self.t, self.t


class EdgeConflict(Conflict):  # conflicts.py:56-80
    def __init__(self, agent_ids, q_from_l, q_to_l, t_from, t_to):
        super().__init__()
        self.agent_ids = agent_ids
        self.agent_id_to_q_from = {a: q for a, q in zip(agent_ids, q_from_l)}
        self.agent_id_to_q_to = {a: q for a, q in zip(agent_ids, q_to_l)}
        self.t_from, self.t_to = t_from, t_to

    def get_t_range(self):
        return self.t_from, self.t_to


class PointConflict(Conflict):  # conflicts.py:83-103
    def __init__(self, agent_ids, q_l, p_l, t_from, t_to):
        super().__init__()
        self.agent_ids = agent_ids
        self.agent_id_to_p = {a: p for a, p in zip(agent_ids, p_l)}
        self.agent_id_to_q = {a: q for a, q in zip(agent_ids, q_l)}
        self.t_from, self.t_to = t_from, t_to

    def get_t_range(self):
        return self.t_from, self.t_to


def global_pad_paths(path_l: Sequence[torch.Tensor], start_time_l: Sequence[int]) -> List[torch.Tensor]:
    """multi_agent_utils.py:120-143: repeat the start state before an agent's start time and its last state afterwards."""
    if len(path_l) == 0:
        return []
    max_t = max(len(p) + start_time_l[i] for i, p in enumerate(path_l))
    out = []
    for i, p in enumerate(path_l):
        if len(p) + start_time_l[i] < max_t:
            p = torch.cat([p, p[-1].repeat(max_t - len(p) - start_time_l[i], 1)])
        if start_time_l[i] > 0:
            p = torch.cat([p[0].repeat(start_time_l[i], 1), p])
        out.append(p)
    return out


def _run(base, cand, agent_id, densify, margin, want_dense):
    lib = _lib.lib()
    R, T, _ = base.shape
    n_cand = 1 if cand is None else cand.shape[0]
    Td = (T - 1) * densify + 1
    dev = base.device
    coll = torch.empty(n_cand, Td, R, R, dtype=torch.uint8, device=dev)
    count = torch.empty(n_cand, dtype=torch.int32, device=dev)
    dense = torch.empty(n_cand, R, Td, 2, dtype=torch.float32, device=dev) if want_dense else None
    _lib.check(lib.mmdk_get_conflicts(_lib.ptr(base), _lib.ptr(cand), n_cand, int(agent_id), R, T, int(densify), float(margin),
                                      _lib.ptr(coll), _lib.ptr(count), _lib.ptr(dense), _lib.stream_ptr()))
    return coll, count, dense


def count_conflicts_batched(best_path_l, start_time_l, agent_id, candidate_paths, robot_radius=0.05, densify=1):
    """len(get_conflicts(state with robot `agent_id` following candidate k)) for every k at once (cbs.py:446-458) ->
    (counts int32 [n_cand] on the device, collision tensor uint8 [n_cand, T_dense, R, R]).
    best_path_l: per-robot [H_i, >=2] paths (the entry of `agent_id` is ignored), candidate_paths: [n_cand, H, >=2]."""
    R = len(best_path_l)
    cands = [c for c in candidate_paths]
    padded = global_pad_paths([p if i != agent_id else cands[0] for i, p in enumerate(best_path_l)], start_time_l)
    base = torch.stack([p[..., :2] for p in padded]).to(torch.float32).contiguous()
    cand_pad = torch.stack([global_pad_paths([c if i == agent_id else best_path_l[i] for i in range(R)], start_time_l)[agent_id][..., :2]
                            for c in cands]).to(torch.float32).contiguous()
    coll, count, _ = _run(base, cand_pad, agent_id, densify, 2.1 * robot_radius, False)
    return count, coll


def get_conflicts(best_path_l, start_time_l, conflict_types=(PointConflict,), robot_radius=0.05):
    """CBS.get_conflicts (cbs.py:166-246) for one joint state: list of conflicts in torch.nonzero order (t, a, b)."""
    if len(best_path_l) == 0:
        return []
    padded = global_pad_paths(list(best_path_l), start_time_l)
    paths_pos_l = [p[..., :2].to(torch.float32) for p in padded]
    densify = 2 if EdgeConflict in conflict_types else 1
    base = torch.stack(paths_pos_l).contiguous()
    coll, _, dense = _run(base, None, -1, densify, 2.1 * robot_radius, True)
    idx = torch.nonzero(coll[0].int())                      # [n, 3] rows (t_dense, a, b), row-major like the reference
    dense = dense[0]                                        # [R, T_dense, 2]
    conflicts = []
    for t_dense, a, b in idx.tolist():
        t_from, t_to = floor(t_dense / densify), ceil(t_dense / densify)
        mid = (dense[a, t_dense] + dense[b, t_dense]) / 2   # check_rr_collisions' collision point (robot_planar_disk.py:197)
        if VertexConflict in conflict_types and t_from == t_to:
            conflicts.append(VertexConflict([a, b], [paths_pos_l[a][t_from], paths_pos_l[b][t_from]], int(t_from)))
        if EdgeConflict in conflict_types and t_from != t_to:
            conflicts.append(EdgeConflict([a, b], q_from_l=[paths_pos_l[a][t_from], paths_pos_l[b][t_from]],
                                          q_to_l=[paths_pos_l[a][t_to], paths_pos_l[b][t_to]], t_from=t_from, t_to=t_to))
        if PointConflict in conflict_types:
            conflicts.append(PointConflict([a, b], p_l=[dense[a, t_dense], dense[b, t_dense]], q_l=[mid, mid],
                                           t_from=int(t_from), t_to=int(t_to)))
    return conflicts


def soft_constraints_from_paths(best_path_l, start_time_l, agent_id, radius=2.4 * 0.05, T_agent=None):
    """CBS.create_soft_constraints_from_other_agents_paths (cbs.py:468-508) without the per-waypoint Python loop: the best
    paths of the OTHER agents become one soft MultiPointConstraint for `agent_id` -- entry (q = position of agent j at its
    waypoint t, range (t_a, t_a + 1), radius) with t_a = t + start_time[j] - start_time[agent_id], kept iff 1 <= t_a <= T_agent
    -- built with a handful of tensor ops on the device, in the reference's order (agents ascending, then t ascending), so
    the bucketed arrays the step kernel reads are identical.  best_path_l: per agent [H_j, >= 2] (positions first).
    Returns [] or [MultiPointConstraint] like the reference."""
    from .planners import MultiPointConstraint
    if len(best_path_l) == 0:
        return []
    if T_agent is None:   # cbs.py:494-497
        T_agent = (best_path_l[agent_id].shape[0] - 1) if agent_id < len(best_path_l) else None
    qs, ts = [], []
    for j, path in enumerate(best_path_l):
        if j == agent_id:
            continue
        Tj = (path.shape[0] - 1) if T_agent is None else T_agent
        t_a = torch.arange(path.shape[0], device=path.device) + int(start_time_l[j]) - int(start_time_l[agent_id])
        keep = (t_a >= 1) & (t_a <= Tj)
        qs.append(path[keep][:, :2])
        ts.append(t_a[keep])
    if not qs:
        return []
    q = torch.cat(qs, 0)
    if q.shape[0] == 0:
        return []
    t = torch.cat(ts, 0).to(torch.float32)
    c = MultiPointConstraint(q_l=q, t_range_l=torch.stack((t, t + 1), -1), radius_l=torch.full((q.shape[0],), float(radius)),
                             is_soft=True)
    return [c]
