"""Pins oracle/port.py against the LIVE reference (only where /root/reference exists: the build container)."""
import pytest
import torch

from oracle import port, ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_build
    P = port.make_unet_params(seed=0)
    r = ref_build.build_reference("EnvConveyor2D", 25, P)
    r["P"] = P
    return r


def test_unet_and_schedule_bit_exact(ref):
    s = port.make_schedule(25)
    for k, v in s.items():
        assert torch.equal(getattr(ref["model"], k), v), k
    x = torch.randn(4, 64, 4, generator=torch.Generator().manual_seed(0))
    t = torch.full((4,), 9, dtype=torch.long)
    with torch.no_grad():
        assert torch.equal(ref["model"].model(x, t, None), port.unet_forward(ref["P"], x, t))


def test_unet_option1_and_attention_bit_exact():
    from oracle import ref_build
    for kw in (dict(dim_mults=(1, 2, 4, 8)), dict(dim_mults=(1, 2, 4), self_attention=True)):
        P = port.make_unet_params(seed=1, **kw)
        r = ref_build.build_reference("EnvEmpty2D", 25, P, **kw)
        x = torch.randn(2, 64, 4, generator=torch.Generator().manual_seed(0))
        t = torch.full((2,), 3, dtype=torch.long)
        with torch.no_grad():
            assert torch.allclose(r["model"].model(x, t, None), port.unet_forward(P, x, t), rtol=0, atol=1e-6)


def test_guide_and_chain_bit_exact(ref):
    from oracle import ref_build
    sdf, grad = port.build_sdf_grid("EnvConveyor2D")
    assert torch.equal(sdf, ref["env"].grid_map_sdf_obj_fixed.sdf_tensor)
    assert torch.equal(grad, ref["env"].grid_map_sdf_obj_fixed.grad_sdf_tensor)
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    gs = port.GuideSpec(port.GridSDF(sdf, grad), norm)
    x = torch.randn(8, 64, 4, generator=torch.Generator().manual_seed(1)) * 0.6
    assert torch.equal(ref["guide"](x), gs(x))
    K = 4
    noise = torch.randn(27, K, 64, 4, generator=torch.Generator().manual_seed(2))
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, -0.6]), torch.tensor([0.8, 0.6]), norm)
    c_ref = ref_build.reference_run_inference(ref, hc, K, noise)
    c_port = port.run_inference(port.DiffusionModel(ref["P"], 25), hc, K, noise, guide=gs)
    assert torch.equal(c_ref, c_port)


def test_lockstep_oracle_is_a_composition_of_reference_calls(ref):
    """port.lockstep_sample == per-step reference ddpm_sample_fn with a CostConstraint from the others."""
    from oracle import ref_build
    from mmd.models.diffusion_models.sample_functions import ddpm_sample_fn, apply_hard_conditioning
    R, K, T = 2, 3, 25
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    sdf, grad = port.build_sdf_grid("EnvConveyor2D")
    guides = [port.GuideSpec(port.GridSDF(sdf, grad), norm) for _ in range(R)]
    starts, goals = port.get_start_goal_pos_circle(R, 0.3)
    hcs = [port.hard_conds_from_start_goal(s, g, norm) for s, g in zip(starts, goals)]
    noise = torch.randn(R, T + 2, K, 64, 4, generator=torch.Generator().manual_seed(3))
    out = port.lockstep_sample(port.DiffusionModel(ref["P"], T), guides, hcs, K, noise)
    # the same thing, spelled with reference objects
    xs = [apply_hard_conditioning(noise[r, 0].clone(), port.repeat_hard_conds(hcs[r], K)) for r in range(R)]
    k = 1
    for i in reversed(range(-1, T)):
        t = torch.full((K,), i, dtype=torch.long)
        reps = [ref["dataset"].unnormalize_trajectories(xs[r][0:1].clone())[0, :, :2] for r in range(R)]
        new = []
        for r in range(R):
            q = reps[1 - r]
            hh = torch.arange(64).float()
            cc = ref_build.make_cost_constraint(ref, q, torch.stack((hh, hh + 1), -1), torch.full((64,), 0.12))
            ref["guide"].add_extra_costs([cc], [2e-2])
            with ref_build.scripted_noise([noise[r, k]]):
                x, _ = ddpm_sample_fn(ref["model"], xs[r], port.repeat_hard_conds(hcs[r], K), None, t,
                                      guide=ref["guide"], n_guide_steps=20, t_start_guide=13,
                                      noise_std_extra_schedule_fn=lambda _: 0.5)
            ref["guide"].reset_extra_costs()
            new.append(apply_hard_conditioning(x, port.repeat_hard_conds(hcs[r], K)))
        xs = new
        k += 1
    assert torch.equal(torch.stack(xs), out)


def _crossing_paths(R=5, H=64, seed=0, spread=0.5):
    """Straight lines through the origin region + noise: several robots meet near the centre (conflicts exist)."""
    g = torch.Generator().manual_seed(seed)
    ang = torch.arange(R, dtype=torch.float32) * (2 * 3.14159265 / R)
    start = torch.stack((torch.cos(ang), torch.sin(ang)), -1) * spread
    t = torch.linspace(0, 1, H)[None, :, None]
    pos = start[:, None, :] * (1 - 2 * t) + 0.004 * torch.randn(R, H, 2, generator=g)
    return [torch.cat((p, torch.zeros(H, 2)), -1) for p in pos]


@pytest.mark.parametrize("kinds", [("point",), ("vertex", "edge"), ("vertex", "edge", "point")])
def test_get_conflicts_matches_reference(kinds):
    """oracle.port.get_conflicts against the reference's own CBS.get_conflicts (cbs.py:166-246) called unbound on a stub
    `self` (the method reads only reference_robot, start_time_l and conflict_type_to_constraint_types)."""
    import types
    from mmd.common.conflicts import EdgeConflict, PointConflict, VertexConflict
    from mmd.planners.multi_agent.cbs import CBS
    from torch_robotics.robots import RobotPlanarDisk
    paths = _crossing_paths()
    start_times = [0, 2, 0, 1, 0]
    paths = [p[: 64 - s] for p, s in zip(paths, start_times)]   # ragged: different start times, same end time
    robot = RobotPlanarDisk(tensor_args={"device": torch.device("cpu"), "dtype": torch.float32})
    cmap = {}
    if "point" in kinds: cmap[PointConflict] = []
    if "vertex" in kinds: cmap[VertexConflict] = []
    if "edge" in kinds: cmap[EdgeConflict] = []
    stub = types.SimpleNamespace(reference_robot=robot, start_time_l=start_times, conflict_type_to_constraint_types=cmap)
    state = types.SimpleNamespace(path_bl=[p[None] for p in paths], ix_best_path_in_batch_l=[0] * len(paths))
    ref = CBS.get_conflicts(stub, state)
    out = port.get_conflicts(paths, start_times, want_vertex="vertex" in kinds, want_edge="edge" in kinds,
                             want_point="point" in kinds)
    assert len(ref) == len(out) and len(ref) > 0
    for r, o in zip(ref, out):
        assert r.agent_ids == [o[1], o[2]]
        if isinstance(r, VertexConflict):
            assert o[0] == "vertex" and r.t == o[3]
        elif isinstance(r, EdgeConflict):
            assert o[0] == "edge" and (r.t_from, r.t_to) == (o[3], o[4])
        else:
            assert o[0] == "point" and (r.t_from, r.t_to) == (o[3], o[4])
            assert torch.equal(r.agent_id_to_p[o[1]], o[5]) and torch.equal(r.agent_id_to_q[o[1]], o[7])


def test_smooth_trajs_matches_reference():
    from mmd.common.trajectory_utils import smooth_trajs
    x = torch.randn(6, 64, 4, generator=torch.Generator().manual_seed(1))
    assert torch.equal(smooth_trajs(x), port.smooth_trajs(x))


@pytest.mark.parametrize("env_name", ["EnvHighways2D", "EnvEmpty2D", "EnvConveyor2D"])
def test_reference_cost_objects_lower_like_ours(env_name):
    """INTEGRATION.md section A claims the reference's OWN CostComposite / CostCollision / CostGPTrajectory /
    CostConstraint / PlanningTask objects lower by duck typing.  Build them with the live reference, hand them to
    mmd_b200's GuideManagerTrajectoriesWithVelocity and compare the lowered mmdk_guide_env (every scalar field and the
    packed SDF grid, bit for bit) and the bucketed constraint arrays with the lowering of mmd_b200's own objects."""
    import ctypes as C
    from oracle import ref_build
    import mmd_b200 as M
    from mmd_b200 import _lib
    from mmd_b200.guides import ConstraintSet
    r = ref_build.build_reference(env_name, 25, None, with_model=False)
    cpu = torch.device("cpu")
    g_ref = M.GuideManagerTrajectoriesWithVelocity(r["dataset"], r["guide"].cost, clip_grad=True, tensor_args=r["tensor_args"])
    ta = {"device": cpu, "dtype": torch.float32}
    env = M.envs.get_env(env_name + "ExtraObjects", tensor_args=ta)
    robot = M.RobotPlanarDisk(tensor_args=ta)
    task = M.PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.05, tensor_args=ta)
    dataset = M.TrajectoryDataset(env, robot, task, *port.DEFAULT_NORMALIZER_LIMITS, tensor_args=ta)
    costs = [M.CostCollision(robot, 64, field=f, sigma_coll=1.0, tensor_args=ta) for f in task.get_collision_fields()]
    costs.append(M.CostGPTrajectory(robot, 64, 5.0 / 64, sigma_gp=1.0, tensor_args=ta))
    comp = M.CostComposite(robot, 64, costs, weights_cost_l=[2e-2] * (len(costs) - 1) + [8e-2], tensor_args=ta)
    g_own = M.GuideManagerTrajectoriesWithVelocity(dataset, comp, clip_grad=True, tensor_args=ta)
    e_ref, keep_ref = g_ref.lower_env(cpu)
    e_own, keep_own = g_own.lower_env(cpu)
    for name, ctype in _lib.GuideEnv._fields_:
        if name == "grid_dev":
            continue
        a, b = getattr(e_ref, name), getattr(e_own, name)
        a = list(a) if hasattr(a, "__len__") else a
        b = list(b) if hasattr(b, "__len__") else b
        assert a == b, (name, a, b)
    assert (e_ref.grid_dev is None) == (e_own.grid_dev is None)
    grids_ref = [k for k in keep_ref if torch.is_tensor(k) and k.dim() == 3]
    grids_own = [k for k in keep_own if torch.is_tensor(k) and k.dim() == 3]
    assert len(grids_ref) == len(grids_own)
    for a, b in zip(grids_ref, grids_own):
        assert torch.equal(a, b)
    # CostConstraint objects of the reference -> the same bucketed arrays
    qs = torch.rand(30, 2, generator=torch.Generator().manual_seed(1)) * 2 - 1
    hh = torch.randint(0, 60, (30,), generator=torch.Generator().manual_seed(2)).float()
    rng = torch.stack((hh, hh + 3), -1)
    rad = torch.full((30,), 0.12)
    cc_ref = ref_build.make_cost_constraint(r, qs, rng, rad, True)
    cc_own = M.CostConstraint(robot, 64, q_l=list(qs), traj_range_l=rng.tolist(), radius_l=rad.tolist(), is_soft=True, tensor_args=ta)
    s_ref, s_own = ConstraintSet([cc_ref], [2e-2], 64, cpu), ConstraintSet([cc_own], [2e-2], 64, cpu)
    assert torch.equal(s_ref.bucket_ptr, s_own.bucket_ptr) and torch.equal(s_ref.cons, s_own.cons)


def test_soft_constraints_from_other_agents_matches_reference(ref):
    """oracle.port.create_soft_constraints_from_other_agents_paths == the reference's CBS method (cbs.py:468-508), called
    unbound on a minimal stand-in for the CBS object / SearchState."""
    import types
    from oracle import ref_shim
    ref_shim.install()
    from mmd.planners.multi_agent.cbs import CBS
    g = torch.Generator().manual_seed(3)
    K = 3
    path_bl = [torch.rand(K, 64, 4, generator=g) * 2 - 1 for _ in range(4)]
    best = [1, 0, 2, 1]
    starts = [0, 4, 0, 2]
    fake = types.SimpleNamespace(reference_robot=ref["robot"], start_time_l=starts)
    state = types.SimpleNamespace(path_bl=path_bl, ix_best_path_in_batch_l=best)
    for agent in range(4):
        out = CBS.create_soft_constraints_from_other_agents_paths(fake, state, agent)
        assert len(out) == 1
        q, rng, rad = port.create_soft_constraints_from_other_agents_paths([path_bl[j][best[j]] for j in range(4)], starts, agent)
        assert torch.equal(torch.stack(out[0].get_q_l()), q)
        assert torch.equal(torch.tensor(out[0].get_t_range_l(), dtype=torch.float32), rng)
        assert torch.allclose(torch.tensor(out[0].radius_l), rad) and out[0].is_soft


@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_ensemble_loop_bit_exact(direction):
    """port.ensemble_p_sample_loop == the reference's DiffusionsEnsemble.p_sample_loop (diffusion_ensemble.py:56-106) on scripted
    noise, for a robot crossing the 1x2 tile grid in either direction (tile transforms (0,0),(2,0) or (2,0),(0,0)): every
    recorded chain frame, i.e. including the reference's aliasing of the recorded frames with the in-place stitch."""
    from math import ceil
    from oracle import ref_build
    ref_shim.install()
    from mmd.models.diffusion_models.diffusion_ensemble import DiffusionsEnsemble
    from mmd.models.diffusion_models.sample_functions import ddpm_sample_fn
    T, K = 25, 3
    P = port.make_unet_params(seed=0)
    tiles = {m: ref_build.build_reference("EnvEmptyNoWait2D", T, P, cutoff_margin=0.01) for m in (0, 1)}
    tr = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    if direction == "backward":
        tr = {0: tr[1], 1: tr[0]}
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    s0 = norm.normalize(torch.tensor([-0.7, 0.1, 0.0, 0.0]))
    g1 = norm.normalize(torch.tensor([0.6, -0.2, 0.0, 0.0]))
    hard = {0: {0: s0}, 1: {63: g1}}
    cross = {(0, 1): (63, 0)}
    g = torch.Generator().manual_seed(5)
    noise = {m: torch.randn(T + 2, K, 64, 4, generator=g) for m in (0, 1)}
    # the reference's draw order: x_T per tile, then per step, per tile (diffusion_ensemble.py:68-95)
    script = [noise[0][0], noise[1][0]] + [noise[m][k] for k in range(1, T + 2) for m in (0, 1)]
    ens = DiffusionsEnsemble({m: tiles[m]["model"] for m in (0, 1)}, tr)
    skw = [dict(guide=tiles[m]["guide"], n_guide_steps=20, t_start_guide=ceil(0.5 * T), noise_std_extra_schedule_fn=lambda x: 0.5)
           for m in (0, 1)]
    hc_ref = {m: {k: v.reshape(1, -1).repeat(K, 1) for k, v in h.items()} for m, h in hard.items()}
    with ref_build.scripted_noise(script), ref_build.quiet():
        _, chains = ens.p_sample_loop((K, 64, 4), hc_ref, dict(cross), n_diffusion_steps=T, return_chain=True, sample_fn=ddpm_sample_fn,
                                      n_diffusion_steps_without_noise=1, sample_kwargs=skw)
    sdf, grad = port.build_sdf_grid("EnvEmptyNoWait2D")
    guides = {m: port.GuideSpec(port.GridSDF(sdf, grad), norm, cutoff_margin=0.01) for m in (0, 1)}
    kw = {m: dict(guide=guides[m], n_guide_steps=20, t_start_guide=ceil(0.5 * T), noise_std=0.5) for m in (0, 1)}
    model = port.DiffusionModel(P, T)
    mine = port.ensemble_p_sample_loop({0: model, 1: model}, {m: port.repeat_hard_conds(h, K) for m, h in hard.items()}, dict(cross),
                                       tr, noise, T, 1, kw)
    for m in (0, 1):
        assert torch.equal(chains[m].transpose(0, 1), mine[m]), (direction, m)


def test_split_cost_constraints_to_tasks_matches_reference():
    """MPDEnsemble.split_cost_constraints_to_tasks (mpd_ensemble.py:431-507): same tiles, same object order (hard objects, then
    soft), same entries as the reference method run on the reference's own CostConstraint objects -- and the installed
    constraints (waypoint ranges - task_id * 64, positions - tile transform, :516-517) agree too."""
    import contextlib
    import io
    import types
    import mmd_b200 as M
    from oracle import ref_build
    ref_shim.install()
    from mmd.planners.single_agent.mpd_ensemble import MPDEnsemble as RefEns
    r = ref_build.build_reference("EnvEmptyNoWait2D", 25, None, with_model=False)
    task_stub = types.SimpleNamespace(infer_task_id_from_q_idx=lambda q_idx: (int(q_idx // 64), None))
    ref_self = types.SimpleNamespace(task=task_stub, robot=r["robot"], n_support_points=64, tensor_args=r["tensor_args"])
    ta = {"device": torch.device("cpu"), "dtype": torch.float32}
    our_self = types.SimpleNamespace(task=task_stub, robot=M.RobotPlanarDisk(tensor_args=ta), n_support_points=64, tensor_args=ta)
    g = torch.Generator().manual_seed(3)
    specs = []   # (qs, ranges, radii, is_soft): entries in both tiles, hard and soft, interleaved
    for is_soft, n in ((True, 7), (False, 4), (True, 3)):
        qs = torch.rand(n, 2, generator=g) * 4 - 1
        h0 = torch.randint(0, 127, (n,), generator=g)
        specs.append((qs, torch.stack((h0, h0 + 2), -1).float(), torch.rand(n, generator=g) * 0.2 + 0.05, is_soft))
    ref_cc = [ref_build.make_cost_constraint(r, qs, rng, rad, is_soft=s) for qs, rng, rad, s in specs]
    our_cc = [M.CostConstraint(our_self.robot, 64, q_l=list(qs), traj_range_l=[tuple(int(v) for v in t) for t in rng.tolist()],
                               radius_l=rad.tolist(), is_soft=s, tensor_args=ta) for qs, rng, rad, s in specs]
    with contextlib.redirect_stdout(io.StringIO()):
        a = RefEns.split_cost_constraints_to_tasks(ref_self, ref_cc)
    b = M.MPDEnsemble.split_cost_constraints_to_tasks(our_self, our_cc)
    assert list(a.keys()) == list(b.keys())
    transforms = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    for task_id in a:
        assert [c.is_soft for c in a[task_id]] == [c.is_soft for c in b[task_id]]
        for ca, cb in zip(a[task_id], b[task_id]):
            assert torch.equal(ca.qs, cb.qs) and torch.equal(ca.traj_ranges.float(), cb.traj_ranges.float())
            assert torch.allclose(torch.as_tensor(ca.radii, dtype=torch.float32), torch.as_tensor(cb.radii, dtype=torch.float32), atol=0, rtol=0)
            # installation (mpd_ensemble.py:516-517)
            ra, qa = ca.traj_ranges - task_id * 64, ca.qs - transforms[task_id]
            rb, qb = cb.traj_ranges - task_id * 64, cb.qs - transforms[task_id]
            assert torch.equal(ra.float(), rb.float()) and torch.equal(qa, qb)


def test_combine_trajs_matches_reference():
    """PlanningTaskEnsemble.combine_trajs (tasks_ensemble.py:162-238): tile chains moved into the global frame and concatenated,
    a sample is free iff it is free in every tile, costs / best trajectory recomputed on the combined free set -- compared field
    by field with the reference method on CPU (the reference's `trajs_final_coll` indexes the wrong dimension, :186: not compared)."""
    import types
    import mmd_b200 as M
    from oracle import ref_build
    ref_shim.install()
    from torch_robotics.tasks.tasks_ensemble import PlanningTaskEnsemble as RefPTE
    r = {m: ref_build.build_reference("EnvEmptyNoWait2D", 25, None, with_model=False, cutoff_margin=0.01) for m in (0, 1)}
    tr = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    ta = {"device": torch.device("cpu"), "dtype": torch.float32}
    ref_self = types.SimpleNamespace(transforms=tr, robot=r[0]["robot"], tensor_args=ta)
    ref_self.transform_q = lambda task_id, q: RefPTE.transform_q(ref_self, task_id, q)
    ours = M.PlanningTaskEnsemble.__new__(M.PlanningTaskEnsemble)
    ours.transforms, ours.horizon, ours.tensor_args = tr, 64, ta
    ours.robot = M.RobotPlanarDisk(tensor_args=ta)
    ours.robot.dt = r[0]["robot"].dt
    g = torch.Generator().manual_seed(11)
    K, S = 6, 5
    t = torch.linspace(0, 1, 64)[None, :, None]
    def chain():
        a, b = torch.rand(K, 1, 2, generator=g) * 1.6 - 0.8, torch.rand(K, 1, 2, generator=g) * 1.6 - 0.8
        pos = a * (1 - t) + b * t + 0.02 * torch.randn(K, 64, 2, generator=g)
        x = torch.cat((pos, torch.randn(K, 64, 2, generator=g) * 0.1), -1)
        return torch.stack([x + 0.01 * s for s in range(S)])
    for coll in ({0: [1, 4], 1: [4]}, {0: [], 1: []}, {0: [0, 1, 2, 3, 4, 5], 1: []}):
        res_ref, res_our = {}, {}
        for m in (0, 1):
            it = chain()
            ci = torch.tensor(coll[m], dtype=torch.long)
            fi = torch.tensor([i for i in range(K) if i not in coll[m]], dtype=torch.long)
            base = dict(trajs_iters=it, trajs_final=it[-1], trajs_final_coll=it[-1][ci] if len(ci) else None,
                        trajs_final_free=it[-1][fi] if len(fi) else None, t_total=0.5)
            res_ref[m] = dict(base, trajs_final_coll_idxs=ci, trajs_final_free_idxs=fi)
            res_our[m] = dict({k: (v.clone() if torch.is_tensor(v) else v) for k, v in base.items()},
                              trajs_final_coll_idxs=ci[:, None].clone(), trajs_final_free_idxs=fi[:, None].clone())
        a = RefPTE.combine_trajs(ref_self, res_ref)
        b = ours.combine_trajs(res_our)
        assert torch.equal(a["trajs_iters"], b["trajs_iters"])
        assert a["trajs_final_coll_idxs"].tolist() == b["trajs_final_coll_idxs"].tolist()
        assert a["trajs_final_free_idxs"].tolist() == b["trajs_final_free_idxs"].tolist()
        assert (a["success_free_trajs"], a["fraction_free_trajs"]) == (b["success_free_trajs"], b["fraction_free_trajs"])
        if a["success_free_trajs"]:
            assert torch.equal(a["trajs_final_free"], b["trajs_final_free"])
            for k in ("cost_smoothness_trajs_final_free", "cost_path_length_trajs_final_free", "cost_all_trajs_final_free",
                      "variance_waypoint_trajs_final_free", "cost_best_free_traj", "traj_final_free_best"):
                assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-7), k
            assert int(a["idx_best_traj"]) == int(b["idx_best_traj"])
