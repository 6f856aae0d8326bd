"""Host-side logic of mmd_b200 that needs no GPU: schedule buffers, env grids, constraint bucketing, lowering."""
import math

import pytest
import torch

import mmd_b200 as M
from mmd_b200.guides import ConstraintSet
from mmd_b200.unet import sinusoidal_table
from oracle import port
from tests.helpers import LIMITS


def test_schedule_buffers_equal_oracle():
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    m = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=25, predict_epsilon=True)
    for k, v in port.make_schedule(25).items():
        assert torch.equal(getattr(m, k), v), k
    sc = m.step_scalars(-1, 20, 0.5)
    assert sc.add_noise == 0 and sc.n_guide_steps == 20 and sc.do_posterior == 1
    assert m.step_scalars(7, 0, 0.5).add_noise == 1


def test_state_dict_keys_match_reference_layout():
    for kw in (dict(dim_mults=(1, 2, 4)), dict(dim_mults=(1, 2, 4, 8)), dict(dim_mults=(1, 2, 4), self_attention=True)):
        unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, **kw)
        want = port.unet_param_shapes(**kw)
        got = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
        assert got == want
    P = port.make_unet_params(seed=0)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    unet.load_state_dict(P, strict=True)
    m = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=25, predict_epsilon=True)
    assert "model.downs.0.0.blocks.0.block.0.weight" in m.state_dict() and "betas" in m.state_dict()


def test_sinusoidal_table_equals_reference_expression():
    tab = sinusoidal_table(25)
    for t in (0, 3, 24):
        assert torch.equal(tab[t], port.sinusoidal_pos_emb(torch.tensor([float(t)]))[0])


def test_env_grids_equal_oracle():
    for name in ["EnvEmpty2D", "EnvConveyor2D", "EnvHighways2D", "EnvDropRegion2D", "EnvEmptyNoWait2D"]:
        e = M.envs.get_env(name + "ExtraObjects")
        sdf, grad = port.build_sdf_grid(name)
        assert torch.equal(e.grid_map_sdf_obj_fixed.sdf_tensor, sdf)
        assert torch.equal(e.grid_map_sdf_obj_fixed.grad_sdf_tensor, grad)
        assert len(e.get_df_obj_list()) == 2


def test_constraint_bucketing():
    H = 64
    qs = torch.tensor([[0.1, 0.2], [0.3, 0.4], [0.5, 0.6]])
    rng = torch.tensor([[3.0, 4.0], [2.0, 6.0], [3.0, 4.0]])
    rad = torch.tensor([0.1, 0.2, 0.3])

    class C:
        pass
    c = C(); c.qs, c.traj_ranges, c.radii = qs, rng, rad
    empty = C(); empty.qs, empty.traj_ranges, empty.radii = torch.zeros(1, 2), torch.tensor([[70.0, 71.0]]), torch.ones(1)
    cs = ConstraintSet([c, empty, c], [0.02, 0.2, 0.02], H, torch.device("cpu"))
    bp = cs.bucket_ptr
    assert bp.shape == (3, H + 1)
    assert cs.cons.shape[0] == 12  # 1 + 4 + 1 entries per copy of c, none for `empty`
    # waypoint 3 of object 0 holds constraints 0, 1, 2 in index order
    e = cs.cons[bp[0, 3]:bp[0, 4]]
    assert torch.equal(e[:, 2], torch.tensor([0.1, 0.2, 0.3]))
    assert bp[0, 2] == 0 and bp[0, 3] == 1 and bp[0, 7] == 6 and bp[0, H] == 6
    assert bp[1, 0] == 6 and bp[1, H] == 6 and bp[2, 0] == 6 and bp[2, H] == 12


def test_lower_env_values():
    ta = {"device": torch.device("cpu"), "dtype": torch.float32}
    env = M.envs.get_env("EnvHighways2DExtraObjects")
    robot = M.RobotPlanarDisk(tensor_args=ta)
    task = M.PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.05, tensor_args=ta)
    ds = M.TrajectoryDataset(env, robot, task, *LIMITS)
    costs = [M.CostCollision(robot, 64, field=f, sigma_coll=1.0) for f in task.get_collision_fields()]
    costs.append(M.CostGPTrajectory(robot, 64, 5.0 / 64, sigma_gp=1.0))
    comp = M.CostComposite(robot, 64, costs, weights_cost_l=[2e-2, 2e-2, 2e-2, 8e-2])
    g = M.GuideManagerTrajectoriesWithVelocity(ds, comp, clip_grad=True)
    e, keep = g.lower_env(torch.device("cpu"))
    o = port.GuideSpec(None, port.LimitsNormalizer(*LIMITS))
    assert e.margin == float(o.margin[0])
    assert (e.ws_min[0], e.ws_max[1]) == (float(o.ws_min[0]), float(o.ws_max[1]))
    assert e.gp_q11 == float(o.q_inv[0, 0]) and e.gp_q12 == float(o.q_inv[0, 2]) and e.gp_q22 == float(o.q_inv[2, 2])
    assert (e.nx, e.ny) == (400, 400) and abs(e.w_smooth - 8e-2) < 1e-8 and e.max_grad_norm == 1.0
    assert e.norm_range[2] == 4.0 and e.norm_min[2] == -2.0


def test_executor_selection():
    assert M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4)).resolve_precision() == "f16x3"
    from mmd_b200 import _lib
    u3 = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    assert u3.native_mode() == _lib.UNET_F16X3 and u3.tensor_core_layers_supported()
    # the 4-level default of scripts/train_diffusion/train.py:35 runs on the per-layer tcgen05 executor (256-channel layers as
    # two 128-channel slices), same numerics class
    u4 = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4, 8))
    assert u4.resolve_precision() == "f16x3" and not u4.tensor_core_supported() and u4.native_mode() == _lib.UNET_F16X3_LAYERS
    assert len(u4._layer_program()) == 4 * 4 + 3 + 4 + 3 * 5 + 2
    # shapes no tensor-core tiling covers fall back to the exact native executor; a refused build is remembered
    u5 = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=48, dim_mults=(1, 2))
    assert u5.resolve_precision() == "fp32" and u5.native_mode() == _lib.UNET_FP32
    u3._note_rejection(_lib.UNET_F16X3)
    assert u3.native_mode() == _lib.UNET_F16X3_LAYERS and u3.resolve_precision() == "f16x3"
    u3._note_rejection(_lib.UNET_F16X3_LAYERS)
    assert u3.resolve_precision() == "fp32"
    assert M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4),
                          self_attention=True).resolve_precision() == "f16x3"   # LinearAttention runs on the tcgen05 executor too
    u = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision="fp32")
    assert u.resolve_precision() == "fp32" and u.resolve_precision("f16x3") == "f16x3"


def test_savgol_matrix_equals_scipy():
    """mmd_b200.smoothing restates scipy's savgol_filter(mode='interp') as one matrix (no scipy at run time)."""
    import numpy as np
    from scipy.signal import savgol_filter
    from mmd_b200.smoothing import savgol_matrix
    for n, w, p in [(64, 10, 2), (64, 11, 2), (20, 10, 2), (64, 9, 3), (10, 10, 2)]:
        assert np.abs(savgol_matrix(n, w, p) - savgol_filter(np.eye(n), w, p, axis=0)).max() < 1e-12, (n, w, p)


def test_global_pad_paths_equals_oracle():
    import mmd_b200 as M
    paths = [torch.randn(60, 4), torch.randn(64, 4), torch.randn(61, 4)]
    starts = [4, 0, 1]
    for a, b in zip(M.global_pad_paths(paths, starts), port.global_pad_paths(paths, starts)):
        assert torch.equal(a, b)


def test_parent_load_state_dict_invalidates_the_native_handle():
    """GaussianDiffusionModel.load_state_dict (mpd.py:167 pattern) recurses through _load_from_state_dict and
    never calls TemporalUnet.load_state_dict; the packed device weights must still be dropped -- by the post hook for any
    load (also assign=True, which replaces the Parameter objects), by the version counters for in-place updates."""
    import unittest.mock as mock
    import mmd_b200 as M
    u = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    m = M.GaussianDiffusionModel(model=u, variance_schedule="exponential", n_diffusion_steps=25, predict_epsilon=True)

    def fake_release(self):
        self._handle, self._handle_key, self._plist = None, None, None

    with mock.patch.object(type(u), "_release", fake_release):
        for kw in ({}, {"assign": True}):
            u._handle, u._handle_key, u._plist = object(), 1, []
            m.load_state_dict(m.state_dict(), **kw)
            assert u._handle is None and u._plist is None, kw
    # in-place update -> the key the handle is compared with changes
    plist = list(u.parameters())
    k0 = tuple((p.data_ptr(), p._version) for p in plist)
    with torch.no_grad():
        plist[3].add_(1.0)
    assert tuple((p.data_ptr(), p._version) for p in plist) != k0


def test_soft_constraints_from_paths_equals_the_reference_loops():
    """f4: vectorised construction == cbs.py:468-508 restated with its loops, ragged start times included, and the
    resulting CostConstraint buckets identically."""
    import mmd_b200 as M
    from mmd_b200.conflicts import soft_constraints_from_paths
    from mmd_b200.guides import ConstraintSet
    g = torch.Generator().manual_seed(3)
    paths = [torch.rand(64, 4, generator=g) * 2 - 1 for _ in range(5)]
    starts = [0, 3, 0, 7, 1]
    for agent in range(5):
        ref = port.create_soft_constraints_from_other_agents_paths(paths, starts, agent)
        out = soft_constraints_from_paths(paths, starts, agent)
        assert len(out) == 1 and out[0].is_soft
        q, rng, rad = ref
        assert torch.equal(out[0].get_q_l(), q) and torch.equal(torch.as_tensor(out[0].get_t_range_l()), rng)
        assert torch.allclose(torch.as_tensor(out[0].radius_l), rad)
        ta = {"device": torch.device("cpu"), "dtype": torch.float32}
        robot = M.RobotPlanarDisk(tensor_args=ta)
        c_vec = M.CostConstraint(robot, 64, q_l=out[0].get_q_l(), traj_range_l=out[0].get_t_range_l(), radius_l=out[0].radius_l,
                                 is_soft=True, tensor_args=ta)
        c_ref = M.CostConstraint(robot, 64, q_l=list(q), traj_range_l=rng.tolist(), radius_l=rad.tolist(), is_soft=True, tensor_args=ta)
        a, b = ConstraintSet([c_vec], [2e-2], 64, torch.device("cpu")), ConstraintSet([c_ref], [2e-2], 64, torch.device("cpu"))
        assert torch.equal(a.bucket_ptr, b.bucket_ptr) and torch.equal(a.cons, b.cons)
    assert soft_constraints_from_paths([], [], 0) == []


def test_cross_cond_entries_match_the_reference_vectors():
    """mmdk_cross_cond entries (include/mmdk.h) carry the rel / bnd vectors of apply_cross_conditioning (sample_functions.py:17-31)
    for both traversal directions of a 1x2 tile grid, restricted to a batch-row range."""
    from mmd_b200.diffusion import cross_cond_entries
    for tr in ({0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}, {0: torch.tensor([2.0, 0.0]), 1: torch.tensor([0.0, 0.0])},
               {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([0.0, -2.0])}):
        (cc,) = cross_cond_entries({(0, 1): (63, 0)}, tr, 4, 128, 192)
        rel = torch.cat([tr[1] - tr[0], torch.zeros(2)])
        bnd = rel / torch.norm(rel, keepdim=True)
        bnd[bnd == 0] = 1e6
        assert (cc.m1, cc.m2, cc.ind1, cc.ind2, cc.row_lo, cc.row_hi) == (0, 1, 63, 0, 128, 192)
        assert list(cc.rel) == rel.tolist() and list(cc.bnd) == [float(torch.tensor(v, dtype=torch.float32)) for v in bnd.tolist()]
        # the oracle's stitch with these vectors == the reference expression
        x = {0: torch.randn(3, 64, 4), 1: torch.randn(3, 64, 4)}
        y = port.apply_cross_conditioning({m: v.clone() for m, v in x.items()}, {(0, 1): (63, 0)}, tr)
        r, b = torch.tensor(list(cc.rel)), torch.tensor(list(cc.bnd))
        a1 = torch.min(x[1][:, 0, :] + r, b)
        assert torch.equal(y[0][:, 63, :], a1) and torch.equal(y[1][:, 0, :], torch.max(a1 - r, -b))


@pytest.mark.parametrize("dim_mults", [(1, 2, 4), (1, 2, 4, 8), (1, 2)])
def test_layer_program_mirrors_the_oracle_forward(dim_mults):
    """TemporalUnet._layer_program() (the shapes the executor-selection logic reasons about) lists exactly the ops the oracle's
    forward executes (temporal_unet.py:121-174), in order, with the same output channels and lengths."""
    P = port.make_unet_params(seed=1, dim_mults=dim_mults)
    taps = {}
    port.unet_forward(P, torch.randn(2, 64, 4), torch.zeros(2, dtype=torch.long), taps=taps)
    u = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=dim_mults)
    prog = u._layer_program()
    assert prog[-1][0] == "final" and len(prog) == len(taps["ops"]) + 1
    for (kind, srcs, res_srcs, cout, L), (name, act) in zip(prog, taps["ops"]):
        l_out = L // 2 if kind == "down" else (2 * L if kind == "up" else L)
        assert (act.shape[1], act.shape[2]) == (cout, l_out), (name, kind, cout, L)
        # a block carries the RTB's 1x1 residual conv exactly where the checkpoint has one
        if kind == "block" and name.endswith(".blocks.1"):
            assert bool(res_srcs) == (name[:-len(".blocks.1")] + ".residual_conv.weight" in P), name
