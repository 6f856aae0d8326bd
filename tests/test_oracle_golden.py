"""Pins oracle/port.py to the golden vectors dumped from the REAL reference (oracle/gen_golden.py).  Runs anywhere."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import port
from oracle.gen_golden import ENVS, golden_inputs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
INP = golden_inputs()
T = 25


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def params():
    return port.make_unet_params(seed=0)


@pytest.fixture(scope="module")
def grids():
    return {e: port.build_sdf_grid(e) for e in ENVS}


def test_schedule_bit_exact():
    s = port.make_schedule(T)
    for k, v in s.items():
        assert torch.equal(v, t(G[f"sched/{k}"])), k


@pytest.mark.parametrize("env", ENVS)
def test_sdf_grids_bit_exact(grids, env):
    sdf, grad = grids[env]
    assert hashlib.sha256(sdf.numpy().tobytes()).digest() == G[f"{env}/sdf_sha"].tobytes()
    assert hashlib.sha256(grad.numpy().tobytes()).digest() == G[f"{env}/grad_sha"].tobytes()
    assert torch.equal(sdf[::16, ::16], t(G[f"{env}/sdf_sub"]))


def test_unet_eps(params):
    for tt in (0, 12, 24):
        eps = port.unet_forward(params, INP["unet_x"], torch.full((2,), tt, dtype=torch.long))
        assert torch.allclose(eps, t(G[f"unet/eps_t{tt}"]), rtol=0, atol=1e-6)


@pytest.mark.parametrize("env", ENVS)
def test_guide_gradients(grids, env):
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    gs = port.GuideSpec(port.GridSDF(*grids[env]), norm)
    assert torch.allclose(gs(INP["guide_x"]), t(G[f"{env}/guide"]), rtol=0, atol=1e-7)
    assert torch.allclose(gs(INP["guide_x_small"]), t(G[f"{env}/guide_small"]), rtol=0, atol=1e-7)
    gs.extra = [port.Constraint(INP["cons_q"], INP["cons_rng"], INP["cons_rad"], True),
                port.Constraint(INP["hard_q"], INP["hard_rng"], INP["hard_rad"], False)]
    assert torch.allclose(gs(INP["guide_x"]), t(G[f"{env}/guide_cons"]), rtol=0, atol=1e-7)


def test_full_guided_chain(params, grids):
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    gs = port.GuideSpec(port.GridSDF(*grids["EnvHighways2D"]), norm)
    gs.extra = [port.Constraint(INP["cons_q"], INP["cons_rng"], INP["cons_rad"], True)]
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), norm)
    chain = port.run_inference(port.DiffusionModel(params, T), hc, 6, INP["chain_noise"], guide=gs)
    assert chain.shape[0] == T + 2
    assert torch.allclose(chain[-1], t(G["chain/final"]), rtol=0, atol=1e-5)
    assert torch.allclose(chain[[0, 1, 13, 25]], t(G["chain/frames"]), rtol=0, atol=1e-5)


def test_integer_outputs(grids):
    grid = port.GridSDF(*grids["EnvHighways2D"])
    assert torch.equal(grid.cell_index(INP["cell_pts"]), t(G["cell_idx"]))
    coll, mid = port.check_rr_collisions(INP["rr_pos"])
    assert torch.equal(coll, t(G["rr/coll"]))
    assert torch.equal(torch.isnan(mid), torch.isnan(t(G["rr/mid"])))
    free_idx, wp = port.get_trajs_free_idxs(INP["cls_trajs"], grid)
    assert torch.equal(wp, t(G["cls/wp"]))
    assert torch.equal(free_idx, t(G["cls/free_idxs"]))


def test_cross_conditioning():
    g = torch.Generator().manual_seed(99)
    x = {0: torch.randn(5, 64, 4, generator=g), 1: torch.randn(5, 64, 4, generator=g)}
    tr = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    xr = port.apply_cross_conditioning(x, {(0, 1): (-1, 0)}, tr)
    assert torch.equal(xr[0], t(G["cross/x0"])) and torch.equal(xr[1], t(G["cross/x1"]))
