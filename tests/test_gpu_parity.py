"""-m gpu parity tests: mmd_b200 (CUDA, through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances: fp32 UNet (FFMA accumulation order differs from oneDNN) 2e-5 relative per forward; guide gradients 1e-5
absolute (they are clipped to unit norm); whole chains 1e-3 relative L2 per trajectory (north star) -- measured values
are printed.  Integer outputs (cell indices, rr collisions, free masks) are compared bit-exact.
"""
import math
import os

import pytest
import torch

from oracle import port
from tests.helpers import (LIMITS, build_oracle, build_product, max_err, random_constraints, rel_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(dev):
    o = build_oracle("EnvHighways2D", T=25)
    p = build_product(dev, "EnvHighways2D", T=25, P=o["P"])
    return o, p


def test_cond_table_matches_time_mlp(pair, dev):
    import ctypes as C
    from mmd_b200 import _lib
    o, p = pair
    P = o["P"]
    import torch.nn.functional as F
    for t in (0, 7, 24):
        temb = port.sinusoidal_pos_emb(torch.tensor([float(t)]))
        temb = F.linear(temb, P["time_mlp.encoder.1.weight"], P["time_mlp.encoder.1.bias"])
        temb = F.linear(F.mish(temb), P["time_mlp.encoder.3.weight"], P["time_mlp.encoder.3.bias"])
        ref = F.linear(F.mish(temb), P["downs.0.0.cond_mlp.1.weight"], P["downs.0.0.cond_mlp.1.bias"])[0]
        n = C.c_int()
        h = p["unet"].native()
        _lib.check(_lib.lib().mmdk_unet_cond_row(h, t, None, C.byref(n), _lib.stream_ptr()))
        row = torch.empty(n.value, device=dev)
        _lib.check(_lib.lib().mmdk_unet_cond_row(h, t, _lib.ptr(row), C.byref(n), _lib.stream_ptr()))
        assert max_err(row[:32], ref) < 1e-5


@pytest.mark.parametrize("B", [1, 5, 7])
def test_unet_fp32_matches_oracle(pair, dev, B):
    o, p = pair
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 64, 4, generator=g)
    for t in (0, 11, 24):
        ref = port.unet_forward(o["P"], x, torch.full((B,), t, dtype=torch.long))
        out = p["unet"].forward_t(x.to(dev), t, precision="fp32")
        e = rel_err(out, ref)
        print(f"unet fp32 B={B} t={t} rel_err={e:.3e}")
        assert e < 2e-5


@pytest.mark.parametrize("B", [3, 7, 20, 200, 1100])
def test_unet_tensor_core_matches_oracle(pair, dev, B):
    """tcgen05 executor (FP16 hi/lo split, 3 MMAs, fp32 accumulate in TMEM): every layer op through the debug tap,
    then eps.  Expected ~2e-6 (CPU emulation of the operand format); bar 5e-5.  The batch sizes cover
    the three tilings of the persistent executor: 1 sample per tile (B <= 148), 3 (B <= 444) and 7."""
    import ctypes as C
    from mmd_b200 import _lib
    o, p = pair
    x = torch.randn(B, 64, 4, generator=torch.Generator().manual_seed(100 + B))
    for t in (3, 24):
        taps = {}
        ref = port.unet_forward(o["P"], x, torch.full((B,), t, dtype=torch.long), taps=taps)
        # production mode hands activations over inside shared memory; the tap mode also stores them: same eps, bit for bit
        fast = p["unet"].forward_t(x.to(dev), t, precision="f16x3").clone()
        h = p["unet"].native()
        _lib.check(_lib.lib().mmdk_unet_debug_keep_activations(h, 1))
        out = p["unet"].forward_t(x.to(dev), t, precision="f16x3")
        torch.cuda.synchronize()
        assert torch.equal(fast, out)
        worst = 0.0
        for j, (name, act) in enumerate(taps["ops"]):
            c, l = C.c_int(), C.c_int()
            _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, None, C.byref(c), C.byref(l), None, _lib.stream_ptr()))
            assert (c.value, l.value) == (act.shape[1], act.shape[2]), name
            buf = torch.empty(B, c.value, l.value, device=dev)
            _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, _lib.ptr(buf), C.byref(c), C.byref(l), None, _lib.stream_ptr()))
            e = rel_err(buf, act)
            worst = max(worst, e)
            assert e < 5e-5, f"op {j} ({name}) rel_err {e:.3e}"
        _lib.check(_lib.lib().mmdk_unet_debug_keep_activations(h, 0))
        e = rel_err(out, ref)
        print(f"unet f16x3 B={B} t={t}: eps rel_err={e:.3e}, worst layer {worst:.3e}")
        assert e < 5e-5


def test_weights_reloaded_through_the_parent_module(dev):
    """mpd.py:167 pattern AFTER a forward has built the native handle: diffusion_model.load_state_dict(ckpt) must reach the
    device copy of the weights, as must an in-place parameter update."""
    o = build_oracle("EnvEmpty2D", T=25)
    p = build_product(dev, "EnvEmpty2D", T=25, P=o["P"])
    x = torch.randn(4, 64, 4, generator=torch.Generator().manual_seed(0))
    t = torch.full((4,), 7, dtype=torch.long)
    for prec in ("fp32", "f16x3"):
        p["unet"].forward_t(x.to(dev), 7, precision=prec)                 # handle + executor state exist now
    P2 = port.make_unet_params(seed=5)
    missing, unexpected = p["model"].load_state_dict({f"model.{k}": v for k, v in P2.items()}, strict=False)
    assert not unexpected
    ref2 = port.unet_forward(P2, x, t)
    for prec in ("fp32", "f16x3"):
        assert rel_err(p["unet"].forward_t(x.to(dev), 7, precision=prec), ref2) < 5e-5, prec
    with torch.no_grad():
        p["unet"].final_conv[1].bias.add_(0.25)
    P3 = {k: v.clone() for k, v in P2.items()}
    P3["final_conv.1.bias"] += 0.25
    assert rel_err(p["unet"].forward_t(x.to(dev), 7, precision="f16x3"), port.unet_forward(P3, x, t)) < 5e-5


@pytest.mark.parametrize("B", [5, 200, 1100])
def test_unet_linear_attention(dev, B):
    """self_attention=True: Residual(PreNorm(LinearAttention)) after every level (layers.py:177-229).  fp32 executor, and the
    persistent tcgen05 executor: to_qkv / to_out 1x1 convs as tensor-core GEMMs (PreNorm folded into the weights and undone
    per row in the epilogue), softmax / context / context^T q in a CUDA-core op between them; all three tilings."""
    import mmd_b200 as M
    P = port.make_unet_params(seed=2, self_attention=True)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4), self_attention=True)
    unet.load_state_dict(P, strict=True)
    unet = unet.to(dev)
    assert unet.resolve_precision("auto") == "f16x3"
    x = torch.randn(B, 64, 4, generator=torch.Generator().manual_seed(8))
    for t in (0, 17):
        ref = port.unet_forward(P, x, torch.full((B,), t, dtype=torch.long))
        for prec, bar in (("fp32", 2e-5), ("f16x3", 5e-5)):
            if prec == "fp32" and B > 5:
                continue
            out = unet.forward_t(x.to(dev), t, precision=prec)
            e = rel_err(out, ref)
            print(f"unet {prec} + linear attention B={B} t={t} rel_err={e:.3e}")
            assert e < bar


@pytest.mark.parametrize("B", [3, 20, 300])
def test_unet_dim_mults_option1(dev, B):
    """dim_mults (1,2,4,8) (temporal_unet.py:17-20; the default of scripts/train_diffusion/train.py:35): exact fp32 executor,
    and the per-layer tcgen05 executor -- 256-channel conv blocks at L = 8 as two 128-channel slices (4 GroupNorm groups each),
    the 512-channel concat block streaming its two source images through shared memory in two phases.  Every layer tapped."""
    import ctypes as C
    from mmd_b200 import _lib
    o = build_oracle("EnvEmpty2D", T=25, dim_mults=(1, 2, 4, 8))
    p = build_product(dev, "EnvEmpty2D", T=25, P=o["P"], dim_mults=(1, 2, 4, 8))
    unet = p["unet"]
    assert unet.resolve_precision("auto") == "f16x3" and unet.native_mode("auto") == _lib.UNET_F16X3_LAYERS
    x = torch.randn(B, 64, 4, generator=torch.Generator().manual_seed(3))
    for t in (4, 19):
        taps = {}
        ref = port.unet_forward(o["P"], x, torch.full((B,), t, dtype=torch.long), taps=taps)
        if B <= 20:
            assert rel_err(unet.forward_t(x.to(dev), t, precision="fp32"), ref) < 2e-5
        out = unet.forward_t(x.to(dev), t, precision="f16x3")
        h = unet.native()
        worst = 0.0
        for j, (name, act) in enumerate(taps["ops"]):            # every op but the final 1x1 conv (eps, checked below)
            c, l = C.c_int(), C.c_int()
            _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, None, C.byref(c), C.byref(l), None, _lib.stream_ptr()))
            assert (c.value, l.value) == (act.shape[1], act.shape[2]), name
            buf = torch.empty(B, c.value, l.value, device=dev)
            _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, _lib.ptr(buf), C.byref(c), C.byref(l), None, _lib.stream_ptr()))
            e = rel_err(buf, act)
            worst = max(worst, e)
            assert e < 5e-5, f"op {j} ({name}) rel_err {e:.3e}"
        e = rel_err(out, ref)
        print(f"unet (1,2,4,8) f16x3 (per-layer tcgen05) B={B} t={t}: eps rel_err={e:.3e}, worst layer {worst:.3e}")
        assert e < 5e-5
    # 'auto' picks the same executor, and a chain runs on it
    assert torch.equal(unet.forward_t(x.to(dev), 19, precision="auto"), out)


def test_dim_mults_option1_chain(dev):
    """The 4-level network through the whole reverse chain on its tensor-core (per-layer tcgen05) executor: run_inference
    (native chain, issued directly) against the oracle with the GP term off, and the same planner call through the
    graph-captured batched sampler (MultiRobotSampler independent mode: PDL launches inside a CUDA graph) == run_inference
    bit for bit."""
    import mmd_b200 as M
    T, K = 25, 8
    o = build_oracle("EnvHighways2D", T=T, dim_mults=(1, 2, 4, 8), w_smooth=0.0)
    p = build_product(dev, "EnvHighways2D", T=T, P=o["P"], dim_mults=(1, 2, 4, 8), w_smooth=0.0, precision="auto")
    assert p["unet"].native_mode("auto") == M._lib.UNET_F16X3_LAYERS
    noise = torch.randn(T + 2, K, 64, 4, generator=torch.Generator().manual_seed(18))
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), o["norm"])
    ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
    kw = dict(n_samples=K, horizon=64, return_chain=True, sample_fn=M.ddpm_sample_fn, n_guide_steps=20,
              t_start_guide=math.ceil(0.5 * T), noise_std_extra_schedule_fn=lambda x: 0.5, n_diffusion_steps_without_noise=1)
    hcd = {k: v.to(dev) for k, v in hc.items()}
    chain = p["model"].run_inference(None, hcd, guide=p["guide"], noise=noise.to(dev), **kw)
    fin = _per_traj(chain[-1], ref[-1])
    print(f"chain (1,2,4,8) f16x3 K={K} T={T} w_smooth=0: final rel L2 median={float(fin.median()):.2e} max={float(fin.max()):.2e}")
    assert torch.isfinite(chain).all()
    assert float(fin.median()) < 1e-4 and float(fin.quantile(0.9)) < 1e-3 and float(fin.max()) < 2e-2
    smp = M.MultiRobotSampler(p["model"], p["guide"], n_guide_steps=20, t_start_guide=math.ceil(0.5 * T), noise_std=0.5,
                              n_diffusion_steps_without_noise=1)
    for _ in range(2):   # capture, then replay
        out, ch = smp.sample([hcd], K, noise=noise.to(dev)[None],
                             mode="independent", return_chain=True)
        assert torch.equal(ch[0], chain)


def test_graph_replay_follows_weight_reloads_and_state_eviction(dev):
    """The batched sampler replays a captured CUDA graph that bakes in device pointers of the UNet handle (weights, cond table)
    and of the executor state of its batch size (activation images).  Both can be destroyed and re-created: a weight reload
    rebuilds the handle, and forwards at many other batch sizes evict the bounded per-batch-size state cache.  The graph key
    carries their unique ids, so a stale graph is never replayed."""
    import mmd_b200 as M
    T, K = 25, 8
    o = build_oracle("EnvEmpty2D", T=T, w_smooth=0.0)
    p = build_product(dev, "EnvEmpty2D", T=T, P=o["P"], w_smooth=0.0, precision="f16x3")
    hc = {k: v.to(dev) for k, v in port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), o["norm"]).items()}
    noise = torch.randn(T + 2, K, 64, 4, device=dev, generator=torch.Generator(device=dev).manual_seed(4))   # step-major: stable pointer
    mk = lambda: M.MultiRobotSampler(p["model"], p["guide"], n_guide_steps=20, t_start_guide=math.ceil(0.5 * T), noise_std=0.5,
                                     n_diffusion_steps_without_noise=1)
    smp = mk()
    run = lambda s: s.sample([hc], K, noise=noise, mode="independent").clone()
    a1, a2 = run(smp), run(smp)            # capture, replay
    assert torch.equal(a1, a2)
    # (1) executor states of other batch sizes push this one out of the bounded cache; it is rebuilt with new buffers
    for B in (3, 5, 6, 7, 9, 10):
        p["unet"].forward_t(torch.randn(B, 64, 4, device=dev), 3, precision="f16x3")
    torch.cuda.synchronize()
    junk = [torch.full((1 << 20,), float("nan"), device=dev) for _ in range(64)]   # recycle freed blocks with poison
    a3 = run(smp)
    assert torch.equal(a3, a1), "replayed a graph over an evicted executor state"
    del junk
    # (2) new weights through the parent module: the native handle is rebuilt (often at the same host address)
    P2 = port.make_unet_params(seed=5)
    p["model"].load_state_dict({f"model.{k}": v for k, v in P2.items()}, strict=False)
    b1 = run(smp)
    fresh = run(mk())
    assert not torch.equal(b1, a1), "replayed a graph over the old weights"
    assert torch.equal(b1, fresh)


def test_cell_index_bit_exact(pair, dev):
    import ctypes as C
    from mmd_b200 import _lib
    o, p = pair
    g = torch.Generator().manual_seed(1)
    pts = torch.rand(20000, 2, generator=g) * 2.4 - 1.2  # includes out-of-bounds
    edges = (torch.arange(0, 401).float() / 400 * 2 - 1)  # points straddling cell edges
    e2 = torch.stack((edges, edges.flip(0)), -1)
    pts = torch.cat((pts, e2, torch.nextafter(e2, torch.tensor(2.0)), torch.nextafter(e2, torch.tensor(-2.0))))
    ref = o["guide"].grid.cell_index(pts)
    env, keep = p["guide"].lower_env(dev)
    out = torch.empty(pts.shape[0], 2, dtype=torch.int32, device=dev)
    pd = pts.to(dev).contiguous()
    _lib.check(_lib.lib().mmdk_cell_index(C.byref(env), _lib.ptr(pd), pts.shape[0], _lib.ptr(out), _lib.stream_ptr()))
    assert torch.equal(out.cpu(), ref.to(torch.int32))


@pytest.mark.parametrize("K,scale", [(16, 0.4), (16, 0.9), (40, 0.5), (128, 0.45)])
def test_guide_grad_matches_oracle(pair, dev, K, scale):
    """One GuideManager.forward evaluation incl. extra CostConstraint objects; scale 0.9 makes the global clip fire,
    K=40/128 exercises the cluster (DSMEM flag exchange) path."""
    import mmd_b200 as M
    o, p = pair
    g = torch.Generator().manual_seed(K)
    x = torch.randn(K, 64, 4, generator=g) * scale
    qs, rng, rad = random_constraints(200, seed=K)
    hard_q, hard_rng, hard_rad = torch.tensor([[0.1, 0.2]]), torch.tensor([[10., 15.]]), torch.tensor([0.12])
    o["guide"].extra = [port.Constraint(qs, rng, rad, True), port.Constraint(hard_q, hard_rng, hard_rad, False)]
    ref, parts = o["guide"](x, return_parts=True)
    o["guide"].extra = []
    ta = p["ta"]
    cc = [M.CostConstraint(p["robot"], 64, q_l=list(qs), traj_range_l=rng.tolist(), radius_l=rad.tolist(),
                           is_soft=True, tensor_args=ta),
          M.CostConstraint(p["robot"], 64, q_l=list(hard_q), traj_range_l=hard_rng.tolist(), radius_l=hard_rad.tolist(),
                           is_soft=False, tensor_args=ta)]
    p["guide"].add_extra_costs(cc, [2e-2, 2e-1])
    try:
        out, raw = p["guide"](x.to(dev), return_raw=True)
    finally:
        p["guide"].reset_extra_costs()
    for k, (g_raw, _) in enumerate(parts):
        scale_k = float(g_raw.abs().max()) + 1e-12
        assert max_err(raw[k], g_raw) / scale_k < 2e-5, f"raw gradient of cost {k}"
    e = max_err(out, ref)
    print(f"guide K={K} scale={scale} max_abs_err={e:.3e} (|grad| max {float(ref.abs().max()):.3f})")
    assert e < 1e-6


@pytest.mark.parametrize("env_name", ["EnvEmpty2D", "EnvConveyor2D", "EnvDropRegion2D"])
def test_guide_other_envs(dev, env_name):
    o = build_oracle(env_name, T=25)
    p = build_product(dev, env_name, T=25, P=o["P"])
    x = torch.randn(12, 64, 4, generator=torch.Generator().manual_seed(5)) * 0.5
    assert max_err(p["guide"](x.to(dev)), o["guide"](x)) < 1e-6


def test_guide_gradient_steps_block(pair, dev):
    """20 fused guide steps (sample_functions.py:89-107)."""
    import mmd_b200 as M
    o, p = pair
    x = torch.randn(16, 64, 4, generator=torch.Generator().manual_seed(9)) * 0.5
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), o["norm"])
    ref = port.guide_gradient_steps(x.clone(), port.repeat_hard_conds(hc, 16), o["guide"], 20)
    out = M.guide_gradient_steps(x.to(dev), {k: v.to(dev) for k, v in hc.items()}, p["guide"], 20)
    e = rel_err(out, ref)
    print(f"20 guide steps rel_err={e:.3e}")
    assert e < 1e-5


def _chain_problem(dev, K, T, out_scale, w_smooth, precision="fp32"):
    import mmd_b200 as M
    o = build_oracle("EnvHighways2D", T=T, out_scale=out_scale, w_smooth=w_smooth)
    p = build_product(dev, "EnvHighways2D", T=T, P=o["P"], w_smooth=w_smooth, precision=precision)
    noise = torch.randn(T + 2, K, 64, 4, generator=torch.Generator().manual_seed(18))
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), o["norm"])
    qs, rng, rad = random_constraints(120, seed=3)
    o["guide"].extra = [port.Constraint(qs, rng, rad, True)]
    ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
    cc = M.CostConstraint(p["robot"], 64, q_l=list(qs), traj_range_l=rng.tolist(), radius_l=rad.tolist(), is_soft=True,
                          tensor_args=p["ta"])
    p["guide"].add_extra_costs([cc], [2e-2])
    return o, p, noise, hc, ref


def _per_traj(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return (a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1)


@pytest.mark.parametrize("K,T,out_scale", [(8, 25, 1.0), (8, 25, 0.05), (24, 50, 1.0)])
def test_run_inference_chain_free_running(dev, K, T, out_scale):
    """Whole reverse chain through GaussianDiffusionModel.run_inference on the oracle's noise, free running.
    With the reference's default smoothness weight (8e-2, clipped GP gradient) the guided phase is CHAOTIC: a 1e-7
    relative perturbation of eps moves the oracle's own final trajectories by ~10% ('chaos').  The
    whole-chain bar (1e-3 per trajectory, north star) is therefore asserted (a) on the unguided prefix with default
    weights and (b) on the complete guided chain with the GP term off; the default-weight guided chain is covered
    step by step in test_run_inference_chain_teacher_forced."""
    import mmd_b200 as M
    kw = dict(n_samples=K, horizon=64, return_chain=True, sample_fn=M.ddpm_sample_fn, n_guide_steps=20,
              t_start_guide=math.ceil(0.5 * T), noise_std_extra_schedule_fn=lambda x: 0.5,
              n_diffusion_steps_without_noise=1)
    for w_smooth in (0.0, 8e-2):
        o, p, noise, hc, ref = _chain_problem(dev, K, T, out_scale, w_smooth)
        chain = p["model"].run_inference(None, {k: v.to(dev) for k, v in hc.items()}, guide=p["guide"],
                                         noise=noise.to(dev), **kw)
        assert chain.shape == ref.shape
        assert torch.isfinite(chain).all()
        n_unguided = 1 + (T - math.ceil(0.5 * T))  # frames produced before the first guided step
        pre = _per_traj(chain[n_unguided], ref[n_unguided])
        fin = _per_traj(chain[-1], ref[-1])
        print(f"chain K={K} T={T} out_scale={out_scale} w_smooth={w_smooth}: end of unguided prefix rel L2 "
              f"median={float(pre.median()):.2e} max={float(pre.max()):.2e}; "
              f"final rel L2 max={float(fin.max()):.2e} median={float(fin.median()):.2e}")
        # the first reverse steps multiply eps differences by c1*coef1 ~ 1e3 on the few elements whose x0 is not
        # saturated by the clamp, so isolated trajectories sit above the median
        assert float(pre.median()) < 1e-4 and float(pre.quantile(0.9)) < 1e-3 and float(pre.max()) < 2e-2
        if w_smooth == 0.0:
            # discontinuities (nearest SDF cell, hinge, in/out of a constraint radius) can flip on a 1-ulp difference
            # and move one waypoint by <= weight (2e-2): SURVEY hard part (b) -> 1e-3 bar on >= 90% of the
            # trajectories, every trajectory within a handful of single-branch flips
            assert float(fin.quantile(0.9)) < 1e-3 and float(fin.median()) < 1e-4
            assert float(fin.max()) < 2e-2
        else:
            assert float(fin.max()) < 0.6  # chaos envelope only -- not a parity claim


@pytest.mark.parametrize("K", [16, 128])
def test_run_inference_chain_free_running_tensor_core(dev, K):
    """The default (tcgen05, FP16 hi/lo) executor on the complete guided chain with the GP term off, T=100 (bench
    schedule).  K=128 (the bench's samples per robot) makes the north-star statement a real percentile.  The bar is the
    REFERENCE's own: on this very problem the oracle perturbed by 1e-7 / 1.3e-6 / 3e-6 relative in eps keeps only 92.2% /
    93.8% / 93.0% of its 128 trajectories within 1e-3 of itself (median 7e-7 .. 2e-6, max 2e-2; measured with
    tests/test_oracle_self_sensitivity.py::test_free_running_k128, MMD_SLOW=1) -- the rest are trajectories where a 1-ulp
    difference flipped a nearest-cell / hinge / in-radius branch somewhere in 51 x 20 guide evaluations.  Measured on
    B200: 88.3% (113 of 128; binomial sigma at p = 0.93 is 2.3%), median 5e-6."""
    import mmd_b200 as M
    T = 100
    o, p, noise, hc, ref = _chain_problem(dev, K, T, 1.0, 0.0, "f16x3")
    chain = p["model"].run_inference(None, {k: v.to(dev) for k, v in hc.items()}, guide=p["guide"], noise=noise.to(dev),
                                     n_samples=K, horizon=64, return_chain=True, sample_fn=M.ddpm_sample_fn,
                                     n_guide_steps=20, t_start_guide=math.ceil(0.5 * T),
                                     noise_std_extra_schedule_fn=lambda x: 0.5, n_diffusion_steps_without_noise=1)
    fin = _per_traj(chain[-1], ref[-1])
    frac = float((fin < 1e-3).float().mean())
    print(f"chain f16x3 K={K} T={T} w_smooth=0: final rel L2 median={float(fin.median()):.2e} "
          f"p90={float(fin.quantile(0.9)):.2e} max={float(fin.max()):.2e}; {100 * frac:.1f}% of the trajectories below 1e-3")
    assert float(fin.median()) < 1e-4 and float(fin.max()) < 1e-1
    # K=16 cannot resolve a 90th percentile (two flipped trajectories = 12.5%); K=128: the oracle's own 92-94% minus 3 sigma
    assert frac >= (0.85 if K >= 64 else 0.75)


@pytest.mark.parametrize("K,T,precision", [(8, 25, "fp32"), (16, 50, "fp32"), (8, 25, "f16x3"), (16, 100, "f16x3")])
def test_run_inference_chain_teacher_forced(dev, K, T, precision):
    """Every reverse step of the default-weight guided chain, one at a time: x_k of the oracle goes in, x_{k+1} must
    come out (ddpm_sample_fn + the trailing hard conditioning).  Bar: 1e-3 relative L2 per trajectory on every unguided
    step; on the guided steps the same bar for all but isolated (step, trajectory) pairs -- the ORACLE perturbed by the
    executor's eps error (1.3e-6 / 3e-6 relative) deviates from itself in exactly this way: median 5e-7, 2 of 416 pairs
    above 1e-3, max 1e-2 (tests/test_oracle_self_sensitivity.py computes that on the CPU): one flipped branch (SDF cell,
    hinge, constraint radius) is amplified by the 20 clipped GP steps of the same timestep."""
    import mmd_b200 as M
    o, p, noise, hc, ref = _chain_problem(dev, K, T, 1.0, 8e-2, precision)
    hcd = {k: v.to(dev).reshape(1, -1).repeat(K, 1) for k, v in hc.items()}
    t_start = math.ceil(0.5 * T)
    errs = []
    k = 1
    for i in reversed(range(-1, T)):
        x_in = ref[k - 1].to(dev)
        t = torch.full((K,), i, dtype=torch.long)
        x_out, _ = M.ddpm_sample_fn(p["model"], x_in, hcd, None, t, guide=p["guide"], n_guide_steps=20,
                                    t_start_guide=t_start, noise_std_extra_schedule_fn=lambda x: 0.5,
                                    noise=noise[k].to(dev))
        x_out = M.apply_hard_conditioning(x_out, hcd)
        e = _per_traj(x_out, ref[k])
        if i >= t_start:
            assert float(e.max()) < 1e-3, f"unguided step t={i}: {float(e.max())}"
        errs.append(e)
        k += 1
    e = torch.stack(errs)
    guided = e[T - t_start:]
    n_out = int((guided > 1e-3).sum())
    print(f"teacher-forced {precision} K={K} T={T}: all steps median {float(e.median()):.2e}; guided steps p99 "
          f"{float(guided.flatten().quantile(0.99)):.2e} max {float(guided.max()):.2e}, pairs above 1e-3: {n_out} of {guided.numel()}")
    assert float(e.median()) < 1e-5
    assert n_out <= max(2, int(0.015 * guided.numel())) and float(guided.max()) < 5e-2


def test_run_local_inference(dev):
    import mmd_b200 as M
    T, K = 25, 8
    o = build_oracle("EnvConveyor2D", T=T, w_smooth=0.0)  # free running: GP term off (chaos)
    p = build_product(dev, "EnvConveyor2D", T=T, P=o["P"], w_smooth=0.0)
    g = torch.Generator().manual_seed(4)
    seed = torch.randn(K, 64, 4, generator=g) * 0.3
    noise = torch.randn(3 + 2, K, 64, 4, generator=g)
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, -0.6]), torch.tensor([0.8, 0.6]), o["norm"])
    ref = port.run_local_inference(o["model"], seed, 3, 3, hc, K, noise, guide=o["guide"])
    out = p["model"].run_local_inference(seed.to(dev), 3, 3, None, {k: v.to(dev) for k, v in hc.items()}, n_samples=K,
                                         horizon=64, return_chain=True, sample_fn=M.ddpm_sample_fn, guide=p["guide"],
                                         n_guide_steps=20, t_start_guide=math.ceil(0.5 * T),
                                         noise_std_extra_schedule_fn=lambda x: 0.5, n_diffusion_steps_without_noise=1,
                                         noise=noise.to(dev))
    assert out.shape == ref.shape
    assert rel_err(out[-1], ref[-1]) < 1e-4


def _lockstep_report(name, chain, ref, frames, T):
    """chain/ref: [R, n_frames, K, H, D] at the same chain frames.  Envelope: the unguided prefix is
    held to the 1e-3 bar; once the robots couple through the unit-direction repulsion the ORACLE perturbed by 1.3e-6 in eps
    deviates from itself by median 2.4e-4 / p90 3.3e-3 / max 5.2e-3 at the end of a T=25 chain."""
    R, nF, K = chain.shape[:3]
    errs = {}
    for fi, f in enumerate(frames):
        e = _per_traj(chain[:, fi].reshape(R * K, 64, 4), ref[:, fi].reshape(R * K, 64, 4))
        errs[f] = e
        print(f"lockstep {name} frame {f}: per-trajectory rel L2 median={float(e.median()):.2e} p90={float(e.quantile(0.9)):.2e} "
              f"max={float(e.max()):.2e}")
    n_unguided = 1 + (T - math.ceil(0.5 * T))
    pre = errs[n_unguided]
    assert float(pre.median()) < 1e-4 and float(pre.quantile(0.9)) < 1e-3 and float(pre.max()) < 2e-2
    first = errs[n_unguided + 1]     # first guided frame: one guided step after a (nearly) identical state
    assert float(first.median()) < 1e-4 and float(first.quantile(0.9)) < 1e-3
    fin = errs[frames[-1]]
    assert float(fin.median()) < 1e-3 and float(fin.quantile(0.9)) < 1e-2 and float(fin.max()) < 1e-1


@pytest.mark.parametrize("R,K,T,env_name", [(3, 4, 25, "EnvEmpty2D"), (2, 8, 50, "EnvEmpty2D")])
def test_lockstep_matches_oracle(dev, R, K, T, env_name):
    """Lock-step fleet against the oracle's composition of reference calls, whole chain, GP term off.
    (2, 8, 50, Empty) is BASELINE.json config 1 exactly."""
    import mmd_b200 as M
    o = build_oracle(env_name, T=T, w_smooth=0.0)
    p = build_product(dev, env_name, T=T, P=o["P"], w_smooth=0.0)
    starts, goals = port.get_start_goal_pos_circle(R, 0.3)  # close together so the peer term is active
    hcs = [port.hard_conds_from_start_goal(s, g, o["norm"]) for s, g in zip(starts, goals)]
    noise = torch.randn(R, T + 2, K, 64, 4, generator=torch.Generator().manual_seed(7))
    guides = [port.GuideSpec(o["guide"].grid, o["norm"], w_smooth=0.0) for _ in range(R)]
    ref = port.lockstep_sample(o["model"], guides, hcs, K, noise, return_chain=True)
    n_unguided = 1 + (T - math.ceil(0.5 * T))
    frames = [n_unguided, n_unguided + 1, T + 1]
    for hash_on in (False, True):
        smp = M.MultiRobotSampler(p["model"], p["guide"], use_peer_hash=hash_on)
        _, chain = smp.sample([{k: v for k, v in hc.items()} for hc in hcs], K, noise=noise.to(dev), mode="lockstep",
                              return_chain=True)
        _lockstep_report(f"R={R} K={K} T={T} hash={hash_on}", chain[:, frames].cpu(), ref[:, frames], frames, T)


@pytest.mark.parametrize("case", ["config2", "config3"])
def test_lockstep_baseline_configs_golden(dev, case):
    """BASELINE.json config 2 (Empty, 6 robots x 32 samples, T=100) and config 3 (Conveyor, 10 robots, T=100; 32 of the 64
    samples) against golden frames of the CPU oracle (tests/golden/lockstep_*.npz, made by oracle/gen_golden_lockstep.py:
    the oracle needs minutes for these).  Default (tcgen05) executor, lock-step through the fused publication.
      * teacher forced: from the oracle's state at the end of the unguided prefix, the first two guided lock-step steps
        must reproduce the oracle's next frames (no accumulated drift: the per-step statement);
      * free running: the unguided prefix to the 1e-3 bar, the guided frames and the final trajectories inside the
        envelope the ORACLE ITSELF shows under a 3e-6 perturbation of eps (`self_err` in the fixture)."""
    import numpy as np
    import mmd_b200 as M
    path = os.path.join(os.path.dirname(__file__), "golden", f"lockstep_{case}.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated")
    g = np.load(path)
    env_name, R, K, T = str(g["env"]), int(g["R"]), int(g["K"]), int(g["T"])
    frames = [int(f) for f in g["frames"]]
    self_err = torch.from_numpy(g["self_err"])
    ref = torch.from_numpy(g["chain"])
    o = build_oracle(env_name, T=T, w_smooth=0.0)
    p = build_product(dev, env_name, T=T, P=o["P"], w_smooth=0.0, precision="f16x3")
    starts, goals = port.get_start_goal_pos_circle(R, float(g["radius"]))
    hcs = [port.hard_conds_from_start_goal(s, gl, o["norm"]) for s, gl in zip(starts, goals)]
    noise = torch.randn(R, T + 2, K, 64, 4, generator=torch.Generator().manual_seed(7))
    assert torch.equal(ref[:, 0], torch.stack([port.apply_hard_conditioning(noise[r, 0].clone(), port.repeat_hard_conds(hcs[r], K))
                                               for r in range(R)]))   # the fixture belongs to these inputs
    hcd = [{k: v for k, v in hc.items()} for hc in hcs]
    smp = M.MultiRobotSampler(p["model"], p["guide"])
    n_unguided = frames[1]                             # chain frame after the FIRST guided step (k = 51, t = 49)
    assert n_unguided == 1 + (T - math.ceil(0.5 * T))
    # ---- teacher forced: restart from the oracle's frame 51; the next step is k = 52 (t = 48) and draws noise[52] -------
    t_rest = T - n_unguided                            # remaining timesteps: 49 (t = 48 .. 0, plus the noise-free step)
    _, sub = smp.sample(hcd, K, noise=noise[:, n_unguided:].to(dev), mode="lockstep", return_chain=True,
                        x_init=ref[:, 1].to(dev), n_diffusion_steps=t_rest)
    for j in (1, 2):
        e = _per_traj(sub[:, j].reshape(R * K, 64, 4), ref[:, 1 + j].reshape(R * K, 64, 4))
        print(f"lockstep {case} teacher-forced, {j} guided step(s): per-trajectory rel L2 median={float(e.median()):.2e} "
              f"p90={float(e.quantile(0.9)):.2e} max={float(e.max()):.2e}")
        assert float(e.median()) < 1e-4 and float((e < 1e-3).float().mean()) >= (0.9 if j == 1 else 0.8)
    # ---- free running ---------------------------------------------------------------------------------------------------
    _, chain = smp.sample(hcd, K, noise=noise.to(dev), mode="lockstep", return_chain=True)
    for fi in range(1, len(frames)):
        e = _per_traj(chain[:, frames[fi]].reshape(R * K, 64, 4), ref[:, fi].reshape(R * K, 64, 4))
        se = self_err[fi]
        print(f"lockstep {case} frame {frames[fi]}: per-trajectory rel L2 median={float(e.median()):.2e} p90={float(e.quantile(0.9)):.2e} "
              f"max={float(e.max()):.2e}   (oracle vs oracle(eps * (1 + 3e-6 N)): median={float(se[0]):.2e} p90={float(se[1]):.2e} "
              f"max={float(se[2]):.2e})")
        if fi == 1:     # one guided step after the unguided prefix: the north-star bar
            assert float(e.median()) < 1e-4 and float(e.quantile(0.9)) < 1e-3 and float(e.max()) < 2e-2
        else:           # guided frames: inside (a small multiple of) the reference's own sensitivity envelope
            assert float(e.median()) < 5.0 * float(se[0]) + 1e-5
            assert float(e.quantile(0.9)) < 5.0 * float(se[1]) + 1e-4
            assert float(e.max()) < 0.2


def test_peer_term_equals_constraint_object(pair, dev):
    """The lock-step peer table must act exactly like the CostConstraint the oracle builds from it (one guide
    evaluation, 5 groups at once, each skipping its own row) and publish_peers must equal the reference unnormalise."""
    import ctypes as C
    from mmd_b200 import _lib
    from mmd_b200.diffusion import lower_for_step
    o, p = pair
    R, K = 5, 6
    g = torch.Generator().manual_seed(11)
    x = torch.randn(R * K, 64, 4, generator=g) * 0.3
    x[7] *= 4.0  # group 1 leaves [-1, 1]: its clip flag fires, the others' do not
    xd = x.to(dev)
    env, grp, keep = lower_for_step(p["guide"], R, K, 64, dev, [None] * R, None)
    peers = torch.empty(R, 64, 2, device=dev)
    _lib.check(_lib.lib().mmdk_publish_peers(C.byref(env), R, K, 64, 0, _lib.ptr(xd), _lib.ptr(peers), _lib.stream_ptr()))
    reps = torch.stack([o["norm"].unnormalize(x[r * K:(r + 1) * K].clone())[0, :, :2] for r in range(R)])
    assert torch.equal(peers.cpu(), reps)
    peer_self = torch.arange(R, dtype=torch.int32, device=dev)
    # a fleet larger than this rank's groups (other GPUs' robots): 200 rows = 102 KB > 48 KB exercises the opt-in
    # large shared-memory staging of the table
    extra = torch.rand(195, 64, 2, generator=g) * 1.6 - 0.8
    reps = torch.cat((reps, extra), 0)
    peers = reps.to(dev).contiguous()
    n_all = reps.shape[0]
    env, grp, keep = lower_for_step(p["guide"], R, K, 64, dev, [None] * R, None, peers, peer_self, 0.12, 2e-2)
    grad = torch.empty_like(xd)
    _lib.check(_lib.lib().mmdk_guide_grad(C.byref(env), C.byref(grp), 64, _lib.ptr(xd), _lib.ptr(grad), None, 0,
                                          _lib.stream_ptr()))
    hh = torch.arange(64, dtype=torch.float32).repeat(n_all - 1)
    for r in range(R):
        qs = torch.cat([reps[j] for j in range(n_all) if j != r], 0)
        o["guide"].extra = [port.Constraint(qs, torch.stack((hh, hh + 1), -1), torch.full((qs.shape[0],), 0.12), True, 2e-2)]
        ref = o["guide"](x[r * K:(r + 1) * K])
        o["guide"].extra = []
        assert max_err(grad[r * K:(r + 1) * K], ref) < 1e-6, r


def test_check_rr_collisions_bit_exact(pair, dev):
    o, p = pair
    g = torch.Generator().manual_seed(2)
    pos = torch.rand(64, 12, 2, generator=g) * 0.6 - 0.3
    # adversarial pairs on a dyadic lattice around the 0.105 threshold (exact squares -> unambiguous in fp32)
    lat = torch.arange(0, 12).float() * (27.0 / 256.0)  # 0.10546875 spacing: just above 2.1 r
    pos[0, :, 0], pos[0, :, 1] = lat, 0.0
    pos[1, :, 0], pos[1, :, 1] = lat * (26.0 / 27.0), 0.0  # 0.1015625 spacing: just below
    ref_c, ref_m = port.check_rr_collisions(pos)
    out_c, out_m = p["robot"].check_rr_collisions(pos.to(dev))
    assert torch.equal(out_c.cpu(), ref_c)
    assert torch.equal(torch.isnan(out_m.cpu()), torch.isnan(ref_m))
    assert torch.equal(torch.nan_to_num(out_m.cpu()), torch.nan_to_num(ref_m))


def test_classify_trajs_matches_oracle(pair, dev):
    o, p = pair
    g = torch.Generator().manual_seed(6)
    K = 96
    t = torch.linspace(0, 1, 64)[None, :, None]
    a = torch.rand(K, 1, 2, generator=g) * 1.8 - 0.9
    b = torch.rand(K, 1, 2, generator=g) * 1.8 - 0.9
    pos = a * (1 - t) + b * t + 0.02 * torch.randn(K, 64, 2, generator=g)
    pos[:5] *= 1.3  # some leave the joint limits
    trajs = torch.cat((pos, 0.1 * torch.randn(K, 64, 2, generator=g)), -1)
    free_idx, coll = port.get_trajs_free_idxs(trajs, o["guide"].grid)
    free, cost, wp = p["task"].classify(trajs.to(dev), return_waypoints=True)
    assert torch.equal(wp.cpu().bool(), coll)
    assert torch.equal(torch.argwhere(free.cpu().bool()).reshape(-1), free_idx)
    ref_cost = port.compute_path_length(trajs) + port.compute_smoothness(trajs)
    assert rel_err(cost, ref_cost) < 1e-5
    assert 0 < free_idx.numel() < K


def test_mpd_planner_call_surface(dev):
    """MPD.__call__(start, goal, constraints_l) (mpd.py:306-405): chain equals the oracle's run_inference on the same
    noise (GP term off for the free-running comparison), PlannerOutput fields are populated, best index is a free
    trajectory with minimal path length + smoothness."""
    import mmd_b200 as M
    T, K = 25, 16
    o = build_oracle("EnvConveyor2D", T=T, w_smooth=0.0)
    # exact fp32 executor for the chain-vs-oracle comparison (K=16: two branch flips already move the 90th percentile
    # with the FP16-split executor, see test_run_inference_chain_free_running_tensor_core)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision="fp32")
    unet.load_state_dict(o["P"], strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True)
    start, goal = torch.tensor([-0.8, -0.6]), torch.tensor([0.8, 0.6])
    planner = M.MPD("EnvConveyor2D-RobotPlanarDisk", "mmd", start, goal, n_samples=K, model=model, device="cuda:0",
                    weight_grad_cost_smoothness=0.0, seed=18)
    qs, rng, rad = random_constraints(40, seed=5)
    cons = [M.MultiPointConstraint([q.to(dev) for q in qs], rng.tolist(), rad.tolist(), is_soft=True)]
    noise = torch.randn(T + 2, K, 64, 4, generator=torch.Generator().manual_seed(3))
    hc = port.hard_conds_from_start_goal(start, goal, o["norm"])
    o["guide"].extra = [port.Constraint(qs, rng, rad, True)]
    ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
    o["guide"].extra = []
    cc = [M.CostConstraint(planner.robot, 64, q_l=c.get_q_l(), traj_range_l=c.get_t_range_l(), radius_l=c.radius_l,
                           is_soft=True, tensor_args=planner.tensor_args) for c in cons]
    chain, t_sampling, t_post = planner.run_constrained_inference(cc, noise=noise.to(dev))   # the reference's 3-tuple
    assert t_sampling > 0 and t_post == 0.0
    e = _per_traj(chain[-1], ref[-1])
    print(f"MPD chain vs oracle: median={float(e.median()):.2e} max={float(e.max()):.2e}")
    assert float(e.median()) < 1e-4 and float(e.quantile(0.9)) < 1e-3
    out = planner(start.to(dev), goal.to(dev), constraints_l=cons)
    assert out.trajs_iters.shape == (T + 2, K, 64, 4) and out.trajs_final.shape == (K, 64, 4)
    n_free = 0 if out.trajs_final_free is None else out.trajs_final_free.shape[0]
    n_coll = 0 if out.trajs_final_coll is None else out.trajs_final_coll.shape[0]
    assert n_free + n_coll == K
    if n_free:
        free_idx, _ = port.get_trajs_free_idxs(out.trajs_iters[-1].cpu(), o["guide"].grid)
        assert torch.equal(out.trajs_final_free_idxs.reshape(-1).cpu(), free_idx)
        assert int(out.idx_best_traj) in free_idx.tolist()
        assert out.cost_smoothness.shape == (n_free,) and out.cost_path_length.shape == (n_free,)
        assert torch.allclose(out.cost_all, out.cost_smoothness + out.cost_path_length, rtol=1e-5)
        assert out.fraction_free_trajs == n_free / K and out.success_free_trajs == 1
    with pytest.raises(ValueError):
        planner(goal.to(dev), start.to(dev))


def test_plan_batch_equals_sequential_calls(dev):
    """plan_batch (the batched entry for cbs.py:316-324 / :390-430) must return exactly what the planners return when
    called one after the other from the same RNG state: same chains bit for bit, same free sets, same best trajectory."""
    import mmd_b200 as M
    T, K, R = 25, 8, 3
    o = build_oracle("EnvHighways2D", T=T)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    unet.load_state_dict(o["P"], strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True)
    starts, goals = port.get_start_goal_pos_circle(R, 0.6)
    planners = [M.MPD("EnvHighways2D-RobotPlanarDisk", "mmd", starts[r], goals[r], n_samples=K, model=model, device="cuda:0")
                for r in range(R)]
    qs, rng, rad = random_constraints(30, seed=9)
    cons = [None, [M.MultiPointConstraint([q.to(dev) for q in qs], rng.tolist(), rad.tolist(), is_soft=True)], None]
    torch.manual_seed(123)
    seq = [planners[r](starts[r].to(dev), goals[r].to(dev), constraints_l=cons[r]) for r in range(R)]
    torch.manual_seed(123)
    bat = M.plan_batch(planners, cons)
    for r in range(R):
        assert torch.equal(seq[r].trajs_iters, bat[r].trajs_iters), f"robot {r}"
        assert torch.equal(seq[r].trajs_final, bat[r].trajs_final)
        assert torch.equal(seq[r].trajs_final_free_idxs, bat[r].trajs_final_free_idxs)
        assert (seq[r].idx_best_traj is None) == (bat[r].idx_best_traj is None)
        if seq[r].idx_best_traj is not None:
            assert int(seq[r].idx_best_traj) == int(bat[r].idx_best_traj)


def test_diffusions_ensemble_two_tiles(dev):
    """DiffusionsEnsemble.run_inference (diffusion_ensemble.py:56-106,224-268): two tiles stitched by cross conditioning
    (63 -> 0 through the (2, 0) tile transform).  Final frames and the recorded chain vs the oracle (chain frames of later
    tiles alias in the reference, oracle/port.py ensemble_p_sample_loop: reproduced by mmdk_run_chain_ensemble)."""
    import mmd_b200 as M
    T, K = 25, 6
    o = build_oracle("EnvEmptyNoWait2D", T=T, w_smooth=0.0)
    p0 = build_product(dev, "EnvEmptyNoWait2D", T=T, P=o["P"], w_smooth=0.0)
    p1 = build_product(dev, "EnvEmptyNoWait2D", T=T, P=o["P"], w_smooth=0.0)
    transforms = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    norm = o["norm"]
    s0 = norm.normalize(torch.tensor([-0.7, 0.1, 0.0, 0.0]))
    g1 = norm.normalize(torch.tensor([0.6, -0.2, 0.0, 0.0]))
    hard = {0: {0: s0}, 1: {63: g1}}
    cross = {(0, 1): (-1, 0)}
    g = torch.Generator().manual_seed(21)
    noise = {m: torch.randn(T + 2, K, 64, 4, generator=g) for m in (0, 1)}
    guides = {m: port.GuideSpec(o["guide"].grid, norm, w_smooth=0.0) for m in (0, 1)}
    kw = {m: dict(guide=guides[m], n_guide_steps=20, t_start_guide=13, noise_std=0.5) for m in (0, 1)}
    ref = port.ensemble_p_sample_loop({0: o["model"], 1: o["model"]},
                                      {m: port.repeat_hard_conds(h, K) for m, h in hard.items()}, cross, transforms,
                                      noise, T, 1, kw)
    ens = M.DiffusionsEnsemble({0: p0["model"], 1: p1["model"]}, {m: t.to(dev) for m, t in transforms.items()})
    skw = [dict(guide=p0["guide"], n_guide_steps=20, t_start_guide=13, noise_std_extra_schedule_fn=lambda x: 0.5),
           dict(guide=p1["guide"], n_guide_steps=20, t_start_guide=13, noise_std_extra_schedule_fn=lambda x: 0.5)]
    chains = ens.run_inference(None, {m: {k: v.to(dev) for k, v in h.items()} for m, h in hard.items()}, cross,
                               n_samples=K, return_chain=True, sample_kwargs=skw, n_diffusion_steps_without_noise=1,
                               noise={m: n.to(dev) for m, n in noise.items()})
    for m in (0, 1):
        assert chains[m].shape == ref[m].shape == (T + 2, K, 64, 4)
        e = _per_traj(chains[m][-1], ref[m][-1])
        print(f"ensemble tile {m}: median={float(e.median()):.2e} max={float(e.max()):.2e}")
        assert float(e.median()) < 1e-4 and float(e.max()) < 2e-2
    # every recorded frame, including the reference's aliasing: frame k of tile 1 carries the stitched first waypoint of
    # reverse step k + 1 (the reference records x[1] itself and stitches in place after tile 0's step)
    stitch = (chains[1][:-1, :, 0, :].cpu() - ref[1][:-1, :, 0, :]).abs().max()
    print(f"ensemble chain frames: stitched waypoint of the recorded frames vs oracle max abs {float(stitch):.2e}")
    assert float(stitch) < 2e-3
    for k in range(0, T + 2, 5):
        for m in (0, 1):
            e = _per_traj(chains[m][k], ref[m][k])
            assert float(e.median()) < 1e-4 and float(e.max()) < 2e-2, (k, m)


def test_mpd_ensemble_planner(dev):
    """MPDEnsemble (mpd_ensemble.py:65-640): two tiles, start in tile 0 / goal in tile 1 (global frame), one soft
    MultiPointConstraint with an entry in each tile -> split per tile (waypoint range - task_id * 64, position - tile
    transform).  Sampling vs the oracle's ensemble loop with the hand-split constraints; then the planner call surface and
    the warm-started local inference (run_constrained_local_inference, diffusion_ensemble.py:270-312)."""
    import mmd_b200 as M
    T, K = 25, 6
    o = build_oracle("EnvEmptyNoWait2D", T=T, w_smooth=0.0, cutoff_margin=0.01)
    models = {}
    for j in (0, 1):
        unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision="fp32")
        unet.load_state_dict(o["P"], strict=True)
        models[j] = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True)
    transforms = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    start, goal = torch.tensor([-0.7, 0.1]), torch.tensor([2.6, -0.2])
    planner = M.MPDEnsemble(("EnvEmptyNoWait2D-RobotPlanarDisk",) * 2, transforms, "mmd", start, goal, n_samples=K, models=models,
                            device="cuda:0", weight_grad_cost_smoothness=0.0, seed=18)
    cons = [M.MultiPointConstraint([torch.tensor([-0.3, 0.0]).to(dev), torch.tensor([2.3, 0.1]).to(dev)], [(10, 12), (84, 86)],
                                   [0.2, 0.2], is_soft=True)]
    cc = [M.CostConstraint(planner.robot, 64, q_l=c.get_q_l(), traj_range_l=c.get_t_range_l(), radius_l=c.radius_l, is_soft=True,
                           tensor_args=planner.tensor_args) for c in cons]
    split = planner.split_cost_constraints_to_tasks(cc)
    assert sorted(split) == [0, 1] and all(len(v) == 1 for v in split.values())
    assert split[1][0].traj_ranges.tolist() == [[84.0, 86.0]]      # shifted only when installed (mpd_ensemble.py:516)
    # oracle with the constraints split by hand
    norm = o["norm"]
    hard = {0: {0: norm.normalize(torch.tensor([-0.7, 0.1, 0.0, 0.0]))}, 1: {63: norm.normalize(torch.tensor([0.6, -0.2, 0.0, 0.0]))}}
    cross = {(0, 1): (63, 0)}
    g = torch.Generator().manual_seed(33)
    noise = {m: torch.randn(T + 2, K, 64, 4, generator=g) for m in (0, 1)}
    guides = {m: port.GuideSpec(o["guide"].grid, norm, w_smooth=0.0, cutoff_margin=0.01) for m in (0, 1)}
    guides[0].extra = [port.Constraint(torch.tensor([[-0.3, 0.0]]), torch.tensor([[10.0, 12.0]]), torch.tensor([0.2]), True, 2e-2)]
    guides[1].extra = [port.Constraint(torch.tensor([[0.3, 0.1]]), torch.tensor([[20.0, 22.0]]), torch.tensor([0.2]), True, 2e-2)]
    kw = {m: dict(guide=guides[m], n_guide_steps=20, t_start_guide=13, noise_std=0.5) for m in (0, 1)}
    ref = port.ensemble_p_sample_loop({0: o["model"], 1: o["model"]}, {m: port.repeat_hard_conds(h, K) for m, h in hard.items()},
                                      cross, transforms, noise, T, 1, kw)
    for m in (0, 1):
        for k, v in hard[m].items():
            got = {kk % 64: vv for kk, vv in planner.hard_conds[m].items()}   # the goal is keyed -1 as in the reference
            assert torch.allclose(got[k].cpu(), v, atol=1e-6), "tile-frame hard conditions"
    chains, _, _ = planner.run_constrained_inference(cc, noise={m: n.to(dev) for m, n in noise.items()})
    for m in (0, 1):
        e = _per_traj(chains[m][-1], ref[m][-1])
        print(f"MPDEnsemble tile {m}: median={float(e.median()):.2e} max={float(e.max()):.2e}")
        assert float(e.median()) < 1e-4 and float(e.max()) < 2e-2
    out = planner(start.to(dev), goal.to(dev), constraints_l=cons)
    assert out.trajs_iters.shape == (T + 2, K, 128, 4) and out.trajs_final.shape == (K, 128, 4)
    last = out.trajs_iters[-1]
    assert float((last[:, 63, :2] - last[:, 64, :2]).abs().max()) < 1e-5          # tiles stitched in the global frame
    assert torch.allclose(last[:, 0, :2].cpu(), start.expand(K, 2), atol=1e-5) and torch.allclose(last[:, -1, :2].cpu(), goal.expand(K, 2), atol=1e-5)
    n_free = 0 if out.trajs_final_free is None else out.trajs_final_free.shape[0]
    assert n_free + out.trajs_final_coll_idxs.numel() == K
    # warm start: the previous result as experience
    class Exp:
        path_b = last.clone()
    q_noise = torch.randn(K, 128, 4, generator=g)
    nz = {m: torch.randn(3 + 2, K, 64, 4, generator=g) for m in (0, 1)}
    chains_l, _, _ = planner.run_constrained_local_inference([], Exp, noise={m: n.to(dev) for m, n in nz.items()}, q_noise=q_noise.to(dev))
    # oracle: q_sample of the whole path by tile 0's schedule, then the ensemble loop from the tile slices
    t3 = torch.full((K,), 3, dtype=torch.long)
    noised = o["model"].q_sample(last.cpu(), t3, q_noise)
    nz_o = {}
    for m in (0, 1):
        x0 = noised[:, m * 64:(m + 1) * 64].clone()
        x0[:, :, :2] -= transforms[m]
        nz_o[m] = torch.cat((x0[None], nz[m][1:]), 0)
    for gd in guides.values():
        gd.extra = []
    ref_l = port.ensemble_p_sample_loop({0: o["model"], 1: o["model"]}, {m: port.repeat_hard_conds(h, K) for m, h in hard.items()},
                                        cross, transforms, nz_o, 3, 1, kw)
    for m in (0, 1):
        assert chains_l[m].shape == (3 + 2, K, 64, 4)
        e = _per_traj(chains_l[m][-1], ref_l[m][-1])
        print(f"MPDEnsemble local inference tile {m}: median={float(e.median()):.2e} max={float(e.max()):.2e}")
        assert float(e.median()) < 1e-4 and float(e.max()) < 2e-2


def test_plan_batch_ensemble_equals_sequential_calls(dev):
    """plan_batch on MPDEnsemble planners (BASELINE config 5: one multi-tile planner per robot, robots crossing the 1x2 tile
    grid in both directions, inference_multi_agent.py:205-222) == the same planners called one after the other from the same
    RNG state, bit for bit; one robot carries a soft constraint that is split over both tiles."""
    import mmd_b200 as M
    T, K = 25, 6
    o = build_oracle("EnvEmptyNoWait2D", T=T, cutoff_margin=0.01)
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
    unet.load_state_dict(o["P"], strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True)
    models = {0: model, 1: model}
    fwd = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    bwd = {0: torch.tensor([2.0, 0.0]), 1: torch.tensor([0.0, 0.0])}
    jobs = [(fwd, [-0.7, 0.1], [2.6, -0.2]), (bwd, [2.5, 0.4], [-0.6, 0.3]), (fwd, [-0.5, -0.6], [2.4, 0.5])]
    planners = [M.MPDEnsemble(("EnvEmptyNoWait2D-RobotPlanarDisk",) * 2, tr, "mmd", torch.tensor(s), torch.tensor(g), n_samples=K,
                              models=models, device="cuda:0") for tr, s, g in jobs]
    cons = [None, None, [M.MultiPointConstraint([torch.tensor([-0.3, 0.0]).to(dev), torch.tensor([2.3, 0.1]).to(dev)],
                                                [(10, 12), (84, 86)], [0.2, 0.2], is_soft=True)]]
    torch.manual_seed(77)
    seq = [planners[r](torch.tensor(jobs[r][1]).to(dev), torch.tensor(jobs[r][2]).to(dev), constraints_l=cons[r]) for r in range(3)]
    torch.manual_seed(77)
    bat = M.plan_batch(planners, cons)
    for r in range(3):
        assert seq[r].trajs_iters.shape == (T + 2, K, 128, 4)
        assert torch.equal(seq[r].trajs_iters, bat[r].trajs_iters), f"robot {r}"
        assert torch.equal(seq[r].trajs_final, bat[r].trajs_final)
        assert torch.equal(seq[r].trajs_final_free_idxs, bat[r].trajs_final_free_idxs)
    # the guides are left clean (constraints are per call)
    assert all(len(g.extra_cost_l) == 0 for p in planners for g in p.guides.values())
    fast = M.plan_batch(planners, cons, rng="batched")
    assert fast[0].trajs_iters.shape == (T + 2, K, 128, 4) and bool(torch.isfinite(fast[0].trajs_iters).all())


@pytest.mark.parametrize("n_peers", [31, 32, 63, 64, 65, 128, 256])
def test_peer_table_sizes_cross_smem_boundary(pair, dev, n_peers):
    """Round-1 N=2 crash: at K=128 the brute-force path stages the [n_peers,64,2] table in shared memory; 64 rows made
    the dynamic size exactly 48 KiB (+128 B static) without the opt-in.  Sweep the boundary with the cluster (K=128)
    launch shape, on both peer paths (brute force in shared memory, spatial hash), and pin a slice against the oracle."""
    import ctypes as C
    from mmd_b200 import _lib
    from mmd_b200.diffusion import lower_for_step
    from mmd_b200.guides import PeerHash
    o, p = pair
    R, K = 2, 128
    g = torch.Generator().manual_seed(n_peers)
    x = (torch.randn(R * K, 64, 4, generator=g) * 0.3).clamp(-0.95, 0.95)
    xd = x.to(dev)
    peers_h = torch.rand(n_peers, 64, 2, generator=g) * 1.2 - 0.6
    peers_h[5] = torch.rand(64, 2, generator=g) * 3.0 - 1.5   # one robot wandering outside the hash box (clamped cells)
    peers = peers_h.to(dev).contiguous()
    peer_self = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    out = {}
    for name, ph in (("brute", None), ("hash", PeerHash(n_peers, 64, 0.12, dev).build(peers))):
        env, grp, keep = lower_for_step(p["guide"], R, K, 64, dev, [None] * R, None, peers, peer_self, 0.12, 2e-2, ph)
        grad = torch.empty_like(xd)
        _lib.check(_lib.lib().mmdk_guide_grad(C.byref(env), C.byref(grp), 64, _lib.ptr(xd), _lib.ptr(grad), None, 0,
                                              _lib.stream_ptr()))
        torch.cuda.synchronize()
        out[name] = grad.cpu()
    assert torch.isfinite(out["brute"]).all()
    assert max_err(out["hash"], out["brute"]) < 1e-6   # same peers contribute; only the summation order may differ
    hh = torch.arange(64, dtype=torch.float32).repeat(n_peers - 1)
    for r in range(R):
        qs = torch.cat([peers_h[j] for j in range(n_peers) if j != r], 0)
        o["guide"].extra = [port.Constraint(qs, torch.stack((hh, hh + 1), -1), torch.full((qs.shape[0],), 0.12), True, 2e-2)]
        sl = slice(r * K, r * K + 2)   # no element leaves [-1, 1]: the clip decision of the slice equals the group's
        assert float(x[r * K:(r + 1) * K].abs().max()) < 1.0
        ref = o["guide"](x[sl])
        o["guide"].extra = []
        assert max_err(out["hash"][sl], ref) < 1e-6, (n_peers, r)


@pytest.mark.parametrize("kinds", [("point",), ("vertex", "edge", "point")])
def test_get_conflicts_bit_exact(dev, kinds):
    """CBS.get_conflicts on the device (densification + pairwise test in one kernel): same conflicts, same order, same
    positions as the oracle (cbs.py:166-246), ragged start times included."""
    import mmd_b200 as M
    from mmd_b200.conflicts import EdgeConflict, PointConflict, VertexConflict
    from tests.test_oracle_vs_reference import _crossing_paths
    paths = _crossing_paths(R=7, seed=3)
    starts = [0, 2, 0, 1, 0, 3, 0]
    paths = [p[: 64 - s] for p, s in zip(paths, starts)]
    types = tuple({"point": PointConflict, "vertex": VertexConflict, "edge": EdgeConflict}[k] for k in kinds)
    ref = port.get_conflicts(paths, starts, want_vertex="vertex" in kinds, want_edge="edge" in kinds, want_point="point" in kinds)
    out = M.get_conflicts([p.to(dev) for p in paths], starts, conflict_types=types)
    assert len(ref) == len(out) and len(ref) > 0
    for r, o in zip(ref, out):
        assert o.agent_ids == [r[1], r[2]]
        if r[0] == "vertex":
            assert isinstance(o, VertexConflict) and o.t == r[3]
        elif r[0] == "edge":
            assert isinstance(o, EdgeConflict) and (o.t_from, o.t_to) == (r[3], r[4])
        else:
            assert isinstance(o, PointConflict) and (o.t_from, o.t_to) == (r[3], r[4])
            assert torch.equal(o.agent_id_to_p[r[1]].cpu(), r[5]) and torch.equal(o.agent_id_to_p[r[2]].cpu(), r[6])
            assert torch.equal(o.agent_id_to_q[r[1]].cpu(), r[7])


def test_count_conflicts_batched_equals_per_sample_calls(dev):
    """cbs.py:446-458: one get_conflicts call per candidate sample of one agent == one batched kernel call."""
    import mmd_b200 as M
    from tests.test_oracle_vs_reference import _crossing_paths
    R, K = 6, 40
    paths = _crossing_paths(R=R, seed=5)
    starts = [0] * R
    g = torch.Generator().manual_seed(9)
    cands = paths[2][None] + torch.cat((0.05 * torch.randn(K, 64, 2, generator=g), torch.zeros(K, 64, 2)), -1)
    counts, _ = M.count_conflicts_batched([p.to(dev) for p in paths], starts, 2, cands.to(dev))
    ref = []
    for k in range(K):
        pl = list(paths)
        pl[2] = cands[k]
        ref.append(len(port.get_conflicts(pl, starts)))
    assert counts.cpu().tolist() == ref and max(ref) > min(ref)


def test_smooth_trajs_matches_scipy(dev):
    import mmd_b200 as M
    x = torch.randn(33, 64, 4, generator=torch.Generator().manual_seed(2))
    ref = port.smooth_trajs(x)          # scipy savgol_filter in fp32, exactly the reference's call
    out = M.smooth_trajs(x.to(dev))
    assert max_err(out, ref) < 5e-6     # scipy's own fp32 rounding is ~1.5e-6 against the exact operator
