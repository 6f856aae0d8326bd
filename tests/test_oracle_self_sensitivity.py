"""How sensitive is the REFERENCE arithmetic itself to a perturbation of eps at the level of a change of fp32 summation
order?  (CPU only; computes the envelopes DESIGN.md section 6 quotes and the GPU parity tests budget for.)

The oracle (bit-exact with the reference, tests/test_oracle_vs_reference.py) is run against a copy of itself whose UNet
output is multiplied by (1 + delta * N(0,1)):
  * free running, default weights: the guided chain is chaotic (GP gradient ~1e4 clipped to unit norm, x 8e-2, x 20
    iterations per step) -- final trajectories move by percent for delta = 1e-7;
  * free running, GP term off: the same perturbation stays at 1e-6;
  * teacher forced (every step restarted from the unperturbed x_k), default weights: almost every (step, trajectory)
    pair agrees to ~1e-5, but isolated pairs jump to 1e-3..1e-2 (one branch flip -- nearest SDF cell, hinge, in/out of a
    constraint radius, the normaliser's global clip -- amplified by the clipped GP steps of the same timestep).
The GPU tests therefore assert medians and bounded outlier FRACTIONS, not a hard per-pair maximum."""
import math
import os

import pytest
import torch

from oracle import port
from tests.helpers import build_oracle, random_constraints


def _problem(K, T, w_smooth, delta, seed=18):
    o = build_oracle("EnvHighways2D", T=T, w_smooth=w_smooth)
    qs, rng, rad = random_constraints(120, seed=3)
    o["guide"].extra = [port.Constraint(qs, rng, rad, True)]
    noise = torch.randn(T + 2, K, 64, 4, generator=torch.Generator().manual_seed(seed))
    hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), o["norm"])
    g = torch.Generator().manual_seed(7)

    def noisy_unet(x, t):
        e = port.unet_forward(o["P"], x, t)
        return e * (1.0 + delta * torch.randn(e.shape, generator=g))

    pert = port.DiffusionModel(o["P"], T, unet_fn=noisy_unet)
    return o, pert, noise, hc


def _per_traj(a, b):
    return (a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1)


def test_free_running_chain_is_chaotic_only_with_the_gp_term():
    K, T = 8, 25
    out = {}
    for w in (8e-2, 0.0):
        o, pert, noise, hc = _problem(K, T, w, 1e-7)
        ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
        alt = port.run_inference(pert, hc, K, noise, guide=o["guide"])
        e = _per_traj(alt[-1], ref[-1])
        out[w] = (float(e.median()), float(e.max()))
        print(f"oracle vs oracle(eps * (1 + 1e-7 N)), free running K={K} T={T} w_smooth={w}: final rel L2 median "
              f"{out[w][0]:.2e} max {out[w][1]:.2e}")
    assert out[8e-2][0] > 1e-3            # default weights: percent-level divergence from a 1e-7 perturbation
    assert out[0.0][0] < 1e-5 and out[0.0][1] < 1e-3


def test_teacher_forced_steps_have_isolated_outliers():
    """delta = 3e-6 is the measured eps error of the FP16-split tensor-core executor, 1.3e-6 of the fp32 one."""
    K, T = 16, 50
    for delta in (1.3e-6, 3e-6):
        o, pert, noise, hc = _problem(K, T, 8e-2, delta)
        ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
        hcr = port.repeat_hard_conds(hc, K)
        errs = []
        k = 1
        for i in reversed(range(-1, T)):
            t = torch.full((K,), i, dtype=torch.long)
            x = port.ddpm_sample_fn(pert, ref[k - 1].clone(), hcr, t, noise[k], guide=o["guide"], n_guide_steps=20,
                                    t_start_guide=math.ceil(0.5 * T), noise_std=0.5)
            x = port.apply_hard_conditioning(x, hcr)
            errs.append(_per_traj(x, ref[k]))
            k += 1
        e = torch.stack(errs)                      # [steps, K]
        guided = e[T - math.ceil(0.5 * T):]
        frac = float((guided > 1e-3).float().mean())
        print(f"oracle self-sensitivity, teacher forced, delta={delta:.1e}: guided steps median {float(guided.median()):.2e} "
              f"p99 {float(guided.flatten().quantile(0.99)):.2e} max {float(guided.max()):.2e}; pairs above 1e-3: "
              f"{int((guided > 1e-3).sum())} of {guided.numel()} ({100 * frac:.2f}%)")
        assert float(e.median()) < 1e-4
        assert frac < 0.05


@pytest.mark.skipif(os.environ.get("MMD_SLOW") != "1", reason="minutes of CPU time: set MMD_SLOW=1")
def test_free_running_k128():
    """The percentile the GPU test test_run_inference_chain_free_running_tensor_core[128] is held to.  Measured (16 cores):
    delta 1e-7: median 7.29e-07 p90 2.10e-04 max 2.20e-02, 92.2% below 1e-3; 1.3e-6: 1.02e-06 / 3.86e-05 / 9.68e-03, 93.8%;
    3e-6: 1.86e-06 / 1.72e-04 / 2.18e-02, 93.0%."""
    K, T = 128, 100
    for delta in (1e-7, 1.3e-6, 3e-6):
        o, pert, noise, hc = _problem(K, T, 0.0, delta)
        ref = port.run_inference(o["model"], hc, K, noise, guide=o["guide"])
        alt = port.run_inference(pert, hc, K, noise, guide=o["guide"])
        e = _per_traj(alt[-1], ref[-1])
        frac = float((e < 1e-3).float().mean())
        print(f"oracle self-sensitivity free running K=128 T=100 GP off delta={delta:.1e}: median {float(e.median()):.2e} "
              f"p90 {float(e.quantile(0.9)):.2e} max {float(e.max()):.2e}, {100 * frac:.1f}% below 1e-3")
        assert 0.85 < frac < 0.99
