"""Golden vectors of the lock-step composition at BASELINE.json's config-2 / config-3 shapes (test infrastructure).

The CPU oracle needs 1.5 - 4 minutes per case, too slow for the GPU suite, so the results are committed as fixtures:
    python -m oracle.gen_golden_lockstep            ->  tests/golden/lockstep_config{2,3}.npz
Inputs are fully determined by seeds (weights: port.make_unet_params(seed=0); noise: torch.Generator().manual_seed(7);
starts/goals: port.get_start_goal_pos_circle(R, 0.3)); the file keeps chain frames 0 (x_T), the end of the unguided
prefix, the first two guided frames and the final frame of every robot.  GP term off (free-running chains are chaotic with
it).  config 3 is generated at K = 32 samples (BASELINE: 64) to bound the generation time.

Besides the reference frames the file keeps `self_err` [n_frames, 3]: (median, p90, max) per-trajectory relative L2 between the
oracle and the oracle whose UNet output is multiplied by (1 + 3e-6 N(0,1)) -- the error level of the FP16-split tensor-core
executor -- i.e. how far the REFERENCE arithmetic itself moves under that perturbation at this fleet size: the lock-step
coupling (unit-direction repulsion between representative paths, 20 x 51 evaluations) amplifies it by orders of magnitude."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from tests.helpers import build_oracle  # noqa: E402

CASES = {"config2": ("EnvEmpty2D", 6, 32, 100, 0.3), "config3": ("EnvConveyor2D", 10, 32, 100, 0.45)}   # last: start/goal circle radius


def frames_of(T):
    n_unguided = 1 + (T - math.ceil(0.5 * T))
    return [0, n_unguided, n_unguided + 1, n_unguided + 2, T + 1]


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    for name, (env, R, K, T, radius) in CASES.items():
        o = build_oracle(env, T=T, w_smooth=0.0)
        starts, goals = port.get_start_goal_pos_circle(R, radius)
        hcs = [port.hard_conds_from_start_goal(s, g, o["norm"]) for s, g in zip(starts, goals)]
        noise = torch.randn(R, T + 2, K, 64, 4, generator=torch.Generator().manual_seed(7))
        guides = [port.GuideSpec(o["guide"].grid, o["norm"], w_smooth=0.0) for _ in range(R)]
        chain = port.lockstep_sample(o["model"], guides, hcs, K, noise, return_chain=True)   # [R, T+2, K, H, D]
        fr = frames_of(T)
        gp = torch.Generator().manual_seed(11)

        def noisy_unet(x, t):
            e = port.unet_forward(o["P"], x, t)
            return e * (1.0 + 3e-6 * torch.randn(e.shape, generator=gp))

        pert = port.DiffusionModel(o["P"], T, unet_fn=noisy_unet)
        chain_p = port.lockstep_sample(pert, guides, hcs, K, noise, return_chain=True)
        self_err = []
        for f in fr:
            a, b = chain_p[:, f].reshape(R * K, -1), chain[:, f].reshape(R * K, -1)
            e = (a - b).norm(dim=1) / b.norm(dim=1)
            self_err.append([float(e.median()), float(e.quantile(0.9)), float(e.max())])
        np.savez_compressed(os.path.join(out_dir, f"lockstep_{name}.npz"), frames=np.array(fr),
                            chain=chain[:, fr].numpy().astype(np.float32), env=env, R=R, K=K, T=T, radius=radius,
                            self_err=np.array(self_err))
        print(name, "self_err (median, p90, max) per frame", fr, self_err)
        print(name, "done", chain.shape)


if __name__ == "__main__":
    main()
