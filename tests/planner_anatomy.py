"""Where a single planner call spends its time (development tool, not a bench line).
usage: planner_anatomy.py [K]   -- MPD.__call__ on EnvHighways2D, T=100, K samples."""
import sys
import time

import torch

sys.path.insert(0, ".")
import mmd_b200 as M  # noqa: E402
from oracle import port  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = 100
dev = torch.device("cuda:0")
unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True).to(dev)
s, g = port.get_start_goal_pos_circle(4, 0.6)
pl = M.MPD("EnvHighways2D-RobotPlanarDisk", "mmd", s[0], g[0], n_samples=K, model=model, device="cuda:0")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


ms_call, _ = timed(lambda: pl(s[0].to(dev), g[0].to(dev)))
ms_chain, (chain, _, _) = timed(lambda: pl.run_constrained_inference([]))
noise = torch.randn(T + 2, K, 64, 4, device=dev)
ms_chain_given_noise, _ = timed(lambda: pl.run_constrained_inference([], noise=noise))
ms_finish, _ = timed(lambda: pl._finish(chain, None, 0.0))
x = torch.randn(K, 64, 4, device=dev)
ms_unet, _ = timed(lambda: [unet.forward_t(x, 5) for _ in range(T + 1)])
# the same chain without guide iterations: what the 20 guide steps x 51 guided reverse steps cost
kw = dict(pl.sample_fn_kwargs)
kw["n_guide_steps"] = 0
ms_noguide, _ = timed(lambda: model.run_inference(None, pl.hard_conds, n_samples=K, horizon=64, return_chain=True,
                                                   sample_fn=M.ddpm_sample_fn, **kw, n_diffusion_steps_without_noise=1, noise=noise))
print(f"K={K} T={T}: planner call {ms_call:.2f} ms = chain {ms_chain:.2f} (with the noise given: {ms_chain_given_noise:.2f}; "
      f"{T + 1} UNet forwards alone: {ms_unet:.2f}; chain without guide iterations: {ms_noguide:.2f}) + post-processing {ms_finish:.2f}")
