"""Generates tests/golden/reference_golden.npz by running the REAL reference (imported from /root/reference through
oracle/ref_shim.py) on seeded inputs.  Run in the build container only:  python -m oracle.gen_golden

The reference ships no tests or golden vectors for this path, so these outputs ARE the pin: the
`-m "not gpu"` suite checks oracle/port.py against them everywhere (also on the GPU box, where the reference tree does
not exist), and tests/test_oracle_vs_reference.py re-checks the port against the live reference where it does.
Inputs are regenerated from the seeds recorded here (torch CPU generators are deterministic), outputs are stored.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import port, ref_build  # noqa: E402

ENVS = ["EnvEmpty2D", "EnvEmptyNoWait2D", "EnvConveyor2D", "EnvHighways2D", "EnvDropRegion2D"]


def golden_inputs():
    """Seeded inputs shared by the generator and the tests."""
    g = torch.Generator().manual_seed(1234)
    inp = {}
    inp["unet_x"] = torch.randn(2, 64, 4, generator=g)
    inp["guide_x"] = torch.randn(4, 64, 4, generator=g) * 0.6   # partly outside [-1,1]: the global clip fires
    inp["guide_x_small"] = torch.randn(4, 64, 4, generator=g) * 0.25  # inside: no clip
    inp["cons_q"] = torch.rand(60, 2, generator=g) * 2 - 1
    hh = torch.randint(0, 64, (60,), generator=g).float()
    inp["cons_rng"] = torch.stack((hh, hh + 1), -1)
    inp["cons_rad"] = torch.full((60,), 0.12)
    inp["hard_q"] = torch.tensor([[0.1, 0.2]])
    inp["hard_rng"] = torch.tensor([[10.0, 15.0]])
    inp["hard_rad"] = torch.tensor([0.12])
    inp["chain_noise"] = torch.randn(27, 6, 64, 4, generator=g)  # T=25: 27 draws, K=6
    inp["cell_pts"] = torch.rand(512, 2, generator=g) * 2.4 - 1.2
    inp["rr_pos"] = torch.rand(8, 6, 2, generator=g) * 0.5 - 0.25
    t = torch.linspace(0, 1, 64)[None, :, None]
    a = torch.rand(24, 1, 2, generator=g) * 1.8 - 0.9
    b = torch.rand(24, 1, 2, generator=g) * 1.8 - 0.9
    pos = a * (1 - t) + b * t + 0.02 * torch.randn(24, 64, 2, generator=g)
    inp["cls_trajs"] = torch.cat((pos, 0.1 * torch.randn(24, 64, 2, generator=g)), -1)
    return inp


def main():
    torch.manual_seed(0)
    inp = golden_inputs()
    out = {}
    T = 25  # the exponential schedule is only finite for some T (25, 47, 50, 55, 100, ...): beta_T rounds to >= 1 otherwise
    P = port.make_unet_params(seed=0)
    for env_name in ENVS:
        ref = ref_build.build_reference(env_name, T, P)
        grid = ref["env"].grid_map_sdf_obj_fixed
        out[f"{env_name}/sdf_sub"] = grid.sdf_tensor[::16, ::16].numpy()
        out[f"{env_name}/grad_sub"] = grid.grad_sdf_tensor[::16, ::16].numpy()
        out[f"{env_name}/sdf_sha"] = np.frombuffer(hashlib.sha256(grid.sdf_tensor.numpy().tobytes()).digest(), dtype=np.uint8)
        out[f"{env_name}/grad_sha"] = np.frombuffer(hashlib.sha256(grid.grad_sdf_tensor.numpy().tobytes()).digest(), dtype=np.uint8)
        # guide without / with extra constraints, clip firing / not firing
        out[f"{env_name}/guide"] = ref["guide"](inp["guide_x"]).numpy()
        out[f"{env_name}/guide_small"] = ref["guide"](inp["guide_x_small"]).numpy()
        cc = ref_build.make_cost_constraint(ref, inp["cons_q"], inp["cons_rng"], inp["cons_rad"], True)
        hard = ref_build.make_cost_constraint(ref, inp["hard_q"], inp["hard_rng"], inp["hard_rad"], False)
        ref["guide"].add_extra_costs([cc, hard], [2e-2, 2e-1])
        out[f"{env_name}/guide_cons"] = ref["guide"](inp["guide_x"]).numpy()
        ref["guide"].reset_extra_costs()
        if env_name == "EnvHighways2D":
            # schedule buffers
            for k in port.make_schedule(T):
                out[f"sched/{k}"] = getattr(ref["model"], k).numpy()
            # UNet eps at three timesteps
            with torch.no_grad():
                for t in (0, 12, 24):
                    out[f"unet/eps_t{t}"] = ref["model"].model(inp["unet_x"], torch.full((2,), t, dtype=torch.long), None).numpy()
            # full guided chain with one soft constraint object (K=6, T=10)
            norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
            hc = port.hard_conds_from_start_goal(torch.tensor([-0.8, 0.0]), torch.tensor([0.8, 0.1]), norm)
            chain = ref_build.reference_run_inference(ref, hc, 6, inp["chain_noise"], extra_costs=[cc], extra_weights=[2e-2])
            out["chain/final"] = chain[-1].numpy()
            out["chain/frames"] = chain[[0, 1, 13, 25]].numpy()
            # integer outputs
            X = inp["cell_pts"]
            idx = ((X - grid.limits[0]) / grid.map_dim * grid.cmap_dim).floor().to(torch.int)
            mx = torch.tensor(grid.points_for_sdf.shape[:-1]).type_as(idx) - 1
            out["cell_idx"] = idx.clamp(torch.zeros_like(mx), mx).numpy()
            coll, mid = ref["robot"].check_rr_collisions(inp["rr_pos"])
            out["rr/coll"] = coll.numpy()
            out["rr/mid"] = mid.numpy()
            with ref_build.quiet():
                _, _, _, free_idxs, wp = ref["task"].get_trajs_collision_and_free(inp["cls_trajs"], return_indices=True)
            out["cls/free_idxs"] = free_idxs.reshape(-1).numpy()
            out["cls/wp"] = wp.numpy()
    # ensemble cross conditioning (sample_functions.py:17-31)
    ref_build.ref_shim.install()
    from mmd.models.diffusion_models.sample_functions import apply_cross_conditioning
    g = torch.Generator().manual_seed(99)
    x = {0: torch.randn(5, 64, 4, generator=g), 1: torch.randn(5, 64, 4, generator=g)}
    tr = {0: torch.tensor([0.0, 0.0]), 1: torch.tensor([2.0, 0.0])}
    xr = apply_cross_conditioning({k: v.clone() for k, v in x.items()}, {(0, 1): (-1, 0)}, tr)
    out["cross/x0"], out["cross/x1"] = xr[0].numpy(), xr[1].numpy()
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
