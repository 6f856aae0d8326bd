"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU): the lock-step fleet sharded
over the ranks must reproduce the single-process result of the same fleet on the same noise, for both exchange paths
(peer-memory publication fused into the step kernel, and the NCCL all-gather loop)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmd_b200 as M  # noqa: E402
from oracle import port  # noqa: E402
from tests.helpers import build_oracle, build_product  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T, K, per = 25, 8, 3
    R_total = per * world
    o = build_oracle("EnvHighways2D", T=T, w_smooth=0.0)
    p = build_product(dev, "EnvHighways2D", T=T, P=o["P"], w_smooth=0.0, precision="f16x3")
    starts, goals = port.get_start_goal_pos_circle(R_total, 0.45)
    hcs = [port.hard_conds_from_start_goal(starts[r], goals[r], o["norm"]) for r in range(R_total)]
    hcs = [{k: v.to(dev) for k, v in hc.items()} for hc in hcs]
    noise = torch.randn(R_total, T + 2, K, 64, 4, generator=torch.Generator().manual_seed(5)).to(dev)
    mine = slice(rank * per, (rank + 1) * per)
    kw = dict(n_guide_steps=20, noise_std=0.5, n_diffusion_steps_without_noise=1)
    whole = M.MultiRobotSampler(p["model"], p["guide"], **kw).sample(hcs, K, noise=noise, mode="lockstep")
    res = {}
    for name in ("p2p", "nccl"):
        s = M.MultiRobotSampler(p["model"], p["guide"], exchange=name, **kw)
        for _ in range(2):   # second call replays the captured graph (p2p) on fresh sequence numbers
            out = s.sample(hcs[mine], K, noise=noise[mine], mode="lockstep", robot_offset=rank * per, n_robots_total=R_total)
        torch.cuda.synchronize()
        ws = next(iter(s._ws.values()))
        if name == "p2p":
            assert ws["ex"] is not None and ws["ex"].world == world, "peer-memory exchange not in use"
            assert not ws["ex"].failed(), "a wait for a peer publication timed out"
        res[name] = out
        err = float((out - whole[mine]).abs().max())
        print(f"rank {rank} {name}: max |sharded - single process| = {err:.3e}, bit-equal {torch.equal(out, whole[mine])}", flush=True)
        assert torch.isfinite(out).all()
        assert err < 1e-5, (name, err)
    assert torch.equal(res["p2p"], res["nccl"])
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
