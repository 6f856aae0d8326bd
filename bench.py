#!/usr/bin/env python
"""bench.py -- denoised trajectories/s of the guided-diffusion sampler (BASELINE.json metric).

Workload (config.workload): EnvHighways2D, 32 robots x 128 samples x T=100 DDPM steps (+1 noise-free step), horizon 64,
full guidance (SDF collision + workspace border + GP smoothness + lock-step inter-robot soft constraints, 20 guide
steps for t < 50), UNet dim_mults (1,2,4) with seeded random-init weights, synthetic SmallCircle starts/goals
(mmd/config/mmd_experiment_configs.py:142-156).  One "step" = one complete reverse chain for the whole batch.
N > 1 (torchrun, one rank per GPU): weak scaling, 32 robots per GPU, ONE lock-step fleet of 32 N robots whose
representative paths are all-gathered over NCCL once per guided timestep.

  value  device-timed, inputs (noise, hard conditions) resident in HBM before the timed region
  e2e    through the public API (MultiRobotSampler.sample) with HOST inputs: pinned start/goal states go H2D, noise is
         drawn on the device like the reference does (torch.randn), final trajectories come back D2H, every step
  --impl reference   the CPU oracle (oracle/port.py == the reference's own arithmetic, bit-exact on CPU) on the host
         cores, bounded sample, extrapolated (see cpu_baseline.sample)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV, R_PER_GPU, K, T, H, D = "EnvHighways2D", 32, 128, 100, 64, 4
UNET_FLOP_PER_SAMPLE = 36_495_360  # BASELINE.md section 2 (torch FlopCounterMode on the reference module)
N_GUIDE, N_EXTRA, NOISE_STD = 20, 1, 0.5


def small_circle(num_agents):
    """EnvHighways2DRobotPlanarDiskSmallCircle (mmd/config/mmd_experiment_configs.py:142-156)."""
    from oracle import port
    s, g = port.get_start_goal_pos_circle(min(num_agents, 10), 0.45)
    if num_agents > 10:
        s2, g2 = port.get_start_goal_pos_circle(num_agents - 10, 0.65)
        s, g = s + s2, g + g2
    return s, g


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, idx):
        self.idx, self.rows, self.p = idx, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        time.sleep(0.05)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle on the host cores, bounded sample
# ----------------------------------------------------------------------------------------------------------------------
def cpu_sample(n_robots_total, cores, n_guided=1, n_unguided=2):
    """Times robot 0 (K samples) of the lock-step fleet for a few reverse steps and extrapolates to the whole chain:
    robots run one after another on the CPU exactly as the reference plans them (cbs.py:316-324), so
    traj/s = K / (50 t_unguided + 51 t_guided)."""
    from oracle import port
    torch.set_num_threads(cores)
    P = port.make_unet_params(seed=0)
    sdf, grad = port.build_sdf_grid(ENV)
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    guide = port.GuideSpec(port.GridSDF(sdf, grad), norm)
    model = port.DiffusionModel(P, T)
    starts, goals = small_circle(n_robots_total)
    g = torch.Generator().manual_seed(18)
    hc = port.repeat_hard_conds(port.hard_conds_from_start_goal(starts[0], goals[0], norm), K)
    x = port.apply_hard_conditioning(torch.randn(K, H, D, generator=g), hc)
    # peers: straight-line paths of the other robots (positions only matter for the cost of the evaluation)
    tt = torch.linspace(0, 1, H)[:, None]
    qs = torch.cat([starts[j][None] * (1 - tt) + goals[j][None] * tt for j in range(1, n_robots_total)], 0)
    hh = torch.arange(H, dtype=torch.float32).repeat(n_robots_total - 1)
    guide.extra = [port.Constraint(qs, torch.stack((hh, hh + 1), -1), torch.full((qs.shape[0],), 0.12), True, 2e-2)]
    t_start = math.ceil(0.5 * T)

    def step(t_i):
        t0 = time.perf_counter()
        port.ddpm_sample_fn(model, x, hc, torch.full((K,), t_i, dtype=torch.long), torch.randn(K, H, D, generator=g),
                            guide=guide, n_guide_steps=N_GUIDE, t_start_guide=t_start, noise_std=NOISE_STD)
        return time.perf_counter() - t0

    step(T - 1)  # warm-up (thread pool, allocator)
    tu = min(step(T - 1 - i) for i in range(n_unguided))
    tg = min(step(10 + i) for i in range(n_guided))
    n_g = t_start + N_EXTRA
    n_u = T - t_start
    per_robot = n_u * tu + n_g * tg
    return K / per_robot, tu, tg, per_robot


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_total = R_PER_GPU * args.gpus
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        v, tu, tg, per_robot = cpu_sample(n_total, cores)
        vals.append(v)
    v = sum(vals) / len(vals)
    sample = (f"robot 0 of {n_total} (K={K}) with {n_total - 1} peer paths ({(n_total - 1) * H} soft vertex constraints): "
              f"{tu * 1e3:.0f} ms per unguided reverse step, {tg:.2f} s per guided step (20 guide evaluations), extrapolated to "
              f"{T - math.ceil(0.5 * T)} unguided + {math.ceil(0.5 * T) + N_EXTRA} guided steps; robots run sequentially on "
              f"the CPU as in cbs.py:316-324, so traj/s = K / per-robot chain time ({per_robot:.0f} s)")
    line = {"impl": "reference", "metric": "denoised trajectories/sec", "value": v, "unit": "trajectories/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_total * K / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": f"{ENV} SmallCircle, {R_PER_GPU} robots x {K} samples per GPU ({R_PER_GPU * n_gpus} robots total, one "
                        f"lock-step fleet), T={T}+{N_EXTRA} reverse steps, horizon {H}, {N_GUIDE} guide steps for t<{math.ceil(0.5 * T)}",
            "robots_per_gpu": R_PER_GPU, "samples": K, "ddpm_steps": T, "horizon": H, "unet_dim_mults": [1, 2, 4],
            "parallelism": f"robots sharded over {n_gpus} GPU(s); all-gather [R,64,2] per guided step",
            "l2": "per-chain inputs (noise 414 MB) exceed the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    import mmd_b200 as M
    from oracle import port  # weights/starts generators only (synthetic data), never on the timed path
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ta = {"device": dev, "dtype": torch.float32}
    env = M.envs.get_env(ENV + "ExtraObjects", tensor_args=ta)
    robot = M.RobotPlanarDisk(tensor_args=ta)
    task = M.PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.05, tensor_args=ta)
    dataset = M.TrajectoryDataset(env, robot, task, *port.DEFAULT_NORMALIZER_LIMITS, tensor_args=ta)
    unet = M.TemporalUnet(n_support_points=H, state_dim=D, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision=args.precision)
    unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True).to(dev)
    costs = [M.CostCollision(robot, H, field=f, sigma_coll=1.0, tensor_args=ta) for f in task.get_collision_fields()]
    costs.append(M.CostGPTrajectory(robot, H, 5.0 / H, sigma_gp=1.0, tensor_args=ta))
    comp = M.CostComposite(robot, H, costs, weights_cost_l=[2e-2, 2e-2, 2e-2, 8e-2], tensor_args=ta)
    guide = M.GuideManagerTrajectoriesWithVelocity(dataset, comp, clip_grad=True, tensor_args=ta)
    sampler = M.MultiRobotSampler(model, guide, n_guide_steps=N_GUIDE, noise_std=NOISE_STD,
                                  n_diffusion_steps_without_noise=N_EXTRA)
    R_total = R_PER_GPU * world
    starts, goals = small_circle(R_total)
    norm = dataset.normalizer
    mine = range(rank * R_PER_GPU, (rank + 1) * R_PER_GPU)
    sg_host = torch.stack([torch.stack((starts[r], goals[r])) for r in mine]).pin_memory()  # [R,2,2]

    def hard_conds_from(sg):
        out = []
        for r in range(R_PER_GPU):
            s = torch.cat((sg[r, 0], torch.zeros(2, device=sg.device)))
            g = torch.cat((sg[r, 1], torch.zeros(2, device=sg.device)))
            out.append({0: norm.normalize(s), H - 1: norm.normalize(g)})
        return out

    n_steps = T + N_EXTRA
    gen = torch.Generator(device=dev).manual_seed(18 + rank)
    noise = torch.randn(n_steps + 1, R_PER_GPU * K, H, D, device=dev, generator=gen)  # step-major: one contiguous frame per step
    hcs_dev = hard_conds_from(sg_host.to(dev))
    kw = dict(mode="lockstep", robot_offset=rank * R_PER_GPU, n_robots_total=R_total)

    # UNet launch durations: the chain is ONE replayed CUDA graph, so host events cannot bracket a forward; the persistent
    # executor stamps %globaltimer per CTA at start/end of every launch instead (mmdk_unet_debug_stamps); for the other
    # executors the forwards are timed by a short instrumented pass with CUDA events on the launching stream
    import ctypes as C
    from mmd_b200 import _lib
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    STAMP_SLOTS = 256
    stamps = torch.zeros(STAMP_SLOTS, n_sm, 2, dtype=torch.int64, device=dev)
    unet.ensure_time_table(T)
    fused = unet.resolve_precision(args.precision) == "f16x3"
    if fused:
        _lib.check(_lib.lib().mmdk_unet_debug_stamps(unet.native(), _lib.ptr(stamps), STAMP_SLOTS, n_sm))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- value: inputs resident ------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        sampler.sample(hcs_dev, K, noise=noise, **kw)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    stamps.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = sampler.sample(hcs_dev, K, noise=noise, **kw)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    if fused:
        grid = min((R_PER_GPU * K + 6) // 7, n_sm)
        st = stamps[:, :grid].cpu()
        live = (st[:, :, 0] > 0).all(dim=1) & (st[:, :, 1] > 0).all(dim=1)
        dur_ns = (st[:, :, 1].max(dim=1).values - st[:, :, 0].min(dim=1).values)[live].double()
        unet_ms = (dur_ns / 1e6).tolist()          # the last replay's forwards (slots are rewritten by every replay)
        unet_forwards_timed = len(unet_ms)
    else:
        xx, oo = torch.randn(R_PER_GPU * K, H, D, device=dev), torch.empty(R_PER_GPU * K, H, D, device=dev)
        evs = []
        for i in range(n_steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            unet.forward_t(xx, max(T - 1 - i, 0), precision=args.precision, out=oo)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        unet_ms = [a.elapsed_time(b) for a, b in evs]
        unet_forwards_timed = len(unet_ms)
    unet_avg_ms = sum(unet_ms) / len(unet_ms)
    finite = bool(torch.isfinite(out).all())

    # ---- e2e: host inputs, device RNG, D2H of the result --------------------------------------------------------------
    result_host = torch.empty(R_PER_GPU, K, H, D).pin_memory()

    def e2e_step():
        sg = sg_host.to(dev, non_blocking=True)
        res = sampler.sample(hard_conds_from(sg), K, noise=None, generator=gen, **kw)
        result_host.copy_(res, non_blocking=True)

    e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        pk, pk_kind = peaks()
        ms_step = ms_total / args.steps
        n_traj = R_total * K
        value = n_traj / (ms_step / 1e3)
        B = R_PER_GPU * K
        achieved = B * UNET_FLOP_PER_SAMPLE / (unet_avg_ms / 1e3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        n_unet_launches = 2 if fused else (31 if args.precision == "f16x3_layers" else 1)   # pack_input + persistent forward | pack_input + 30 layer launches | fp32 kernel
        n_guided = math.ceil(0.5 * T) + N_EXTRA
        # per reverse step: UNet launches + ddpm_step_kernel; per guided lock-step step: publish_peers + build_peer_hash
        launches_per_chain = n_steps * (n_unet_launches + 1) + (2 * n_guided if R_total > 1 else 0)
        cb_v, tu, tg, per_robot = (None, None, None, None)
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb_v, tu, tg, per_robot = cpu_sample(R_total, cores)
            cpu = {"value": cb_v, "unit": "trajectories/s", "cores": cores, "kind": "port",
                   "sample": f"oracle/port.py (bit-exact restatement of the reference on CPU), robot 0 of {R_total}, K={K}, "
                             f"{R_total - 1} peer paths: {tu * 1e3:.0f} ms per unguided step, {tg:.2f} s per guided step, "
                             f"extrapolated to 50 unguided + 51 guided steps ({per_robot:.0f} s per robot, robots sequential)"}
        line = {
            "metric": "denoised trajectories/sec", "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else args.precision,
            "data": "synthetic", "config": workload_config(world), "finite": finite,
            "e2e": {"value": n_traj / (ms_e2e / args.steps / 1e3), "unit": "trajectories/s",
                    "h2d_bytes_per_step": int(sg_host.numel() * 4), "d2h_bytes_per_step": int(result_host.numel() * 4),
                    "note": "noise drawn on the device by torch.randn as the reference does"},
            "gpu_launches": launches_per_chain * args.steps,
            "roofline": {"bound": "tensor", "kernel": "TemporalUnet forward (" + args.precision + (": unet_fused_kernel, one persistent tcgen05 launch)" if fused else ")"), "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step); "
                                        f"dense TF32 is nominally half of it",
                         "algorithmic_flop_per_launch": B * UNET_FLOP_PER_SAMPLE, "avg_launch_ms": unet_avg_ms,
                         "unet_share_of_step": unet_avg_ms * n_steps / ms_step,
                         "timing": ("%globaltimer stamps written by every CTA of every forward of the last timed chain (the chain is one "
                                    "replayed CUDA graph): max(end) - min(start) per launch" if fused else
                                    "CUDA events around every forward of an instrumented pass on the launching stream"),
                         "forwards_timed": unet_forwards_timed},
            "cpu_baseline": cpu, "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mmd_b200", choices=["mmd_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MMD_UNET_PRECISION", "f16x3"), choices=["fp32", "f16x3", "f16x3_layers"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
