#!/usr/bin/env python
"""bench.py -- denoised trajectories/s of the guided-diffusion sampler (BASELINE.json metric).

Default workload = BASELINE.json config 4: EnvHighways2D, 32 robots x 128 samples x T=100 DDPM steps (+1 noise-free step),
horizon 64, full guidance (SDF collision + workspace border + GP smoothness + lock-step inter-robot soft constraints, 20 guide
steps for t < 50), UNet dim_mults (1,2,4) with seeded random-init weights, synthetic SmallCircle starts/goals
(mmd/config/mmd_experiment_configs.py:142-156).  One "step" = one complete reverse chain for the whole batch.  --config 1|2|3
select the smaller BASELINE.json configurations (parity-test cases; their lines are kept under profiles/); --config 5 is the
MPDEnsemble 1x2 multi-tile configuration (one multi-tile planner per robot, all robots served by mmd_b200.plan_batch).

N > 1 (torchrun, one rank per GPU), --scaling:
  strong (default; BASELINE.json: "32 robots x 128 samples ... batch sharded across 8xB200"): the SAME fleet, robots sharded
         contiguously over the ranks (32 / N robots per GPU)
  weak   32 robots per GPU in ONE lock-step fleet of 32 N robots
Either way the fleet is one lock-step fleet: at the end of every reverse step each rank's step kernel stores its robots'
representative paths straight into the peer tables of all ranks over NVLink (mmd_b200/exchange.py) -- no host-side collective
on the data path, one captured CUDA graph per rank.

  value  device-timed, inputs (noise, hard conditions) resident in HBM before the timed region
  e2e    through the public API (MultiRobotSampler.sample) with HOST inputs: pinned start/goal states go H2D, noise is
         drawn on the device like the reference does (torch.randn), final trajectories come back D2H, every step
  e2e_planner  (N=1) the reference's own call pattern: one MPD planner __call__ per robot, sequentially (cbs.py:316-324)
  --impl reference   the CPU oracle (oracle/port.py == the reference's own arithmetic, bit-exact on CPU) on the host
         cores, bounded sample, extrapolated (see cpu_baseline.sample)
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, D = 64, 4
UNET_FLOP_PER_SAMPLE = 36_495_360  # BASELINE.md section 2 (torch FlopCounterMode on the reference module)
N_GUIDE, N_EXTRA, NOISE_STD = 20, 1, 0.5

# BASELINE.json configs (1-based).  R = robots of the fleet, K = samples per robot.
CONFIGS = {
    1: dict(env="EnvEmpty2D", R=2, K=8, T=50, name="config 1: Empty map, 2 robots x 8 samples, 50 DDPM steps"),
    2: dict(env="EnvEmpty2D", R=6, K=32, T=100, name="config 2: Empty map, 6 robots x 32 samples, 100 steps"),
    3: dict(env="EnvConveyor2D", R=10, K=64, T=100, name="config 3: Conveyor map, 10 robots x 64 samples, 100 steps"),
    4: dict(env="EnvHighways2D", R=32, K=128, T=100, name="config 4: Highways map, 32 robots x 128 samples, 100 steps"),
    5: dict(env="EnvEmptyNoWait2D", R=16, K=64, T=100, tiles=2,
            name="config 5: MPDEnsemble 1x2 multi-tile, 16 robots x 64 samples, 100 steps"),
}


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
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
def cpu_sample(cfg, n_robots_total, budget_s=25.0):
    """Times robot 0 (K samples) of the lock-step fleet for a few reverse steps and extrapolates to the whole chain:
    robots run one after another on the CPU exactly as the reference plans them (cbs.py:316-324), so
    traj/s = K / (n_unguided t_unguided + n_guided t_guided).
    Thread policy: torch intra-op threads swept over {8, 16, 32, ... , all cores}; for every
    setting the MEDIAN of 3 guided steps and of 3 unguided steps (after one warm-up step); the best setting is reported."""
    from oracle import port
    env, K, T = cfg["env"], cfg["K"], cfg["T"]
    n_tiles = cfg.get("tiles", 1)
    P = port.make_unet_params(seed=0)
    sdf, grad = port.build_sdf_grid(env)
    norm = port.LimitsNormalizer(*port.DEFAULT_NORMALIZER_LIMITS)
    guide = port.GuideSpec(port.GridSDF(sdf, grad), norm, cutoff_margin=0.01 if n_tiles > 1 else 0.05)   # mpd_ensemble.py:139 / mpd.py:127
    model = port.DiffusionModel(P, T)
    starts, goals = small_circle(n_robots_total)
    g = torch.Generator().manual_seed(18)
    hc = port.repeat_hard_conds(port.hard_conds_from_start_goal(starts[0], goals[0], norm), K)
    x = port.apply_hard_conditioning(torch.randn(K, H, D, generator=g), hc)
    # peers: straight-line paths of the other robots (positions only matter for the cost of the evaluation); the multi-tile
    # configuration is the CBS root pattern (one unconstrained planner call per robot, cbs.py:316-324)
    if n_robots_total > 1 and n_tiles == 1:
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

    cores = os.cpu_count() or 1
    sweep = sorted({c for c in (8, 16, 32, 64, 128, cores) if c <= cores} | {cores})
    n_g, n_u = t_start + N_EXTRA, T - t_start
    best = None
    t_begin = time.perf_counter()
    tried = []
    for nt in sweep:
        torch.set_num_threads(nt)
        step(T - 1)  # warm-up (thread pool, allocator)
        tu = statistics.median(step(T - 1 - i) for i in range(3))
        tg = statistics.median(step(10 + i) for i in range(3))
        per_robot = n_tiles * (n_u * tu + n_g * tg)   # an ensemble steps every tile model once per reverse step (diffusion_ensemble.py:86-101)
        tried.append((nt, round(tg, 3)))
        if best is None or per_robot < best[3]:
            best = (nt, tu, tg, per_robot)
        if time.perf_counter() - t_begin > budget_s:
            break
    nt, tu, tg, per_robot = best
    n_peer = (n_robots_total - 1) if n_tiles == 1 else 0
    sample = (f"oracle/port.py (bit-exact restatement of the reference on CPU), robot 0 of {n_robots_total} (K={K}"
              f"{', one tile of %d; per-robot time = %d x one tile' % (n_tiles, n_tiles) if n_tiles > 1 else ''}) with "
              f"{n_peer} peer paths = {n_peer * H} soft vertex constraints: median of 3 -> "
              f"{tu * 1e3:.0f} ms per unguided reverse step, {tg:.2f} s per guided step (20 guide evaluations), best of the thread "
              f"sweep {tried} (threads, guided s) = {nt} threads; extrapolated to {n_u} unguided + {n_g} guided steps; robots run "
              f"sequentially on the CPU as in cbs.py:316-324, so traj/s = K / per-robot chain time ({per_robot:.1f} s)")
    return K / per_robot, nt, sample, per_robot


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n_total = cfg["R"] * (args.gpus if args.scaling == "weak" else 1)
    v, nt, sample, per_robot = cpu_sample(cfg, n_total)
    line = {"impl": "reference", "metric": "denoised trajectories/sec", "value": v, "unit": "trajectories/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_total * cfg["K"] / v,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args.gpus, args.scaling),
            "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": nt, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(cfg, n_gpus, scaling):
    if cfg.get("tiles", 1) > 1:
        T = cfg["T"]
        return {"workload": f"BASELINE.json {cfg['name']}; tiles {cfg['env']} with transforms (0,0),(2,0) (inference_multi_agent.py:146-149), "
                            f"{cfg['R']} MPDEnsemble planners (skeletons alternating tile 0->1 / 1->0) x {cfg['K']} samples, each trajectory = "
                            f"{cfg['tiles']} x {H} waypoints stitched by cross conditioning; T={T}+{N_EXTRA} reverse steps x {cfg['tiles']} tile "
                            f"models per step, {N_GUIDE} guide steps for t<{math.ceil(0.5 * T)}; CBS root pattern (no inter-robot constraints)",
                "robots_total": cfg["R"], "robots_per_gpu": cfg["R"], "samples": cfg["K"], "ddpm_steps": T, "horizon": H,
                "tiles": cfg["tiles"], "unet_dim_mults": [1, 2, 4],
                "parallelism": "single GPU; all planner calls of the root loop in one batch per tile model (mmd_b200.plan_batch)",
                "l2": "chain frames + noise (2 tiles x 102 frames x 1 MiB x 2) exceed the 126 MB L2; no explicit flush"}
    R_total = cfg["R"] * (n_gpus if scaling == "weak" else 1)
    per = (R_total + n_gpus - 1) // n_gpus
    T = cfg["T"]
    return {"workload": f"BASELINE.json {cfg['name']}; {cfg['env']} SmallCircle, {R_total} robots x {cfg['K']} samples in one lock-step "
                        f"fleet ({scaling} scaling: {per} robots per GPU), T={T}+{N_EXTRA} reverse steps, horizon {H}, {N_GUIDE} guide "
                        f"steps for t<{math.ceil(0.5 * T)}",
            "robots_total": R_total, "robots_per_gpu": per, "samples": cfg["K"], "ddpm_steps": T, "horizon": H,
            "unet_dim_mults": [1, 2, 4],
            "parallelism": f"robots sharded over {n_gpus} GPU(s); representative paths [R,64,2] published into every rank's peer "
                           f"table by the step kernel (NVLink stores + release/acquire flags) once per guided step",
            "l2": "per-chain inputs (noise frames, 4 MiB each, 101 per chain) exceed the 126 MB L2 at config 4; no explicit flush"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    import mmd_b200 as M
    from mmd_b200 import _lib
    from oracle import port  # weights/starts generators only (synthetic data), never on the timed path
    cfg = CONFIGS[args.config]
    ENV, K, T = cfg["env"], cfg["K"], cfg["T"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R_total = cfg["R"] * (world if args.scaling == "weak" else 1)
    lo, hi = M.shard_robots(R_total, world, rank)
    R_local = hi - lo
    if R_local < 1:
        raise SystemExit(f"rank {rank} has no robots: {R_total} robots over {world} ranks")
    ta = {"device": dev, "dtype": torch.float32}
    env = M.envs.get_env(ENV + "ExtraObjects", tensor_args=ta)
    robot = M.RobotPlanarDisk(tensor_args=ta)
    task = M.PlanningTask(env=env, robot=robot, ws_limits=env.limits, obstacle_cutoff_margin=0.05, tensor_args=ta)
    dataset = M.TrajectoryDataset(env, robot, task, *port.DEFAULT_NORMALIZER_LIMITS, tensor_args=ta)
    unet = M.TemporalUnet(n_support_points=H, state_dim=D, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision=args.precision)
    unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True).to(dev)
    costs = [M.CostCollision(robot, H, field=f, sigma_coll=1.0, tensor_args=ta) for f in task.get_collision_fields()]
    weights = [2e-2] * len(costs) + [8e-2]
    costs.append(M.CostGPTrajectory(robot, H, 5.0 / H, sigma_gp=1.0, tensor_args=ta))
    comp = M.CostComposite(robot, H, costs, weights_cost_l=weights, tensor_args=ta)
    guide = M.GuideManagerTrajectoriesWithVelocity(dataset, comp, clip_grad=True, tensor_args=ta)
    sampler = M.MultiRobotSampler(model, guide, n_guide_steps=N_GUIDE, noise_std=NOISE_STD,
                                  n_diffusion_steps_without_noise=N_EXTRA, exchange=args.exchange)
    starts, goals = small_circle(R_total)
    norm = dataset.normalizer
    sg_host = torch.stack([torch.stack((starts[r], goals[r])) for r in range(lo, hi)]).pin_memory()  # [R,2,2]

    def hard_conds_from(sg):
        out = []
        for r in range(R_local):
            s = torch.cat((sg[r, 0], torch.zeros(2, device=sg.device)))
            g = torch.cat((sg[r, 1], torch.zeros(2, device=sg.device)))
            out.append({0: norm.normalize(s), H - 1: norm.normalize(g)})
        return out

    n_steps = T + N_EXTRA
    B = R_local * K
    gen = torch.Generator(device=dev).manual_seed(18 + rank)
    noise = torch.randn(n_steps + 1, B, H, D, device=dev, generator=gen)  # step-major: one contiguous frame per step
    hcs_dev = hard_conds_from(sg_host.to(dev))
    kw = dict(mode="lockstep", robot_offset=lo, n_robots_total=R_total)

    # UNet launch durations: the chain is ONE replayed CUDA graph, so host events cannot bracket a forward; the persistent
    # executor stamps %globaltimer per CTA at start/end of every launch instead (mmdk_unet_debug_stamps); for the other
    # executors the forwards are timed by a short instrumented pass with CUDA events on the launching stream
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    STAMP_SLOTS = 256
    stamps = torch.zeros(STAMP_SLOTS, n_sm, 2, dtype=torch.int64, device=dev)
    unet.ensure_time_table(T)
    fused = unet.native_mode(args.precision) == _lib.UNET_F16X3
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
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = sampler.sample(hcs_dev, K, noise=noise, **kw)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    if fused:
        spt = 1 if B <= n_sm else (3 if B <= 3 * n_sm else 7)   # samples per tile the executor picks (unet_fused.cu build_fused)
        grid = min((B + spt - 1) // spt, n_sm)
        st = stamps[:, :grid].cpu()
        live = (st[:, :, 0] > 0).all(dim=1) & (st[:, :, 1] > 0).all(dim=1)
        dur_ns = (st[:, :, 1].max(dim=1).values - st[:, :, 0].min(dim=1).values)[live].double()
        unet_ms = (dur_ns / 1e6).tolist()          # the last replay's forwards (slots are rewritten by every replay)
    else:
        xx, oo = torch.randn(B, H, D, device=dev), torch.empty(B, H, D, device=dev)
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
    ws = next(iter(sampler._ws.values()))
    exchange_failed = bool(ws["ex"].failed()) if ws.get("ex") is not None else False
    exchange_kind = ("peer memory (fused into ddpm_step_kernel)" if ws.get("ex") is not None else
                     ("nccl all-gather loop" if world > 1 else "none"))

    # ---- e2e: host inputs, device RNG, D2H of the result --------------------------------------------------------------
    result_host = torch.empty(R_local, K, H, D).pin_memory()

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

    # ---- e2e_planner (N=1): the reference's call pattern, one planner call per robot ------------------------------------
    planner = None
    if world == 1 and not args.no_planner:
        planner = planner_bench(M, model, ENV, ta, starts, goals, K, min(R_total, args.planner_calls))

    if rank == 0:
        pk, pk_kind = peaks()
        ms_step = ms_total / args.steps
        n_traj = R_total * K
        value = n_traj / (ms_step / 1e3)
        achieved = B * UNET_FLOP_PER_SAMPLE / (unet_avg_ms / 1e3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        n_unet_launches = 2 if fused else (31 if args.precision == "f16x3_layers" else 1)   # pack_input + persistent forward | pack_input + 30 layer launches | fp32 kernel
        n_guided = math.ceil(0.5 * T) + N_EXTRA
        lock = R_total > 1
        per_guided = 0
        if lock:
            per_guided = (1 if world > 1 else 0) + (1 if ws["peer_hash"] is not None else 0)   # wait_peers, build_peer_hash
        # per reverse step: UNet launches + ddpm_step_kernel (publication fused); once per chain: the first publication
        launches_per_chain = n_steps * (n_unet_launches + 1) + n_guided * per_guided + (1 if lock else 0)
        traffic, traffic_source = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp) and args.config == 4 and fused and world == 1:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_launch"]
            traffic_source = tj["source"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cb_v, nt, sample, _ = cpu_sample(cfg, R_total)
            cpu = {"value": cb_v, "unit": "trajectories/s", "cores": nt, "kind": "port", "sample": sample,
                   "host_cores": os.cpu_count()}
        line = {
            "metric": "denoised trajectories/sec", "value": value, "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16x3",
            "data": "synthetic", "config": workload_config(cfg, world, args.scaling), "finite": finite,
            "e2e": {"value": n_traj / (ms_e2e / args.steps / 1e3), "unit": "trajectories/s",
                    "h2d_bytes_per_step": int(sg_host.numel() * 4), "d2h_bytes_per_step": int(result_host.numel() * 4),
                    "note": "noise drawn on the device by torch.randn as the reference does"},
            "e2e_planner": planner,
            "gpu_launches": launches_per_chain * args.steps,
            "exchange": {"kind": exchange_kind, "timed_out": exchange_failed},
            "roofline": {"bound": "tensor", "kernel": "TemporalUnet forward (" + args.precision + (": unet_fused_kernel, one persistent tcgen05 launch)" if fused else ")"), "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                         "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step); "
                                        f"the kernel runs kind::f16 MMAs at the bf16 rate and executes 3x the algorithmic FLOP (FP16 hi/lo split)",
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


def run_gpu_ensemble(args):
    """BASELINE config 5: one MPDEnsemble (1x2 tiles) planner per robot.  value = the batched multi-tile chain alone (noise
    resident, CUDA events); e2e = mmd_b200.plan_batch(planners) from host start/goal states to the planners' smoothed final
    trajectories on the host; e2e_planner = the reference's own pattern, the planners called one after the other."""
    import mmd_b200 as M
    from mmd_b200 import _lib, planners as PL
    from oracle import port
    cfg = CONFIGS[args.config]
    if args.gpus != 1 or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("config 5 is benchmarked on one GPU (independent planner calls: more GPUs = independent replicas)")
    K, T, R, NT = cfg["K"], cfg["T"], cfg["R"], cfg["tiles"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    unet = M.TemporalUnet(n_support_points=H, state_dim=D, unet_input_dim=32, dim_mults=(1, 2, 4), unet_precision=args.precision)
    unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
    model = M.GaussianDiffusionModel(model=unet, variance_schedule="exponential", n_diffusion_steps=T, predict_epsilon=True).to(dev)
    models = {j: model for j in range(NT)}   # the reference loads the same checkpoint once per tile (inference_multi_agent.py:205-222)
    tile_tf = [torch.tensor([2.0 * j, 0.0]) for j in range(NT)]
    s_loc, g_loc = port.get_start_goal_pos_circle(R, 0.8)
    planners, sg = [], []
    for r in range(R):
        skel = list(range(NT)) if r % 2 == 0 else list(reversed(range(NT)))
        tr = {j: tile_tf[skel[j]] for j in range(NT)}
        start, goal = s_loc[r] + tile_tf[skel[0]], g_loc[r] + tile_tf[skel[-1]]
        planners.append(M.MPDEnsemble((cfg["env"] + "-RobotPlanarDisk",) * NT, tr, "mmd", start, goal, n_samples=K, models=models,
                                      device=str(dev)))
        sg.append(torch.stack((start, goal)))
    sg_host = torch.stack(sg).pin_memory()
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    STAMP_SLOTS = 512
    stamps = torch.zeros(STAMP_SLOTS, n_sm, 2, dtype=torch.int64, device=dev)
    unet.ensure_time_table(T)
    fused = unet.native_mode(args.precision) == _lib.UNET_F16X3
    if fused:
        _lib.check(_lib.lib().mmdk_unet_debug_stamps(unet.native(), _lib.ptr(stamps), STAMP_SLOTS, n_sm))
    result_host = torch.empty(R, K, NT * H, D).pin_memory()

    def step():
        s_g = sg_host.to(dev, non_blocking=True)
        outs = M.plan_batch(planners, None, rng="batched", starts_goals=s_g)
        for r, o in enumerate(outs):
            result_host[r].copy_(o.trajs_final, non_blocking=True)
        return outs

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    stamps.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chain_ms = []
    e0.record()
    for _ in range(args.steps):
        outs = step()
        chain_ms.append(PL.last_batch_chain_ms())
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clk = clocks.stop()
    ms_chain = sum(chain_ms) / len(chain_ms)
    B = R * K
    n_steps = T + N_EXTRA
    if fused:
        spt = 1 if B <= n_sm else (3 if B <= 3 * n_sm else 7)
        grid = min((B + spt - 1) // spt, n_sm)
        st = stamps[:, :grid].cpu()
        live = (st[:, :, 0] > 0).all(dim=1) & (st[:, :, 1] > 0).all(dim=1)
        dur_ns = (st[:, :, 1].max(dim=1).values - st[:, :, 0].min(dim=1).values)[live].double()
        unet_ms = (dur_ns / 1e6).tolist()
    else:
        unet_ms = [float("nan")]
    unet_avg_ms = sum(unet_ms) / len(unet_ms)
    # the reference's pattern: sequential planner calls
    torch.cuda.synchronize()
    planners[0](planners[0].start_state_pos, planners[0].goal_state_pos)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for p in planners:
        p(p.start_state_pos, p.goal_state_pos)
    torch.cuda.synchronize()
    dt_seq = time.perf_counter() - t0
    pk, pk_kind = peaks()
    achieved = B * UNET_FLOP_PER_SAMPLE / (unet_avg_ms / 1e3) / 1e12
    cpu = None
    if not args.no_cpu_baseline:
        cb_v, nt, sample, _ = cpu_sample(cfg, R)
        cpu = {"value": cb_v, "unit": "trajectories/s", "cores": nt, "kind": "port", "sample": sample, "host_cores": os.cpu_count()}
    n_guided = math.ceil(0.5 * T) + N_EXTRA
    line = {
        "metric": "denoised trajectories/sec", "value": B / (ms_chain / 1e3), "unit": "trajectories/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_chain, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f16x3", "data": "synthetic",
        "config": workload_config(cfg, 1, args.scaling),
        "finite": bool(all(torch.isfinite(o.trajs_final).all() for o in outs)),
        "note": f"one trajectory = {NT} tiles x {H} waypoints = {NT} UNet forwards per reverse step; tile-trajectories/s = {NT} x value",
        "e2e": {"value": B / (ms_e2e / 1e3), "unit": "trajectories/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(sg_host.numel() * 4), "d2h_bytes_per_step": int(result_host.numel() * 4),
                "note": "mmd_b200.plan_batch(planners): noise drawn on the device, batched chain, per-planner post-processing "
                        "(unnormalise, classify, stitch tiles, best trajectory, Savitzky-Golay), final trajectories D2H"},
        "e2e_planner": {"value": B / dt_seq, "unit": "trajectories/s", "calls": R, "ms_per_call": 1e3 * dt_seq / R,
                        "note": "sequential MPDEnsemble.__call__(start, goal) -> PlannerOutput (cbs.py:316-324)"},
        # per tile step: pack_input + persistent forward + ddpm_step + one cross-conditioning launch per transform set (2)
        "gpu_launches": args.steps * (n_steps * NT * ((2 if fused else 1) + 1 + 2) + 2),
        "roofline": {"bound": "tensor", "kernel": "TemporalUnet forward (" + args.precision + ")", "achieved": achieved,
                     "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                     "traffic": None, "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained",
                     "algorithmic_flop_per_launch": B * UNET_FLOP_PER_SAMPLE, "avg_launch_ms": unet_avg_ms,
                     "unet_share_of_step": unet_avg_ms * n_steps * NT / ms_chain, "forwards_timed": len(unet_ms),
                     "timing": "%globaltimer stamps written by every CTA of every forward of the last timed chain"},
        "cpu_baseline": cpu, "clocks": clk,
    }
    print(json.dumps(line))


def planner_bench(M, model, env_name, ta, starts, goals, K, n_calls):
    """The reference's call pattern (cbs.py:316-324): one MPD planner object per robot (mpd.py:64-88), called one after the
    other; each call = full guided chain for K samples + the planner's post-processing."""
    dev = ta["device"]
    try:
        planners = [M.MPD(model_id=f"{env_name}-RobotPlanarDisk", start_state_pos=starts[r].to(dev), goal_state_pos=goals[r].to(dev),
                          model=model, n_samples=K, device=str(dev)) for r in range(n_calls)]
    except Exception as e:  # the planner object is optional for the bench line; say why it is missing
        return {"value": None, "unavailable": f"{type(e).__name__}: {e}"}
    planners[0](starts[0].to(dev), goals[0].to(dev))   # warm-up: executor state for B = K

    def one_pass():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_free = 0
        for r in range(n_calls):
            out = planners[r](starts[r].to(dev), goals[r].to(dev))
            n_free += 0 if out.trajs_final_free is None else int(out.trajs_final_free.shape[0])
        torch.cuda.synchronize()
        return time.perf_counter() - t0, n_free

    dt_first, _ = one_pass()      # every planner's FIRST call (lowering of its own guide / task objects)
    dt, n_free = one_pass()       # steady state: CBS calls each planner many times (cbs.py:390-430)
    # the same calls served in one batch (mmd_b200.plan_batch: bit-identical results to the sequential calls)
    M.plan_batch(planners)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    M.plan_batch(planners)
    torch.cuda.synchronize()
    dt_b = time.perf_counter() - t0
    return {"value": n_calls * K / dt, "unit": "trajectories/s", "calls": n_calls, "ms_per_call": 1e3 * dt / n_calls,
            "ms_per_first_call": 1e3 * dt_first / n_calls,
            "collision_free_trajectories": n_free,
            "batched": {"value": n_calls * K / dt_b, "unit": "trajectories/s", "ms_total": 1e3 * dt_b,
                        "note": "mmd_b200.plan_batch(planners): the same planner calls as ONE batched chain + per-planner post-processing"},
            "note": "sequential MPD.__call__(start, goal) -> PlannerOutput, K samples each, no inter-robot constraints (CBS root "
                    "calls), host wall clock including the planner's post-processing (unnormalise the chain, collision "
                    "classification, best-trajectory selection, Savitzky-Golay smoothing)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mmd_b200", choices=["mmd_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MMD_UNET_PRECISION", "f16x3"), choices=["fp32", "f16x3", "f16x3_layers"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=os.environ.get("MMD_BENCH_SCALING", "strong"), choices=["strong", "weak"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-planner", action="store_true")
    ap.add_argument("--planner-calls", type=int, default=32)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 untimed chains
    if args.impl == "reference":
        run_reference(args)
    elif CONFIGS[args.config].get("tiles", 1) > 1:
        run_gpu_ensemble(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
