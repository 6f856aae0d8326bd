"""CPU oracle: a plain-PyTorch restatement of the reference's guided-diffusion sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing in mmd_b200/ may import this module; it is the
checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product path is the CUDA library behind include/mmdk.h.

Pinning status: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against the reference ITSELF:
tests/test_oracle_vs_reference.py imports /root/reference (through oracle/ref_shim.py,
only where that tree exists) and checks every function below against the reference
function it restates; oracle/gen_golden.py dumps reference outputs into tests/golden/
so the pin travels to the GPU box.

Every function cites the reference file:line it follows.  Paths are relative to
/root/reference; TR = deps/torch_robotics/torch_robotics, MPB =
deps/motion_planning_baselines/mp_baselines.

Conventions: x is [B, H, D] fp32 (D = 4: x, y, vx, vy), normalised to [-1, 1] by a
LimitsNormalizer.  Noise is always passed in explicitly (SURVEY hard part d): noise[0]
is the x_T draw of p_sample_loop (diffusion_model_base.py:191) and noise[1 + k] is the
randn_like draw of the k-th call of ddpm_sample_fn (sample_functions.py:74).
"""
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants (mmd/config/mmd_params.py:28-63, TR/robots/robot_planar_disk.py:59-73)
# ----------------------------------------------------------------------------------------------
ROBOT_RADIUS = 0.05
HORIZON = 64
STATE_DIM = 4
UNET_DIM_MULTS = {0: (1, 2, 4), 1: (1, 2, 4, 8)}  # temporal_unet.py:17-20


# ----------------------------------------------------------------------------------------------
# D1: variance schedule and the 11 buffers (diffusion_model_base.py:69-105, helpers.py:29-49)
# ----------------------------------------------------------------------------------------------
def exponential_beta_schedule(n_steps, beta_start=1e-4, beta_end=1.0):
    x = torch.linspace(0, n_steps, n_steps)
    beta_start = torch.tensor(beta_start)
    beta_end = torch.tensor(beta_end)
    a = 1 / n_steps * torch.log(beta_end / beta_start)
    return beta_start * torch.exp(a * x)


def cosine_beta_schedule(n_steps, s=0.008, a_min=0, a_max=0.999):
    steps = n_steps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.tensor(np.clip(betas, a_min=a_min, a_max=a_max), dtype=torch.float32)


def make_schedule(n_steps, variance_schedule="exponential") -> Dict[str, torch.Tensor]:
    if variance_schedule == "exponential":
        betas = exponential_beta_schedule(n_steps)
    elif variance_schedule == "cosine":
        betas = cosine_beta_schedule(n_steps)
    else:
        raise NotImplementedError(variance_schedule)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, axis=0)
    ac_prev = torch.cat([torch.ones(1), ac[:-1]])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(torch.clamp(post_var, min=1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


# ----------------------------------------------------------------------------------------------
# U0-U7: TemporalUnet forward, functional over a reference-keyed state dict
# (temporal_unet.py:23-174, layers.py:197-398)
# ----------------------------------------------------------------------------------------------
def unet_param_shapes(state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4), time_emb_dim=32,
                      self_attention=False) -> Dict[str, tuple]:
    """Names/shapes of TemporalUnet.state_dict() for conditioning_type=None (temporal_unet.py:66-119)."""
    dims = [state_dim] + [unet_input_dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    sh = {}
    sh["time_mlp.encoder.1.weight"] = (128, 32)
    sh["time_mlp.encoder.1.bias"] = (128,)
    sh["time_mlp.encoder.3.weight"] = (time_emb_dim, 128)
    sh["time_mlp.encoder.3.bias"] = (time_emb_dim,)

    def rtb(prefix, ci, co):
        for j, (a, b) in enumerate([(ci, co), (co, co)]):
            sh[f"{prefix}.blocks.{j}.block.0.weight"] = (b, a, 5)
            sh[f"{prefix}.blocks.{j}.block.0.bias"] = (b,)
            sh[f"{prefix}.blocks.{j}.block.2.weight"] = (b,)
            sh[f"{prefix}.blocks.{j}.block.2.bias"] = (b,)
        sh[f"{prefix}.cond_mlp.1.weight"] = (co, time_emb_dim)
        sh[f"{prefix}.cond_mlp.1.bias"] = (co,)
        if ci != co:
            sh[f"{prefix}.residual_conv.weight"] = (co, ci, 1)
            sh[f"{prefix}.residual_conv.bias"] = (co,)

    def attn(prefix, c):
        sh[f"{prefix}.fn.norm.g"] = (1, c, 1)
        sh[f"{prefix}.fn.norm.b"] = (1, c, 1)
        sh[f"{prefix}.fn.fn.to_qkv.weight"] = (384, c, 1)
        sh[f"{prefix}.fn.fn.to_out.weight"] = (c, 128, 1)
        sh[f"{prefix}.fn.fn.to_out.bias"] = (c,)

    n = len(in_out)
    for i, (ci, co) in enumerate(in_out):
        rtb(f"downs.{i}.0", ci, co)
        rtb(f"downs.{i}.1", co, co)
        if self_attention:
            attn(f"downs.{i}.2", co)
        if i < n - 1:
            sh[f"downs.{i}.4.conv.weight"] = (co, co, 3)
            sh[f"downs.{i}.4.conv.bias"] = (co,)
    mid = dims[-1]
    rtb("mid_block1", mid, mid)
    if self_attention:
        attn("mid_attn", mid)
    rtb("mid_block2", mid, mid)
    for i, (ci, co) in enumerate(reversed(in_out[1:])):
        rtb(f"ups.{i}.0", co * 2, ci)
        rtb(f"ups.{i}.1", ci, ci)
        if self_attention:
            attn(f"ups.{i}.2", ci)
        sh[f"ups.{i}.4.conv.weight"] = (ci, ci, 4)  # ConvTranspose1d: [C_in, C_out, k]
        sh[f"ups.{i}.4.conv.bias"] = (ci,)
    sh["final_conv.0.block.0.weight"] = (unet_input_dim, unet_input_dim, 5)
    sh["final_conv.0.block.0.bias"] = (unet_input_dim,)
    sh["final_conv.0.block.2.weight"] = (unet_input_dim,)
    sh["final_conv.0.block.2.bias"] = (unet_input_dim,)
    sh["final_conv.1.weight"] = (state_dim, unet_input_dim, 1)
    sh["final_conv.1.bias"] = (state_dim,)
    return sh


def make_unet_params(seed=0, out_scale=1.0, **cfg) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoints ship with the reference).

    Fan-in-scaled uniform like torch's default init; GroupNorm/LayerNorm affine perturbed
    away from (1, 0) so the affine path is exercised.  `out_scale` scales the final 1x1 conv
    (a small value keeps x0 out of the +-1 clamp and exercises the unsaturated regime).
    """
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k, s in unet_param_shapes(**cfg).items():
        if ".block.2." in k or k.endswith("norm.g") or k.endswith("norm.b"):
            base = 1.0 if (k.endswith("weight") or k.endswith(".g")) else 0.0
            p[k] = base + 0.1 * torch.randn(s, generator=g)
        else:
            fan_in = int(np.prod(s[1:])) if len(s) > 1 else None
            if fan_in is None:  # bias: use the matching weight's fan-in
                w = p[k[:-4] + "weight"]
                fan_in = int(np.prod(w.shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            p[k] = (torch.rand(s, generator=g) * 2 - 1) * bound
    p["final_conv.1.weight"] = p["final_conv.1.weight"] * out_scale
    p["final_conv.1.bias"] = p["final_conv.1.bias"] * out_scale
    return p


def sinusoidal_pos_emb(t, dim=32):  # layers.py:246-258
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def _conv1d_block(p, pre, x):  # layers.py:279-296 (Conv1d k5 p2 -> GroupNorm(8) -> Mish)
    w = p[pre + ".block.0.weight"]
    x = F.conv1d(x, w, p[pre + ".block.0.bias"], padding=w.shape[-1] // 2)
    x = F.group_norm(x, group_norm_n_groups(w.shape[0]), p[pre + ".block.2.weight"], p[pre + ".block.2.bias"], 1e-5)
    return F.mish(x)


def group_norm_n_groups(n_channels, target=8):  # layers.py:392-398
    if n_channels < target:
        return 1
    for n in range(target, target + 10):
        if n_channels % n == 0:
            return n
    return 1


def _rtb(p, pre, x, c, ops=None):  # layers.py:346-358
    h = _conv1d_block(p, pre + ".blocks.0", x)
    h = h + F.linear(F.mish(c), p[pre + ".cond_mlp.1.weight"], p[pre + ".cond_mlp.1.bias"])[:, :, None]
    if ops is not None:
        ops.append((pre + ".blocks.0", h))
    h = _conv1d_block(p, pre + ".blocks.1", h)
    if pre + ".residual_conv.weight" in p:
        res = F.conv1d(x, p[pre + ".residual_conv.weight"], p[pre + ".residual_conv.bias"])
    else:
        res = x
    if ops is not None:
        ops.append((pre + ".blocks.1", h + res))
    return h + res


def _linear_attention(p, pre, x, heads=4, dim_head=32):  # layers.py:177-229 Residual(PreNorm(LinearAttention))
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    xn = (x - mean) / (var + 1e-5).sqrt() * p[pre + ".fn.norm.g"] + p[pre + ".fn.norm.b"]
    qkv = F.conv1d(xn, p[pre + ".fn.fn.to_qkv.weight"]).chunk(3, dim=1)
    b, _, n = x.shape
    q, k, v = [t.reshape(b, heads, dim_head, n) for t in qkv]
    q = q * dim_head ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q).reshape(b, heads * dim_head, n)
    return F.conv1d(out, p[pre + ".fn.fn.to_out.weight"], p[pre + ".fn.fn.to_out.bias"]) + x


def unet_forward(p: Dict[str, torch.Tensor], x, t, taps: Optional[dict] = None):
    """TemporalUnet.forward(x, t, context=None) (temporal_unet.py:121-174).  x [B,H,D], t [B] -> eps [B,H,D]."""
    temb = sinusoidal_pos_emb(t.to(torch.float32))
    temb = F.linear(temb, p["time_mlp.encoder.1.weight"], p["time_mlp.encoder.1.bias"])
    temb = F.linear(F.mish(temb), p["time_mlp.encoder.3.weight"], p["time_mlp.encoder.3.bias"])
    n_levels = len([k for k in p if k.startswith("downs.") and k.endswith(".0.blocks.0.block.0.weight")])
    attn = any(".fn.fn.to_qkv." in k for k in p)
    x = x.transpose(1, 2)
    hs = []
    ops = None  # per-op activations in layer-program order (debug taps of the executors)
    if taps is not None:
        ops = taps.setdefault("ops", [])
    for i in range(n_levels):
        x = _rtb(p, f"downs.{i}.0", x, temb, ops)
        x = _rtb(p, f"downs.{i}.1", x, temb, ops)
        if attn:
            x = _linear_attention(p, f"downs.{i}.2", x)
        hs.append(x)
        if i < n_levels - 1:
            x = F.conv1d(x, p[f"downs.{i}.4.conv.weight"], p[f"downs.{i}.4.conv.bias"], stride=2, padding=1)
            if ops is not None:
                ops.append((f"downs.{i}.4", x))
        if taps is not None:
            taps[f"down{i}"] = x
    x = _rtb(p, "mid_block1", x, temb, ops)
    if attn:
        x = _linear_attention(p, "mid_attn", x)
    x = _rtb(p, "mid_block2", x, temb, ops)
    if taps is not None:
        taps["mid"] = x
    for i in range(n_levels - 1):
        x = torch.cat((x, hs.pop()), dim=1)
        x = _rtb(p, f"ups.{i}.0", x, temb, ops)
        x = _rtb(p, f"ups.{i}.1", x, temb, ops)
        if attn:
            x = _linear_attention(p, f"ups.{i}.2", x)
        x = F.conv_transpose1d(x, p[f"ups.{i}.4.conv.weight"], p[f"ups.{i}.4.conv.bias"], stride=2, padding=1)
        if ops is not None:
            ops.append((f"ups.{i}.4", x))
        if taps is not None:
            taps[f"up{i}"] = x
    x = _conv1d_block(p, "final_conv.0", x)
    if ops is not None:
        ops.append(("final_conv.0", x))
    x = F.conv1d(x, p["final_conv.1.weight"], p["final_conv.1.bias"])
    return x.transpose(1, 2)


# ----------------------------------------------------------------------------------------------
# G3: LimitsNormalizer (mmd/datasets/normalization.py:150-168)
# ----------------------------------------------------------------------------------------------
class LimitsNormalizer:
    def __init__(self, mins, maxs):
        self.mins = torch.as_tensor(mins, dtype=torch.float32)
        self.maxs = torch.as_tensor(maxs, dtype=torch.float32)

    def normalize(self, x):
        x = (x - self.mins) / (self.maxs - self.mins)
        return 2 * x - 1

    def unnormalize(self, x, eps=1e-4):
        if x.max() > 1 + eps or x.min() < -1 - eps:  # global, data dependent
            x = torch.clip(x, -1, 1)
        x = (x + 1) / 2.0
        return x * (self.maxs - self.mins) + self.mins


DEFAULT_NORMALIZER_LIMITS = ([-1.0, -1.0, -2.0, -2.0], [1.0, 1.0, 2.0, 2.0])  # SURVEY 8d synthetic normaliser


# ----------------------------------------------------------------------------------------------
# G6 + environments: analytic SDFs -> 400x400 grid (+ autograd gradient grid)
# (TR/environments/grid_map_sdf.py:9-114, primitives.py:108-117, 312-333, env_*_2d.py tables)
# ----------------------------------------------------------------------------------------------
# Obstacle tables: (box centres, box sizes) of the single fixed ObjectField of each env; all
# objects sit at identity pose and every "extra objects" field is an EMPTY MultiSphereField.
ENV_BOXES = {
    "EnvEmpty2D": ([], []),  # env_empty_2d.py
    "EnvEmptyNoWait2D": ([], []),  # env_empty_nowait_2d.py
    "EnvConveyor2D": ([[0.0, 0.0], [0.0, 0.35], [0.0, -0.35]],  # env_conveyor_2d.py:47-68
                      [[0.8, 0.1], [1.0, 0.1], [1.0, 0.1]]),
    "EnvHighways2D": ([[0, 0.0], [0.0, 0.875], [0.0, -0.875], [0.875, 0.0], [-0.875, 0.0],  # env_highways_2d.py:48-79
                       [0.875, 0.875], [0.875, -0.875], [-0.875, 0.875], [-0.875, -0.875]],
                      [[0.5, 0.5], [0.5, 0.25], [0.5, 0.25], [0.25, 0.5], [0.25, 0.5],
                       [0.25, 0.25], [0.25, 0.25], [0.25, 0.25], [0.25, 0.25]]),
    "EnvDropRegion2D": ([[0.4, 0.4], [-0.4, 0.4], [0.4, -0.4], [-0.4, -0.4]],  # env_drop_region_2d.py:47-98
                        [[0.4, 0.4], [0.4, 0.4], [0.4, 0.4], [0.4, 0.4]]),
}
ENV_HAS_SPHERE_FIELD = {"EnvEmpty2D": True, "EnvEmptyNoWait2D": True, "EnvConveyor2D": True,
                        "EnvHighways2D": True, "EnvDropRegion2D": False}


def rounded_box_sdf(x, centers, sizes):  # primitives.py:312-333
    half = sizes / 2.0
    radius = torch.min(sizes, dim=-1)[0] * 0.15
    d = torch.abs(x.unsqueeze(-2) - centers.unsqueeze(0))
    q = d - half.unsqueeze(0) + radius.unsqueeze(0).unsqueeze(-1)
    max_q = torch.amax(q, dim=-1)
    sdfs = torch.minimum(max_q, torch.zeros_like(max_q)) + torch.linalg.norm(torch.relu(q), dim=-1) - radius.unsqueeze(0)
    return torch.min(sdfs, dim=-1)[0]


def env_signed_distance(env_name, x):
    """ObjectField.compute_signed_distance_impl at identity pose (primitives.py:554-572): min over primitive fields;
    an empty MultiSphereField returns ones (primitives.py:108-110)."""
    centers, sizes = ENV_BOXES[env_name]
    fields = []
    if ENV_HAS_SPHERE_FIELD[env_name]:
        fields.append(torch.ones_like(x[..., 0]))
    if len(centers):
        fields.append(rounded_box_sdf(x, torch.tensor(centers, dtype=torch.float32), torch.tensor(sizes, dtype=torch.float32)))
    return torch.min(torch.stack(fields, dim=-1), dim=-1)[0]


def build_sdf_grid(env_name, cell_size=0.005, limits=((-1.0, -1.0), (1.0, 1.0))):
    """GridMapSDF.precompute_sdf (grid_map_sdf.py:34-59): nodes at linspace(lo, hi, n) (NOT cell centres),
    value = analytic SDF, gradient = autograd of the summed SDF.  Returns sdf [n,n], grad [n,n,2]."""
    lo = torch.tensor(limits[0], dtype=torch.float32)
    hi = torch.tensor(limits[1], dtype=torch.float32)
    n = torch.ceil(torch.abs(hi - lo) / cell_size).long()
    bx = torch.linspace(lo[0], hi[0], int(n[0]))
    by = torch.linspace(lo[1], hi[1], int(n[1]))
    pts = torch.stack(torch.meshgrid(bx, by, indexing="ij"), dim=-1)
    sdf_rows, grad_rows = [], []
    for i in range(pts.shape[0]):  # row by row, as the reference does
        row = pts[i].clone().requires_grad_(True)
        s = env_signed_distance(env_name, row)
        # torch.autograd.functional.jacobian returns zeros when the output does not depend on the input (empty envs)
        g = torch.autograd.grad(s.sum(), row, allow_unused=True)[0] if s.requires_grad else None
        sdf_rows.append(s.detach())
        grad_rows.append(torch.zeros_like(row) if g is None else g)
    return torch.stack(sdf_rows), torch.stack(grad_rows)


class GridSDF:
    """GridMapSDF.get_sdf (grid_map_sdf.py:84-114)."""

    def __init__(self, sdf, grad, limits=((-1.0, -1.0), (1.0, 1.0))):
        self.sdf, self.grad = sdf, grad
        self.lo = torch.tensor(limits[0], dtype=torch.float32)
        self.map_dim = torch.abs(torch.tensor(limits[1], dtype=torch.float32) - self.lo)
        self.cmap_dim = torch.tensor(sdf.shape[:2], dtype=torch.long)

    def cell_index(self, X):
        idx = ((X - self.lo) / self.map_dim * self.cmap_dim).floor().to(torch.int)
        max_idx = (self.cmap_dim - 1).to(torch.int)
        return idx.clamp(torch.zeros_like(max_idx), max_idx)

    def __call__(self, X):
        idx = self.cell_index(X).detach()
        q = torch.unbind(idx, dim=-1)
        s = self.sdf[q].clone()
        g = self.grad[q]
        s += (X * g).sum(-1) - (X.detach() * g).sum(-1)  # value-neutral surrogate that routes autograd to g
        return s


# ----------------------------------------------------------------------------------------------
# G1-G11: the guide (guides.py:152-253, MPB/planners/costs/cost_functions.py, factors/*.py,
# TR/torch_planning_objectives/fields/distance_fields.py)
# ----------------------------------------------------------------------------------------------
class Constraint:
    """One CostConstraint object (cost_functions.py:275-326): n vertex constraints, differentiated / clipped /
    weighted as ONE cost term."""

    def __init__(self, qs, traj_ranges, radii, is_soft=True, weight=None):
        self.qs = torch.as_tensor(qs, dtype=torch.float32).reshape(-1, 2)
        self.traj_ranges = torch.as_tensor(traj_ranges, dtype=torch.float32).reshape(-1, 2)
        self.radii = torch.as_tensor(radii, dtype=torch.float32).reshape(-1)
        self.is_soft = is_soft
        self.weight = weight if weight is not None else (2e-2 if is_soft else 2e-1)  # mmd_params.py:42-43

    def eval(self, trajs):  # cost_functions.py:297-326 (materialises [n, B, H, 2] exactly like the reference)
        q_pos = trajs[..., :2]
        s, e = self.traj_ranges[:, 0], self.traj_ranges[:, 1]
        mask = torch.arange(q_pos.shape[1]).unsqueeze(0).unsqueeze(0)
        mask = (mask >= s.view(-1, 1, 1)) & (mask < e.view(-1, 1, 1))
        q_pos = q_pos.unsqueeze(0).expand(self.qs.shape[0], -1, -1, -1)
        q_masked = q_pos * mask.unsqueeze(-1)
        d = torch.norm(q_masked - self.qs.view(-1, 1, 1, 2), dim=-1)
        d = torch.where(d > self.radii.view(-1, 1, 1), torch.zeros_like(d), d)
        return (self.radii.view(-1, 1, 1) - d).sum(dim=-1).sum()


class GuideSpec:
    """Everything GuideManagerTrajectoriesWithVelocity.forward reads (mpd.py:215-265)."""

    def __init__(self, grid: Optional[GridSDF], normalizer: LimitsNormalizer,
                 cutoff_margin=0.05, dt=5.0 / 64, ws_limits=((-1.0, -1.0), (1.0, 1.0)),
                 w_collision=2e-2, w_smooth=8e-2, max_grad_norm=1.0, sigma_coll=1.0, sigma_gp=1.0):
        self.grid = grid
        self.normalizer = normalizer
        # distance_fields.py:115: fp32 link-margin tensor (robot_planar_disk.py:68) + python cutoff margin
        self.margin = torch.tensor([ROBOT_RADIUS * 1.1], dtype=torch.float32) + cutoff_margin
        # tasks.py:82-84: workspace boundaries scaled by 1.08
        self.ws_min = torch.tensor(ws_limits[0], dtype=torch.float32) * 1.08
        self.ws_max = torch.tensor(ws_limits[1], dtype=torch.float32) * 1.08
        self.dt = dt
        self.w_collision, self.w_smooth = w_collision, w_smooth
        self.max_grad_norm = max_grad_norm
        self.sigma_coll, self.sigma_gp = sigma_coll, sigma_gp
        self.extra: List[Constraint] = []
        # GPFactor (gp_factor.py:34-50)
        eye, z = torch.eye(2), torch.zeros(2, 2)
        self.phi = torch.cat((torch.cat((eye, dt * eye), 1), torch.cat((z, eye), 1)), 0)
        qc = eye / sigma_gp ** 2
        m1, m2, m3 = 12.0 * (dt ** -3.0) * qc, -6.0 * (dt ** -2.0) * qc, 4.0 * (dt ** -1.0) * qc
        self.q_inv = torch.cat((torch.cat((m1, m2), -1), torch.cat((m2, m3), -1)), -2)

    # -- individual costs -----------------------------------------------------------------------
    def _field_cost(self, sdf_stack):  # distance_fields.py:110-129 + field_factor.py:24-48
        c = torch.relu(-(sdf_stack - self.margin))  # [B, H-1, n_sdfs, 1]
        c = c.max(-2)[0].sum(-1)  # max over sdfs, sum over links (1 link)
        c = torch.clamp(c, min=-0.02)
        return (1.0 / self.sigma_coll ** 2) * c.sum(1)  # cost_functions.py:190-191

    def cost_collision_objects(self, xu):  # CostCollision(field=CollisionObjectDistanceField)
        p = xu[:, 1:, :2].unsqueeze(-2)  # FieldFactor traj_range [1, None]; identity FK adds the link dim
        flat = p.reshape(-1, 2)
        dfs = [self.grid(flat).view(p.shape[:-1]), torch.ones_like(flat[..., 0]).view(p.shape[:-1])]  # grid + empty extra
        return self._field_cost(torch.stack(dfs, dim=-2))

    def cost_collision_border(self, xu):  # CollisionWorkspaceBoundariesDistanceField (distance_fields.py:354-367)
        p = xu[:, 1:, :2].unsqueeze(-2)
        smin = p - self.ws_min
        smin = torch.sign(smin) * torch.abs(smin)
        smax = self.ws_max - p
        smax = torch.sign(smax) * torch.abs(smax)
        sd = torch.cat((smin, smax), dim=-1).transpose(-2, -1)
        return self._field_cost(sd)

    def cost_gp(self, xu):  # CostGPTrajectory.eval (cost_functions.py:532-542)
        s1, s2 = xu[:, :-1].unsqueeze(-1), xu[:, 1:].unsqueeze(-1)
        err = s2 - self.phi @ s1
        c = err.transpose(2, 3) @ self.q_inv.reshape(1, 1, 4, 4) @ err
        return c.sum(1).squeeze()

    def clip(self, g):  # guides.py:247-253
        n = torch.linalg.norm(g + 1e-6, dim=-1, keepdims=True)
        return torch.clip(n, 0.0, self.max_grad_norm) / n * g

    def cost_terms(self, xu):
        terms = []
        if self.grid is not None:
            terms.append((self.cost_collision_objects(xu), self.w_collision))
        terms.append((self.cost_collision_border(xu), self.w_collision))
        terms.append((self.cost_gp(xu), self.w_smooth))
        for c in self.extra:
            terms.append((c.eval(xu), c.weight))
        return terms

    def __call__(self, x_normalized, return_parts=False):  # guides.py:180-226
        x = x_normalized.clone()
        with torch.enable_grad():
            x.requires_grad_(True)
            xu = self.normalizer.unnormalize(x)
            # (guides.py:189-191: an interpolated copy is computed here and never consumed -- SURVEY G1)
            grad = 0
            parts = []
            for cost, w in self.cost_terms(xu):
                g = torch.autograd.grad([cost.sum()], [xu], retain_graph=True)[0]
                gc = self.clip(g)
                gc[..., 0, :] = 0.0
                gc[..., -1, :] = 0.0
                grad = grad + w * gc
                parts.append((g, gc))
        grad = -1.0 * grad
        return (grad, parts) if return_parts else grad


# ----------------------------------------------------------------------------------------------
# D2-D8: sampling (sample_functions.py:8-107, diffusion_model_base.py:126-211, 321-433)
# ----------------------------------------------------------------------------------------------
def apply_hard_conditioning(x, conditions):  # sample_functions.py:8-14
    for t, val in conditions.items():
        x[:, t, :] = val.clone()
    return x


def extract(a, t, x_shape):  # sample_functions.py:34-37
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


class DiffusionModel:
    """GaussianDiffusionModel restricted to inference (diffusion_model_base.py:48-433)."""

    def __init__(self, unet_params, n_diffusion_steps=100, variance_schedule="exponential", predict_epsilon=True,
                 unet_fn=None):
        self.p = unet_params
        self.n_diffusion_steps = n_diffusion_steps
        self.predict_epsilon = predict_epsilon
        self.sched = make_schedule(n_diffusion_steps, variance_schedule)
        self.unet_fn = unet_fn or (lambda x, t: unet_forward(self.p, x, t))

    def p_mean_variance(self, x, t):  # :126-160
        s = self.sched
        eps = self.unet_fn(x, t)
        if self.predict_epsilon:
            x_recon = extract(s["sqrt_recip_alphas_cumprod"], t, x.shape) * x - \
                      extract(s["sqrt_recipm1_alphas_cumprod"], t, x.shape) * eps
        else:
            x_recon = eps
        x_recon.clamp_(-1.0, 1.0)
        mean = extract(s["posterior_mean_coef1"], t, x.shape) * x_recon + extract(s["posterior_mean_coef2"], t, x.shape) * x
        return mean, extract(s["posterior_log_variance_clipped"], t, x.shape)

    def q_sample(self, x_start, t, noise):  # :425-433
        s = self.sched
        return extract(s["sqrt_alphas_cumprod"], t, x_start.shape) * x_start + \
               extract(s["sqrt_one_minus_alphas_cumprod"], t, x_start.shape) * noise


@torch.no_grad()
def ddpm_sample_fn(model: DiffusionModel, x, hard_conds, t, noise, guide=None, n_guide_steps=1,
                   t_start_guide=float("inf"), noise_std=1.0):  # sample_functions.py:41-86
    t_single = int(t[0])
    if t_single < 0:
        t = torch.zeros_like(t)
    x, logvar = model.p_mean_variance(x, t)
    std = torch.exp(0.5 * logvar)
    if guide is not None and t_single < t_start_guide:
        x = guide_gradient_steps(x, hard_conds, guide, n_guide_steps)
    noise = noise.clone()
    noise[t == 0] = 0
    return x + std * noise * noise_std


def guide_gradient_steps(x, hard_conds, guide, n_guide_steps):  # sample_functions.py:89-107
    for _ in range(n_guide_steps):
        x = x + guide(x)
        x = apply_hard_conditioning(x, hard_conds)
    return x


@torch.no_grad()
def p_sample_loop(model: DiffusionModel, hard_conds, noise, n_diffusion_steps, n_extra=1, x_init=None,
                  **sample_kwargs):  # diffusion_model_base.py:163-211
    """noise: [n_diffusion_steps + n_extra + 1, B, H, D] (noise[0] unused when x_init is given)."""
    x = noise[0].clone() if x_init is None else x_init
    B = x.shape[0]
    x = apply_hard_conditioning(x, hard_conds)
    chain = [x]
    k = 1
    for i in reversed(range(-n_extra, n_diffusion_steps)):
        t = torch.full((B,), i, dtype=torch.long)
        x = ddpm_sample_fn(model, x, hard_conds, t, noise[k], **sample_kwargs)
        x = apply_hard_conditioning(x, hard_conds)
        chain.append(x)
        k += 1
    return torch.stack(chain, dim=0)  # [steps + 2, B, H, D], the layout run_inference returns (:345)


def repeat_hard_conds(hard_conds, n_samples):  # diffusion_model_base.py:327-329
    return {k: v.reshape(1, -1).repeat(n_samples, 1) for k, v in hard_conds.items()}


def run_inference(model, hard_conds, n_samples, noise, guide=None, n_guide_steps=20, t_start_guide=None,
                  noise_std=0.5, n_extra=1):  # diffusion_model_base.py:321-351 with mpd.py:299-304,416-424
    if t_start_guide is None:
        t_start_guide = math.ceil(0.5 * model.n_diffusion_steps)  # mmd_params.py:37, mpd.py:267
    return p_sample_loop(model, repeat_hard_conds(hard_conds, n_samples), noise, model.n_diffusion_steps, n_extra,
                         guide=guide, n_guide_steps=n_guide_steps, t_start_guide=t_start_guide, noise_std=noise_std)


def run_local_inference(model, seed_trajectory_b, n_noising_steps, n_denoising_steps, hard_conds, n_samples, noise,
                        **kw):  # diffusion_model_base.py:353-421; noise[0] is the q_sample draw (:383-384)
    B = seed_trajectory_b.shape[0]
    t = torch.full((B,), n_noising_steps, dtype=torch.long)
    x_init = model.q_sample(seed_trajectory_b, t, noise[0])
    kw.setdefault("t_start_guide", math.ceil(0.5 * model.n_diffusion_steps))
    kw.setdefault("n_guide_steps", 20)
    kw.setdefault("noise_std", 0.5)
    n_extra = kw.pop("n_extra", 1)
    return p_sample_loop(model, repeat_hard_conds(hard_conds, n_samples), noise, n_denoising_steps, n_extra,
                         x_init=x_init, **kw)


# ----------------------------------------------------------------------------------------------
# D9: multi-tile ensemble (diffusion_ensemble.py:56-106, sample_functions.py:17-31)
# ----------------------------------------------------------------------------------------------
def apply_cross_conditioning(x, conditions, transforms):
    for (m1, m2), (ind1, ind2) in conditions.items():
        rel = transforms[m2] - transforms[m1]
        if x[m1].shape[2] > rel.shape[0]:
            rel = torch.cat([rel, torch.zeros(x[m1].shape[2] - rel.shape[0])])
        boundary = rel / torch.norm(rel, keepdim=True)
        boundary[boundary == 0] = 1e6
        x[m1][:, ind1, :] = torch.min(x[m2][:, ind2, :] + rel, boundary)
        x[m2][:, ind2, :] = torch.max(x[m1][:, ind1, :] - rel, -boundary)
    return x


@torch.no_grad()
def ensemble_p_sample_loop(models: Dict[int, DiffusionModel], hard_conds: Dict[int, dict], cross_conds, transforms,
                           noise: Dict[int, torch.Tensor], n_diffusion_steps, n_extra, sample_kwargs: Dict[int, dict]):
    """noise[m]: [n_steps + n_extra + 1, B, H, D] per tile m; draw order of the reference is tile-major within a
    step (diffusion_ensemble.py:68-95), which explicit per-tile noise makes irrelevant."""
    x = {}
    for m in models:
        x[m] = noise[m][0].clone()
        x[m] = apply_hard_conditioning(x[m], hard_conds.setdefault(m, {}))
    x = apply_cross_conditioning(x, cross_conds, transforms)
    chains = {m: [x[m]] for m in models}  # NB: holds references; later in-place stitching aliases (reference quirk)
    B = next(iter(x.values())).shape[0]
    k = 1
    for i in reversed(range(-n_extra, n_diffusion_steps)):
        t = torch.full((B,), i, dtype=torch.long)
        for m in models:
            x[m] = ddpm_sample_fn(models[m], x[m], hard_conds[m], t, noise[m][k], **sample_kwargs[m])
            x[m] = apply_hard_conditioning(x[m], hard_conds[m])
            x = apply_cross_conditioning(x, cross_conds, transforms)
        for m in models:
            chains[m].append(x[m])
        k += 1
    return {m: torch.stack(v, dim=0) for m, v in chains.items()}


# ----------------------------------------------------------------------------------------------
# 8(e): lock-step multi-robot composition of reference functions
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def lockstep_sample(model: DiffusionModel, guides: Sequence[GuideSpec], hard_conds_l: Sequence[dict], n_samples, noise,
                    n_guide_steps=20, t_start_guide=None, noise_std=0.5, n_extra=1, radius=2.4 * ROBOT_RADIUS,
                    weight=2e-2, rep_index=0, robots: Optional[Sequence[int]] = None, return_chain=False):
    """All robots denoise simultaneously.  Before each reverse step every robot r receives, as ONE soft CostConstraint,
    the current unnormalised positions [H, 2] of the representative sample (`rep_index`) of every other robot, one
    vertex constraint per waypoint with range (h, h + 1), radius 2.4 r (mmd_params.py:52, cbs.py:468-508), then runs
    the reference's own per-timestep function ddpm_sample_fn.  noise: [R, T + n_extra + 1, K, H, D].
    `robots`: subset of robot indices to advance each step (bounded CPU-baseline samples); the others are frozen."""
    R = len(guides)
    T = model.n_diffusion_steps
    if t_start_guide is None:
        t_start_guide = math.ceil(0.5 * T)
    hcs = [repeat_hard_conds(hc, n_samples) for hc in hard_conds_l]
    xs = [apply_hard_conditioning(noise[r, 0].clone(), hcs[r]) for r in range(R)]
    chains = [[x] for x in xs]
    active = list(range(R)) if robots is None else list(robots)
    H = xs[0].shape[1]
    k = 1
    for i in reversed(range(-n_extra, T)):
        t = torch.full((n_samples,), i, dtype=torch.long)
        reps = [guides[r].normalizer.unnormalize(xs[r][rep_index:rep_index + 1].clone())[0, :, :2] for r in range(R)]
        new = list(xs)
        for r in active:
            others = [q for j, q in enumerate(reps) if j != r]
            if others:
                qs = torch.cat(others, 0)
                hh = torch.arange(H, dtype=torch.float32).repeat(len(others))
                guides[r].extra = [Constraint(qs, torch.stack((hh, hh + 1), -1), torch.full((qs.shape[0],), radius),
                                              is_soft=True, weight=weight)]
            else:
                guides[r].extra = []
            x = ddpm_sample_fn(model, xs[r], hcs[r], t, noise[r, k], guide=guides[r], n_guide_steps=n_guide_steps,
                               t_start_guide=t_start_guide, noise_std=noise_std)
            new[r] = apply_hard_conditioning(x, hcs[r])
            guides[r].extra = []
        xs = new
        if return_chain:
            for r in range(R):
                chains[r].append(xs[r])
        k += 1
    if return_chain:
        return torch.stack([torch.stack(c, 0) for c in chains], 0)  # [R, T+2, K, H, D]
    return torch.stack(xs, 0)


# ----------------------------------------------------------------------------------------------
# I1-I4: integer outputs
# ----------------------------------------------------------------------------------------------
def check_rr_collisions(robot_q, radius=ROBOT_RADIUS):  # robot_planar_disk.py:173-203
    margin = 2.1 * radius
    p1, p2 = robot_q.unsqueeze(-2), robot_q.unsqueeze(-3)
    n = torch.norm(p1 - p2, dim=-1)
    coll = n < margin
    coll = coll & ~torch.eye(coll.shape[-1], dtype=coll.dtype)
    pts = (p1 + p2) / 2
    pts = pts * coll.unsqueeze(-1)
    pts[~coll.unsqueeze(-1).expand_as(pts)] = float("nan")
    return coll, pts


def interpolate_traj_via_points(trajs, num_interpolation=5):  # TR/trajectory/utils.py:73-86
    H, D = trajs.shape[-2:]
    alpha = torch.linspace(0, 1, num_interpolation + 2).type_as(trajs)[1:num_interpolation + 1]
    alpha = alpha.view((1,) * len(trajs.shape[:-1]) + (-1, 1))
    out = trajs[..., 0:H - 1, None, :] * alpha + trajs[..., 1:H, None, :] * (1 - alpha)
    return out.view(trajs.shape[:-2] + (-1, D))


def trajs_collision_mask(trajs, grid: Optional[GridSDF], ws_limits=((-1.0, -1.0), (1.0, 1.0)), cutoff_margin=0.05,
                         radius=ROBOT_RADIUS, num_interpolation=5):
    """PlanningTask.compute_collision(field_type='occupancy', margin=radius) on the interpolated trajectory
    (tasks.py:190-254, distance_fields.py:318-326): a waypoint collides if any sdf < margin (objects: grid and the
    constant-1 extra field; border: the four wall distances at 1.08 x limits).  Returns bool [B, (H-1)*n_interp]."""
    ti = interpolate_traj_via_points(trajs, num_interpolation)
    p = ti[..., :2]
    coll = torch.zeros(p.shape[:-1], dtype=torch.bool)
    if grid is not None:
        coll = coll | (grid(p.reshape(-1, 2)).view(p.shape[:-1]) < radius)
    ws_min = torch.tensor(ws_limits[0]) * 1.08
    ws_max = torch.tensor(ws_limits[1]) * 1.08
    sd = torch.cat((p - ws_min, ws_max - p), dim=-1)
    coll = coll | (sd < radius).any(-1)
    return coll


def get_trajs_free_idxs(trajs, grid, q_min=(-1.0, -1.0), q_max=(1.0, 1.0), **kw):
    """Index set of free trajectories as get_trajs_collision_and_free computes it (tasks.py:236-311): no collision on
    any interpolated waypoint AND all support positions inside the joint limits."""
    coll = trajs_collision_mask(trajs, grid, **kw)
    free = torch.logical_not(coll).all(dim=-1)
    pos = trajs[..., :2]
    inside = torch.logical_and(pos >= torch.tensor(q_min), pos <= torch.tensor(q_max)).all(-1).all(-1)
    return torch.argwhere(free & inside).reshape(-1), coll


def compute_path_length(trajs):  # TR/trajectory/metrics.py:7-16
    return torch.linalg.norm(torch.diff(trajs[..., :2], dim=-2), dim=-1).sum(-1)


def compute_smoothness(trajs):  # TR/trajectory/metrics.py:31-39
    return torch.linalg.norm(torch.diff(trajs[..., 2:4], dim=-2), dim=-1).sum(-1)


def select_best(trajs_unnormalized, grid, **kw):  # mpd.py:356-382
    free_idxs, _ = get_trajs_free_idxs(trajs_unnormalized, grid, **kw)
    if free_idxs.numel() == 0:
        return None, free_idxs
    tf = trajs_unnormalized[free_idxs]
    cost = compute_path_length(tf) + compute_smoothness(tf)
    return int(free_idxs[torch.argmin(cost)]), free_idxs


# ----------------------------------------------------------------------------------------------
# synthetic start / goal generators (mmd/common/multi_agent_utils.py:146-179, device-free restatement)
# ----------------------------------------------------------------------------------------------
def get_start_goal_pos_circle(num_agents, radius=0.8):  # :146-155
    start = [torch.tensor([radius * np.cos(2 * torch.pi * i / num_agents), radius * np.sin(2 * torch.pi * i / num_agents)],
                          dtype=torch.float32) for i in range(num_agents)]
    goal = [torch.tensor([radius * np.cos(2 * torch.pi * i / num_agents + torch.pi),
                          radius * np.sin(2 * torch.pi * i / num_agents + torch.pi)], dtype=torch.float32)
            for i in range(num_agents)]
    return start, goal


def hard_conds_from_start_goal(start_pos, goal_pos, normalizer: LimitsNormalizer, horizon=HORIZON):
    """TrajectoryDataset.get_hard_conditions(normalize=True) (mmd/datasets/trajectories.py:216-239)."""
    s = normalizer.normalize(torch.cat((start_pos, torch.zeros_like(start_pos))))
    g = normalizer.normalize(torch.cat((goal_pos, torch.zeros_like(goal_pos))))
    return {0: s, horizon - 1: g}


# ---------------------------------------------------------------------------------------------------------------------
# conflict detection (mmd/planners/multi_agent/cbs.py:166-246) and post-smoothing (mmd/common/trajectory_utils.py:31-69)
# ---------------------------------------------------------------------------------------------------------------------
def global_pad_paths(path_l, start_time_l):  # mmd/common/multi_agent_utils.py:120-143
    if len(path_l) == 0:
        return []
    max_t = max(len(p) + start_time_l[i] for i, p in enumerate(path_l))
    out = []
    for i, p in enumerate(path_l):
        if len(p) + start_time_l[i] < max_t:
            p = torch.cat([p, p[-1].repeat(max_t - len(p) - start_time_l[i], 1)])
        if start_time_l[i] > 0:
            p = torch.cat([p[0].repeat(start_time_l[i], 1), p])
        out.append(p)
    return out


def densify_trajs(trajs, n_points_interp=10):  # trajectory_utils.py:54-69 (Python loops, kept verbatim in structure)
    out = []
    for traj in trajs:
        d = []
        for i in range(traj.shape[0] - 1):
            d.append(traj[i])
            for j in range(1, n_points_interp):
                d.append(traj[i] + j * (traj[i + 1] - traj[i]) / n_points_interp)
        d.append(traj[-1])
        out.append(torch.stack(d))
    return out


def get_conflicts(best_path_l, start_time_l, want_vertex=False, want_edge=False, want_point=True, radius=ROBOT_RADIUS):
    """cbs.py:166-246 -> list of tuples in the reference's order:
    ("vertex", a, b, t) / ("edge", a, b, t_from, t_to) / ("point", a, b, t_from, t_to, p_a, p_b, midpoint)."""
    import math
    best_path_l = global_pad_paths(list(best_path_l), start_time_l)
    paths_pos_l = [p[..., :2] for p in best_path_l]
    if len(paths_pos_l) == 0:
        return []
    factor = 2 if want_edge else 1
    dense_l = densify_trajs(paths_pos_l, factor)
    b = torch.stack(dense_l).permute(1, 0, 2)
    coll, pts = check_rr_collisions(b, radius)
    out = []
    for t_dense, a, bb in torch.nonzero(coll.int()).tolist():
        t_from, t_to = math.floor(t_dense / factor), math.ceil(t_dense / factor)
        if want_vertex and t_from == t_to:
            out.append(("vertex", a, bb, t_from))
        if want_edge and t_from != t_to:
            out.append(("edge", a, bb, t_from, t_to))
        if want_point:
            out.append(("point", a, bb, t_from, t_to, dense_l[a][t_dense], dense_l[bb][t_dense], pts[t_dense, a, bb]))
    return out


def create_soft_constraints_from_other_agents_paths(best_path_l, start_time_l, agent_id, radius=2.4 * ROBOT_RADIUS):
    """cbs.py:468-508, loops kept as in the reference.  Returns (q [n, 2], t_range [n, 2], radii [n]) or None."""
    if len(best_path_l) == 0:
        return None
    q_l, t_range_l, radius_l = [], [], []
    for other in range(len(best_path_l)):
        if other != agent_id:
            path = best_path_l[other]
            for t_other in range(0, len(path), 1):
                t_agent = t_other + start_time_l[other] - start_time_l[agent_id]
                if agent_id >= len(best_path_l):
                    T_agent = len(path) - 1
                else:
                    T_agent = len(best_path_l[agent_id]) - 1
                if 1 <= t_agent <= T_agent:
                    q_l.append(path[t_other][:2])
                    t_range_l.append((t_agent, t_agent + 1))
                    radius_l.append(radius)
    if not q_l:
        return None
    return torch.stack(q_l), torch.tensor(t_range_l, dtype=torch.float32), torch.tensor(radius_l)


def smooth_trajs(trajs, window_size=10, poly_order=2):  # trajectory_utils.py:31-38 (scipy on the host, as the reference)
    from scipy.signal import savgol_filter
    return torch.tensor(savgol_filter(trajs.cpu().numpy(), window_size, poly_order, axis=1))
