"""TemporalUnet with the reference's constructor signature and state_dict key layout, executed by libmmdk.

Mirrors mmd/models/diffusion_models/temporal_unet.py:23-174 and mmd/models/layers/layers.py:177-398 as a PARAMETER
CONTAINER only: the nn.Module tree exists so that `load_state_dict` accepts the reference's checkpoints
(`downs.0.0.blocks.0.block.0.weight`, ...); `forward` never runs a torch op on the activations -- it hands the
weights to mmdk_unet_create once and calls mmdk_unet_forward (hand-written sm_100a kernels).
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib

UNET_DIM_MULTS = {0: (1, 2, 4), 1: (1, 2, 4, 8)}  # temporal_unet.py:17-20


def group_norm_n_groups(n_channels, target_n_groups=8):  # layers.py:392-398
    if n_channels < target_n_groups:
        return 1
    for n_groups in range(target_n_groups, target_n_groups + 10):
        if n_channels % n_groups == 0:
            return n_groups
    return 1


class _Conv1dBlock(nn.Module):  # layers.py:279-296 -> keys block.0.* (Conv1d) and block.2.* (GroupNorm)
    def __init__(self, ci, co, k=5):
        super().__init__()
        self.block = nn.Sequential(nn.Conv1d(ci, co, k, padding=k // 2), nn.Identity(),
                                   nn.GroupNorm(group_norm_n_groups(co), co), nn.Identity(), nn.Identity())


class _ResidualTemporalBlock(nn.Module):  # layers.py:326-358
    def __init__(self, ci, co, cond_dim):
        super().__init__()
        self.blocks = nn.ModuleList([_Conv1dBlock(ci, co), _Conv1dBlock(co, co)])
        self.cond_mlp = nn.Sequential(nn.Identity(), nn.Linear(cond_dim, co), nn.Identity())
        self.residual_conv = nn.Conv1d(ci, co, 1) if ci != co else nn.Identity()


class _Resample(nn.Module):  # layers.py:261-276 -> key conv.*
    def __init__(self, dim, up):
        super().__init__()
        self.conv = nn.ConvTranspose1d(dim, dim, 4, 2, 1) if up else nn.Conv1d(dim, dim, 3, 2, 1)


class _LayerNorm(nn.Module):  # layers.py:197-207
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1))


class _LinearAttention(nn.Module):  # layers.py:210-229
    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.to_qkv = nn.Conv1d(dim, heads * dim_head * 3, 1, bias=False)
        self.to_out = nn.Conv1d(heads * dim_head, dim, 1)


class _PreNorm(nn.Module):  # layers.py:186-194
    def __init__(self, dim):
        super().__init__()
        self.fn = _LinearAttention(dim)
        self.norm = _LayerNorm(dim)


class _Residual(nn.Module):  # layers.py:177-183
    def __init__(self, dim):
        super().__init__()
        self.fn = _PreNorm(dim)


class _TimeEncoder(nn.Module):  # layers.py:232-243
    def __init__(self, dim, dim_out):
        super().__init__()
        self.encoder = nn.Sequential(nn.Identity(), nn.Linear(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim_out))


def sinusoidal_table(n_steps, dim=32):
    """SinusoidalPosEmb (layers.py:246-258) for t = 0..n_steps-1, evaluated with the reference's own expression so
    the device never re-derives sin/cos (copy tables from torch to avoid libm drift)."""
    x = torch.arange(n_steps)
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half) * -emb)
    emb = x[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1).to(torch.float32).contiguous()


class TemporalUnet(nn.Module):
    def __init__(self, n_support_points=None, state_dim=None, unet_input_dim=32, dim_mults=(1, 2, 4, 8),
                 time_emb_dim=32, self_attention=False, conditioning_embed_dim=4, conditioning_type=None,
                 attention_num_heads=2, attention_dim_head=32, unet_precision="auto", **kwargs):
        super().__init__()
        if conditioning_type not in (None, "None"):
            # temporal_unet.py:44-58: 'concatenate'/'attention'/'default' need context models no planner builds
            raise NotImplementedError("only conditioning_type=None is on the sampling path (mpd.py:154-159,210)")
        self.state_dim = state_dim
        self.n_support_points = n_support_points
        self.unet_input_dim = unet_input_dim
        self.dim_mults = tuple(dim_mults)
        self.time_emb_dim = time_emb_dim
        self.self_attention = bool(self_attention)
        self.conditioning_type = None
        self.unet_precision = unet_precision

        dims = [state_dim, *[unet_input_dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.time_mlp = _TimeEncoder(32, time_emb_dim)
        cond_dim = time_emb_dim
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n_res = len(in_out)
        for ind, (di, do) in enumerate(in_out):
            is_last = ind >= (n_res - 1)
            self.downs.append(nn.ModuleList([
                _ResidualTemporalBlock(di, do, cond_dim), _ResidualTemporalBlock(do, do, cond_dim),
                _Residual(do) if self_attention else nn.Identity(), None,
                _Resample(do, up=False) if not is_last else nn.Identity()]))
        mid = dims[-1]
        self.mid_block1 = _ResidualTemporalBlock(mid, mid, cond_dim)
        self.mid_attn = _Residual(mid) if self_attention else nn.Identity()
        self.mid_attention = nn.Identity()
        self.mid_block2 = _ResidualTemporalBlock(mid, mid, cond_dim)
        for ind, (di, do) in enumerate(reversed(in_out[1:])):
            self.ups.append(nn.ModuleList([
                _ResidualTemporalBlock(do * 2, di, cond_dim), _ResidualTemporalBlock(di, di, cond_dim),
                _Residual(di) if self_attention else nn.Identity(), None, _Resample(di, up=True)]))
        self.final_conv = nn.Sequential(_Conv1dBlock(unet_input_dim, unet_input_dim), nn.Conv1d(unet_input_dim, state_dim, 1))

        self._handle = None
        self._handle_key = None
        self._keepalive = None
        self._plist = None
        self._time_table_steps = 1024
        # a load_state_dict through ANY ancestor module reaches this hook (nn.Module.load_state_dict runs the post hooks of
        # every module it visits), including assign=True loads that replace the Parameter objects
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._invalidate())

    def tensor_core_supported(self):
        """Whether the tcgen05 executor covers this network shape (unet_fused.cu build_fused): every level's (rows of a
        7-sample tile) x channels fits the epilogue's register tiling.  LinearAttention blocks are covered (QKV / out
        projections as tensor-core GEMMs)."""
        if self.state_dim > 8:
            return False
        L = self.n_support_points
        for i, m in enumerate(self.dim_mults):
            C = self.unet_input_dim * m
            n_mt = (7 * ((L >> i) + 2) + 127) // 128
            if C not in (32, 64, 128) or n_mt * C not in (64, 128):
                return False
        # the output-side ops of build_tc: up-path blocks run with the channel count of the level above, the final
        # conv block (unet_input_dim channels) and the final 1x1 conv (N = 16) run at full resolution
        for i in range(1, len(self.dim_mults)):
            C_up = self.unet_input_dim * self.dim_mults[i - 1]
            n_mt = (7 * ((L >> i) + 2) + 127) // 128
            if n_mt * C_up not in (64, 128):
                return False
        n_mt0 = (7 * (L + 2) + 127) // 128
        if n_mt0 * 16 != 64 or n_mt0 * self.unet_input_dim not in (64, 128):
            return False
        return True

    def _layer_program(self):
        """(kind, [main source channel counts], [1x1 residual-conv source channel counts], cout, L) of every conv op in
        execution order (temporal_unet.py:60-119,121-174; layers.py:326-358): the shapes the native builders see.  The
        SECOND conv block of a ResidualTemporalBlock also evaluates the block's residual 1x1 conv (when cin != cout)."""
        n = len(self.dim_mults)
        dims = [self.state_dim] + [self.unet_input_dim * m for m in self.dim_mults]
        L = [self.n_support_points >> i for i in range(n)]
        ops = []

        def rtb(srcs, cout, Li):
            ops.append(("block", list(srcs), [], cout, Li))
            ops.append(("block", [cout], list(srcs) if sum(srcs) != cout else [], cout, Li))

        for i in range(n):
            rtb([dims[i]], dims[i + 1], L[i])
            rtb([dims[i + 1]], dims[i + 1], L[i])
            if i < n - 1:
                ops.append(("down", [dims[i + 1]], [], dims[i + 1], L[i]))
        rtb([dims[n]], dims[n], L[n - 1])
        rtb([dims[n]], dims[n], L[n - 1])
        for j in range(n - 1, 0, -1):
            rtb([dims[j + 1], dims[j + 1]], dims[j], L[j])
            rtb([dims[j]], dims[j], L[j])
            ops.append(("up", [dims[j]], [], dims[j], L[j]))
        ops.append(("block", [dims[1]], [], dims[1], L[0]))
        ops.append(("final", [dims[1]], [], self.state_dim, L[0]))
        return ops

    def tensor_core_layers_supported(self):
        """Whether the per-layer tcgen05 executor (unet_tc.cu build_tc) covers this network shape.  Wider than the persistent
        executor: 256-channel levels (dim_mults (1,2,4,8), the default of scripts/train_diffusion/train.py:35) run as two
        128-channel slices, and ops whose source images do not fit shared memory together (the 512-channel concat blocks at
        L = 8) stream them through one buffer, one image per input phase."""
        if self.state_dim > 8 or self.self_attention:
            return False
        for kind, srcs, res_srcs, cout, Li in self._layer_program():
            rows = 2 + 128 * ((7 * (Li + 2) + 127) // 128) + 2
            n_mt = (rows - 4) // 128
            N = 16 if kind == "final" else cout
            split = 1
            if kind == "block" and cout == 256:
                N, split = 128, 2
            if N not in (16, 32, 64, 128):
                return False
            NV = n_mt * N
            if not ((NV == 128 and N in (32, 64, 128)) or (NV == 64 and N in (64, 32, 16))):
                return False
            if split > 1 and NV != 128:
                return False
            if kind == "block" and (cout < 8 or cout % 8 != 0):
                return False
            src_bytes = [((16 if c == self.state_dim else c) * rows * 4 + 127) // 128 * 128 for c in list(srcs) + list(res_srcs)]
            scratch = (5 * N + n_mt * 128 * 8 * 2 + 2 * 7 * 8) * 4
            fixed = 3 * 32 * N * 4 + (scratch + 15) // 16 * 16 + 80
            if sum(src_bytes) + fixed > 232448:     # streamed: one source image per input phase
                if kind != "block" or max(src_bytes) + fixed > 232448 or len(src_bytes) > 4:
                    return False
        return True

    def resolve_precision(self, precision=None):
        """'auto' (default): a tensor-core executor ("f16x3": FP16 hi/lo split, ~3e-6 relative on eps) when one covers the shape,
        else the exact fp32 CUDA-core executor.  All are native sm_100a kernels; native_mode() picks the executor."""
        p = precision or self.unet_precision
        if p == "auto" and getattr(self, "_tc_rejected", 0) >= 2:
            return "fp32"   # both native tensor-core builders refused this shape (ValueError): exact executor from then on
        if p == "auto":
            p = "f16x3" if (self.tensor_core_supported() or self.tensor_core_layers_supported()) else "fp32"
        return p

    def _note_rejection(self, mode):
        """A native tensor-core builder refused the shape: 1 = persistent executor out, >= 2 = per-layer executor out too."""
        r = getattr(self, "_tc_rejected", 0)
        if mode == _lib.UNET_F16X3_LAYERS or not self.tensor_core_layers_supported():
            r |= 2
        if mode == _lib.UNET_F16X3:
            r |= 1
        self._tc_rejected = r

    def native_mode(self, precision=None):
        """mmdk_unet_forward mode for a precision: "f16x3" runs on the persistent whole-forward executor when it covers the
        network shape, else on the per-layer tcgen05 executor (same numerics)."""
        p = self.resolve_precision(precision)
        if p in ("f16x3", "tc"):
            rejected = getattr(self, "_tc_rejected", 0)
            if (not self.tensor_core_supported() or (rejected & 1)) and self.tensor_core_layers_supported() and not (rejected & 2):
                return _lib.UNET_F16X3_LAYERS
            return _lib.UNET_F16X3
        return _lib.UNET_MODES[p]

    # -- native handle ------------------------------------------------------------------------------------------
    def _invalidate(self):
        self._release()

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().mmdk_unet_destroy(self._handle)
        self._handle = None
        self._handle_key = None
        self._keepalive = None
        self._plist = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out

    def ensure_time_table(self, n_steps):
        if n_steps > self._time_table_steps:
            self._time_table_steps = int(n_steps)
            self._invalidate()

    def native(self):
        """mmdk_unet handle for the current weights/device (built lazily, rebuilt after load_state_dict / .to())."""
        lib = _lib.lib()
        # the Parameter objects are cached with the handle (walking nn.Module.parameters() costs ~100 us per call, more than
        # a small-batch forward); _apply / load_state_dict drop the cache, in-place updates are caught by the versions below
        if getattr(self, "_plist", None) is None:
            self._plist = list(self.parameters())
        dev = self._plist[0].device
        if dev.type != "cuda":
            raise _lib.MMDKError("TemporalUnet parameters must live on a CUDA device (no CPU path)")
        # keyed on every parameter's version counter too: load_state_dict through a PARENT module
        # (GaussianDiffusionModel.load_state_dict recurses via _load_from_state_dict and never calls this class's
        # override) and in-place updates (p.data.copy_, optimiser steps) bump `_version`, so stale packed weights on
        # the device can never be used silently
        key = (dev, self._time_table_steps, tuple([(p.data_ptr(), p._version) for p in self._plist]))
        if self._handle is not None and self._handle_key == key:
            return self._handle
        self._release()
        sd = {k: v.detach().to(torch.float32).contiguous() for k, v in self.state_dict().items()}
        sd["__sincos_table__"] = sinusoidal_table(self._time_table_steps).to(dev)
        names = list(sd.keys())
        cfg = _lib.UnetConfig()
        cfg.state_dim, cfg.horizon, cfg.unet_input_dim = self.state_dim, self.n_support_points, self.unet_input_dim
        cfg.n_levels = len(self.dim_mults)
        for i, m in enumerate(self.dim_mults):
            cfg.dim_mults[i] = m
        cfg.time_emb_dim = self.time_emb_dim
        cfg.n_diffusion_steps = self._time_table_steps
        cfg.self_attention = int(self.self_attention)
        c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        c_ptrs = (C.c_void_p * len(names))(*[sd[n].data_ptr() for n in names])
        c_numel = (C.c_int64 * len(names))(*[sd[n].numel() for n in names])
        out = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.mmdk_unet_create(C.byref(cfg), len(names), c_names, c_ptrs, c_numel, _lib.stream_ptr(), C.byref(out)))
        self._handle, self._handle_key, self._keepalive = out, key, sd
        self._plist = list(self.parameters())
        return out

    # -- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, time, context=None, precision=None):
        """x: [B, H, D] cuda fp32; time: [B] long with one shared value (make_timesteps, diffusion_model_base.py:27)."""
        if context is not None:
            raise NotImplementedError("context conditioning is not on the sampling path (mpd.py:210 context=None)")
        t = int(time[0]) if torch.is_tensor(time) else int(time)
        return self.forward_t(x, t, precision)

    @torch.no_grad()
    def forward_t(self, x, t, precision=None, out=None):
        lib = _lib.lib()
        x = x.contiguous()
        assert x.dtype == torch.float32 and x.dim() == 3
        assert x.shape[1] == self.n_support_points and x.shape[2] == self.state_dim
        self.ensure_time_table(t + 1)
        h = self.native()
        if out is None:
            out = torch.empty_like(x)
        while True:
            mode = self.native_mode(precision)
            try:
                _lib.check(lib.mmdk_unet_forward(h, mode, _lib.ptr(x), x.shape[0], t, _lib.ptr(out), _lib.stream_ptr()))
                break
            except ValueError:
                # 'auto' may only ever fall back to another NATIVE executor (persistent tcgen05 -> per-layer tcgen05 -> exact
                # fp32; never a CPU / torch path), and only when a builder rejects the network shape; an explicit precision
                # request fails loudly
                if (precision or self.unet_precision) != "auto" or mode == _lib.UNET_FP32:
                    raise
                self._note_rejection(mode)
        return out
