"""UNet forward timing per executor (not a bench line: development tool).  usage: tc_bench.py [B] [modes...]
MMD_DIM_MULTS=1,2,4,8 selects the 4-level network (per-layer tcgen05 executor vs the exact fp32 executor)."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import mmd_b200 as M  # noqa: E402
from oracle import port  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
modes = sys.argv[2:] or ["f16x3_layers", "f16x3"]
dev = torch.device("cuda:0")
DM = tuple(int(v) for v in os.environ.get("MMD_DIM_MULTS", "1,2,4").split(","))
P = port.make_unet_params(seed=0, dim_mults=DM)
unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=DM)
unet.load_state_dict(P, strict=True)
unet = unet.to(dev)
x = torch.randn(B, 64, 4, device=dev)
ref = None
for mode in modes:
    out = torch.empty_like(x)
    for _ in range(3):
        unet.forward_t(x, 5, precision=mode, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        unet.forward_t(x, 5, precision=mode, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fin = bool(torch.isfinite(out).all())
    if ref is None:
        ref = out.clone()
    err = float((out - ref).norm() / ref.norm())
    print(f"dim_mults={DM} {mode:14s} B={B}: {ms:.3f} ms/forward  ({B * 36495360 / ms / 1e9:.1f} TFLOP/s algorithmic)  finite={fin}  rel diff vs first mode {err:.2e}")
