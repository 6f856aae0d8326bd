"""Launch anatomy of the persistent UNet executor (development tool): %globaltimer stamps of every CTA vs CUDA-event time."""
import sys

import torch

sys.path.insert(0, ".")
import mmd_b200 as M  # noqa: E402
from mmd_b200 import _lib  # noqa: E402
from oracle import port  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
unet = unet.to(dev)
x = torch.randn(B, 64, 4, device=dev)
out = torch.empty_like(x)
for _ in range(3):
    unet.forward_t(x, 5, precision="f16x3", out=out)
n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
N = 20
stamps = torch.zeros(N, n_sm, 2, dtype=torch.int64, device=dev)
_lib.check(_lib.lib().mmdk_unet_debug_stamps(unet.native(), _lib.ptr(stamps), N, n_sm))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    unet.forward_t(x, 5, precision="f16x3", out=out)
e1.record()
torch.cuda.synchronize()
st = stamps.cpu()
import os
spt = int(os.environ.get('MMDK_FUSED_ST', 0)) or (1 if B <= n_sm else (3 if B <= 3 * n_sm else 7))
grid = min((B + spt - 1) // spt, n_sm)
st = st[:, :grid]
dur = (st[:, :, 1].max(1).values - st[:, :, 0].min(1).values).double() / 1e3
skew = (st[:, :, 0].max(1).values - st[:, :, 0].min(1).values).double() / 1e3
per_cta = (st[:, :, 1] - st[:, :, 0]).double() / 1e3
gap = (st[1:, :, 0].min(1).values - st[:-1, :, 1].max(1).values).double() / 1e3
print(f"B={B} grid={grid}: event time per forward {e0.elapsed_time(e1) / N * 1e3:.1f} us; kernel duration (stamps) median {float(dur.median()):.1f} us; "
      f"CTA start skew median {float(skew.median()):.1f} us; per-CTA lifetime median {float(per_cta.median()):.1f} / max {float(per_cta.max()):.1f} us; "
      f"gap between consecutive forwards (end -> next start) median {float(gap.median()):.1f} us")
