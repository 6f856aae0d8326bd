"""Per-item timeline of CTA 0 of the persistent UNet executor (development tool; see mmdk_unet_debug_timeline)."""
import ctypes as C
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
for _ in range(3):
    unet.forward_t(x, 5, precision="f16x3")
n_items = 33 * 4
dbg = torch.zeros(n_items, 16, dtype=torch.int64, device=dev)
lib = _lib.lib()
_lib.check(lib.mmdk_unet_debug_timeline(unet.native(), -2, _lib.ptr(dbg), _lib.stream_ptr()))
unet.forward_t(x, 5, precision="f16x3")
torch.cuda.synchronize()
_lib.check(lib.mmdk_unet_debug_timeline(unet.native(), -2, None, _lib.stream_ptr()))
d = dbg.cpu()
t0 = int(d[0, 0])
print("item op par | load_issue  in_ready(+load)  mma_issued(+mma)  [w_stall]  acc_ready  epi_start..epi_done(+epi) | period")
prev_done = None
for k in range(n_items):
    if int(d[k, 0]) == 0:
        break
    j, par = (k // 2) % 33, k % 2
    li, ir, mi, ar, ed, ws, es = [int(d[k, i]) for i in (0, 1, 2, 3, 5, 6, 7)]
    e = [int(d[k, i]) for i in (8, 9, 10, 11, 12, 13)]
    st = max(es, ar)
    phases = (f" epi: tmem {e[0] - st:5d} stats {e[1] - e[0]:5d} bar {e[2] - e[1]:5d} reduce+bar {e[3] - e[2]:5d} norm {e[4] - e[3]:5d} "
              f"fence {e[5] - e[4]:5d}") if e[0] else ""
    print(f"{k:4d} {j:2d} {par} | {li - t0:8d} {ir - t0:8d} (+{ir - li:5d}) {mi - t0:8d} (+{mi - ir:5d}) [{ws:5d}] {ar - t0:8d} "
          f"{st - t0:8d}..{ed - t0:8d} (+{ed - st:5d}) | {'' if prev_done is None else ed - prev_done}{phases}")
    prev_done = ed
print("total cycles CTA 0:", int(d[:k, 5].max()) - t0)
