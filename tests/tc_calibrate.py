"""Calibration of the tensor-pipe counter (run under ncu, see profiles/README.md): pure tcgen05.mma loops."""
import sys

import torch

sys.path.insert(0, ".")
from mmd_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda:0")
n_ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 148
for N in (32, 64, 128, 256):
    out = torch.zeros(n_ctas, 2, dtype=torch.int64, device=dev)
    for _ in range(2):
        _lib.check(lib.mmdk_debug_mma_calibrate(N, 4096, n_ctas, _lib.ptr(out), _lib.stream_ptr()))
    torch.cuda.synchronize()
    cyc = out[:, 0].double()
    per = cyc / 4096
    print(f"N={N}: cycles per MMA (M=128,K=16) median {float(per.median()):.1f} min {float(per.min()):.1f} max {float(per.max()):.1f}"
          f"  -> {2 * 128 * N * 16 / float(per.median()):.0f} FLOP/cycle/SM (issue model: {N / 2:.0f} cycles)")
