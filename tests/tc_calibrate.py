"""Calibration of the tensor-pipe counter (run under ncu, see profiles/README.md): pure tcgen05.mma loops."""
import sys

import torch

sys.path.insert(0, ".")
from mmd_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda:0")
n_ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 148
for n_acc in (1, 2, 4):
    for N in (16, 32, 64, 128, 256):
        if n_acc > 2 and N > 128:
            continue
        out = torch.zeros(n_ctas, 2, dtype=torch.int64, device=dev)
        for _ in range(2):
            _lib.check(lib.mmdk_debug_mma_calibrate(N, 4096, n_ctas, n_acc, _lib.ptr(out), _lib.stream_ptr()))
        torch.cuda.synchronize()
        per = out[:, 0].double() / 4096
        print(f"N={N:3d} accumulators={n_acc}: cycles per MMA (M=128,K=16) median {float(per.median()):.1f} min {float(per.min()):.1f} "
              f"max {float(per.max()):.1f} -> {2 * 128 * N * 16 / float(per.median()):.0f} FLOP/cycle/SM (math model: {N / 2:.0f} cycles)")
