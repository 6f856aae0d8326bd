"""Timing of one guided ddpm_step_kernel launch vs fleet size (development tool): 32 groups x 128 samples on this GPU, peer
table of n_peers rows, brute-force scan vs spatial hash."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import mmd_b200 as M  # noqa: E402
from mmd_b200 import _lib  # noqa: E402
from mmd_b200.diffusion import lower_for_step  # noqa: E402
from mmd_b200.guides import PeerHash  # noqa: E402
from tests.helpers import build_product  # noqa: E402

dev = torch.device("cuda:0")
p = build_product(dev, "EnvHighways2D", T=100)
R, K, H = 32, 128, 64
g = torch.Generator().manual_seed(0)
x0 = (torch.randn(R * K, H, 4, generator=g) * 0.4).clamp(-0.95, 0.95).to(dev)
eps = torch.randn(R * K, H, 4, generator=g).to(dev) * 0.1
noise = torch.randn(R * K, H, 4, generator=g).to(dev)
sc = p["model"].step_scalars(10, 20, 0.5, True)
lib = _lib.lib()
for n_peers in (32, 64, 128, 256):
    peers = (torch.rand(n_peers, H, 2, generator=g) * 1.6 - 0.8).to(dev).contiguous()
    peer_self = torch.arange(R, dtype=torch.int32, device=dev)
    for name in ("brute", "hash"):
        ph = PeerHash(n_peers, H, 0.12, dev) if name == "hash" else None
        env, grp, keep = lower_for_step(p["guide"], R, K, H, dev, [None] * R, None, peers, peer_self, 0.12, 2e-2, ph)
        x = x0.clone()
        def run():
            if ph is not None:
                ph.build(peers)
            _lib.check(lib.mmdk_ddpm_step(C.byref(env), C.byref(grp), C.byref(sc), H, _lib.ptr(x), _lib.ptr(eps), _lib.ptr(noise), None,
                                          _lib.stream_ptr()))
        for _ in range(3):
            x.copy_(x0); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            x.copy_(x0)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"n_peers={n_peers:4d} {name:5s}: guided step (+hash build) median {ts[len(ts) // 2] * 1e3:.0f} us", flush=True)
