import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from oracle import port
from tests.helpers import build_product
from mmd_b200 import _lib
dev = torch.device('cuda:0')
P = port.make_unet_params(seed=0)
p = build_product(dev, "EnvEmpty2D", T=25, P=P)
B = 4096
x = torch.randn(B, 64, 4, device=dev)
u = p["unet"]
for _ in range(3): u.forward_t(x, 7, precision="f16x3")
h = u.native(); lib = _lib.lib()
n_tiles = (B + 6) // 7
names = {0: 'd0.0.b0 (4->32 @64)', 3: 'd0.1.b1 (32->32 @64)', 12: 'd2.1.b0 (128->128 @16)', 15: 'mid1.b1 (128 @16, res id)', 18: 'u0.0.b0 (256->64 @16)', 25: 'u1.1.b0 (32->32 @32)', 29: 'final 1x1'}
for op in [3, 12, 15, 18, 25, 29]:
    dbg = torch.zeros(n_tiles, 16, dtype=torch.int64, device=dev)
    _lib.check(lib.mmdk_unet_debug_timeline(h, op, _lib.ptr(dbg), _lib.stream_ptr()))
    u.forward_t(x, 7, precision="f16x3"); torch.cuda.synchronize()
    d = dbg.cpu()
    t0 = d[:, 0]
    rel = (d[:, 1:12] - t0[:, None]).float()
    # choose tiles that were in the first wave on their SM: smallest t0 per smid
    print(f"op {op} {names.get(op)}: median cycles since CTA start [prod:depwait_done, prod:done, mma:in_full, mma:last_commit, epi:before_acc_wait, epi:acc_ready, epi:pass1_done, epi:bar1, epi:stats_done, epi:end, all:after_sync]")
    print('   ', [int(v) for v in rel.median(0).values.tolist()])
    life = (d[:, 11] - d[:, 0]).float()
    print('    CTA lifetime median %.0f p90 %.0f cycles' % (life.median(), life.quantile(0.9)))
    # per-SM span: first start to last end
    sm = d[:, 15]
    spans = []
    for s_ in sm.unique()[:20]:
        m = sm == s_
        spans.append(int(d[m, 11].max() - d[m, 0].min()))
    print('    per-SM span (first 20 SMs) median', sorted(spans)[len(spans)//2], 'tiles/SM', float(m.sum()))
_lib.check(lib.mmdk_unet_debug_timeline(h, -1, None, _lib.stream_ptr()))
