import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from oracle import port
from tests.helpers import build_oracle, build_product, rel_err
from mmd_b200 import _lib
dev = torch.device('cuda:0')
o = build_oracle("EnvEmpty2D", T=25); p = build_product(dev, "EnvEmpty2D", T=25, P=o["P"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 9
x = torch.randn(B, 64, 4, generator=torch.Generator().manual_seed(5))
taps = {}
ref = port.unet_forward(o["P"], x, torch.full((B,), 7, dtype=torch.long), taps=taps)
out = p["unet"].forward_t(x.to(dev), 7, precision="f16x3")
torch.cuda.synchronize()
h = p["unet"].native()
c, l = C.c_int(), C.c_int()
_lib.check(_lib.lib().mmdk_unet_debug_tap(h, -1, None, C.byref(c), C.byref(l), None, _lib.stream_ptr()))
buf = torch.empty(B, c.value, l.value, device=dev)
_lib.check(_lib.lib().mmdk_unet_debug_tap(h, -1, _lib.ptr(buf), C.byref(c), C.byref(l), None, _lib.stream_ptr()))
print('input image rel', rel_err(buf[:, :4], x.transpose(1, 2)), 'pad max', float(buf[:, 4:].abs().max()))
for j, (name, act) in enumerate(taps["ops"]):
    _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, None, C.byref(c), C.byref(l), None, _lib.stream_ptr()))
    buf = torch.empty(B, c.value, l.value, device=dev)
    _lib.check(_lib.lib().mmdk_unet_debug_tap(h, j, _lib.ptr(buf), C.byref(c), C.byref(l), None, _lib.stream_ptr()))
    print(j, name, tuple(act.shape), 'rel_err %.3e' % rel_err(buf, act), 'nan' if not torch.isfinite(buf).all() else '')
print('eps rel_err %.3e' % rel_err(out, ref))
