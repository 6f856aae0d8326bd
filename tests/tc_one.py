"""Few forwards of the persistent UNet executor (ncu target)."""
import sys
import torch
sys.path.insert(0, ".")
import mmd_b200 as M
from oracle import port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
dev = torch.device("cuda:0")
unet = M.TemporalUnet(n_support_points=64, state_dim=4, unet_input_dim=32, dim_mults=(1, 2, 4))
unet.load_state_dict(port.make_unet_params(seed=0), strict=True)
unet = unet.to(dev)
x = torch.randn(B, 64, 4, device=dev)
for _ in range(4):
    unet.forward_t(x, 5, precision=mode)
torch.cuda.synchronize()
