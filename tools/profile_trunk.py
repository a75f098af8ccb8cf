"""A few launches of the dominant kernel alone for `ncu --set full`: the generator trunk on the metric-config clip batch
[B,32,32,64] (bench.py `roofline`).  Default: the one-launch row-linearised kernel (`-k regex:conv3x3_lin`, 33 layers);
TP_PER_LAYER=1: the per-layer persistent kernel (`-k regex:conv3x3_tc`).  TP_N / TP_H / TP_L select the shape."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import kernels as K  # noqa: E402
n, h, L = int(os.environ.get("TP_N", 296)), int(os.environ.get("TP_H", 32)), int(os.environ.get("TP_L", 33))
x = (torch.randn(n, h, 32, 64, device="cuda") * 0.5).to(torch.bfloat16)
a, b = torch.zeros_like(x), torch.zeros_like(x)
if os.environ.get("TP_PER_LAYER"):
    w1 = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
    w2 = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
    bz = torch.zeros(64, device="cuda")
    for _ in range(4):
        K.conv3x3_tc(x, w1, bz, a, cout=64, act=1)
        K.conv3x3_tc(a, w2, bz, x, cout=64, act=0, res=x)
else:
    ws = torch.cat([K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64) for _ in range(L)]).contiguous()
    bs = torch.zeros(L * 64, device="cuda")
    plan = ([(0, 1, -1, 1)] + [(1, 2, -1, 1), (2, 1, 1, 0)] * 20)[:L]
    for _ in range(3):
        K.conv3x3_lin_chain(x, a, b, ws, bs, plan)
torch.cuda.synchronize()
