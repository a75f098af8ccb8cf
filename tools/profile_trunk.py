"""A few launches of the dominant kernel alone -- the 3x3 64->64 res-block layer on the metric-config clip batch
[B,32,32,64] (bench.py `roofline`) -- for `ncu --set full -k regex:conv3x3_tc`.  TP_N / TP_H select the shape."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import kernels as K  # noqa: E402
n, h = int(os.environ.get("TP_N", 296)), int(os.environ.get("TP_H", 32))
x = (torch.randn(n, h, h, 64, device="cuda") * 0.5).to(torch.bfloat16)
y = torch.empty_like(x)
w1 = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
w2 = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
b = torch.zeros(64, device="cuda")
for _ in range(4):
    K.conv3x3_tc(x, w1, b, y, cout=64, act=1)
    K.conv3x3_tc(y, w2, b, x, cout=64, act=0, res=x)
torch.cuda.synchronize()
