"""A few launches of the fused warp + space-to-depth kernel on the bench's HBM-sized batch (32 x 1024x1024 HR) for ncu."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import kernels as K  # noqa: E402
n, h = int(os.environ.get("WP_N", 32)), 256
pre = torch.rand(n, 4 * h, 4 * h, 3, device="cuda")
yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h, device="cuda"), torch.linspace(0, 6.28, h, device="cuda"), indexing="ij")
flow = torch.stack((1.5 + 0.5 * torch.sin(yy + xx), -0.75 + 0.5 * torch.cos(yy - xx)), dim=-1).expand(n, h, h, 2).contiguous()
dst = torch.zeros(n, h, h, 64, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    K.warp_s2d_fused(pre, flow, dst, 0)
torch.cuda.synchronize()
