"""A/B of the two versions of the fused warp + space-to-depth kernel (TECO_WARP_V2 = 0 / default, read per call) on the bench's
HBM-sized batch (32 x 1024x1024 HR, smooth and rough motion) and on the metric-config shape (296 x 128x128 HR):
time per launch, GB/s against the 18.5 B/HR-pixel algorithmic traffic, and the largest difference between the outputs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import kernels as K  # noqa: E402


def case(name, n, h, rough):
    pre = torch.rand(n, 4 * h, 4 * h, 3, device="cuda")
    yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h, device="cuda"), torch.linspace(0, 6.28, h, device="cuda"), indexing="ij")
    flow = torch.stack((1.5 + 0.5 * torch.sin(yy + xx), -0.75 + 0.5 * torch.cos(yy - xx)), dim=-1)
    if rough:
        flow = flow + 3.0 * (torch.rand(h, h, 2, device="cuda") - 0.5)
    flow = flow.expand(n, h, h, 2).contiguous()
    outs = {}
    for v in ("0", "1"):
        os.environ["TECO_WARP_V2"] = v
        dst = torch.zeros(n, h, h, 64, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            K.warp_s2d_fused(pre, flow, dst, 0, in_scale=0.5, in_shift=0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.warp_s2d_fused(pre, flow, dst, 0, in_scale=0.5, in_shift=0.5)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100.0
        gbs = n * 16 * h * h * 18.5 / (us * 1e-6) / 1e9
        outs[v] = dst.float()
        print("%-28s v%s: %8.1f us  %7.1f GB/s" % (name, "2" if v == "1" else "1", us, gbs), flush=True)
    d = (outs["0"] - outs["1"]).abs().max().item()
    print("%-28s max |v1 - v2| = %.3e   (pad channels zero: %s)" % (name, d, float(outs["1"][..., 48:].abs().max()) == 0.0), flush=True)


case("32x1024x1024 smooth", 32, 256, False)
case("32x1024x1024 rough", 32, 256, True)
case("296x128x128 smooth", 296, 32, False)
case("7x148x148 (ragged tiles)", 7, 37, True)
os.environ.pop("TECO_WARP_V2", None)
