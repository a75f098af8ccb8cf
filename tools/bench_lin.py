"""Timing + per-strip stamps of the row-linearised multi-layer kernel (teco_conv3x3_lin_tc) on the metric-config batch."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import _ffi, kernels as K  # noqa: E402

n, h = int(os.environ.get("BL_N", 296)), int(os.environ.get("BL_H", 32))
x = (torch.randn(n, h, 32, 64, device="cuda") * 0.5).to(torch.bfloat16)
a, b = torch.zeros_like(x), torch.zeros_like(x)
for L in [int(v) for v in os.environ.get("BL_L", "1,3,33").split(",")]:
    ws = torch.cat([K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64) for _ in range(L)]).contiguous()
    bs = torch.zeros(L * 64, device="cuda")
    plan = ([(0, 1, -1, 1)] + [(1, 2, -1, 1), (2, 1, 1, 0)] * 16)[:L]
    fn = lambda: K.conv3x3_lin_chain(x, a, b, ws, bs, plan)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    flop = 2.0 * n * h * 32 * 576 * 64 * L
    print("lin chain N=%d H=%d L=%2d: %8.1f us/launch  %6.2f us/layer  %7.1f TFLOP/s" % (n, h, L, us, us / L, flop / us / 1e6), flush=True)
    if L == int(os.environ.get("BL_STAMP_L", 1)):
        for flags in [int(v) for v in os.environ.get("BL_STAMP_FLAGS", "0").split(",")]:
            os.environ["TECO_LIN_DBG"] = str(flags)
            buf = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
            _ffi.call("teco_debug_timing", _ffi.ptr(buf))
            fn(); torch.cuda.synchronize()
            _ffi.call("teco_debug_timing", _ffi.ptr(None))
            os.environ["TECO_LIN_DBG"] = "0"
            st = buf.view(148, 64).cpu().double()
            t0 = st[:, 0:1]
            cyc, ns = (st[:, 3] - st[:, 0]).median().item(), (st[:, 2] - st[:, 1]).median().item()
            print("dbg flags %d: CTA lifetime %.0f cycles = %.1f us -> SM clock %.0f MHz" % (flags, cyc, ns / 1e3, cyc / ns * 1e3))
            el = (st[:, 4:64] - t0).view(148, 15, 4)
            print("strip: mma_start  mma_issued  epi_start  epi_end   [median cycles since CTA start]")
            for i in range(15):
                print("  %2d: " % i + "  ".join("%7.0f" % el[:, i, k].median().item() for k in range(4)))
if os.environ.get("BL_BISECT"):
    # developer bisection (TECO_LIN_DBG): which part of the strip pipeline slows the MMA stream
    ws = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
    bs = torch.zeros(64, device="cuda")
    for flags in (0, 1, 2, 3, 4, 7, 10, 15):
        os.environ["TECO_LIN_DBG"] = str(flags)
        fn = lambda: K.conv3x3_lin_chain(x, a, b, ws, bs, [(0, 1, -1, 1)])
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print("dbg flags %2d (1 no store, 2 no epilogue math/staging, 4 no halo TMA, 8 no TMEM loads): %.2f us/layer" % (flags, e0.elapsed_time(e1) * 50), flush=True)
    os.environ["TECO_LIN_DBG"] = "0"
# reference point: the per-layer persistent kernel on the same batch
w1 = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64)
bz = torch.zeros(64, device="cuda")
fn = lambda: (K.conv3x3_tc(x, w1, bz, a, cout=64, act=1), K.conv3x3_tc(a, w1, bz, x, cout=64, act=0, res=x))
fn(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(8):
        fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / 80
print("per-layer conv3x3_tc N=%d: %.2f us/layer  %.1f TFLOP/s" % (n, us, 2.0 * n * h * 32 * 576 * 64 / us / 1e6))
