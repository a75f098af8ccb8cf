"""Micro-benchmark of the tcgen05 conv kernel (teco_conv3x3_tc) alone: CUDA-event timing over back-to-back launches
for a few shapes, algorithmic TFLOP/s and fraction of the measured bf16 peak.  Also the fused warp+s2d kernel GB/s."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tecogan_b200 import kernels as K  # noqa: E402

peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
PEAK_TF, PEAK_GBS = peaks.get("bf16_tflops", 1590.0), peaks.get("hbm_gbs", 6650.0)


def time_us(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


def conv_case(n, h, w, cin, cout, mode=0, graph=True):
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(3, 3, cout, cin, device="cuda") * 0.05 if mode else torch.randn(3, 3, cin, cout, device="cuda") * 0.05
    wpk = K.packed_weight(wt, cin, cout, transpose_layout=bool(mode))
    b = torch.zeros(cout, device="cuda")
    s = 2 if mode else 1
    y = torch.empty(n, s * h, s * w, cout, device="cuda", dtype=torch.bfloat16)
    fn = lambda: K.conv3x3_tc(x, wpk, b, y, cout=cout, act=1, mode=mode)
    if graph:  # take host launch overhead out: 20 launches per graph replay
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        us = time_us(g.replay, reps=10, warm=2) / 20
    else:
        us = time_us(fn)
    flop = 2.0 * n * h * w * 9 * cin * cout
    tf = flop / us / 1e6
    print("conv mode=%d N=%d %dx%d %d->%d: %.2f us/launch  %.1f TFLOP/s  %.3f of measured bf16 peak (%s)"
          % (mode, n, h, w, cin, cout, us, tf, tf / PEAK_TF, "graph" if graph else "eager"), flush=True)


def warp_case(n, h, w, rough=False):
    pre = torch.rand(n, 4 * h, 4 * w, 3, device="cuda")
    if rough:    # white-noise flow, +-12 HR px inside one tile: nothing like a motion field, the direct-gather path
        flow = torch.randn(n, h, w, 2, device="cuda")
    else:        # smooth motion: global translation (1.5, -0.75) LR px plus a low-frequency component
        yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h, device="cuda"), torch.linspace(0, 6.28, w, device="cuda"), indexing="ij")
        flow = torch.stack((1.5 + 0.5 * torch.sin(yy + xx), -0.75 + 0.5 * torch.cos(yy - xx)), dim=-1).expand(n, h, w, 2).contiguous()
    dst = torch.zeros(n, h, w, 64, device="cuda", dtype=torch.bfloat16)
    us = time_us(lambda: K.warp_s2d_fused(pre, flow, dst, 0))
    byts = n * 16 * h * w * (12 + 6) + n * h * w * 8
    print("warp_s2d_fused N=%d LR %dx%d (%s flow): %.2f us  %.0f GB/s algorithmic  %.3f of measured HBM peak"
          % (n, h, w, "rough" if rough else "smooth", us, byts / us / 1e3, byts / us / 1e3 / PEAK_GBS), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "conv"):
        conv_case(1, 128, 128, 64, 64)
        conv_case(1, 128, 128, 64, 64, graph=False)
        conv_case(1, 256, 256, 64, 64)
        conv_case(4, 256, 256, 64, 64)
        conv_case(1, 128, 128, 64, 64, mode=1)
        conv_case(1, 256, 256, 64, 64, mode=1)
        conv_case(1, 512, 512, 64, 16)
        conv_case(1, 64, 64, 128, 128)
        conv_case(1, 32, 32, 256, 256)
    if which in ("all", "warp"):
        warp_case(1, 128, 128)
        warp_case(8, 256, 256)
        warp_case(32, 256, 256)
        warp_case(32, 256, 256, rough=True)
    if which == "one":   # for ncu: a handful of launches of the dominant layer
        x = torch.randn(1, 128, 128, 64, device="cuda").to(torch.bfloat16)
        wpk = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.05, 64, 64)
        y = torch.empty_like(x)
        for _ in range(6):
            K.conv3x3_tc(x, wpk, torch.zeros(64, device="cuda"), y, cout=64, act=1)
        torch.cuda.synchronize()
    if which == "chain":   # wall-clock (globaltimer) timeline of 8 consecutive trunk layers sharing SMs through PDL
        from tecogan_b200 import _ffi
        xin = torch.randn(1, 128, 128, 64, device="cuda").to(torch.bfloat16)
        a, b = torch.zeros_like(xin), torch.zeros_like(xin)
        ws = [K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.03, 64, 64) for _ in range(8)]
        bz = torch.zeros(64, device="cuda")
        def chain():
            for i in range(0, 8, 2):
                K.conv3x3_tc(a, ws[i], bz, b, cout=64, act=1)
                K.conv3x3_tc(b, ws[i + 1], bz, a, cout=64, act=0, res=a)
        chain(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        buf = torch.zeros(8 * 256 * 32, device="cuda", dtype=torch.int64)
        _ffi.call("teco_debug_timing", _ffi.ptr(buf))
        with torch.cuda.graph(g):
            chain()
        g.replay(); torch.cuda.synchronize()
        buf.zero_()
        g.replay(); torch.cuda.synchronize()
        _ffi.call("teco_debug_timing", _ffi.ptr(None))
        st = buf.view(8, 256, 32)[:, :128].cpu().double()
        t0 = st[0, :, 26].min()
        print("layer: cta_start  dep_wait_done  mma_start  mma_issued  epi_done   [ns since first CTA start; median (min..max) over 128 CTAs]")
        for l in range(8):
            print("  %d: " % l + "  ".join("%6.0f (%5.0f..%5.0f)" % ((st[l, :, k] - t0).median().item(), (st[l, :, k] - t0).min().item(),
                                                                  (st[l, :, k] - t0).max().item()) for k in (26, 27, 28, 29, 30)))
    if which == "pstamps":  # per-tile timeline of the persistent kernel (multi-wave layer), median over CTAs
        from tecogan_b200 import _ffi
        n, hh, cout = (int(os.environ.get("PS_N", 4)), int(os.environ.get("PS_H", 256)), int(os.environ.get("PS_COUT", 64)))
        x = torch.randn(n, hh, hh, 64, device="cuda").to(torch.bfloat16)
        wpk = K.packed_weight(torch.randn(3, 3, 64, cout, device="cuda") * 0.05, 64, cout)
        y = torch.empty(n, hh, hh, cout, device="cuda", dtype=torch.bfloat16)
        bz = torch.zeros(cout, device="cuda")
        for _ in range(3):
            K.conv3x3_tc(x, wpk, bz, y, cout=cout, act=1)
        buf = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
        _ffi.call("teco_debug_timing", _ffi.ptr(buf))
        K.conv3x3_tc(x, wpk, bz, y, cout=cout, act=1)
        torch.cuda.synchronize()
        _ffi.call("teco_debug_timing", _ffi.ptr(None))
        st = buf.view(148, 64).cpu().double()
        t0 = st[:, 0:1]
        tl = (st[:, 32:64] - t0).view(148, 8, 4)
        print("tile: halo_seen(MMA start)  mma_issued  acc_seen(epi start)  epi_end   [median cycles since CTA start]")
        for it in range(8):
            print("  %d: " % it + "  ".join("%7.0f" % tl[:, it, k].median().item() for k in range(4)))
        print("teardown median %.0f" % (st[:, 8] - st[:, 0]).median().item())
    if which == "stamps":  # phase timeline of one launch of the dominant layer (clock64 stamps per CTA)
        from tecogan_b200 import _ffi
        x = torch.randn(1, 128, 128, 64, device="cuda").to(torch.bfloat16)
        wpk = K.packed_weight(torch.randn(3, 3, 64, 64, device="cuda") * 0.05, 64, 64)
        y = torch.empty_like(x)
        bz = torch.zeros(64, device="cuda")
        for _ in range(3):
            K.conv3x3_tc(x, wpk, bz, y, cout=64, act=1)
        buf = torch.zeros(128 * 32, device="cuda", dtype=torch.int64)
        _ffi.call("teco_debug_timing", _ffi.ptr(buf))
        K.conv3x3_tc(x, wpk, bz, y, cout=64, act=1)
        torch.cuda.synchronize()
        _ffi.call("teco_debug_timing", _ffi.ptr(None))
        st = buf.view(128, 32).cpu().double()
        rel = (st - st[:, 0:1])
        names = {1: "setup", 9: "weights_issued", 10: "pdl_wait_done(prod)", 25: "pdl_wait_done(epi)", 2: "halo_landed",
                 16: "w0", 17: "w1", 18: "w2", 19: "w3", 20: "w4", 21: "w5", 22: "w6", 23: "w7", 24: "w8", 5: "mma_issued",
                 6: "acc_ready", 11: "first_ldtm", 12: "epi_it0", 13: "epi_it1", 14: "epi_it2", 15: "epi_it3", 7: "epi_done", 8: "teardown"}
        print("cycles since CTA start (median / max over 128 CTAs):")
        for i, nme in sorted(names.items(), key=lambda kv: rel[:, kv[0]].median().item()):
            print("  %-20s %8.0f %8.0f" % (nme, rel[:, i].median().item(), rel[:, i].max().item()))
