#!/usr/bin/env python
"""bench.py -- headline benchmark of the TecoGAN recurrent video-SR hot path on B200.

Metric (BASELINE.json): HR frames/s, 4x SR.  Workload at N GPUs (BASELINE.json configs[1]): every rank streams
its own synthetic 120-frame clip 128x128 -> 512x512 through the full recurrence (fnet -> upscale/warp/s2d ->
generator, N=16 res-blocks, random-init weights), bf16 tensor-core convolutions.  A "step" = one 120-frame clip
per rank.  `value` = frames of all ranks / max-over-ranks device time with the LR clip resident in HBM;
`e2e` = the same through the public engine API from pinned HOST frames (H2D per frame, uint8 HR D2H per frame).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # CPU restatement of the reference (oracle/) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LR_H = LR_W = 128
CLIP_FRAMES = 120
NUM_RESBLOCK = 16
LOOKAHEAD = os.environ.get("TECO_BENCH_LOOKAHEAD", "1") != "0"   # fnet of frame t+1 concurrent with generator of frame t
# algorithmic MACs per LR pixel: generator N=16 + fnet
MACS_PER_LR_PX = 1420992 + 126720
RESBLOCK_CONV_FLOP = 2.0 * LR_H * LR_W * 576 * 64      # one 3x3 64->64 layer at 128x128 (the dominant kernel)


def synthetic_clip(frames, h, w, seed):
    """Smooth-noise video translated by (1.5,-0.75) px/frame, values in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, h // 4 + 100, w // 4 + 100, generator=g)
    big = torch.nn.functional.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False).clamp(0, 1)[0]
    out = torch.empty(frames, h, w, 3)
    for t in range(frames):
        oy, ox = 100.0 + 1.5 * t, 100.0 - 0.75 * t
        y0, x0 = int(oy), int(ox)
        fy, fx = oy - y0, ox - x0
        p = big[:, y0:y0 + h + 1, x0:x0 + w + 1]
        fr = ((1 - fy) * (1 - fx) * p[:, :h, :w] + (1 - fy) * fx * p[:, :h, 1:w + 1]
              + fy * (1 - fx) * p[:, 1:h + 1, :w] + fy * fx * p[:, 1:h + 1, 1:w + 1])
        out[t] = fr.permute(1, 2, 0)
    return out.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads this process may really use: scheduler affinity, capped by the cgroup CPU quota and by 32
    (more threads than that only thrash on the small convolutions of a 128x128 frame)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        pass
    return max(1, min(n, 32))


def cpu_reference_fps(max_frames, threads, budget_s=12.0):
    """The CPU restatement of the reference path (oracle/, 'port'): the full recurrence (main.py:253-268 order) on the
    first frames of the same synthetic clip, frame by frame until `budget_s` seconds or `max_frames` frames.
    Returns (fps, seconds, frames)."""
    from oracle import teco_oracle as O
    torch.set_num_threads(threads)
    clip = synthetic_clip(max_frames, LR_H, LR_W, seed=0)
    from tecogan_b200.init_params import xavier_params
    pw = xavier_params(1234, NUM_RESBLOCK)          # the same seeded weights the CUDA arm runs (TF variable names)
    pg = {k: v for k, v in pw.items() if k.startswith("generator/")}
    pf = {k: v for k, v in pw.items() if k.startswith("fnet/")}
    h, w = LR_H, LR_W
    with torch.no_grad():
        O.inference_sequence(pg, pf, [clip[0], clip[1]], NUM_RESBLOCK)   # warm the thread pool / allocator
        pre_inputs = torch.zeros(1, h, w, 3)
        pre_gen = torch.zeros(1, 4 * h, 4 * w, 3)
        pre_warp = torch.zeros(1, 4 * h, 4 * w, 3)
        t0 = time.perf_counter()
        n = 0
        for i in range(max_frames):
            cur = clip[i].unsqueeze(0)
            if i != 0:
                flow = O.upscale_four(O.fnet(pf, torch.cat((pre_inputs, cur), dim=-1)) * 4.0)
                pre_warp = O.dense_image_warp(pre_gen, flow)
            out = O.generator_F(pg, torch.cat((cur, O.space_to_depth4(pre_warp)), dim=-1), NUM_RESBLOCK)
            pre_inputs, pre_gen = cur, O.deprocess(out)
            n += 1
            if time.perf_counter() - t0 > budget_s and n >= 3:
                break
        dt = time.perf_counter() - t0
    return n / dt, dt, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_threads()
    vals, sample = [], 0
    for _ in range(max(1, min(args.steps, 2))):
        fps, dt, sample = cpu_reference_fps(48, cores)
        vals.append(fps)
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": "HR frames/sec (4x SR)", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * CLIP_FRAMES / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 4x SR inference 128x128->512x512, 120-frame synthetic clip", "num_resblock": NUM_RESBLOCK},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "first %d frames of the 120-frame clip, full recurrence, torch-CPU fp32 restatement "
                                   "(TensorFlow 1.x reference is not installable here)" % sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from tecogan_b200 import _ffi, config, variables as V
    from tecogan_b200 import kernels as K
    from tecogan_b200.engine import InferenceEngine
    from tecogan_b200.init_params import xavier_params   # product-side seeded init; oracle/ is only the cpu_baseline leg

    # count our C-ABI kernel launches
    counter = {"n": 0}
    orig_call = _ffi.call

    def counting_call(name, *a):
        counter["n"] += 1
        return orig_call(name, *a)
    _ffi.call = counting_call
    import tecogan_b200.kernels, tecogan_b200.engine, tecogan_b200.tc_nets
    for m in (tecogan_b200.kernels, tecogan_b200.engine, tecogan_b200.tc_nets):
        m.call = counting_call

    config.set_precision("bf16")
    st = V.set_default_store(V.VariableStore())
    st.load(xavier_params(1234, NUM_RESBLOCK))
    eng = InferenceEngine(LR_H, LR_W, NUM_RESBLOCK, batch=1, use_graph=True)

    clip_host = synthetic_clip(CLIP_FRAMES, LR_H, LR_W, seed=rank).pin_memory()
    clip_dev = clip_host.cuda()
    out_host = torch.empty((CLIP_FRAMES, 4 * LR_H, 4 * LR_W, 3), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def run_clip_resident():
        eng.reset()
        for t in range(CLIP_FRAMES):                             # look-ahead: fnet(t, t+1) overlaps generator(t)
            eng.step(clip_dev[t], next_lr=clip_dev[t + 1] if (LOOKAHEAD and t + 1 < CLIP_FRAMES) else None)

    def run_clip_e2e():
        eng.reset()
        for t in range(CLIP_FRAMES):
            # pinned host -> device inside; with look-ahead frame t+1 is the one uploaded (each frame exactly once)
            eng.step(clip_host[t], next_lr=clip_host[t + 1] if (LOOKAHEAD and t + 1 < CLIP_FRAMES) else None)
            out_host[t].copy_(eng.out_u8[0], non_blocking=True)  # uint8 HR frame back to pinned host
        torch.cuda.current_stream().synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        total_ms = 0.0
        for _ in range(steps):
            flush.fill_(1)                                       # evict L2 between timed steps (not timed)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms

    def note(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    note("engine ready; warm-up")
    for _ in range(args.warmup):
        run_clip_resident()
    run_clip_e2e()
    cpre = counter["n"]
    eng._frame_next()                      # one eager steady-state frame, only to COUNT our C-ABI kernel launches
    launches_per_frame = counter["n"] - cpre
    torch.cuda.synchronize()
    barrier()
    c0 = counter["n"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    ms_res = timed(run_clip_resident, args.steps)
    note("resident leg: %.1f ms/clip" % (ms_res / args.steps))
    barrier()
    wall = time.perf_counter() - wall0
    ms_e2e = timed(run_clip_e2e, args.steps)
    barrier()
    note("e2e leg: %.1f ms/clip" % (ms_e2e / args.steps))
    clocks = sampler.stop()
    # python-side C-ABI calls during the timed region (frame 0 of each clip runs eagerly, the rest replay the graph)
    eager_calls = counter["n"] - c0

    # --- dominant kernel: the 3x3 64->64 tcgen05 layer at 128x128, timed alone with CUDA events on this stream
    g = eng.gen
    c1, c2 = g.l_res[0]
    reps = 64

    def trunk_pairs():
        for i in range(reps // 2):
            K.conv3x3_tc(g.a, c1.wpk, c1.bias, g.b, cout=64, act=1)
            K.conv3x3_tc(g.b, c2.wpk, c2.bias, g.a, cout=64, act=0, res=g.a)
    trunk_pairs()
    torch.cuda.synchronize()
    kg = torch.cuda.CUDAGraph()          # graph replay: device time of the launches, not ctypes/Python overhead
    with torch.cuda.graph(kg):
        trunk_pairs()
    kg.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        kg.replay()
    e1.record()
    torch.cuda.synchronize()
    k_us = e0.elapsed_time(e1) * 1000.0 / (5 * reps)

    # --- HBM-bound kernel of the path: fused upscale_four + warp + space-to-depth feedback, on a batch larger than L2
    # (32 clips of 256x256 LR -> 1024x1024 HR: 403 MB read + 201 MB written), algorithmic 18.5 B per HR pixel
    wn, wh = 32, 256
    w_pre = torch.rand(wn, 4 * wh, 4 * wh, 3, device="cuda")
    yy, xx = torch.meshgrid(torch.linspace(0, 6.28, wh, device="cuda"), torch.linspace(0, 6.28, wh, device="cuda"), indexing="ij")
    w_flow = torch.stack((1.5 + 0.5 * torch.sin(yy + xx), -0.75 + 0.5 * torch.cos(yy - xx)), dim=-1).expand(wn, wh, wh, 2).contiguous()
    w_dst = torch.zeros(wn, wh, wh, 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        K.warp_s2d_fused(w_pre, w_flow, w_dst, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        K.warp_s2d_fused(w_pre, w_flow, w_dst, 0)
    e1.record()
    torch.cuda.synchronize()
    warp_us = e0.elapsed_time(e1) * 1000.0 / 5
    warp_bytes = wn * (4 * wh) * (4 * wh) * 18.5
    del w_pre, w_flow, w_dst

    t = torch.tensor([ms_res, ms_e2e], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])
    frames_total = CLIP_FRAMES * args.steps * world
    value = frames_total / (ms_res / 1000.0)
    e2e = frames_total / (ms_e2e / 1000.0)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops", 1590.0))
        peak_src = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s"
        ach_tf = RESBLOCK_CONV_FLOP / (k_us * 1e-6) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_tc_traffic.json")))["traffic_bytes_per_launch"]
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        warp_gbs = warp_bytes / (warp_us * 1e-6) / 1e9
        graph_launches = launches_per_frame   # every C-ABI call in the frame path launches exactly one kernel
        line = {
            "metric": "HR frames/sec (4x SR)", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: 4x SR inference 128x128->512x512, 120-frame synthetic clip per GPU",
                       "num_resblock": NUM_RESBLOCK, "clips_per_gpu": 1, "frames_per_step": CLIP_FRAMES,
                       "l2": "flushed by a 256 MiB write between timed steps; inside a step the recurrence's own "
                             "working set is what it is (frames depend on each other)",
                       "weights": "seeded random init (xavier, res-block/output weights x0.5)", "cuda_graph": True, "fnet_lookahead": LOOKAHEAD},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": CLIP_FRAMES * LR_H * LR_W * 3 * 4,
                    "d2h_bytes_per_step": CLIP_FRAMES * 16 * LR_H * LR_W * 3, "result": "uint8 HR frames (save_img quantisation)"},
            "gpu_launches": int(graph_launches * CLIP_FRAMES * args.steps),
            "roofline": {"bound": "tensor", "kernel": "conv3x3_tc_kernel (3x3 64->64 @128x128, res-block layer)",
                         "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         "peak_source": peak_src, "us_per_launch": k_us, "flop_per_launch": RESBLOCK_CONV_FLOP,
                         "traffic": traffic,
                         "whole_frame": {"algorithmic_gflop_per_frame": 2e-9 * MACS_PER_LR_PX * LR_H * LR_W,
                                         "achieved_tflops": 2e-12 * MACS_PER_LR_PX * LR_H * LR_W * value / world,
                                         "frac_of_sustained": 2e-12 * MACS_PER_LR_PX * LR_H * LR_W * value / world
                                         / float(peaks.get("bf16_tflops_sustained", 1400.0))}},
            "roofline_hbm": {"bound": "hbm", "kernel": "warp_s2d_fused_kernel (upscale_four + dense_image_warp + space_to_depth), "
                                                         "32 x 1024x1024 HR frames, smooth motion field (translation + low-frequency), working set 604 MB > L2",
                             "achieved": warp_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": warp_gbs / peak_gbs,
                             "us_per_launch": warp_us, "algorithmic_bytes_per_launch": warp_bytes,
                             "peak_source": "measured copy (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"},
            "wall_s_resident_leg": wall, "python_abi_calls_in_timed_region": eager_calls,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            fps, dt, nfr = cpu_reference_fps(48, cores)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "first %d frames of the same clip, full recurrence, torch-CPU fp32 oracle (%.1f s)" % (nfr, dt)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
