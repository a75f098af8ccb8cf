"""A/B of the number of frame pairs per fnet pass in ClipEngine (fnet_pairs) on the metric-config clip batch: ms per step and
bit-identity with one pair per pass.  python tools/ab_fnet.py [pairs ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tecogan_b200 import config, variables as V  # noqa: E402
from tecogan_b200.engine import ClipEngine  # noqa: E402
from tecogan_b200.init_params import xavier_params  # noqa: E402

B, T, LR = int(os.environ.get("TECO_PROF_CLIPS", 296)), 10, 32
pairs = [int(a) for a in sys.argv[1:]] or [1, 3, 9]
config.set_precision("bf16")
V.set_default_store(V.VariableStore()).load(xavier_params(1234, 16))
clips = bench.synthetic_clips(T, B, LR, LR, seed=0).cuda()
ref = None
for p in pairs:
    eng = ClipEngine(LR, LR, T, 16, batch=B, fnet_pairs=p)
    eng.clip_in.copy_(clips)
    for _ in range(3):
        eng.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.replay()
    e1.record()
    torch.cuda.synchronize()
    out = eng.clip_u8.clone()
    if ref is None:
        ref = out
    print("fnet_pairs %d: %.3f ms/step  (%d launches)  identical to first: %s" %
          (p, e0.elapsed_time(e1) / 5, eng.launches, bool(torch.equal(out, ref))), flush=True)
    del eng
    torch.cuda.empty_cache()
