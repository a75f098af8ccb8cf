"""One steady-state frame of the bench workload (128x128 -> 512x512, N=16) launched eagerly inside a
cudaProfilerStart/Stop range, for `ncu --profile-from-start off` (launch list / full capture)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tecogan_b200.init_params import xavier_params  # noqa: E402
from tecogan_b200 import config, variables as V  # noqa: E402
from tecogan_b200.engine import InferenceEngine  # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frames_profiled = int(sys.argv[2]) if len(sys.argv) > 2 else 1
config.set_precision("bf16")
V.set_default_store(V.VariableStore()).load(xavier_params(1234, 16))
eng = InferenceEngine(h, h, 16, use_graph=False)
clip = bench.synthetic_clip(8, h, h, 0).cuda()
for t in range(4):
    eng.step(clip[t])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in range(4, 4 + frames_profiled):
    eng.step(clip[t])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
