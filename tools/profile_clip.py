"""One eager pass of the metric-config clip batch (bench.py headline) for `ncu --metrics gpu__time_duration.sum`:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python tools/profile_clip.py
TECO_PROF_CLIPS / TECO_PROF_T / TECO_PROF_LR select the shape (defaults 296 clips x 10 frames, 32x32)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tecogan_b200 import config, variables as V  # noqa: E402
from tecogan_b200.engine import ClipEngine  # noqa: E402
from tecogan_b200.init_params import xavier_params  # noqa: E402

B = int(os.environ.get("TECO_PROF_CLIPS", 296))
T = int(os.environ.get("TECO_PROF_T", 10))
LR = int(os.environ.get("TECO_PROF_LR", 32))
config.set_precision("bf16")
V.set_default_store(V.VariableStore()).load(xavier_params(1234, 16))
eng = ClipEngine(LR, LR, T, 16, batch=B, use_graph=False)
eng.clip_in.copy_(bench.synthetic_clips(T, B, LR, LR, seed=0))
eng.replay()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
