"""Training step timing (BASELINE configs[2]/[3]): FRVSR case 4 and TecoGAN case 3 at B=4, RNN_N=10, 32x32 LR, fp32."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main as M  # noqa: E402
from tecogan_b200 import variables as V  # noqa: E402
from tecogan_b200.init_params import xavier_params  # noqa: E402
from tecogan_b200.lib.dataloader import frvsr_gpu_data_loader  # noqa: E402
from tecogan_b200.lib.Teco import FRVSR, TecoGAN  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "frvsr"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if which == "frvsr":
    F = M.parse_flags(["--mode", "train", "--output_dir", "/tmp/x", "--num_resblock", "10", "--ratio", "-0.01", "--nopingpang",
                       "--learning_rate", "0.00005", "--decay_rate", "1.0", "--stair"])
else:
    F = M.parse_flags(["--mode", "train", "--output_dir", "/tmp/x", "--num_resblock", "16", "--ratio", "0.01", "--pingpang",
                       "--pp_scaling", "0.5", "--vgg_scaling", "0.2", "--learning_rate", "0.00005", "--decay_rate", "1.0", "--stair"])
gan = F.ratio > 0
from tecogan_b200 import config  # noqa: E402
config.set_train_precision(os.environ.get("TECO_TRAIN_PRECISION", "fp32"))
st = V.set_default_store(V.VariableStore())
st.load(xavier_params(1, F.num_resblock, gan, F.vgg_scaling > 0))
dev = torch.device("cuda")
lr, tg = frvsr_gpu_data_loader(M.synthetic_hr_batch(F, 0, 0, dev), F)
Net = TecoGAN(lr, tg, F) if gan else FRVSR(lr, tg, F)
if os.environ.get("TECO_TRAIN_NOGRAPH"):
    Net.train.use_graph = False
for i in range(2):
    r = Net.train()
torch.cuda.synchronize()
if os.environ.get("TECO_TRAIN_PROFILE"):      # one eager step inside a profiler range (ncu --profile-from-start off)
    torch.cuda.profiler.start()
    Net.train()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
t0 = time.perf_counter()
for i in range(steps):
    r = Net.train()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print("%s [%s]: %.1f ms/step  -> %.1f unique HR frames/s (B=%d x RNN_N=%d); losses %s" % (
    which, config.train_precision(), dt * 1e3, F.batch_size * F.RNN_N / dt, F.batch_size, F.RNN_N,
    dict(zip(r["update_list_name"], [round(v, 5) for v in r["update_list"]]))))
