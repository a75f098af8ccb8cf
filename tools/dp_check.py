"""2-GPU data-parallel check (run under torchrun): each rank trains on its own clip shard with one NCCL all-reduce per
step.  Verifies (a) the reduced gradient bucket equals the mean of the two single-rank gradient buckets, (b) parameters
and the with-D decision stay identical on all ranks after several steps."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main as M  # noqa: E402
from tecogan_b200 import variables as V  # noqa: E402
from tecogan_b200.init_params import xavier_params  # noqa: E402
from tecogan_b200.lib.dataloader import frvsr_gpu_data_loader  # noqa: E402
from tecogan_b200.lib.Teco import TecoGAN  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
F = M.parse_flags(["--mode", "train", "--output_dir", "/tmp/x", "--num_resblock", "2", "--ratio", "0.01", "--pingpang",
                   "--pp_scaling", "0.5", "--vgg_scaling", "0.2", "--batch_size", "1", "--RNN_N", "3", "--crop_size", "16"])
P = xavier_params(5, 2, True, True)


def build(r):
    V.set_default_store(V.VariableStore()).load(P)
    lr, tg = frvsr_gpu_data_loader(M.synthetic_hr_batch(F, 0, r, dev), F)
    return TecoGAN(lr, tg, F)


# single-rank references (before the process group exists: allreduce_bucket is a no-op)
solo = []
for r in range(world):
    net = build(r)
    net.train()
    solo.append(net.train.bucket.clone())
mean_bucket = sum(solo) / world

dist.init_process_group("nccl", device_id=dev)
net = build(rank)
out = net.train()
st = net.train
ngrad = st.bucket.numel() - st.n_scalars
got = st.bucket[:ngrad] / world
err = (got - mean_bucket[:ngrad]).abs().max().item() / mean_bucket[:ngrad].abs().max().item()
decisions = [out["with_d"]]
for _ in range(3):
    decisions.append(net.train()["with_d"])
chk = torch.stack([st.opt_g.flat.double().sum(), st.opt_f.flat.double().sum(), st.opt_d.flat.double().sum(),
                   torch.tensor(float(sum(decisions)), device=dev, dtype=torch.float64)])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
if rank == 0:
    print("DP check: reduced-grad rel err vs mean of solo runs %.2e; identical params+decisions on all ranks: %s; decisions %s"
          % (err, same, decisions))
    assert err < 1e-5 and same
dist.destroy_process_group()
