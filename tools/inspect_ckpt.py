"""List a TensorFlow checkpoint without TensorFlow (like inspect_checkpoint): name, dtype enum, shape, min / max / mean.
Usage: python tools/inspect_ckpt.py <V2 prefix | V1 file> [--no-verify] [name-substring]
First thing to run on a real `model/TecoGAN` or `vgg_19.ckpt`: the reader (tecogan_b200/tf_bundle.py) was written from
the format description only, so look at the shapes and value ranges before trusting a restore."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tecogan_b200 import tf_bundle  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
if not args:
    sys.exit(__doc__)
reader = tf_bundle.load_checkpoint(args[0], verify="--no-verify" not in sys.argv)
pat = args[1] if len(args) > 1 else ""
total = 0
for name in reader.keys():
    if pat not in name:
        continue
    shape = reader.shape(name)
    line = "%-90s dtype=%-2d shape=%s" % (name, reader.dtype(name), list(shape))
    try:
        v = reader.get_tensor(name)
        total += v.size
        if v.size and v.dtype.kind in "fiu":
            line += "  min %.4g max %.4g mean %.4g" % (v.min(), v.max(), v.mean())
    except ValueError as e:
        line += "  (%s)" % e
    print(line)
print("%d values in the listed tensors" % total)
