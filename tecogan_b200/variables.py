"""TF1-style variable scopes for the mirror of the reference interface.

The reference keeps its weights in the TensorFlow graph under variable scopes
(`tf.variable_scope('generator')`, main.py:203; names in SURVEY.md Appendix C).  Here a
VariableStore maps the same full names to fp32 CUDA tensors, so `generator_F(inputs, 3,
reuse=False, FLAGS=FLAGS)` can keep its reference signature and a TF checkpoint name map stays trivial.
"""
import contextlib
import math
from collections import OrderedDict

import torch


class VariableStore(OrderedDict):
    """name -> fp32 CUDA tensor.  `version` bumps whenever a tensor object is replaced or updated in place
    by an optimiser, so packed bf16 copies (tecogan_b200/lib/ops.py) know when to re-pack."""

    _next_uid = [0]

    def __init__(self, device="cuda", seed=1234):
        super().__init__()
        VariableStore._next_uid[0] += 1
        self.uid = VariableStore._next_uid[0]      # never reused (id() of a freed store can be): key for packed-weight caches
        self.device = torch.device(device)
        self.gen = torch.Generator().manual_seed(seed)
        self.version = 0

    def load(self, params):
        """Install tensors (e.g. oracle-initialised weights in tests) under their TF names."""
        for k, v in params.items():
            self[k] = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.version += 1

    def touch(self):
        self.version += 1


_state = {"store": None, "scope": []}


def default_store():
    if _state["store"] is None:
        _state["store"] = VariableStore()
    return _state["store"]


def set_default_store(store):
    _state["store"] = store
    return store


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    """tf.variable_scope: pushes a name component.  `reuse` is accepted for signature parity; variables
    are created on first use and shared afterwards (what reuse=True does in the reference)."""
    _state["scope"].append(name)
    try:
        yield "/".join(_state["scope"])
    finally:
        _state["scope"].pop()


def current_scope():
    return "/".join(_state["scope"])


def xavier_uniform(store, shape, fan_in, fan_out):
    """tf.contrib.layers.xavier_initializer() (uniform) -- reference lib/ops.py:40,52,98."""
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    t = (torch.rand(shape, generator=store.gen, dtype=torch.float32) * 2 - 1) * lim
    return t.to(store.device)


def get_variable(local_name, shape, init="xavier", fans=None):
    store = default_store()
    full = "/".join(_state["scope"] + [local_name])
    if full in store:
        v = store[full]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, expected %s" % (full, tuple(v.shape), tuple(shape)))
        return v
    if init == "xavier":
        v = xavier_uniform(store, shape, *fans)
    elif init == "zeros":
        v = torch.zeros(shape, dtype=torch.float32, device=store.device)
    else:
        raise ValueError("unknown initializer " + str(init))
    store[full] = v
    store.version += 1
    return v
