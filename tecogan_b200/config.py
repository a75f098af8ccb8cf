"""Run-time configuration of the compute path (not part of the reference interface)."""
_cfg = {"precision": "bf16"}


def set_precision(p):
    """'bf16': tcgen05 tensor-core convolutions for inference (default); 'fp32': CUDA-core fp32 everywhere."""
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _cfg["precision"] = p


def precision():
    return _cfg["precision"]


def use_tensor_cores():
    return _cfg["precision"] == "bf16"
