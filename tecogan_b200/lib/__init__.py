"""Host-side mirror of the reference's Python interface (lib/ops.py, lib/frvsr.py, lib/Teco.py,
lib/dataloader.py): same function names, argument meaning and error behaviour, torch CUDA tensors
(NHWC) instead of TF graph tensors."""
