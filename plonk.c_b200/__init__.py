"""plonk.c_b200 -- B200-native batched prove/verify path of plonk.c.

The product is the C-ABI shared library built from csrc/ (include/plonk_b200.h) plus the drop-in
reference headers under include/.  This Python package is a thin host-side mirror used by the tests
and bench.py: `workload` (synthetic inputs) and `host` (ctypes binding of the C-ABI; PyTorch supplies
device buffers, streams and torch.distributed only).
"""
from . import workload  # noqa: F401
