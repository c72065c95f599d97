"""en234_fea_b200 — B200-native (sm_100a, FP64) implementation of the EN234_FEA static-step hot path.

Layout: csrc/ (CUDA kernels + C ABI, C++ host driver), api.py (ctypes glue for tests and bench).
"""
from .api import GpuSystem, MultiGpuSystem, library_partition, tie_plan, Model, Partition, Stats, synthetic_block_input, gpu_lib, host_lib, user_element_static, uel  # noqa: F401
