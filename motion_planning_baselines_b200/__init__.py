"""B200-native batched trajectory cost-and-update hot path for the mp_baselines planners.

The CUDA kernels live in csrc/ behind the C ABI of include/mpb.h (libmpb_b200.so); this package
is the Python host mirroring the reference's planner / cost / field interface."""
from . import configs, models  # noqa: F401

__all__ = ['configs', 'models']
