"""Drop-in for ``ub360_utils_cuda`` (FourierGrid/cuda/ub360_utils.cpp:20-22)."""
from unboundednerfpytorch_b200.ops import cumdist_thres  # noqa: F401
