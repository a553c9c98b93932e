"""Drop-in for ``adam_upd_cuda`` (FourierGrid/cuda/adam_upd.cpp:79-86)."""
from unboundednerfpytorch_b200.ops import adam_upd, masked_adam_upd, adam_upd_with_perlr  # noqa: F401
