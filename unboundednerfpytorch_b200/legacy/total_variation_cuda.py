"""Drop-in for ``total_variation_cuda`` (FourierGrid/cuda/total_variation.cpp:22-24)."""
from unboundednerfpytorch_b200.ops import total_variation_add_grad  # noqa: F401
