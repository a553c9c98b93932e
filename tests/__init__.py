"""GPU / CPU parity tests of unboundednerfpytorch_b200 (see tests/conftest.py for the `gpu` marker)."""
