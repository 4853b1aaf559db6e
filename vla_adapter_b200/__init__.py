"""vla_adapter_b200 - sm_100a engine behind VLA-Adapter's batched predict_action path.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of
include/vla_b200.h), `_lib.py` (ctypes binding), `ops.py` (operator-level wrappers used by the
parity tests) and `engine.py` (the host-side mirror of the reference's predict_action interface).
"""
__all__ = ["_lib", "ops"]
