"""ctypes binding of libvla_b200.so (include/vla_b200.h).  There is no fallback: if the shared library is
missing or fails to load, every use raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB = None
MISSING: list[str] = []
LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvla_b200.so"

c_void_p, c_int, c_ll, c_float = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class VlaCfg(C.Structure):
    """Mirror of `vla_cfg` in include/vla_b200.h."""

    _fields_ = [
        ("n_images", C.c_int32),
        ("chunk_len", C.c_int32),
        ("action_dim", C.c_int32),
        ("proprio_dim", C.c_int32),
        ("variant", C.c_int32),
        ("dino_depth", C.c_int32),
        ("siglip_depth", C.c_int32),
        ("llm_layers", C.c_int32),
        ("vocab_size", C.c_int32),
        ("max_batch", C.c_int32),
        ("max_prompt_len", C.c_int32),
        ("causal", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/vla_b200.h declares
SIGNATURES = {
    "vla_create": (c_int, [C.POINTER(VlaCfg), C.POINTER(c_void_p)]),
    "vla_load_tensor": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, c_int, C.POINTER(C.c_int64)]),
    "vla_set_action_stats": (c_int, [c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]),
    "vla_finalize": (c_int, [c_void_p]),
    "vla_predict": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p]),
    "vla_predict_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "vla_predict_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "vla_predict_host_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "vla_set_center_crop": (c_int, [c_void_p, c_float]),
    "vla_op_center_crop_u8": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_float, c_void_p]),
    "vla_set_image_norm": (c_int, [c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vla_segment_timing": (c_int, [c_void_p, c_int]),
    "vla_segment_times": (c_int, [c_void_p, C.POINTER(C.c_float)]),
    "vla_get_tap": (c_int, [c_void_p, C.c_char_p, c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "vla_last_launch_count": (c_ll, [c_void_p]),
    "vla_last_error": (C.c_char_p, [c_void_p]),
    "vla_destroy": (None, [c_void_p]),
    "vla_op_gemm": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_ll,
                            c_int, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "vla_op_layernorm": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                 c_void_p]),
    "vla_op_rmsnorm": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    "vla_op_fold_norm": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vla_op_norm_gemm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_int, c_float, c_int, c_void_p, c_void_p]),
    "vla_op_block_tail": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_int,
                                  c_void_p, c_void_p]),
    "vla_op_attention": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_int, c_void_p]),
    "vla_op_gemm_rope": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vla_op_cross_attention": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_void_p, c_int, c_void_p]),
    "vla_set_attention_impl": (c_int, [c_int]),
    "vla_op_rope": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vla_profile_gemm": (c_int, [c_int]),
    "vla_profile_gemm_read": (c_int, [C.POINTER(C.c_double), C.POINTER(c_ll)]),
    "vla_check_errors": (c_int, [c_void_p, c_void_p]),
    "vla_watchdog_report": (c_int, [C.c_char_p, C.c_size_t]),
    "vla_watchdog_set_timeout_ms": (c_int, [c_int]),
    "vla_watchdog_selftest": (c_int, [c_void_p]),
    "vla_global_error": (C.c_char_p, []),
    "vla_total_launch_count": (c_ll, []),
}


def load() -> C.CDLL:
    """Loads the library (building it is `__graft_entry__.build()`'s job).  Raises if it is absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = Path(os.environ.get("VLA_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the CUDA extension is required (there is no CPU fallback). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` first."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            MISSING.append(name)  # tests/test_abi.py asserts this list is empty
            continue
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, engine=None) -> None:
    if rc == 0:
        return
    lib = load()
    msg = (lib.vla_last_error(engine) if engine else lib.vla_global_error()) or b""
    msg = msg.decode("utf-8", "replace")
    if rc == -4:  # CUDA error: if the device watchdog fired, its records say which barrier was stuck
        buf = C.create_string_buffer(4096)
        if lib.vla_watchdog_report(buf, 4096) > 0 and buf.value.decode("utf-8", "replace") not in msg:
            msg += "\n" + buf.value.decode("utf-8", "replace")
    if rc in (-1, -2, -3):
        raise ValueError(f"libvla_b200: {msg} (status {rc})")
    raise RuntimeError(f"libvla_b200: {msg} (status {rc})")
