"""ctypes binding of librap_b200.so (the C ABI declared in include/rap_b200.h).

There is deliberately no fallback: if the shared library is missing the import fails loudly, and
every compute entry point fails if no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librap_b200.so")


class RapError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(rap_b200 has no CPU / PyTorch fallback)")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _i32p, _f32p, _u8p = C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p
PROTOTYPES = {
    "rap_version": (C.c_int, []),
    "rap_last_error": (C.c_char_p, []),
    "rap_device_error": (C.c_int, [C.POINTER(C.c_uint)]),
    "rap_dit_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int]),
    "rap_dit_destroy": (None, [_vp]),
    "rap_dit_load_weight": (C.c_int, [_vp, C.c_char_p, _f32p, C.c_int64, _vp]),
    "rap_dit_finalize": (C.c_int, [_vp]),
    "rap_dit_num_keys": (C.c_int, [_vp]),
    "rap_dit_key": (C.c_char_p, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "rap_plan_create": (C.c_int, [C.POINTER(_vp), _vp, C.c_int, _i32p, _i32p, C.c_int, C.c_int]),
    "rap_plan_destroy": (None, [_vp]),
    "rap_plan_tokens": (C.c_int, [_vp]),
    "rap_plan_valid_parts": (C.c_int, [_vp]),
    "rap_plan_set_condition": (C.c_int, [_vp, _f32p, _f32p, _f32p, _u8p, _vp]),
    "rap_dit_forward": (C.c_int, [_vp, _f32p, _f32p, _f32p, _f32p, _vp]),
    "rap_sample_euler": (C.c_int, [_vp, _f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _vp]),
    "rap_procrustes_fit": (C.c_int, [_f32p, _f32p, _vp, _vp, C.c_int, C.c_int64, _f32p, _f32p, _vp]),
    "rap_procrustes_rigidify": (C.c_int, [_f32p, _f32p, _vp, C.c_int, C.c_int64, _f32p, _f32p, C.c_float, C.c_float, _vp]),
    "rap_procrustes_fit_sse": (C.c_int, [_f32p, _f32p, _vp, _vp, C.c_int, C.c_int64, _f32p, _f32p, _f32p, _vp]),
    "rap_rigid_sse": (C.c_int, [_f32p, _f32p, _vp, _vp, C.c_int, C.c_int64, _f32p, _f32p, _f32p, _vp]),
    "rap_plan_fit": (C.c_int, [_vp, _f32p, _f32p, _f32p, _f32p, _vp]),
    "rap_plan_rigidify": (C.c_int, [_vp, _f32p, _f32p, _f32p, _vp]),
    "rap_spinnet_create": (C.c_int, [C.POINTER(_vp)]),
    "rap_spinnet_destroy": (None, [_vp]),
    "rap_spinnet_num_keys": (C.c_int, [_vp]),
    "rap_spinnet_key": (C.c_char_p, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "rap_spinnet_load_weight": (C.c_int, [_vp, C.c_char_p, _f32p, C.c_int64, _vp]),
    "rap_spinnet_finalize": (C.c_int, [_vp]),
    "rap_spinnet_extract": (C.c_int, [_vp, _f32p, C.c_int, _vp, _f32p, C.c_int, C.c_float, _f32p, _vp]),
    "rap_voxel_downsample": (C.c_int, [_f32p, C.c_int, C.c_float, _vp, C.POINTER(C.c_int), _vp]),
    "rap_voxel_coverage": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), _vp]),
    "rap_fps_capacity": (C.c_int, []),
    "rap_fps_batched": (C.c_int, [_f32p, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _f32p, _vp]),
    "rap_statistical_outliers": (C.c_int, [_f32p, C.c_int, C.c_int, C.c_double, _vp, C.POINTER(C.c_int), _vp]),
    "rap_segment_nn": (C.c_int, [_f32p, _f32p, _vp, C.c_int, _vp, _vp, _f32p, _vp]),
    "rap_apply_transform": (C.c_int, [_vp, C.c_longlong, C.c_int, C.POINTER(C.c_double), _vp, _vp]),
    "rap_bench_fmha": (C.c_int, [_vp, C.c_int, C.c_int, _i32p, C.c_int, _vp, C.c_int, C.POINTER(C.c_float)]),
    "rap_bench_fmha_variant": (C.c_int, [_vp, C.c_int, C.c_int, _i32p, C.c_int, _vp, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "rap_fmha_schedule": (C.c_int, [_i32p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rap_fmha_schedule_persistent": (C.c_int, [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                               C.POINTER(C.c_int)]),
    "rap_test_spinnet_patches": (C.c_int, [_f32p, C.c_int, _vp, _f32p, C.c_int, C.c_float, _f32p, _vp]),
    "rap_kabsch_host": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rap_launch_count": (C.c_int64, []),
    "rap_profile_begin": (C.c_int, []),
    "rap_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "rap_profile_classes": (C.c_int, []),
    "rap_profile_class_name": (C.c_char_p, [C.c_int]),
    "rap_test_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _f32p, _f32p, _f32p, _vp, _f32p, _vp, _vp]),
    "rap_test_gemm_ln": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _f32p, _f32p, _f32p, _f32p, _f32p, _vp, C.c_int, C.c_float, _vp, _vp]),
    "rap_test_ln_modulate": (C.c_int, [_f32p, _vp, C.c_int, _f32p, _f32p, C.c_int, _vp, C.c_float, _vp]),
    "rap_debug_read": (C.c_int, [_vp, C.c_int, _vp, C.c_int64]),
    "rap_test_fmha": (C.c_int, [_vp, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _vp, _vp, _vp]),
}
for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)           # AttributeError here = header/library mismatch
    _fn.restype, _fn.argtypes = _res, _args


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RapError(f"{what or 'rap_b200'} failed ({rc}): {lib.rap_last_error().decode(errors='replace')}")


def device_error() -> int:
    code = C.c_uint(0)
    check(lib.rap_device_error(C.byref(code)), "rap_device_error")
    return code.value


def ptr(t) -> int | None:
    """Device pointer of a contiguous torch tensor (None passes NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "rap_b200 kernels need contiguous tensors"
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def i32_array(values) -> C.Array:
    vals = [int(v) for v in values]
    return (C.c_int32 * len(vals))(*vals)


class profile:
    """Context manager: per-kernel-class device time of everything the library launches inside the block."""

    def __enter__(self):
        check(lib.rap_profile_begin(), "rap_profile_begin")
        self.result = {}
        return self

    def __exit__(self, *exc):
        n = lib.rap_profile_classes()
        ms, cnt = (C.c_float * n)(), (C.c_int * n)()
        check(lib.rap_profile_end(ms, cnt), "rap_profile_end")
        self.result = {lib.rap_profile_class_name(i).decode(): {"ms": float(ms[i]), "launches": int(cnt[i])} for i in range(n) if cnt[i]}
        return False
