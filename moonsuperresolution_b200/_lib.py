"""ctypes binding of ``libmoonsr.so`` (the C ABI declared in ``include/moonsr.h``).

There is NO fallback: if the shared library is missing, cannot be loaded, or lacks a symbol, importing the compute
path raises.  PyTorch is used by callers only for device memory, streams and ``torch.distributed``; every pointer
that crosses this boundary is a plain address.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmoonsr.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "moonsr.h")

MSR_OK = 0
ARCH = {"spade": 0, "cnn": 1, "pix2pix": 2}
PRECISION = {"fp32": 0, "bf16": 1}
REPEAT_NONE, REPEAT_FIRST, REPEAT_NEXT = 0, 1, 2

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must stay in sync with include/moonsr.h (tests/test_abi.py checks every prototype)
SIGNATURES: Dict[str, tuple] = {
    "msr_version": (_i, []),
    "msr_last_error": (C.c_char_p, []),
    "msr_device_sm_count": (_i, []),
    "msr_profile_enable": (_i, [_i]),
    "msr_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "msr_profile_records": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), _i64, C.POINTER(_i64)]),
    "msr_resize_area4": (_i, [_vp, _i, _i, _vp, _i, _i, _f, _vp]),
    "msr_resize_cubic": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp]),
    "msr_pad_inputs": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "msr_validity_sat": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp]),
    "msr_patch_validity": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    "msr_gather_normalize": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "msr_blend_tile": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i64, _i, _i,
                            _vp]),
    "msr_blend_accumulate": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _i64, _i, _i,
                                  _i, _i, _i, _vp]),
    "msr_blend_finalize": (_i, [_vp, _vp, _vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i64, _vp]),
    "msr_blend_tile_fast": (_i, [_vp, _vp, _i, _vp, _i, _vp, _f, _f, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i64, _i, _i,
                                 _vp]),
    "msr_blend_accumulate_fast": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _i64,
                                       _i, _i, _i, _i, _i, _vp]),
    "msr_tiff_lzw_bound": (_i64, [_i64]),
    "msr_tiff_encode_strips": (_i, [_vp, _i64, _i, _i, _i, _i, _i, _vp, _i64, _vp, _i]),
    "msr_tiff_decode_chunks": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i64, _i, _i, _i, _vp, _i64, _i64, _i]),
    "msr_generator_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "msr_generator_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "msr_generator_finalize": (_i, [_vp]),
    "msr_generator_forward": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "msr_generator_forward_repeat": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "msr_generator_last_launch_count": (_i64, [_vp]),
    "msr_generator_device_bytes": (_i64, [_vp]),
    "msr_generator_read_activation": (_i, [_vp, C.c_char_p, _vp, _i64, C.POINTER(_i64)]),
    "msr_generator_destroy": (_i, [_vp]),
    "msr_op_conv3x3_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "msr_op_conv_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "msr_host_pack_mask_weights": (_i, [_vp, _vp, _i, _vp]),
    "msr_host_pack_phase_weights": (_i, [_vp, _i, _vp, C.POINTER(_i), C.POINTER(_i)]),
    "msr_op_enc1_tc": (_i, [_vp, _i, _vp, _vp, _i, _f, _vp]),
    "msr_op_mask_tc": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "msr_op_phase_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "msr_op_spade_tc": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "msr_debug_tc_counters": (_i, [_vp]),
    "msr_op_conv3x3_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}


class MoonSRError(RuntimeError):
    pass


_lib = None


def header_symbols() -> List[str]:
    """Every function name declared in include/moonsr.h."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(msr_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Loads libmoonsr.so once and types every entry point.  Raises MoonSRError when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MoonSRError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C moonsuperresolution_b200/csrc`).  There is no CPU fallback.")
    try:
        handle = C.CDLL(LIB_PATH)
    except OSError as e:
        raise MoonSRError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise MoonSRError(f"{LIB_PATH} does not export {name}") from e
        fn.restype, fn.argtypes = res, args
    _lib = handle
    return _lib


PROFILE_FAMILIES = ("conv_tc", "conv_f32", "mask_conv", "stats", "elementwise", "dense", "final_conv", "pad",
                    "validity", "gather", "blend", "preprocess")


def profile_enable(on: bool) -> None:
    check(lib().msr_profile_enable(1 if on else 0), "msr_profile_enable")


def profile_read() -> Dict[str, dict]:
    """family -> {"ms", "work" (FLOPs or bytes), "launches"} summed since profile_enable(True)."""
    n = len(PROFILE_FAMILIES)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (_i64 * n)()
    check(lib().msr_profile_read(ms, work, cnt), "msr_profile_read")
    return {f: {"ms": ms[k], "work": work[k], "launches": int(cnt[k])} for k, f in enumerate(PROFILE_FAMILIES)}


def profile_records(family: str, capacity: int = 1 << 16):
    """[(ms, work)] per launch group of ``family`` in launch order."""
    ms, work, cnt = (C.c_double * capacity)(), (C.c_double * capacity)(), _i64(0)
    check(lib().msr_profile_records(PROFILE_FAMILIES.index(family), ms, work, capacity, C.byref(cnt)),
          "msr_profile_records")
    n = min(int(cnt.value), capacity)
    return [(ms[k], work[k]) for k in range(n)]


def last_error() -> str:
    return (lib().msr_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != MSR_OK:
        raise MoonSRError(f"{what or 'libmoonsr call'} failed ({rc}): {last_error()}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / numpy array; None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
