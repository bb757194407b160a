"""ctypes binding of include/maf.h (libmaf_b200.so).  No CPU fallback: importing this module without the built
library raises, and every call that fails raises MafError with the library's message."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmaf_b200.so")

MAF_ERR_INVALID, MAF_ERR_CUDA, MAF_ERR_UNSUPPORTED, MAF_ERR_TOO_SHORT = -1, -2, -3, -4


class MafError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmaf_b200 error {code}: {message}")
        self.code = code
        self.message = message


class ItemDesc(C.Structure):
    _fields_ = [("buf_len", C.c_int64), ("origin", C.c_int64), ("own_len", C.c_int64), ("g0", C.c_int64),
                ("n_global", C.c_int64), ("lo_avail", C.c_int32), ("hi_avail", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("stationary", C.c_int32), ("prop_decrease", C.c_float),
                ("auto_detect_noise", C.c_int32), ("noise_sample_duration_s", C.c_float),
                ("external_noise", C.c_int32), ("peak_normalize", C.c_int32), ("target_lufs", C.c_float),
                ("n_mels", C.c_int32), ("mel_padding", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [("pcm_off", C.c_int64), ("out_off", C.c_int64), ("mel_off", C.c_int64), ("mel_frames", C.c_int32),
                ("n_chunks", C.c_int32)]


# every symbol include/maf.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "maf_version": (C.c_int, []),
    "maf_last_error": (C.c_char_p, []),
    "maf_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "maf_destroy": (None, [_P]),
    "maf_plan_create": (C.c_int, [_P, C.POINTER(Params), C.POINTER(ItemDesc), C.c_int32, C.POINTER(_P)]),
    "maf_plan_destroy": (None, [_P]),
    "maf_plan_workspace_bytes": (C.c_size_t, [_P]),
    "maf_plan_layout": (C.c_int, [_P, C.c_int32, C.POINTER(Layout)]),
    "maf_plan_total_pcm": (C.c_int64, [_P]),
    "maf_plan_total_out": (C.c_int64, [_P]),
    "maf_plan_total_mel": (C.c_int64, [_P]),
    "maf_plan_launches": (C.c_int64, [_P]),
    "maf_plan_profile": (C.c_int, [_P, C.c_int32]),
    "maf_plan_stage_ms": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "maf_set_noise_clips": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "maf_process": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "maf_process_host": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "maf_denoise": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "maf_peak_normalize": (C.c_int, [_P, _P, _P, _P, _P]),
    "maf_loudness": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "maf_logmel": (C.c_int, [_P, _P, _P, _P, _P]),
    "maf_frame_stats": (C.c_int, [_P, _P, _P, _P, _P]),
    "maf_select_clip": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_float, _P, _P]),
    "maf_sharded_denoise": (C.c_int, [_P, _P, _P, _P, _P]),
    "maf_sharded_energy": (C.c_int, [_P, _P, _P]),
    "maf_gate": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_float, _P, _P, _P]),
    "maf_sharded_logmel": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "maf_mel_finalize": (C.c_int, [_P, _P, _P, _P]),
    "maf_plan_buffer": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "maf_plan_ws_layout": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    "maf_plan_den_offset": (C.c_int64, [_P, C.c_int32]),
    "maf_ms_energy": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P]),
    "maf_gather_compact": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P]),
    "maf_sharded_edge_bytes": (C.c_int, []),
    "maf_sharded_pack_edges": (C.c_int, [_P, _P, _P, _P]),
    "maf_sharded_unpack_edges": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "maf_sharded_pack_energy": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "maf_sharded_unpack_energy": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "maf_loudness_interleaved_ws": (C.c_int64, [C.c_int64, C.c_int32]),
    "maf_loudness_interleaved": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_float, _P, _P, _P, _P]),
    "maf_silence_layout": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "maf_silence": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, C.c_int64, C.c_int32, C.c_double, C.c_int32, _P, _P, C.c_int64, _P]),
    "maf_sharded_clip_contrib": (C.c_int, [_P, _P, _P, _P, _P]),
    "maf_set_noise_clips_dev": (C.c_int, [_P, _P, _P]),
    "maf_front_door_out_len": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "maf_front_door": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "maf_mel_windows": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libmaf_b200.so (once).  Raises ImportError when it has not been built -- there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m modular_audio_pipeline_b200.build` "
                          "(needs nvcc); this package has no CPU fallback")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MafError(rc, lib().maf_last_error().decode("utf-8", "replace"))
