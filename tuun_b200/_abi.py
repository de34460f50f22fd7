"""ctypes binding of tuun_b200/libtuun_b200.so — the C ABI declared in include/tuun_b200.h.

There is no fallback: if the library is missing this raises, and every compute entry point
returns TB_ERR_CUDA when no device is usable.
"""
from __future__ import annotations

import ctypes
import os

from .waveform import TbNode

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TUUN_B200_LIB") or os.path.join(_HERE, "libtuun_b200.so")

TB_OK = 0
TB_ERR_INVALID = -1
TB_ERR_UNSUPPORTED = -2
TB_ERR_CUDA = -3
TB_ERR_NOMEM = -4
TB_ERR_STATE = -5

TB_OUT_DEVICE = 1
TB_PARAMS_DEVICE = 2
TB_NO_VOICE_OUT = 4

# every symbol include/tuun_b200.h declares
EXPORTS = [
    "tb_program_create", "tb_program_destroy", "tb_render", "tb_render_mix", "tb_length", "tb_reset",
    "tb_seed_noise", "tb_substitute", "tb_segments_begin", "tb_segments_pass", "tb_segments_states", "tb_segments_fix",
    "tb_segments_end", "tb_stream", "tb_set_stream", "tb_program_get_info", "tb_lane_kernel_times", "tb_lower_check", "tb_last_error", "tb_abi_version",
]


class TbProgramInfo(ctypes.Structure):
    _fields_ = [
        ("n_nodes", ctypes.c_uint32),
        ("n_code_words", ctypes.c_uint32),
        ("n_slots", ctypes.c_uint32),
        ("state_words", ctypes.c_uint32),
        ("tile", ctypes.c_uint32),
        ("threads", ctypes.c_uint32),
        ("smem_bytes", ctypes.c_uint32),
        ("n_params", ctypes.c_uint32),
        ("kernel_launches", ctypes.c_uint64),
        ("lane_launches", ctypes.c_uint64),
        ("lane_smem_bytes", ctypes.c_uint32),
        ("lane_min_voices", ctypes.c_uint32),
        ("lane_capacity", ctypes.c_uint32),
        ("lane_fm_capacity", ctypes.c_uint32),
        ("split_passes", ctypes.c_uint32),
        ("split_segments", ctypes.c_uint32),
        ("split_seg_samples", ctypes.c_uint64),
        ("split_rounds", ctypes.c_uint64),
        ("sequence_parts", ctypes.c_uint32),
        ("split_fm_rounds", ctypes.c_uint32),
        ("sequence_renders", ctypes.c_uint64),
        ("lane_fm_ws_capacity", ctypes.c_uint32),
        ("reserved0", ctypes.c_uint32),
        ("fm_ws_launches", ctypes.c_uint64),
    ]


class TuunB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"tuun_b200 status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.c_void_p
    u32, u64 = ctypes.c_uint32, ctypes.c_uint64
    L.tb_program_create.restype = ctypes.c_int
    L.tb_program_create.argtypes = [ctypes.POINTER(TbNode), u32, P, u32, P, u64, u32, ctypes.c_int,
                                    ctypes.POINTER(P)]
    L.tb_program_destroy.restype = None
    L.tb_program_destroy.argtypes = [P]
    L.tb_render.restype = ctypes.c_int
    L.tb_render.argtypes = [P, P, u32, u32, u64, P, u64, P, u32]
    L.tb_render_mix.restype = ctypes.c_int
    L.tb_render_mix.argtypes = [P, P, u32, u32, u64, P, u64, P, P, u32]
    L.tb_length.restype = ctypes.c_int
    L.tb_length.argtypes = [P, P, u32, u32, u64, P, u32]
    L.tb_seed_noise.restype = ctypes.c_int
    L.tb_seed_noise.argtypes = [P, u64, u64]
    L.tb_substitute.restype = ctypes.c_int
    L.tb_substitute.argtypes = [P, u32, ctypes.c_float, ctypes.POINTER(u32)]
    L.tb_segments_begin.restype = ctypes.c_int
    L.tb_segments_begin.argtypes = [P, P, u32, u32, u32, u64, u32, ctypes.POINTER(u32)]
    L.tb_segments_pass.restype = ctypes.c_int
    L.tb_segments_pass.argtypes = [P, u32, u32, u32, P, u64, u32]
    L.tb_segments_states.restype = ctypes.c_int
    L.tb_segments_states.argtypes = [P, ctypes.POINTER(P), ctypes.POINTER(u64)]
    L.tb_segments_fix.restype = ctypes.c_int
    L.tb_segments_fix.argtypes = [P, u32]
    L.tb_segments_end.restype = ctypes.c_int
    L.tb_segments_end.argtypes = [P]
    L.tb_reset.restype = ctypes.c_int
    L.tb_reset.argtypes = [P]
    L.tb_stream.restype = P
    L.tb_stream.argtypes = [P]
    L.tb_set_stream.restype = ctypes.c_int
    L.tb_set_stream.argtypes = [P, P]
    L.tb_program_get_info.restype = ctypes.c_int
    L.tb_program_get_info.argtypes = [P, ctypes.POINTER(TbProgramInfo)]
    L.tb_lane_kernel_times.restype = ctypes.c_int
    L.tb_lane_kernel_times.argtypes = [P, P, u32, ctypes.POINTER(u32)]
    L.tb_lower_check.restype = ctypes.c_int
    L.tb_lower_check.argtypes = [ctypes.POINTER(TbNode), u32, P, u32, u64, ctypes.POINTER(TbProgramInfo)]
    L.tb_last_error.restype = ctypes.c_char_p
    L.tb_last_error.argtypes = []
    L.tb_abi_version.restype = u32
    _lib = L
    return L


def check(status: int):
    if status != TB_OK:
        raise TuunB200Error(status, lib().tb_last_error().decode("utf-8", "replace"))
