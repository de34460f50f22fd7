"""ctypes binding of oracle/libtuun_oracle.so (TEST INFRASTRUCTURE ONLY, see tuun_oracle.h)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

from tuun_b200.waveform import OpList, TbNode, Waveform, flatten

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtuun_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("tuun_oracle.cpp", "tuun_oracle.h")]
    src.append(os.path.join(_HERE, "..", "include", "tuun_b200.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libtuun_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        P = ctypes.c_void_p
        L.tbo_program_create.restype = ctypes.c_int
        L.tbo_program_create.argtypes = [ctypes.POINTER(TbNode), ctypes.c_uint32, P, ctypes.c_uint32,
                                         P, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(P)]
        L.tbo_program_destroy.argtypes = [P]
        L.tbo_set_params.argtypes = [P, P, ctypes.c_uint32]
        L.tbo_generate.restype = ctypes.c_uint64
        L.tbo_generate.argtypes = [P, P, ctypes.c_uint64]
        L.tbo_length.restype = ctypes.c_uint64
        L.tbo_length.argtypes = [P, ctypes.c_uint64]
        L.tbo_set_state_initial.argtypes = [P]
        L.tbo_initialize_state.argtypes = [P]
        L.tbo_substitute_const.argtypes = [P, ctypes.c_uint32, ctypes.c_float]
        L.tbo_allocations.restype = ctypes.c_uint64
        L.tbo_allocations.argtypes = [P]
        L.tbo_seed_noise.argtypes = [P, ctypes.c_uint64]
        L.tbo_set_clean_tails.restype = None
        L.tbo_set_clean_tails.argtypes = [P, ctypes.c_int]
        L.tbo_set_voice.argtypes = [P, ctypes.c_uint64]
        L.tbo_render_batch.restype = ctypes.c_uint64
        L.tbo_render_batch.argtypes = [P, P, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64,
                                       ctypes.c_uint32, P, ctypes.c_uint64, P, P, ctypes.c_uint32]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data_as(ctypes.c_void_p)


class OracleProgram:
    """initialize_state(waveform) + a Generator (generator.rs:39,68)."""

    def __init__(self, w, sample_rate: int):
        self.ops: OpList = w if isinstance(w, OpList) else flatten(w)
        self.sample_rate = sample_rate
        h = ctypes.c_void_p()
        rc = lib().tbo_program_create(self.ops.nodes, self.ops.n_nodes, _ptr(self.ops.lists),
                                      len(self.ops.lists), _ptr(self.ops.fixed_pool),
                                      len(self.ops.fixed_pool), sample_rate, ctypes.byref(h))
        if rc != 0:
            raise ValueError(f"tbo_program_create failed: {rc}")
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tbo_program_destroy(self._h)
            self._h = None

    def set_params(self, row):
        row = np.ascontiguousarray(row, dtype=np.float32)
        rc = lib().tbo_set_params(self._h, _ptr(row), len(row))
        if rc != 0:
            raise ValueError("bad parameter row")

    def generate(self, out: np.ndarray) -> int:
        assert out.dtype == np.float32 and out.flags.c_contiguous
        return int(lib().tbo_generate(self._h, out.ctypes.data_as(ctypes.c_void_p), out.size))

    def length(self, max_: int) -> int:
        return int(lib().tbo_length(self._h, max_))

    def initialize_state(self):
        lib().tbo_initialize_state(self._h)

    def seed_noise(self, seed: int, voice: int = 0):
        """Seed and voice index of the Noise streams (same convention as tb_seed_noise)."""
        lib().tbo_seed_noise(self._h, ctypes.c_uint64(seed))
        lib().tbo_set_voice(self._h, ctypes.c_uint64(voice))

    def set_clean_tails(self, on: bool = True):
        """Zero scratch tails instead of reading the producers' leftovers (tuun_oracle.cpp Gen::clean_tails)."""
        lib().tbo_set_clean_tails(self._h, 1 if on else 0)

    def substitute_const(self, mark_id: int, value: float) -> int:
        return lib().tbo_substitute_const(self._h, mark_id, value)

    def render(self, n_samples: int, block: int = 1024) -> np.ndarray:
        """The bench shape: generate block by block until n_samples or the waveform ends."""
        out = np.empty(n_samples, dtype=np.float32)
        done = 0
        while done < n_samples:
            want = min(block, n_samples - done)
            got = self.generate(out[done:done + want])
            done += got
            if got < want:
                break
        return out[:done]

    def render_batch(self, params, n_voices, n_samples, block=1024, keep=True, mix=False, threads=1):
        params = None if params is None else np.ascontiguousarray(params, dtype=np.float32)
        n_params = 0 if params is None else params.shape[1]
        out = np.zeros((n_voices, n_samples), dtype=np.float32) if keep else None
        lens = np.zeros(n_voices, dtype=np.uint64)
        mixb = np.zeros(n_samples, dtype=np.float32) if mix else None
        total = lib().tbo_render_batch(self._h, _ptr(params), n_params, n_voices, n_samples, block,
                                       _ptr(out), n_samples, _ptr(lens), _ptr(mixb), threads)
        return out, lens, mixb, int(total)
