"""Host-side mirror of the reference generator interface (src/lib/generator.rs:39-95) over the
C ABI: `initialize_state(w)` gives a stateful program, `Generator(sample_rate).generate(w, out)`
fills `out` and returns the number of samples generated, `length` advances without output.

Beyond the reference's one-waveform calls, `Program.render` renders a whole batch of voices
(identically shaped trees with per-voice constants) in one launch — the shape the B200 path is
built for.  Arrays may be numpy (host) or torch CUDA tensors (device, written in place).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _abi
from .waveform import OpList, Waveform, flatten


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data_as(ctypes.c_void_p)


def lower_check(w) -> _abi.TbProgramInfo:
    """Validate and lower a tree on the host only (tb_lower_check): raises TuunB200Error with the
    status a render would get, else returns the launch geometry.  Needs no device."""
    ops: OpList = w if isinstance(w, OpList) else flatten(w)
    info = _abi.TbProgramInfo()
    _abi.check(_abi.lib().tb_lower_check(ops.nodes, ops.n_nodes, _np_ptr(ops.lists), len(ops.lists),
                                         len(ops.fixed_pool), ctypes.byref(info)))
    return info


class Program:
    """A Waveform with its carried state on the device (`Waveform<M, State>`, generator.rs:37)."""

    def __init__(self, w, sample_rate: int, device: int = -1):
        self.ops: OpList = w if isinstance(w, OpList) else flatten(w)
        self.sample_rate = int(sample_rate)
        h = ctypes.c_void_p()
        L = _abi.lib()
        _abi.check(L.tb_program_create(self.ops.nodes, self.ops.n_nodes, _np_ptr(self.ops.lists),
                                       len(self.ops.lists), _np_ptr(self.ops.fixed_pool),
                                       len(self.ops.fixed_pool), self.sample_rate, device,
                                       ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _abi.lib().tb_program_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> _abi.TbProgramInfo:
        info = _abi.TbProgramInfo()
        _abi.check(_abi.lib().tb_program_get_info(self._h, ctypes.byref(info)))
        return info

    def lane_kernel_times(self, last: int = 64) -> np.ndarray:
        """tb_lane_kernel_times: device milliseconds of the most recent lane-per-voice launches."""
        ms = np.zeros(min(int(last), 64), dtype=np.float32)
        n = ctypes.c_uint32(0)
        _abi.check(_abi.lib().tb_lane_kernel_times(self._h, _np_ptr(ms), len(ms), ctypes.byref(n)))
        return ms[:n.value]

    @property
    def stream(self) -> int:
        return int(_abi.lib().tb_stream(self._h) or 0)

    def set_stream(self, cuda_stream: int):
        _abi.check(_abi.lib().tb_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def seed_noise(self, seed: int, first_voice: int = 0):
        """tb_seed_noise: the seed of the per-node, per-voice Noise streams (generator.rs:113-118)."""
        _abi.check(_abi.lib().tb_seed_noise(self._h, ctypes.c_uint64(seed), ctypes.c_uint64(first_voice)))

    def substitute(self, mark_id: int, value: float) -> int:
        """waveform::substitute(&mut w, &mark_id, &Const(value)) (waveform.rs:396): every Marked node with this
        id now holds Const(value); the stream continues.  Returns the number of nodes replaced."""
        n = ctypes.c_uint32(0)
        _abi.check(_abi.lib().tb_substitute(self._h, mark_id, ctypes.c_float(value), ctypes.byref(n)))
        return int(n.value)

    def reset(self):
        """waveform::set_state(root, Initial) for every voice (waveform.rs:322)."""
        _abi.check(_abi.lib().tb_reset(self._h))

    # -- batch interface --------------------------------------------------------------------
    def render(self, out, params=None, out_len: Optional[np.ndarray] = None, n_samples=None):
        """Render the next `n_samples` of every voice into out[v, :n_samples].

        `out`: 2-D float32, numpy (host) or torch.cuda tensor (device).  `params`: [n_voices,
        n_params] float32 (numpy or torch.cuda) or None.  Returns out_len (uint64 per voice) when
        `out_len` is given or out is a numpy array.

        Device tensors: the program renders on its OWN (non-blocking) CUDA stream, which is not ordered
        against torch's.  Work queued on torch's stream that touches `out` or `params` (a `torch.zeros`
        fill, a copy) must have finished — torch.cuda.synchronize(), or an event — before this call, or
        the program must be put on torch's stream with set_stream(); without host lengths the call
        returns before the samples exist, and a reader on torch's stream has to wait for `stream` too.
        """
        flags = 0
        is_torch = hasattr(out, "data_ptr")
        if is_torch:
            assert out.is_cuda and out.dtype.is_floating_point and out.element_size() == 4
            assert out.dim() == 2 and out.stride(1) == 1
            n_voices, width = out.shape
            stride = out.stride(0)
            optr = ctypes.c_void_p(out.data_ptr())
            flags |= _abi.TB_OUT_DEVICE
        else:
            assert out.dtype == np.float32 and out.ndim == 2 and out.strides[1] == 4
            n_voices, width = out.shape
            stride = out.strides[0] // 4
            optr = out.ctypes.data_as(ctypes.c_void_p)
        n = width if n_samples is None else int(n_samples)
        pptr, n_params = None, 0
        if params is not None:
            if hasattr(params, "data_ptr"):
                assert params.is_cuda and params.is_contiguous() and params.element_size() == 4
                pptr = ctypes.c_void_p(params.data_ptr())
                n_params = params.shape[1]
                flags |= _abi.TB_PARAMS_DEVICE
            else:
                params = np.ascontiguousarray(params, dtype=np.float32)
                pptr = params.ctypes.data_as(ctypes.c_void_p)
                n_params = params.shape[1]
        want_len = out_len is not None or not is_torch
        if want_len and out_len is None:
            out_len = np.zeros(n_voices, dtype=np.uint64)
        _abi.check(_abi.lib().tb_render(self._h, pptr, n_params, n_voices, n, optr, stride,
                                        _np_ptr(out_len) if want_len else None, flags))
        return out_len

    def render_mix(self, mix, n_voices: int, params=None, out=None, out_len: Optional[np.ndarray] = None):
        """tb_render_mix: the tracker's mix loop (tracker.rs:597-642) over a batch.  `mix` is a 1-D
        float32 array (numpy, or torch.cuda when `out`/`params` are device tensors too); with
        out=None the per-voice rows are never materialised for the caller (TB_NO_VOICE_OUT)."""
        flags = 0
        dev = hasattr(mix, "data_ptr")
        n = mix.shape[0]
        if dev:
            assert mix.is_cuda and mix.is_contiguous() and mix.element_size() == 4
            mptr = ctypes.c_void_p(mix.data_ptr())
            flags |= _abi.TB_OUT_DEVICE
        else:
            assert mix.dtype == np.float32 and mix.flags.c_contiguous
            mptr = mix.ctypes.data_as(ctypes.c_void_p)
        optr, stride = None, 0
        if out is None:
            flags |= _abi.TB_NO_VOICE_OUT
        else:
            assert hasattr(out, "data_ptr") == dev, "out and mix must live on the same side"
            if dev:
                optr, stride = ctypes.c_void_p(out.data_ptr()), out.stride(0)
            else:
                optr, stride = out.ctypes.data_as(ctypes.c_void_p), out.strides[0] // 4
        pptr, n_params = None, 0
        if params is not None:
            if hasattr(params, "data_ptr"):
                pptr, n_params = ctypes.c_void_p(params.data_ptr()), params.shape[1]
                flags |= _abi.TB_PARAMS_DEVICE
            else:
                params = np.ascontiguousarray(params, dtype=np.float32)
                pptr, n_params = params.ctypes.data_as(ctypes.c_void_p), params.shape[1]
        if out_len is None and not dev:
            out_len = np.zeros(n_voices, dtype=np.uint64)
        _abi.check(_abi.lib().tb_render_mix(self._h, pptr, n_params, n_voices, n, optr, stride,
                                            _np_ptr(out_len) if out_len is not None else None, mptr, flags))
        return out_len

    def lengths(self, n_voices: int, max_: int, params=None) -> np.ndarray:
        pptr, n_params = None, 0
        if params is not None:
            params = np.ascontiguousarray(params, dtype=np.float32)
            pptr = params.ctypes.data_as(ctypes.c_void_p)
            n_params = params.shape[1]
        lens = np.zeros(n_voices, dtype=np.uint64)
        _abi.check(_abi.lib().tb_length(self._h, pptr, n_params, n_voices, max_, _np_ptr(lens), 0))
        return lens


def initialize_state(w: Waveform, sample_rate: int, device: int = -1) -> Program:
    """generator::initialize_state (generator.rs:39): the returned program starts from Initial."""
    return Program(w, sample_rate, device)


class Generator:
    """`Generator::new(sample_rate)` (generator.rs:68) — one waveform at a time, like the reference."""

    def __init__(self, sample_rate: int, device: int = -1):
        self.sample_rate = int(sample_rate)
        self.device = device

    def initialize_state(self, w: Waveform) -> Program:
        return Program(w, self.sample_rate, self.device)

    def generate(self, w: Program, out: np.ndarray) -> int:
        """Fills out[..n] and returns n; n < len(out) means the waveform finished (generator.rs:76-95)."""
        assert w.sample_rate == self.sample_rate
        if out.size == 0:
            return 0
        lens = w.render(out.reshape(1, -1))
        return int(lens[0])

    def length(self, w: Program, max_: int) -> int:
        """Generator::length (generator.rs:620)."""
        return int(w.lengths(1, max_)[0])
