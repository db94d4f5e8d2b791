"""ctypes helpers over the C ABI (include/kfunca_b200.h) for things the reference's Python surface does not
expose: CUDA-event timing on the library stream, pinned host buffers, async host copies, launch counter."""
from __future__ import annotations

import ctypes

import numpy as np

from . import LIB_PATH

_lib = ctypes.CDLL(LIB_PATH)
_lib.kf_last_error.restype = ctypes.c_char_p


def _ck(status: int) -> None:
    if status != 0:
        raise RuntimeError(_lib.kf_last_error().decode())


class Event:
    """CUDA event recorded on the library's compute stream (kf_event_*)."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        _ck(_lib.kf_event_create(ctypes.byref(self._h)))

    def record(self):
        _ck(_lib.kf_event_record(self._h))
        return self

    def synchronize(self):
        _ck(_lib.kf_event_synchronize(self._h))

    def elapsed_ms(self, end: "Event") -> float:
        ms = ctypes.c_float()
        _ck(_lib.kf_event_elapsed_ms(self._h, end._h, ctypes.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            _lib.kf_event_destroy(self._h)
        except Exception:
            pass


def launch_count() -> int:
    n = ctypes.c_int64()
    _ck(_lib.kf_launch_count(ctypes.byref(n)))
    return int(n.value)


def synchronize() -> None:
    _ck(_lib.kf_synchronize())


def stream_ptr() -> int:
    s = ctypes.c_void_p()
    _ck(_lib.kf_stream(ctypes.byref(s)))
    return int(s.value or 0)


class PinnedBuffer:
    """Page-locked host buffer exposed as a NumPy array (kf_host_alloc_pinned)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = ctypes.c_void_p()
        _ck(_lib.kf_host_alloc_pinned(ctypes.c_size_t(self.nbytes), ctypes.byref(self._p)))
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    @property
    def ptr(self) -> int:
        return int(self._p.value)

    def __del__(self):
        try:
            self.array = None
            _lib.kf_host_free_pinned(self._p)
        except Exception:
            pass


def copy_from_host_async(tensor, pinned: PinnedBuffer) -> None:
    """H2D of a pinned buffer into a contiguous `tensor` on the library stream, no sync (kf_memcpy_h2d_async)."""
    assert tensor.is_contiguous() and tensor.numel() * pinned.dtype.itemsize == pinned.nbytes
    _ck(_lib.kf_memcpy_h2d_async(ctypes.c_void_p(tensor.data_ptr()), ctypes.c_void_p(pinned.ptr), ctypes.c_size_t(pinned.nbytes)))


def copy_to_host_async(pinned: PinnedBuffer, tensor) -> None:
    """D2H of a contiguous `tensor` into a pinned buffer on the library stream, no sync (kf_memcpy_d2h_async)."""
    assert tensor.is_contiguous() and tensor.numel() * pinned.dtype.itemsize == pinned.nbytes
    _ck(_lib.kf_memcpy_d2h_async(ctypes.c_void_p(pinned.ptr), ctypes.c_void_p(tensor.data_ptr()), ctypes.c_size_t(pinned.nbytes)))


def gemm_host(c: PinnedBuffer, a: PinnedBuffer, b: PinnedBuffer, dtype, alpha: float = 1.0, slab_rows: int = 0) -> None:
    """c = alpha * a @ b with all three operands in pinned HOST memory (kf_gemm_host): B is uploaded once, A streams up in
    M-slabs, each slab's product runs on the library stream and comes down while the next slab goes up.  Asynchronous —
    call kf.synchronize() before reading `c`.  `dtype` is the kfunca dtype of the buffers (their NumPy dtype may be a raw
    16-bit view of bf16 data)."""
    M, K = a.shape
    K2, N = b.shape
    assert K == K2 and c.shape == (M, N)
    _ck(_lib.kf_gemm_host(ctypes.c_void_p(a.ptr), ctypes.c_void_p(b.ptr), ctypes.c_void_p(c.ptr), ctypes.c_int64(M), ctypes.c_int64(N),
                          ctypes.c_int64(K), ctypes.c_int(int(dtype)), ctypes.c_float(alpha), ctypes.c_int64(slab_rows)))
