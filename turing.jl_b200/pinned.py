"""Pinned (page-locked) host arrays for the D2H copy of kept draws, allocated through the C ABI
(tnuts_host_alloc) and recycled through a small pool: page-locking 10+ GB costs seconds, the
copy itself a fraction of one."""
import ctypes as C
import threading
import weakref

import numpy as np

from . import _lib

_pool = {}      # nbytes -> [address, ...]
_lock = threading.Lock()
_MAX_POOL_BYTES = 48 << 30
_pool_bytes = 0


def _release(addr, nbytes):
    global _pool_bytes
    with _lock:
        if _pool_bytes + nbytes <= _MAX_POOL_BYTES:
            _pool.setdefault(nbytes, []).append(addr)
            _pool_bytes += nbytes
            return
    _lib.load().tnuts_host_free(C.c_void_p(addr))


def empty(shape, dtype=np.float64):
    """numpy array backed by pinned memory; returned to the pool when garbage-collected"""
    global _pool_bytes
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if n == 0:
        return np.empty(shape, dtype)
    addr = None
    with _lock:
        lst = _pool.get(n)
        if lst:
            addr = lst.pop()
            _pool_bytes -= n
    if addr is None:
        p = C.c_void_p()
        rc = _lib.load().tnuts_host_alloc(C.byref(p), n)
        if rc != 0 or not p.value:
            return np.empty(shape, dtype)  # pageable: slower copy, same result
        addr = p.value
    buf = (C.c_char * n).from_address(addr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, _release, addr, n)
    return arr


def drain():
    global _pool_bytes
    with _lock:
        items = [(a, n) for n, lst in _pool.items() for a in lst]
        _pool.clear()
        _pool_bytes = 0
    for a, _ in items:
        _lib.load().tnuts_host_free(C.c_void_p(a))
