"""ctypes wrapper of oracle/sumtree_oracle.c (TEST INFRASTRUCTURE - see the header of that file)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_longlong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsumtree_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    if not os.path.isfile(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.st_create.restype = c_void_p
    lib.st_create.argtypes = [c_longlong]
    lib.st_destroy.argtypes = [c_void_p]
    lib.st_levels.argtypes = [c_void_p]
    lib.st_level_size.restype = c_longlong
    lib.st_level_size.argtypes = [c_void_p, c_int]
    lib.st_level_ptr.restype = POINTER(c_float)
    lib.st_level_ptr.argtypes = [c_void_p, c_int]
    lib.st_total.restype = c_float
    lib.st_total.argtypes = [c_void_p]
    lib.st_set_range.argtypes = [c_void_p, c_longlong, c_longlong, c_void_p]
    lib.st_update_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    lib.st_sample.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    return lib


class SumTreeOracle:
    def __init__(self, capacity: int):
        self.lib = _load()
        self.h = self.lib.st_create(int(capacity))

    def __del__(self):
        try:
            self.lib.st_destroy(self.h)
        except Exception:
            pass

    @property
    def levels(self) -> int:
        return int(self.lib.st_levels(self.h))

    def level(self, l: int) -> np.ndarray:
        n = int(self.lib.st_level_size(self.h, l))
        return np.ctypeslib.as_array(self.lib.st_level_ptr(self.h, l), shape=(n,)).copy()

    @property
    def total(self) -> float:
        return float(self.lib.st_total(self.h))

    def set_range(self, first: int, values=None, count: int | None = None):
        if values is None:
            self.lib.st_set_range(self.h, int(first), int(count), None)
        else:
            v = np.ascontiguousarray(values, dtype=np.float32)
            self.lib.st_set_range(self.h, int(first), v.size, v.ctypes.data_as(c_void_p))

    def update_batch(self, leaf, prio):
        leaf = np.ascontiguousarray(leaf, dtype=np.int64)
        prio = np.ascontiguousarray(prio, dtype=np.float32)
        self.lib.st_update_batch(self.h, leaf.ctypes.data_as(c_void_p), prio.ctypes.data_as(c_void_p), leaf.size)

    def sample(self, u) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float32)
        out = np.empty(u.size, dtype=np.int64)
        self.lib.st_sample(self.h, u.ctypes.data_as(c_void_p), u.size, out.ctypes.data_as(c_void_p))
        return out
