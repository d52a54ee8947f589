"""C arrays -> Python lists for the response side of ``B200Model.predict`` (``csrc/pylists.cpp`` -> ``lib/libb2fpy.so``).

The reference's model object returns plain lists (``...predict_proba(...)[:, 1].tolist()``, reference
``databricks/src/02-register-model.ipynb:335-337``) because the handler ``json.dumps``-es and re-validates them
(``app/main.py:75-86``).  At 65 536 rows that conversion is the largest single cost of the plugin call, so it is done by a
small CPython helper that RECYCLES float objects whose previous response has been dropped (see the C file).  The helper is
optional: without ``Python.h`` at build time the lists come from ``ndarray.tolist()``.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb2fpy.so")
_lib = None
_tried = False


def _load():
    global _lib, _tried
    if not _tried:
        _tried = True
        if os.path.exists(_PATH) and os.environ.get("B200_PYLISTS", "1") != "0":
            try:
                lib = C.PyDLL(_PATH)  # PyDLL: calls keep the GIL -- these functions create Python objects
                lib.b2f_py_list_new.restype = C.py_object
                lib.b2f_py_list_new.argtypes = [C.c_ssize_t]
                for name in ("b2f_py_list_fill_f64", "b2f_py_list_fill_i32"):
                    fn = getattr(lib, name)
                    fn.restype = C.c_int
                    fn.argtypes = [C.py_object, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_ssize_t]
                lib.b2f_py_pool_stats.restype = None
                lib.b2f_py_pool_stats.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
                _lib = lib
            except (OSError, AttributeError):
                _lib = None
    return _lib


def available() -> bool:
    return _load() is not None


class ListBuilder:
    """A list of ``n`` values filled piecewise from numpy views (float64 or int32, any stride)."""

    def __init__(self, n: int):
        self.n = int(n)
        self._lib = _load()
        self.items = self._lib.b2f_py_list_new(self.n) if self._lib is not None else []

    def fill(self, offset: int, values: np.ndarray) -> None:
        """items[offset : offset + len(values)] = values (a 1-D float64 / int32 view; record fields welcome)."""
        if self._lib is None:
            self.items += values.tolist()
            return
        if values.ndim != 1 or values.dtype not in (np.float64, np.int32):
            raise TypeError("ListBuilder.fill takes 1-D float64 or int32 arrays")
        if len(values) == 0:
            return
        fn = self._lib.b2f_py_list_fill_f64 if values.dtype == np.float64 else self._lib.b2f_py_list_fill_i32
        # PyDLL turns a -1 return with a Python exception set into that exception
        fn(self.items, int(offset), values.ctypes.data, len(values), values.strides[0])


def pool_stats():
    """(recycled, freshly allocated) float objects so far; (0, 0) without the helper."""
    lib = _load()
    if lib is None:
        return 0, 0
    a, b = C.c_longlong(0), C.c_longlong(0)
    lib.b2f_py_pool_stats(C.byref(a), C.byref(b))
    return a.value, b.value
