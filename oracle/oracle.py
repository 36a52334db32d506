"""ctypes binding for the CPU oracle (oracle/maray_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (maray_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmaray_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Returns the path of the shared library."""
    src = os.path.join(_HERE, "maray_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmaray_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.mo_open.restype = ctypes.c_void_p
        L.mo_open.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.mo_close.argtypes = [ctypes.c_void_p]
        L.mo_size.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.mo_is_legacy_layout.argtypes = [ctypes.c_void_p]
        L.mo_tree_nodes.restype = ctypes.c_uint64
        L.mo_tree_nodes.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mo_set_textures.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.mo_fixed_bytes.restype = ctypes.c_size_t
        L.mo_fixed_bytes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.mo_eval.restype = ctypes.c_double
        L.mo_eval.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.mo_step_margin.restype = ctypes.c_double
        L.mo_step_margin.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.mo_render_window.argtypes = [ctypes.c_void_p] + [ctypes.c_uint32] * 4 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.mo_render_rows.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.mo_render.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


class OracleScene:
    """A `.maray` scene opened by the oracle (`maray::open` + `var_fixer::fix_color`)."""

    def __init__(self, maray_bytes: bytes, textures: Optional[Sequence[np.ndarray]] = None):
        L = lib()
        self._bytes = bytes(maray_bytes)
        self._h = L.mo_open(self._bytes, len(self._bytes))
        if not self._h:
            raise ValueError("oracle: not a .maray file in HEAD or legacy layout")
        w, h = ctypes.c_uint32(), ctypes.c_uint32()
        L.mo_size(self._h, ctypes.byref(w), ctypes.byref(h))
        self.size = (w.value, h.value)
        self.legacy = bool(L.mo_is_legacy_layout(self._h))
        if textures:
            self.set_textures(textures)

    def set_textures(self, textures: Sequence[np.ndarray]) -> None:
        """textures: list of uint8 arrays shaped (h, w, 3) -- image::RgbImage layout."""
        n = len(textures)
        arrs = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        for a in arrs:
            assert a.ndim == 3 and a.shape[2] == 3
        ptrs = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        ws = (ctypes.c_uint32 * max(n, 1))(*[a.shape[1] for a in arrs])
        hs = (ctypes.c_uint32 * max(n, 1))(*[a.shape[0] for a in arrs])
        lib().mo_set_textures(self._h, n, ptrs, ws, hs)

    def fixed_bytes(self) -> bytes:
        """`save((size, var_fixer::fix_color(color)))`: the scene as the render loops see it."""
        n = lib().mo_fixed_bytes(self._h, None, 0)
        buf = ctypes.create_string_buffer(n)
        lib().mo_fixed_bytes(self._h, buf, n)
        return buf.raw

    def tree_nodes(self, channel: int) -> int:
        return int(lib().mo_tree_nodes(self._h, channel))

    def eval(self, channel: int, x: float, y: float = 0.0) -> float:
        """`Expr::eval` (reference src/lib.rs:617-620) generalised to a y coordinate."""
        return float(lib().mo_eval(self._h, channel, x, y))

    def step_margin(self, x: float, y: float) -> float:
        """Diagnostic: how close the nearest `step` argument at pixel (x, y) is to zero, in ULP of the larger
        term of the sum it comes from (inf when no step is evaluated).  Tests use it to attribute a
        0<->255 difference to a step flip."""
        return float(lib().mo_step_margin(self._h, x, y))

    def render_window(self, x0: int, x1: int, y0: int, y1: int, threads: int = 0, want_f64: bool = False):
        """RGB8 array (y1-y0, x1-x0, 3) of that window of the image; with want_f64 also the raw
        channel values as float64 (3, y1-y0, x1-x0)."""
        threads = threads or (os.cpu_count() or 1)
        hh, ww = y1 - y0, x1 - x0
        rgb = np.zeros((hh, ww, 3), dtype=np.uint8)
        planes = np.zeros((3, hh, ww), dtype=np.float64) if want_f64 else None
        rc = lib().mo_render_window(self._h, x0, x1, y0, y1, threads, rgb.ctypes.data,
                                    planes.ctypes.data if want_f64 else None)
        if rc == -1:
            raise IndexError("oracle: App id outside Runtime::functions (the reference panics here)")
        if rc != 0:
            raise RuntimeError(f"oracle: render failed ({rc})")
        return (rgb, planes) if want_f64 else rgb

    def render(self, w: Optional[int] = None, h: Optional[int] = None, threads: int = 0) -> np.ndarray:
        w = self.size[0] if w is None else w
        h = self.size[1] if h is None else h
        return self.render_window(0, w, 0, h, threads)

    def render_rows(self, rows: Sequence[int], w: Optional[int] = None, threads: int = 0) -> np.ndarray:
        """RGB8 (len(rows), w, 3): full rows `rows` of the image (any subset, any order); the rows are
        the work items the threads pull, as in the reference's row scheduling."""
        w = self.size[0] if w is None else w
        threads = threads or (os.cpu_count() or 1)
        rows_a = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.zeros((len(rows_a), w, 3), dtype=np.uint8)
        rc = lib().mo_render_rows(self._h, 0, w, rows_a.ctypes.data, len(rows_a), threads, out.ctypes.data, None)
        if rc == -1:
            raise IndexError("oracle: App id outside Runtime::functions (the reference panics here)")
        if rc != 0:
            raise RuntimeError(f"oracle: render failed ({rc})")
        return out

    def close(self) -> None:
        if self._h:
            lib().mo_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
