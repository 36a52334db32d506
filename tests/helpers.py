"""Test-side checkers of the HOST logic (lowering, code generation, bytecode), runnable without a GPU.

Nothing here is part of the product and nothing in the product can reach it:
  * host_jit_run   compiles the CUDA source the NVRTC back end generated as plain C++ with g++
                   (-ffp-contract=off) behind a shim of the few CUDA intrinsics it uses, and runs the
                   kernel body thread by thread.  It checks the generated program text, not the GPU.
  * bytecode_run   a numpy reading of the interpreter's bytecode (csrc/bytecode.hpp).
Both are compared with the oracle in tests/test_host_lowering.py.
"""
from __future__ import annotations

import ctypes
import hashlib
import math
import os
import subprocess
import tempfile

import numpy as np

HOST_SHIM = r"""
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define MR_LIBM_PLAIN 1          /* the host check uses the host libm, like the oracle */
#define MR_PLAIN_FN static inline
#define MR_HOST_TEXT 1
#define MR_DYN_DECL static double mr_dyn_f64[32 * 256];   /* the scratch rows of the batched sin/exp/ln calls */
struct uint3_ { unsigned int x, y, z; };
static uint3_ threadIdx, blockIdx, blockDim;
struct uint4 { unsigned int x, y, z, w; };
static inline double __drcp_rn(double v) { return 1.0 / v; }
static inline double __dsqrt_rn(double v) { return std::sqrt(v); }
static inline unsigned int __double2uint_rz(double v) {
    if (v != v) return 0xffffffffu;   // pessimistic: generated code must not rely on NaN -> 0
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 4294967295u;
    return (unsigned int)v;
}
static inline unsigned char __ldg(const unsigned char* p) { return *p; }
static inline double __ldg(const double* p) { return *p; }
static inline double __longlong_as_double(long long b) { double d; std::memcpy(&d, &b, 8); return d; }
static inline void __syncthreads() {}
using std::fabs; using std::sin; using std::exp; using std::log; using std::fmax; using std::fmin;
"""

HOST_DRIVER = r"""
extern "C" void host_run(unsigned char* out, double* f64_out, unsigned long long plane, const MrTexture* tex,
                         unsigned int p0, unsigned int n, unsigned int W, double* colv, double* rowv) {
    MrParams p;
    p.out = out; p.f64_out = f64_out; p.f64_plane = plane; p.tex = tex; p.p0 = p0; p.n = n; p.W = W;
    p.out_aligned = 0;
    const unsigned int y_first = p0 / W, rows = (p0 + n - 1) / W - y_first + 1;
    p.colv = colv; p.rowv = rowv; p.row_base = y_first; p.rows = rows;
    blockDim.x = 256; blockDim.y = blockDim.z = 1;
#ifdef MR_HOST_HAS_PROLOGUE
    // the two prologue kernels, one "thread" per column / row
    blockIdx.x = 0; blockDim.x = 0x7fffffff;
    for (unsigned int t = 0; t < W; t++) { threadIdx.x = t; maray_pre_x(colv, W, 0u, tex); }
    for (unsigned int t = 0; t < rows; t++) { threadIdx.x = t; maray_pre_y(rowv, rows, y_first, tex); }
    blockDim.x = 256;
#endif
    unsigned int blocks = (n + 255) / 256;
    for (unsigned int b = 0; b < blocks; b++) {
        blockIdx.x = b;
        // two passes: after the first every thread's bytes are in the staging tile, so the second
        // pass's cooperative copy-out sees a complete tile (sequential stand-in for __syncthreads)
        for (int pass = 0; pass < 2; pass++)
            for (unsigned int t = 0; t < 256; t++) { threadIdx.x = t; maray_jit(p); }
    }
}
"""


class _Tex(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("w", ctypes.c_uint32), ("h", ctypes.c_uint32)]


_CACHE = {}


def host_jit_run(source: str, w: int, p0: int, n: int, textures=()):
    """Runs the generated kernel text on the CPU.  Returns (rgb uint8 (n,3), planes float64 (3,n))."""
    key = hashlib.sha256(source.encode()).hexdigest()
    lib = _CACHE.get(key)
    if lib is None:
        d = tempfile.mkdtemp(prefix="maray_hostjit_")
        src = os.path.join(d, "k.cpp")
        with open(src, "w") as f:
            f.write(HOST_SHIM)
            if "maray_pre_x" in source:
                f.write("#define MR_HOST_HAS_PROLOGUE 1\n")
            f.write(source.replace('extern "C" __global__', "static"))
            f.write(HOST_DRIVER)
        so = os.path.join(d, "k.so")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-w", "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.host_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p,
                                 ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        _CACHE[key] = lib
    rgb = np.zeros((n, 3), dtype=np.uint8)
    planes = np.zeros((3, n), dtype=np.float64)
    arrs = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
    tab = (_Tex * max(1, len(arrs)))()
    for i, a in enumerate(arrs):
        tab[i].data = a.ctypes.data
        tab[i].w, tab[i].h = a.shape[1], a.shape[0]
    # hoisting tables, generously sized (the host check does not know the table widths)
    rows = (p0 + n - 1) // w - p0 // w + 1
    colv = np.zeros(4096 * w, dtype=np.float64)
    rowv = np.zeros(4096 * rows, dtype=np.float64)
    lib.host_run(rgb.ctypes.data, planes.ctypes.data, n, ctypes.addressof(tab), p0, n, w, colv.ctypes.data, rowv.ctypes.data)
    return rgb, planes


# ---- bytecode (csrc/bytecode.hpp, version 2) -------------------------------------------------------
(BC_END, BC_MOV, BC_ADD, BC_MUL, BC_MAX, BC_MIN, BC_NEG, BC_ABS, BC_RECIP, BC_SQRT, BC_STEP, BC_SIN, BC_EXP, BC_LN,
 BC_TEX, BC_OUT_R, BC_OUT_G, BC_OUT_B) = range(18)
F_STORE, F_ACC_A, F_SWAP, F_B_CONST, F_FWD_B = 1, 2, 4, 8, 16
F_A_UNI, F_B_UNI, F_ST_UNI = 32, 64, 128


def _vec(fn):
    def safe(v):
        try:
            return fn(v)
        except (ValueError, OverflowError):
            if fn is math.log:
                return -math.inf if v == 0 else math.nan
            if fn is math.exp:
                return math.inf
            return math.nan
    uf = np.frompyfunc(safe, 1, 1)
    return lambda a: uf(a).astype(np.float64)


_sin, _exp, _log = _vec(math.sin), _vec(math.exp), _vec(math.log)      # the host libm, like the oracle


def sem_max(a, b):
    """f64::max as the reference compiles it: NaN ignored, tie returns the first operand."""
    return np.where((b > a) | np.isnan(a), b, a)


def sem_min(a, b):
    return np.where((b < a) | np.isnan(a), b, a)


def as_u8(v):
    v = np.where(np.isnan(v), 0.0, v)
    return np.clip(np.trunc(v), 0, 255).astype(np.uint8)


def _as_u32(v):
    v = np.where(np.isnan(v), 0.0, v)
    return np.clip(np.trunc(v), 0, 4294967295).astype(np.uint64)


def tex_fetch(tex, ch, x, y):
    h, w = tex.shape[0], tex.shape[1]
    xi, yi = _as_u32(x), _as_u32(y)
    ok = ~((x < 0.0) | (y < 0.0)) & (xi < w) & (yi < h)
    out = np.zeros(x.shape, dtype=np.float64)
    out[ok] = tex[yi[ok], xi[ok], ch].astype(np.float64)
    return out


def bytecode_run(code, consts, xs, ys, textures=()):
    """Executes the bytecode for the pixels (xs[i], ys[i]) the way the kernel does, INCLUDING its
    one-instruction-ahead operand fetch: operands of instruction i+1 are read before instruction i
    stores, so a program that forgot an ACC_A / FWD_B mark computes a wrong value here too.
    Returns planes float64 (3, n)."""
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    n = xs.shape[0]
    slots = {0: xs, 1: ys}
    # Row-uniform form (bytecode.hpp BC_F_*_UNI): one word per BLOCK.  The kernel's blocks lie inside one
    # row, so the pixels given here must share one y; the uniform word is read from / written by pixel 0
    # and asserted to be the same for every pixel -- a value wrongly classified as row-uniform fails here.
    uslots = {}
    zero = np.zeros(n)
    acc = np.zeros(n)
    out = np.zeros((3, n))

    def fetch(w):
        fl, a, b = (w >> 8) & 0xFF, (w >> 32) & 0xFFFF, (w >> 48) & 0xFFFF
        fa = np.full(n, uslots.get(a, 0.0)) if fl & F_A_UNI else slots.get(a, zero)
        if fl & F_B_CONST:
            fb = np.full(n, consts[b])
        else:
            fb = np.full(n, uslots.get(b, 0.0)) if fl & F_B_UNI else slots.get(b, zero)
        return fa, fb

    words = [int(w) for w in code]
    va, vb = fetch(words[0])
    with np.errstate(all="ignore"):
        for i, w in enumerate(words):
            nxt = fetch(words[i + 1]) if i + 1 < len(words) else (zero, zero)     # before this instruction's store
            op, fl, dst = w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFFFF
            if op == BC_END:
                break
            f = acc if fl & F_ACC_A else va
            s = acc if fl & F_FWD_B else vb
            x, y = (s, f) if fl & F_SWAP else (f, s)
            if op == BC_MOV: acc = x
            elif op == BC_ADD: acc = x + y
            elif op == BC_MUL: acc = x * y
            elif op == BC_MAX: acc = sem_max(x, y)
            elif op == BC_MIN: acc = sem_min(x, y)
            elif op == BC_NEG: acc = -x
            elif op == BC_ABS: acc = np.abs(x)
            elif op == BC_RECIP: acc = 1.0 / x
            elif op == BC_SQRT: acc = np.sqrt(x)
            elif op == BC_STEP: acc = np.where(x >= 0.0, 1.0, 0.0)
            elif op == BC_SIN: acc = _sin(x)
            elif op == BC_EXP: acc = _exp(x)
            elif op == BC_LN: acc = _log(x)
            elif op == BC_TEX: acc = tex_fetch(textures[dst >> 2], dst & 3, x, y)
            elif op == BC_OUT_R: out[0] = x
            elif op == BC_OUT_G: out[1] = x
            elif op == BC_OUT_B: out[2] = x
            else:
                raise ValueError(f"bad opcode {op}")
            if fl & F_STORE:
                assert op != BC_TEX, "TEX keeps its texture id in the dst field and must not store"
                if fl & F_ST_UNI:
                    assert np.unique(ys).size == 1, "row-uniform bytecode must be run one row at a time"
                    assert bits_equal(acc, np.full(n, acc[0])).all(), "a row-uniform slot got a value that varies along the row"
                    assert dst not in uslots, "row-uniform slots are never recycled"
                    uslots[dst] = float(acc[0])
                else:
                    slots[dst] = acc
            va, vb = nxt
    return out


def bits_equal(a, b):
    """Bit-identical float64 arrays, any NaN matching any NaN."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
