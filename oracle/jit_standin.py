"""CPU baseline no. 2, a stand-in for the reference's WASM JIT (test/benchmark infrastructure only).

The reference's fastest CPU path compiles every channel to WebAssembly and runs it through
wasmer/cranelift, rows pulled by worker threads (`wasm_par_gen_to_image`, reference src/render.rs:102-192;
`Wasm::from_expr`, reference src/wasm.rs:136-158).  wasmer cannot be had here (no Rust toolchain, no
network), so the closest runnable thing is: the SAME straight-line program the NVRTC back
end generates (one statement per value, common sub-expressions shared across the three channels --
already more than the WASM path shares), compiled for the host by g++ -O2 -ffp-contract=off and run by one
thread per core, each pulling rows -- the structure of `wasm_par_gen_to_image`.  It is an upper bound on
what that JIT could do on these cores, reported next to the interpreter port (`cpu_render_sample`) and
labelled as a stand-in.  sin/exp/ln call glibc, like the reference's host imports.

Nothing in the product imports this module.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import tempfile
import time
from typing import Sequence

import numpy as np

_SHIM = r"""
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static thread_local
#define __constant__ static
#define __launch_bounds__(...)
#define MR_LIBM_PLAIN 1
#define MR_PLAIN_FN static inline
#define MR_HOST_TEXT 1
#define MR_DYN_DECL static thread_local double mr_dyn_f64[32 * 256];
struct uint3_ { unsigned int x, y, z; };
static thread_local uint3_ threadIdx, blockIdx, blockDim;
struct uint4 { unsigned int x, y, z, w; };
static inline double __drcp_rn(double v) { return 1.0 / v; }
static inline double __dsqrt_rn(double v) { return std::sqrt(v); }
static inline unsigned int __double2uint_rz(double v) {
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

# Between the device-side text and the kernel: pixels are written straight to the row buffer (no staging
# tile, no cooperative copy-out -- one pass per pixel).
_STORE = r"""
static inline void mr_host_store(const MrParams& p, unsigned int*, double r, double g, double b, bool active, unsigned int j) {
    if (!active) return;
    p.out[3u * j + 0u] = (unsigned char)mr_as_u8(r);
    p.out[3u * j + 1u] = (unsigned char)mr_as_u8(g);
    p.out[3u * j + 2u] = (unsigned char)mr_as_u8(b);
}
#define mr_store_block mr_host_store
"""

_DRIVER = r"""
#undef mr_store_block
// Worker threads pull rows (reference src/render.rs:168-173); a row is rendered 256 pixels at a time
// through the generated function, one "GPU thread" after the other.
extern "C" double standin_render_rows(unsigned char* out, const MrTexture* tex, unsigned int W, const unsigned int* rows,
                                      unsigned int n_rows, unsigned int x_count, unsigned int n_threads) {
    std::atomic<unsigned int> next{0};
    auto work = [&]() {
        blockDim.x = 256; blockDim.y = blockDim.z = 1;
        for (unsigned int i = next.fetch_add(1); i < n_rows; i = next.fetch_add(1)) {
            MrParams p;
            std::memset(&p, 0, sizeof p);
            p.out = out + (size_t)i * x_count * 3; p.tex = tex; p.W = W;
            p.p0 = rows[i] * W; p.n = x_count;
            for (unsigned int b = 0; b * 256 < x_count; b++) {
                blockIdx.x = b;
                for (unsigned int t = 0; t < 256; t++) { threadIdx.x = t; maray_jit(p); }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned int t = 1; t < n_threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return 0.0;
}
"""


class _Tex(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("w", ctypes.c_uint32), ("h", ctypes.c_uint32)]


class JitStandIn:
    """The scene's generated straight-line program, built for the host.  `compile_s` is the g++ time
    (the counterpart of the reference's per-thread wasmer compile, reference src/render.rs:158-165)."""

    def __init__(self, scene_bytes: bytes, textures: Sequence[np.ndarray] = (), opt: str = "-O2"):
        from maray_b200 import CudaRenderer      # the code generator is the product's; no GPU is touched

        saved = {k: os.environ.get(k) for k in ("MARAY_JIT_INLINE_TRANS_BELOW", "MARAY_JIT_SEGMENT_VALUES")}
        os.environ["MARAY_JIT_INLINE_TRANS_BELOW"] = "4000000000"      # plain calls of the host libm
        os.environ["MARAY_JIT_SEGMENT_VALUES"] = "4096"                # keeps g++ linear on huge programs
        try:
            with CudaRenderer(gpus=0) as r:
                r.set_textures(list(textures))
                r.load(scene_bytes)
                self.size = r.size
                os.environ["MARAY_JIT_SOURCE_ONLY"] = "1"
                try:
                    r.compile("nvrtc")
                except Exception:
                    pass                                              # SOURCE_ONLY reports through the error path
                finally:
                    del os.environ["MARAY_JIT_SOURCE_ONLY"]
                source = r.source()
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        cut = source.index('extern "C" __global__')
        first_seg = source.find("__device__ __noinline__ void mr_seg0")
        if first_seg != -1:
            cut = first_seg
        text = _SHIM + source[:cut] + _STORE + source[cut:].replace('extern "C" __global__', "static") + _DRIVER
        self._dir = tempfile.mkdtemp(prefix="maray_standin_")
        src = os.path.join(self._dir, hashlib.sha256(text.encode()).hexdigest()[:16] + ".cpp")
        with open(src, "w") as f:
            f.write(text)
        so = src[:-4] + ".so"
        t0 = time.perf_counter()
        subprocess.check_call(["g++", opt, "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-pthread",
                               "-w", "-o", so, src])
        self.compile_s = time.perf_counter() - t0
        self._lib = ctypes.CDLL(so)
        self._lib.standin_render_rows.restype = ctypes.c_double
        self._lib.standin_render_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        self._arrs = [np.ascontiguousarray(t, dtype=np.uint8) for t in textures]
        self._tab = (_Tex * max(1, len(self._arrs)))()
        for i, a in enumerate(self._arrs):
            self._tab[i].data = a.ctypes.data
            self._tab[i].w, self._tab[i].h = a.shape[1], a.shape[0]

    def render_rows(self, rows: Sequence[int], w: int, x_count: int = 0, threads: int = 0) -> np.ndarray:
        """RGB8 (len(rows), x_count, 3): the first x_count pixels (default: all) of the given rows."""
        x_count = x_count or w
        threads = threads or (os.cpu_count() or 1)
        rows_a = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.zeros((len(rows_a), x_count, 3), dtype=np.uint8)
        self._lib.standin_render_rows(out.ctypes.data, ctypes.addressof(self._tab), w, rows_a.ctypes.data, len(rows_a),
                                      x_count, threads)
        return out


def timed_sample(scene_bytes, textures, w, h, target_s, threads):
    """Same sampling scheme as bench.py's interpreter leg: batches of evenly spread rows, one per thread,
    until ~target_s of wall time is used.  Returns (Mpixel/s, seconds, description, compile seconds)."""
    js = JitStandIn(scene_bytes, textures)
    t0 = time.perf_counter()
    js.render_rows([h // 2], w, min(w, 256), threads=1)
    per_px = (time.perf_counter() - t0) / min(w, 256)
    seg = w if per_px * w <= target_s / 4 else max(256, min(w, int(target_s / 4 / per_px) // 256 * 256))
    npx, batches = 0, 0
    t0 = time.perf_counter()
    while True:
        ys = sorted({int(((i + 0.5) / threads + batches * 0.6180339887) % 1.0 * h) for i in range(threads)})
        js.render_rows(ys, w, seg, threads=threads)
        npx += len(ys) * seg
        batches += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or dt + dt / batches > 1.5 * target_s or npx >= w * h:
            break
    what = "full rows" if seg == w else f"{seg}-pixel row segments"
    sample = f"{npx} pixels of the {w}x{h} frame ({batches} batches of {threads} {what} spread over the frame)"
    return npx / dt / 1e6, dt, sample, js.compile_s
