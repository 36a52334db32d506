"""Compile one workload under several settings of the NVRTC back end's knobs and, when a GPU is
present, time the kernel of each and check that every variant renders the same bytes as the first.

usage: jit_variants.py WORKLOAD[:WxH] "K=V,K=V;K=V;..." [repeats]
  e.g. MARAY_DEEP_VALUES=20000 jit_variants.py deep:1024x1024 "MARAY_JIT_PARALLEL=0;;MARAY_JIT_SEGMENT_VALUES=1024"

With MARAY_JIT_CACHE set, a run on a machine without a GPU fills the cache (NVRTC needs no device), and
the same command on the GPU box then only loads cubins -- compile seconds are not paid in GPU time.
"""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maray_b200 import CudaRenderer, scenes  # noqa: E402
from maray_b200.roofline import fp64_ops_per_pixel  # noqa: E402

spec = sys.argv[1]
variants = sys.argv[2].split(";") if len(sys.argv) > 2 else [""]
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 3
name, _, size = spec.partition(":")
scene, tex, (w, h) = scenes.by_name(name)
if size:
    w, h = (int(v) for v in size.split("x"))

try:
    import torch
    have_gpu = torch.cuda.is_available()
except Exception:
    have_gpu = False

peak = None
first = None
base_env = dict(os.environ)
for var in variants:
    os.environ.clear()
    os.environ.update(base_env)
    for kv in filter(None, var.split(",")):
        k, _, v = kv.partition("=")
        os.environ[k] = v
    with CudaRenderer(gpus=1 if have_gpu else 0) as r:
        r.set_textures(tex)
        r.load(scene)
        t0 = time.time()
        st = r.compile("nvrtc")
        rec = {"variant": var or "(default)", "compile_s": round(time.time() - t0, 2), "cache_hit": st["jit_cache_hit"],
               "units": st["jit_units"], "threads": st["jit_compile_threads"],
               "segments": st["jit_segments"], "frame_slots": st["jit_frame_slots"], "regs": st["jit_registers"],
               "cubin_mb": round(st["jit_cubin_bytes"] / 1e6, 2)}
        if have_gpu:
            if peak is None:
                peak = r.fp64_peak(0)[0]
            ops = fp64_ops_per_pixel(st)
            best = 1e30
            for _ in range(repeats):
                r.render_device(w, h)
                best = min(best, r.stats()["kernel_ms"][0])
            frame = r.render(w, h)
            digest = hashlib.sha256(frame.tobytes()).hexdigest()[:16]
            if base_env.get("JIT_VARIANTS_DUMP"):      # debugging: keep the frames
                import numpy as _np
                _np.save(os.path.join(base_env["JIT_VARIANTS_DUMP"], "frame_%d.npy" % variants.index(var)), frame)
            if first is None:
                first = digest
            rec.update({"kernel_ms": round(best, 3), "mpix_s": round(w * h / best / 1e3, 2),
                        "frac": round(w * h * ops / (best * 1e-3) / peak, 3), "rgb_sha": digest, "same_as_first": digest == first})
        print(json.dumps(rec), flush=True)
