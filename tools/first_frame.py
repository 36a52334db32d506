"""Time to first frame, cold: no cubin in the cache, a fresh handle -- load + compile + render + device->host, wall
clock, for one workload under each back end.  "auto" is what a one-shot caller (the reference's `gen`,
src/lib.rs:1199-1213) gets: the interpreter renders while NVRTC compiles on another thread.

usage: first_frame.py WORKLOAD[:WxH] [backends, comma separated: auto,nvrtc,interp]
"""
import hashlib
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from maray_b200 import CudaRenderer, scenes  # noqa: E402

spec = sys.argv[1]
backends = (sys.argv[2] if len(sys.argv) > 2 else "auto,nvrtc,interp").split(",")
name, _, size = spec.partition(":")
scene, tex, (w, h) = scenes.by_name(name)
if size:
    w, h = (int(v) for v in size.split("x"))

with CudaRenderer(gpus=1) as r:          # CUDA context, module loading etc. are paid once per process, not per scene
    r.load(scenes.sdf(64, 64, 2))
    r.compile("interp")
    r.render(64, 64)

img = np.zeros((h, w, 3), dtype=np.uint8)
for backend in backends:
    os.environ["MARAY_JIT_CACHE"] = tempfile.mkdtemp(prefix="maray_cold_")     # empty: nothing is cached
    t0 = time.perf_counter()
    r = CudaRenderer(gpus=1)
    r.set_textures(tex)
    r.load(scene)
    t1 = time.perf_counter()
    st = r.compile(backend)
    t2 = time.perf_counter()
    rs = r.render_into(img)
    t3 = time.perf_counter()
    after = r.stats()
    rec = {"workload": spec, "backend": backend, "first_frame_s": round(t3 - t0, 4), "load_s": round(t1 - t0, 4),
           "compile_call_s": round(t2 - t1, 4), "render_call_s": round(t3 - t2, 4), "lower_ms": round(st["lower_ms"], 1), "codegen_ms": round(st["codegen_ms"], 1), "nvrtc_ms": round(st["nvrtc_ms"], 1),
           "cache_hit": st["jit_cache_hit"], "compile_threads": st["jit_compile_threads"], "cache_dir": os.environ["MARAY_JIT_CACHE"],
           "rows_by_interpreter": after["tier_rows_interp"], "jit_active_after": after["jit_active"],
           "rgb_sha": hashlib.sha256(img.tobytes()).hexdigest()[:16]}
    t4 = time.perf_counter()
    r.close()                                 # auto: waits for the background build (its cubins go to the cache)
    rec["close_s"] = round(time.perf_counter() - t4, 3)
    print(json.dumps(rec), flush=True)
