"""Times the interpreter kernel under several launch shapes (MARAY_INTERP_SHAPE=block,pixels_per_thread) and
checks that every shape renders the bytes of the NVRTC back end.

usage: interp_sweep.py WORKLOAD[:WxH] ["B,P;B,P;..."] [repeats]
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maray_b200 import CudaRenderer, scenes  # noqa: E402
from maray_b200.roofline import fp64_ops_per_pixel  # noqa: E402

spec = sys.argv[1]
shapes = sys.argv[2].split(";") if len(sys.argv) > 2 and sys.argv[2] else ["", "64,1", "128,1", "256,1", "64,2", "128,2", "256,2", "64,4", "128,4", "256,4"]
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 3
name, _, size = spec.partition(":")
scene, tex, (w, h) = scenes.by_name(name)
if size:
    w, h = (int(v) for v in size.split("x"))

with CudaRenderer(gpus=1) as r:
    r.set_textures(tex)
    r.load(scene)
    r.compile("nvrtc")
    peak = r.fp64_peak(0)[0]
    want = hashlib.sha256(r.render(w, h).tobytes()).hexdigest()[:16]

for shape in shapes:
    if shape:
        os.environ["MARAY_INTERP_SHAPE"] = shape
    else:
        os.environ.pop("MARAY_INTERP_SHAPE", None)
    with CudaRenderer(gpus=1) as r:
        r.set_textures(tex)
        r.load(scene)
        st = r.compile("interp")
        ops = fp64_ops_per_pixel(st)
        best = 1e30
        for _ in range(repeats):
            r.render_device(w, h)
            best = min(best, r.stats()["kernel_ms"][0])
        got = hashlib.sha256(r.render(w, h).tobytes()).hexdigest()[:16]
    print(json.dumps({"workload": spec, "shape_asked": shape or "(auto)", "block": st["interp_block"], "ppt": st["interp_pixels_per_thread"],
                      "wide_slots": st["interp_slots"], "uniform_slots": st["interp_uniform_slots"], "instructions": st["interp_instructions"],
                      "kernel_ms": round(best, 4), "mpix_s": round(w * h / best / 1e3, 2),
                      "frac": round(w * h * ops / (best * 1e-3) / peak, 4), "same_as_nvrtc": got == want}), flush=True)
