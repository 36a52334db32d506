"""Debug: flaky frame difference seen once in call 15 (deep 20k, chain of 7).  Repeat and locate."""
import os, sys, json, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from maray_b200 import CudaRenderer, scenes
os.environ["MARAY_DEEP_VALUES"] = "20000"
scene, tex, _ = scenes.by_name("deep")
w = h = 1024
base = None
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for seg, mb in (("6144", "2"), ("3072", "2"), ("3072", "3"), ("1536", "2")):
        os.environ["MARAY_JIT_CHAIN_SEGMENT_VALUES"] = seg
        os.environ["MARAY_JIT_MIN_BLOCKS"] = mb
        with CudaRenderer(gpus=1) as r:
            r.load(scene)
            st = r.compile("nvrtc")
            for _ in range(3):
                r.render_device(w, h)
            fs = [r.render(w, h) for _ in range(3)]
            d_ptr = r.render_device(w, h)
            fd = np.empty((h, w, 3), np.uint8)
            r.copy_to_host(d_ptr, fd)
            fs.append(fd)
        if base is None:
            base = fs[0]
        for k, f in enumerate(fs):
            d = (f != base).any(axis=2)
            if d.any():
                ys, xs = np.nonzero(d)
                print(json.dumps({"it": it, "seg": seg, "mb": mb, "which": k, "differ": int(d.sum()), "rows": [int(ys.min()), int(ys.max())],
                                  "cols": [int(xs.min()), int(xs.max())], "first": [(int(y), int(x)) for y, x in list(zip(ys, xs))[:8]],
                                  "maxdiff": int(np.abs(f.astype(int) - base.astype(int)).max())}), flush=True)
    print("iteration", it, "done", flush=True)
