"""Debug: the sequence of call 15 (6144/2, 6144/3, 6144/4, then 3072/2) renders a different frame in the fourth variant."""
import os, sys, json, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from maray_b200 import CudaRenderer, scenes
os.environ["MARAY_DEEP_VALUES"] = "20000"
scene, tex, _ = scenes.by_name("deep")
w = h = 1024
base = None
seq = [("6144", "2"), ("6144", "3"), ("6144", "4"), ("3072", "2"), ("3072", "2")]
if len(sys.argv) > 1: seq = [tuple(s.split("/")) for s in sys.argv[1].split(",")]
for seg, mb in seq:
    os.environ["MARAY_JIT_CHAIN_SEGMENT_VALUES"] = seg
    os.environ["MARAY_JIT_MIN_BLOCKS"] = mb
    with CudaRenderer(gpus=1) as r:
        r.load(scene)
        st = r.compile("nvrtc")
        outs = []
        for k in range(3):
            d_ptr = r.render_device(w, h)
            f = np.empty((h, w, 3), np.uint8); r.copy_to_host(d_ptr, f); outs.append(("dev%d" % k, f))
        outs.append(("host", r.render(w, h)))
        d_ptr = r.render_device(w, h)
        f = np.empty((h, w, 3), np.uint8); r.copy_to_host(d_ptr, f); outs.append(("dev_after", f))
    if base is None: base = outs[0][1]
    for name, f in outs:
        d = (f != base).any(axis=2)
        rec = {"seg": seg, "mb": mb, "which": name, "differ": int(d.sum())}
        if d.any():
            ys, xs = np.nonzero(d)
            rec.update({"rows": [int(ys.min()), int(ys.max())], "cols": [int(xs.min()), int(xs.max())], "n_rows": int(len(set(ys.tolist()))),
                        "first": [(int(y), int(x), f[y, x].tolist(), base[y, x].tolist()) for y, x in list(zip(ys, xs))[:5]],
                        "maxdiff": int(np.abs(f.astype(int) - base.astype(int)).max())})
        print(json.dumps(rec), flush=True)
