"""Debug: do chain segmentations of the deep scene render the same bytes?  (call 15 saw one variant differ.)"""
import os, sys, json, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from maray_b200 import CudaRenderer, scenes
os.environ["MARAY_DEEP_VALUES"] = "20000"
scene, tex, _ = scenes.by_name("deep")
w = h = 1024
frames = {}
for libm in ("fast", "glibc"):
    for seg in ("6144", "3072", "1536"):
        for rep in range(2):
            os.environ["MARAY_JIT_CHAIN_SEGMENT_VALUES"] = seg
            with CudaRenderer(gpus=1) as r:
                r.load(scene)
                st = r.compile("nvrtc", libm=libm)
                f = r.render(w, h)
            frames[(libm, seg, rep)] = f
            base = frames[(libm, "6144", 0)]
            d = (f != base).any(axis=2)
            ys, xs = np.nonzero(d)
            print(json.dumps({"libm": libm, "seg": seg, "rep": rep, "segments": st["jit_segments"], "regs": st["jit_registers"],
                              "sha": hashlib.sha256(f.tobytes()).hexdigest()[:12], "differ": int(d.sum()),
                              "where": [(int(y), int(x)) for y, x in list(zip(ys, xs))[:6]],
                              "maxdiff": int(np.abs(f.astype(int) - base.astype(int)).max())}), flush=True)
# exact mode against the oracle at the first differing pixels
from oracle.oracle import OracleScene
o = OracleScene(scene)
bad = frames[("glibc", "3072", 0)] != frames[("glibc", "6144", 0)]
ys, xs = np.nonzero(bad.any(axis=2))
for y, x in list(zip(ys, xs))[:3]:
    want_rgb, _ = o.render_window(int(x), int(x) + 1, int(y), int(y) + 1, want_f64=True)
    print("pixel", y, x, "oracle", want_rgb.ravel().tolist(), "seg6144", frames[("glibc", "6144", 0)][y, x].tolist(), "seg3072", frames[("glibc", "3072", 0)][y, x].tolist())
