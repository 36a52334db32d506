"""Debug: poison the chain frame before every chunk (MARAY_JIT_FRAME_POISON): does any segmentation read a slot before writing it?"""
import os, sys, json, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from maray_b200 import CudaRenderer, scenes
os.environ["MARAY_DEEP_VALUES"] = "20000"
scene, tex, _ = scenes.by_name("deep")
w = h = 1024
base = None
for poison in (False, True):
    if poison: os.environ["MARAY_JIT_FRAME_POISON"] = "1"
    for seg, mb in (("6144", "2"), ("3072", "2"), ("3072", "3"), ("1536", "2"), ("768", "2")):
        os.environ["MARAY_JIT_CHAIN_SEGMENT_VALUES"] = seg
        os.environ["MARAY_JIT_MIN_BLOCKS"] = mb
        with CudaRenderer(gpus=1) as r:
            r.load(scene)
            st = r.compile("nvrtc")
            f1 = r.render(w, h)
            d_ptr = r.render_device(w, h)
            f2 = np.empty((h, w, 3), np.uint8); r.copy_to_host(d_ptr, f2)
            f3 = r.render(w, h)
        if base is None: base = f1
        print(json.dumps({"poison": poison, "seg": seg, "mb": mb, "differ": [int((f != base).any(axis=2).sum()) for f in (f1, f2, f3)]}), flush=True)
