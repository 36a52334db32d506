"""Fills the cubin cache (MARAY_JIT_CACHE, default <repo>/.jitcache) with the large scenes the GPU tests and
bench.py compile, on a machine WITHOUT a GPU (NVRTC needs no device).  The cache directory travels to the
GPU box with the snapshot, so GPU minutes are spent on rendering, not on NVRTC.

usage: python tools/warm_cache.py [names...]      names: chess_1k chess_4k chess_dsl sdf textured deep
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MARAY_JIT_CACHE", os.path.join(ROOT, ".jitcache"))
from maray_b200 import CudaRenderer, scenes  # noqa: E402

names = sys.argv[1:] or ["chess_1k", "chess_4k", "chess_dsl", "sdf", "textured", "deep"]
for name in names:
    if name == "chess_dsl":
        scene, tex = scenes.chess_dsl(3840, 2160), []
    else:
        scene, tex, _ = scenes.by_name(name)
    with CudaRenderer(gpus=0) as r:
        r.set_textures(tex)
        r.load(scene)
        t0 = time.time()
        st = r.compile("nvrtc")
        print(f"{name}: {st['dag_nodes']} values, {st['jit_segments']} segments, {st['jit_units']} units, "
              f"cache_hit={st['jit_cache_hit']}, {time.time() - t0:.1f} s", flush=True)
