"""What a row band costs against its share of the frame (the drain of a short launch), per launch shape.

usage: band_tail.py WORKLOAD "K=V,K=V;K=V;..." [parts=8] [repeats=20]
For every variant: kernel time of the whole frame and of the first 1/parts of its rows (CUDA events on the launching
stream, L2 flushed before every launch), and band / (frame / parts)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maray_b200 import CudaRenderer, scenes  # noqa: E402

name = sys.argv[1]
variants = sys.argv[2].split(";") if len(sys.argv) > 2 else [""]
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 8
repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 20
scene, tex, (w, h) = scenes.by_name(name)
dev = torch.device("cuda:0")
frame = torch.empty(h * w * 3, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def timed(r, y0, y1):
    ts = []
    for _ in range(repeats):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.render_band(w, h, y0, y1, frame.data_ptr() + y0 * w * 3, stream)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


base_env = dict(os.environ)
for var in variants:
    os.environ.clear()
    os.environ.update(base_env)
    for kv in filter(None, var.split(",")):
        k, _, v = kv.partition("=")
        os.environ[k] = v
    with CudaRenderer(device_ids=[0]) as r:
        r.set_textures(tex)
        r.load(scene)
        st = r.compile("nvrtc")
        timed(r, 0, h)
        full = timed(r, 0, h)
        rows = h // parts
        band = timed(r, 0, rows)
        mid = timed(r, h // 2, h // 2 + rows)
        print(json.dumps({"variant": var or "(default)", "regs": st["jit_registers"], "frame_ms": round(full, 4),
                          "band_rows": rows, "band_ms": round(band, 4), "band_mid_ms": round(mid, 4),
                          "band_over_share": round(band / (full / parts), 4)}), flush=True)
