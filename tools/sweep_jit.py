"""Sweep NVRTC back-end knobs on one workload and print kernel time per configuration (GPU box)."""
import itertools, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maray_b200 import CudaRenderer, scenes
from maray_b200.roofline import fp64_ops_per_pixel

workload = sys.argv[1] if len(sys.argv) > 1 else "chess_4k"
scene, tex, (w, h) = scenes.by_name(workload)
grid = {
    "MARAY_JIT_MAXREG": os.environ.get("SWEEP_MAXREG", "0,128,96,80,64").split(","),
    "MARAY_JIT_SEGMENT_VALUES": os.environ.get("SWEEP_SEG", "4096,100000").split(","),
    "MARAY_JIT_BLOCK": os.environ.get("SWEEP_BLOCK", "256,128").split(","),
}
peak = None
for combo in itertools.product(*grid.values()):
    for k, v in zip(grid.keys(), combo):
        os.environ[k] = v
    with CudaRenderer(gpus=1) as r:
        if peak is None:
            peak = r.fp64_peak(0)[0]
        r.set_textures(tex); r.load(scene)
        st = r.compile("nvrtc")
        ops = fp64_ops_per_pixel(st)
        best = 1e9
        for i in range(4):
            r.render_device(w, h)
            best = min(best, r.stats()["kernel_ms"][0])
        print(json.dumps({"cfg": dict(zip(grid.keys(), combo)), "regs": st["jit_registers"], "nvrtc_ms": round(st["nvrtc_ms"]),
                          "kernel_ms": round(best, 3), "mpix_s": round(w * h / best / 1e3, 1),
                          "frac": round(w * h * ops / (best * 1e-3) / peak, 3)}), flush=True)
