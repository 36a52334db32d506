#!/usr/bin/env python
"""Benchmark of the Maray render path on B200: Mpixel/s + FP64-pipe fraction of measured peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--backend nvrtc|interp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on host cores

A "step" renders one whole frame of the workload.  One process per GPU: rank r renders its row band
on its own GPU through the C ABI (maray_cuda_render_band), bands are gathered on rank 0 with one
NCCL gather (the path's only exchange step).  Rank 0 prints ONE JSON line.

  value   whole-frame Mpixel/s, frame left in HBM on rank 0 (device-timed, max over ranks)
  e2e     the same through the reference-facing call with a HOST image buffer: at N=1 the C ABI's
          maray_cuda_render (device->host copy inside the timed region); at N>1 band render +
          gather + rank 0's device->host copy into pinned memory
  roofline  FP64-pipe lane-operations/s achieved (algorithmic ops per pixel from the un-hoisted
          program, maray_b200/roofline.py) against the FP64 issue rate measured on the same GPU
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "chess_4k"
METRIC = "render_throughput"
UNIT = "Mpixel/s"


def committed_dram_traffic(workload: str, backend: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary of this
    workload (profiles/rNN_<workload>_<backend>_ncu_full_summary.txt), or None when there is none."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{workload.replace('_', '')}_{backend}_ncu_full_summary.txt"))):
        total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        found = 0
        for line in open(path):
            m = re.match(r"dram__bytes_(read|write)\.sum \[(\w+)\] = ([0-9.eE+-]+)", line)
            if m:
                total += float(m.group(3)) * unit_scale.get(m.group(2), 1.0)
                found += 1
        if found == 2:
            best = {"bytes": total, "source": os.path.relpath(path, ROOT)}
    return best


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MARAY_BENCH_WORKLOAD", DEFAULT_WORKLOAD),
                    choices=["chess_1k", "sdf", "chess_4k", "textured", "deep"])
    ap.add_argument("--backend", default="nvrtc", choices=["nvrtc", "interp"])
    ap.add_argument("--cpu-sample-s", type=float, default=12.0, help="target seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-jit-standin", action="store_true",
                    help="skip the second CPU baseline (generated straight-line program built with g++, oracle/jit_standin.py)")
    ap.add_argument("--size", default=None, help="WxH override of the workload's frame size (experiments only)")
    return ap.parse_args()


# ---- CPU arm: the reference's algorithm (oracle port) on the host cores ---------------------------
def host_threads() -> int:
    """Threads the CPU arm may use: the affinity mask, capped by a cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def cpu_render_sample(scene_bytes, textures, w, h, target_s, threads, keep=None):
    """Times the oracle (restatement of par_gen_to_image, all host threads) on a bounded sample of the
    workload: batches of evenly spread full rows (one row per thread per batch) until ~target_s of
    wall time is used.  Scenes so expensive that one row would blow the budget are sampled as
    short row segments instead.  Returns (Mpixel/s, seconds, description).  `keep` (a list) receives
    what was rendered as (x0, x1, [rows], rgb array (len(rows), x1-x0, 3)) -- the parity sample."""
    from oracle.oracle import OracleScene

    sc = OracleScene(scene_bytes, textures)
    probe = min(w, 32)
    t0 = time.perf_counter()
    sc.render_window(0, probe, h // 2, h // 2 + 1, threads=1)
    per_px = (time.perf_counter() - t0) / probe            # one thread, one mid-frame pixel
    seg = w if per_px * w <= target_s / 2 else max(1, min(w, int(target_s / 2 / per_px)))
    npx, batches = 0, 0
    t0 = time.perf_counter()
    while True:
        # rows spread over the frame, different every batch
        ys = sorted({int(((i + 0.5) / threads + batches * 0.6180339887) % 1.0 * h) for i in range(threads)})
        if seg == w:
            got = sc.render_rows(ys, w, threads=threads)
        else:
            # `threads` consecutive rows so every thread pulls one row segment
            y0 = min(max(0, ys[len(ys) // 2]), max(0, h - threads))
            got = sc.render_window(0, seg, y0, min(h, y0 + threads), threads=threads)
            ys = list(range(y0, min(h, y0 + threads)))
        if keep is not None:
            keep.append((0, seg, list(ys), got))
        npx += len(ys) * seg
        batches += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or dt + dt / batches > 1.5 * target_s or npx >= w * h:
            break
    sc.close()
    what = "full rows" if seg == w else f"{seg}-pixel row segments"
    sample = f"{npx} pixels of the {w}x{h} frame ({batches} batches of {threads} {what} spread over the frame)"
    return npx / dt / 1e6, dt, sample


def parity_against_oracle(frame, kept, scene_bytes, textures):
    """Compares the oracle pixels the CPU baseline rendered anyway with the same pixels of the GPU frame.
    Bar (BASELINE.json north_star): >= 99.99 % of the sampled pixels identical; a differing channel is within
    1 LSB, or a `step` flip -- attributed by the oracle finding a step argument within 64 ULP of zero."""
    import numpy as np
    from oracle.oracle import OracleScene

    n_px = differ = max_lsb = flips = unexplained = 0
    rows = 0
    sc = None
    for x0, x1, ys, want in kept:
        for i, y in enumerate(ys):
            got = frame[y, x0:x1]
            d = np.abs(got.astype(np.int16) - want[i].astype(np.int16)).max(axis=1)
            n_px += x1 - x0
            rows += 1
            differ += int((d != 0).sum())
            small = d[d <= 1]
            max_lsb = max(max_lsb, int(small.max()) if small.size else 0)
            for x in np.nonzero(d > 1)[0].tolist():
                if sc is None:
                    sc = OracleScene(scene_bytes, textures)
                if sc.step_margin(float(x0 + x), float(y)) <= 64.0:
                    flips += 1
                else:
                    unexplained += 1
    if sc is not None:
        sc.close()
    ok = unexplained == 0 and differ <= max(1, n_px // 10000)
    return {"rows": rows, "pixels": n_px, "differ": differ, "max_lsb": max_lsb, "step_flips": flips,
            "unexplained": unexplained, "ok": ok,
            "bar": ">= 99.99 % identical; others <= 1 LSB or a step flip (oracle: step argument within 64 ULP of 0)"}


def cpu_jit_standin(scene_bytes, textures, w, h, threads, dag_values, target_s=6.0):
    """Second CPU baseline: a stand-in for the reference's WASM JIT -- the generated
    straight-line program compiled for the host with g++ -O2 -ffp-contract=off, rows pulled by `threads`
    workers.  Skipped for programs whose host compile alone would take minutes."""
    if dag_values > 30000:
        return {"skipped": f"{dag_values} values: the g++ build alone would exceed the bench budget"}
    try:
        from oracle.jit_standin import timed_sample
        v, dt, sample, compile_s = timed_sample(scene_bytes, textures, w, h, target_s, threads)
        return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "what": "stand-in for the reference's WASM JIT: the generated straight-line program built with g++ -O2 "
                        "-ffp-contract=off (wasmer is not available), one row-pulling thread per core",
                "sample": sample, "seconds": dt, "compile_s": compile_s}
    except Exception as exc:          # a missing host compiler must not take the bench line down
        return {"skipped": f"{type(exc).__name__}: {exc}"[:200]}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The Rust crate cannot
    be built here (no cargo/rustc), so this times oracle/ -- the C restatement of
    par_gen_to_image -- with all host threads, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from maray_b200 import scenes          # scene generators only: pure Python, loads no native library

    scene_bytes, textures, (w, h) = scenes.by_name(args.workload)
    threads = host_threads()
    per_step_s = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, dt, sample = cpu_render_sample(scene_bytes, textures, w, h, per_step_s, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = sum(dt for _, dt in vals) / len(vals) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "width": w, "height": h},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    # (The JIT stand-in -- the generated straight-line program built with g++ -- needs the product's code
    # generator; it is reported by the GPU arm's cpu_baseline only, so that this arm loads nothing but oracle/.)
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons of one GPU during the timed region, through
    NVML (a few ms per sample; nvidia-smi takes longer than a short timed region lasts)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.ready = threading.Event()
        self._th = None
        self._bits = {}
        self.t0 = self.t1 = None        # the timed region (perf_counter), set by mark()

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            hnd = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(hnd, nv.NVML_CLOCK_SM))
            bits = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            }
            self.ready.set()
            while not self._stop.is_set():
                now = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM))
                mask = 0
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                except Exception:
                    pass
                self.samples.append((now, mhz, mask))
                self._bits = bits
                self._stop.wait(0.002)
        except Exception as exc:   # NVML unavailable: say so instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")
            self.ready.set()

    def start(self):
        """Starts sampling (call before the warm-up so NVML is initialised when the timed region begins)."""
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        self.ready.wait(timeout=10)

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        inside = [s for s in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= s[0] <= self.t1]
        scope = "timed region"
        if not inside and self.samples and self.t0 is not None:
            # region shorter than the sampling period: take the samples that bracket it
            before = [s for s in self.samples if s[0] < self.t0][-1:]
            after = [s for s in self.samples if s[0] > (self.t1 or self.t0)][:1]
            inside = before + after
            scope = "samples bracketing the timed region"
        for _, _, mask in inside:
            for name, bit in self._bits.items():
                if mask & bit:
                    self.reasons.add(name)
        mhz = sorted(s[1] for s in inside)
        med = mhz[len(mhz) // 2] if mhz else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(mhz), "scope": scope}


# ---- GPU arm --------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from maray_b200 import CudaRenderer, bands, scenes
    from maray_b200.roofline import fp64_ops_per_pixel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this render path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene_bytes, textures, (w, h) = scenes.by_name(args.workload)
    if args.size:
        w, h = (int(v) for v in args.size.lower().split("x"))
    r = CudaRenderer(device_ids=[local_rank])
    r.set_textures(textures)
    r.load(scene_bytes)
    t0 = time.perf_counter()
    stats = r.compile(args.backend)
    compile_s = time.perf_counter() - t0
    ops_px = fp64_ops_per_pixel(stats)

    y0, y1 = bands.band(h, world, rank)
    piece = bands.max_band_rows(h, world) * w * 3
    frame = torch.empty(h * w * 3, dtype=torch.uint8, device=dev) if rank == 0 else None
    # at world == 1 the band IS the frame; otherwise a padded band buffer feeds the gather
    band_buf = frame if world == 1 else torch.empty(piece, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        r.render_band(w, h, y0, y1, band_buf.data_ptr(), stream.cuda_stream)
        if world > 1:
            bands.gather_bands(band_buf, frame, w, h, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 issue-rate peak of this GPU (roofline denominator), measured before the timed region.
    peak_nofma, peak_fma = r.fp64_peak(0)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
        flush.zero_()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.mark_start()
    for i in range(args.steps):
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        r.render_band(w, h, y0, y1, band_buf.data_ptr(), stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            bands.gather_bands(band_buf, frame, w, h, rank, world)
        ev[i][1].record(stream)
        flush.zero_()            # L2 flush between steps, outside the event pairs
        if world > 1:
            dist.barrier()       # every step starts together on all ranks
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    value = w * h / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host image buffer, device->host inside the timed region ------------------------
    # The primary figure uses a PAGEABLE buffer: a Rust RgbImage is a plain Vec<u8> (img.as_mut_ptr(),
    # reference src/lib.rs:1210).  A pinned buffer is timed next to it as a note.
    def e2e_with(host):
        host_np = host.numpy() if host is not None else None

        def step_e2e():
            if world == 1:
                r.render_into(host_np)               # the C ABI call a user makes: maray_cuda_render
            else:
                step_device()
                if rank == 0:
                    host.view(-1).copy_(frame, non_blocking=True)
                torch.cuda.synchronize()

        for _ in range(min(args.warmup, 3)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
            if world > 1:
                dist.barrier()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return w * h / float(te[0]) / 1e6

    pageable = torch.zeros((h, w, 3), dtype=torch.uint8) if rank == 0 else None
    e2e_value = e2e_with(pageable)
    e2e_pinned = e2e_with(torch.zeros((h, w, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None)

    if rank == 0:
        # roofline of the dominant kernel (the band kernel): per launch it processes band pixels
        band_px = (y1 - y0) * w
        achieved = band_px * ops_px / (kernel_ms_max * 1e-3) / 1e12
        peak = peak_nofma / 1e12
        traffic = committed_dram_traffic(args.workload, args.backend) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "width": w, "height": h, "backend": args.backend,
                       "parallelism": f"row-bands x{world}, gather to rank 0 (NCCL)" if world > 1 else "1 GPU",
                       "l2": "flushed between steps (256 MiB memset outside the per-step event pairs)",
                       "dag_values": stats["dag_nodes"], "fp64_ops_per_pixel": ops_px},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": w * h * 3,
                    "host_buffer": "pageable (what a Rust Vec<u8>/RgbImage is)", "value_pinned_host_buffer": e2e_pinned,
                    "note": "inputs are pixel coordinates generated on chip; the compiled scene is resident"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "Tlaneop/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": traffic["bytes"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None,
                         "peak_source": "measured live: maray_cuda_fp64_peak DADD/DMUL issue rate (no FMA)",
                         "peak_dfma": peak_fma / 1e12, "kernel_ms": kernel_ms_max,
                         "hbm_bytes_per_launch_algorithmic": band_px * 3},
            "compile": {"backend": args.backend, "lower_ms": stats["lower_ms"], "codegen_ms": stats["codegen_ms"],
                        "nvrtc_ms": stats["nvrtc_ms"], "load_ms": stats["load_ms"], "wall_s": compile_s,
                        "registers": stats["jit_registers"], "segments": stats["jit_segments"],
                        "units": stats["jit_units"], "compile_threads": stats["jit_compile_threads"],
                        "cache_hit": bool(stats["jit_cache_hit"]),
                        "interp_instructions": stats["interp_instructions"], "interp_slots": stats["interp_slots"]},
        }
        if not args.no_cpu_baseline and world == 1:
            threads = host_threads()
            kept = []
            v, dt, sample = cpu_render_sample(scene_bytes, textures, w, h, args.cpu_sample_s, threads, keep=kept)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                    "seconds": dt}
            # the oracle pixels just rendered are the parity sample for the frame the timed region produced
            line["parity"] = parity_against_oracle(frame.cpu().numpy().reshape(h, w, 3), kept, scene_bytes, textures)
            if not args.no_cpu_jit_standin:
                line["cpu_baseline"]["jit_standin"] = cpu_jit_standin(scene_bytes, textures, w, h, threads, stats["dag_nodes"])
        print(json.dumps(line), file=RESULT_OUT, flush=True)
        parity_failed = "parity" in line and not line["parity"]["ok"]
    else:
        parity_failed = False
    r.close()
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        raise SystemExit("bench.py: the GPU frame violates the parity bar against the oracle (see \"parity\" in the line)")


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on stdout when the first communicator is created), so stdout is pointed at stderr for the
    whole run and the result line goes to a private copy of the original descriptor."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    global RESULT_OUT
    RESULT_OUT = _claim_stdout()
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
