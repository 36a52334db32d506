"""Row-band partition of an image over ranks / GPUs, and how the bands of one-process-per-GPU hosts meet.

The reference schedules independent rows (rayon over 0..h, reference src/render.rs:85; a shared row
counter, reference src/render.rs:150-173).  Here rank r owns the contiguous rows [band(h, n, r)) and
renders them on its own GPU.  Three ways for the bands to become one frame, all used by bench.py:
  * device frame on rank 0, no gather: the band kernels store into rank 0's frame over NVLink
    (CudaRenderer.frame_export / frame_import, CUDA IPC);
  * host frame: every rank copies its band into ONE host frame shared by the ranks (SharedHostFrame:
    POSIX shared memory, pinned by every rank), over its own PCIe link;
  * gather_bands: one `gather` collective to rank 0 (NCCL on GPUs, gloo in the CPU tests).
`torch.distributed` is plumbing throughout.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def band(h: int, n: int, r: int) -> Tuple[int, int]:
    """Rows [y0, y1) of rank r of n.  Same split as the C ABI's in-process multi-GPU path
    (csrc/api.cu render_rows_to_frame): y = h*r/n, so bands differ by at most one row."""
    if not (0 <= r < n):
        raise ValueError("rank out of range")
    return (h * r) // n, (h * (r + 1)) // n


def bands(h: int, n: int) -> List[Tuple[int, int]]:
    return [band(h, n, r) for r in range(n)]


def max_band_rows(h: int, n: int) -> int:
    return max(y1 - y0 for y0, y1 in bands(h, n))


def two_parts_on_a_round(y0: int, y1: int, w: int, round_pixels: int) -> List[Tuple[int, int]]:
    """Cuts the band [y0, y1) of a w-wide frame in two so that the first part can travel to the host while the second
    renders, WITHOUT adding a round of blocks: `round_pixels` (stats["jit_round_pixels"]) is what the resident blocks of
    all SMs cover at once, a band of r rounds pays for ceil(r), and two halves of a 10.95-round band would pay for 6
    each.  The cut therefore lies just under floor(r / 2) whole rounds, on a row boundary; bands of two rounds or
    less (or an unknown round size) are cut in the middle.  Empty parts are dropped."""
    rows = y1 - y0
    ym = y0 + (rows + 1) // 2
    if round_pixels > 0 and rows * w > 2 * round_pixels:
        half_rounds = (rows * w // round_pixels) // 2
        ym = y0 + max(1, half_rounds * round_pixels // w)
    return [(a, b) for a, b in ((y0, ym), (ym, y1)) if b > a]


def gather_bands(band_buf, frame, w: int, h: int, rank: int, world: int, group=None) -> None:
    """Gathers every rank's band into `frame` on rank 0.

    band_buf: this rank's uint8 tensor with max_band_rows(h, world)*w*3 elements, the band in its
              first (y1-y0)*w*3 bytes (padding keeps the collective's pieces equal-sized);
    frame:    on rank 0 a uint8 tensor of h*w*3 elements, elsewhere None.
    One `gather` collective; rank 0 then places the pieces (a no-op copy when h % world == 0, because
    then the gather list aliases the frame itself).
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        y0, y1 = band(h, 1, 0)
        frame[: (y1 - y0) * w * 3].copy_(band_buf[: (y1 - y0) * w * 3])
        return
    piece = max_band_rows(h, world) * w * 3
    if rank == 0:
        if h % world == 0:
            pieces = [frame[r * piece:(r + 1) * piece] for r in range(world)]
            dist.gather(band_buf[:piece], pieces, dst=0, group=group)
        else:
            pieces = [torch.empty(piece, dtype=torch.uint8, device=band_buf.device) for _ in range(world)]
            dist.gather(band_buf[:piece], pieces, dst=0, group=group)
            for r in range(world):
                y0, y1 = band(h, world, r)
                frame[y0 * w * 3: y1 * w * 3].copy_(pieces[r][: (y1 - y0) * w * 3])
    else:
        dist.gather(band_buf[:piece], None, dst=0, group=group)


class SharedHostFrame:
    """One h x w x 3 host image shared by the ranks of a one-process-per-GPU job: rank 0 creates a POSIX shared
    memory segment, the others attach to it by name (passed through `torch.distributed`), and every rank copies
    its own band into its own rows -- so a frame leaves N GPUs over N PCIe links instead of one.  `pin()` page-locks
    the mapping in the calling process (cudaHostRegister) so the copies can be asynchronous."""

    def __init__(self, w: int, h: int, rank: int, world: int, group=None):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.w, self.h, self.rank = w, h, rank
        box = [None]
        if rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=max(1, h * w * 3))
            box[0] = self._shm.name
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if rank != 0:
            self._shm = shared_memory.SharedMemory(name=box[0])
            try:    # attaching registers the segment with this process's resource tracker, which would unlink it at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.flat = np.ndarray((h * w * 3,), dtype=np.uint8, buffer=self._shm.buf)
        self._pinned = False

    @property
    def image(self):
        return self.flat.reshape(self.h, self.w, 3)

    def band_view(self, y0: int, y1: int):
        """This rank's rows as a flat torch uint8 tensor over the shared memory."""
        import torch
        return torch.from_numpy(self.flat)[y0 * self.w * 3: y1 * self.w * 3]

    def pin(self) -> int:
        """cudaHostRegister of the mapping; returns the CUDA error code (0 = pinned)."""
        import torch
        rc = torch.cuda.cudart().cudaHostRegister(self.flat.ctypes.data, self.flat.nbytes, 0)
        rc = int(rc[0]) if isinstance(rc, tuple) else int(rc)
        self._pinned = rc == 0
        return rc

    def close(self) -> None:
        import torch
        if self._pinned:
            torch.cuda.cudart().cudaHostUnregister(self.flat.ctypes.data)
            self._pinned = False
        self.flat = None
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()


def host_barrier(tag: str, world: int) -> None:
    """Barrier through the rendezvous store: the waiting ranks hold no GPU.  (An NCCL barrier parks a spinning
    kernel on every waiting rank's GPU; when rank 0 then renders on ALL GPUs from one process -- the in-process
    multi-GPU measurement of bench.py -- that kernel time-slices against rank 0's work.)"""
    import time
    import torch.distributed as dist

    if world == 1:
        return
    store = dist.distributed_c10d._get_default_store()
    key = f"maray_host_barrier_{tag}"
    store.add(key, 1)
    while int(store.add(key, 0)) < world:
        time.sleep(0.002)
