"""Row-band partition of an image over ranks / GPUs, and the band gather for one-process-per-GPU hosts.

The reference schedules independent rows (rayon over 0..h, reference src/render.rs:85; a shared row
counter, reference src/render.rs:150-173).  Here rank r owns the contiguous rows
[band(h, n, r)), renders them on its own GPU, and the bands are gathered on rank 0 -- the only
exchange step of the path.  `torch.distributed` is plumbing: NCCL on GPUs, gloo in the CPU tests.
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
