"""The N>1 path on CPU: row-band partition and the band gather (maray_b200/bands.py) with
torch.distributed's gloo backend at world_size 2 and 3.  The bands are produced by the oracle here
(no GPU); on the GPU box the same gather runs over NCCL with bands from the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maray_b200 import bands, scenes


def test_band_partition_properties():
    for h in (1, 7, 77, 1080, 2160, 8192):
        for n in (1, 2, 3, 4, 8):
            bs = bands.bands(h, n)
            assert bs[0][0] == 0 and bs[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(bs, bs[1:]))             # contiguous, no overlap
            sizes = [y1 - y0 for y0, y1 in bs]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == bands.max_band_rows(h, n)
    with pytest.raises(ValueError):
        bands.band(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, scene, out_path):
    from oracle.oracle import OracleScene

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    y0, y1 = bands.band(h, world, rank)
    piece = bands.max_band_rows(h, world) * w * 3
    band_buf = torch.zeros(piece, dtype=torch.uint8)
    mine = OracleScene(scene).render_window(0, w, y0, y1, threads=1)
    band_buf[: (y1 - y0) * w * 3] = torch.from_numpy(mine.reshape(-1))
    frame = torch.zeros(h * w * 3, dtype=torch.uint8) if rank == 0 else None
    bands.gather_bands(band_buf, frame, w, h, rank, world)
    if rank == 0:
        np.save(out_path, frame.numpy().reshape(h, w, 3))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 40), (2, 41), (3, 40)])
def test_gather_bands_gloo(world, h, tmp_path):
    from oracle.oracle import OracleScene

    w = 48
    scene = scenes.sdf(w, h, 5, seed=4)
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, scene, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), OracleScene(scene).render(w, h))


def _shared_frame_worker(rank, world, port, w, h, scene, out_path):
    from oracle.oracle import OracleScene

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frame = bands.SharedHostFrame(w, h, rank, world)
    y0, y1 = bands.band(h, world, rank)
    mine = OracleScene(scene).render_window(0, w, y0, y1, threads=1)
    frame.band_view(y0, y1).copy_(torch.from_numpy(mine.reshape(-1)))      # every rank writes its own rows
    bands.host_barrier("frame_written", world)                             # store-based: no collective, no device
    if rank == 0:
        np.save(out_path, np.array(frame.image))
    bands.host_barrier("frame_read", world)
    frame.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 41), (3, 40)])
def test_shared_host_frame_gloo(world, h, tmp_path):
    """The end-to-end exchange of bench.py at N > 1: one host frame in POSIX shared memory, every rank fills in
    its own band, a store-based barrier says when it is complete (on the GPU box the rows arrive by
    device->host copies over each GPU's own PCIe link)."""
    from oracle.oracle import OracleScene

    w = 48
    scene = scenes.sdf(w, h, 5, seed=4)
    out = str(tmp_path / "frame.npy")
    mp.spawn(_shared_frame_worker, args=(world, _free_port(), w, h, scene, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), OracleScene(scene).render(w, h))


def test_two_parts_on_a_round():
    """The e2e path of bench.py copies the first part of a band out while the second renders; the cut must not cost
    a round of blocks."""
    import math
    w, rp = 3840, 148 * 640
    for n in (2, 4, 8):
        for r in range(n):
            y0, y1 = bands.band(2160, n, r)
            parts = bands.two_parts_on_a_round(y0, y1, w, rp)
            assert len(parts) == 2 and parts[0][0] == y0 and parts[0][1] == parts[1][0] and parts[1][1] == y1
            whole = math.ceil((y1 - y0) * w / rp)
            assert sum(math.ceil((b - a) * w / rp) for a, b in parts) == whole          # 44, 22, 11 rounds: none added
    # small bands and unknown round sizes: the middle; degenerate bands: no empty part
    assert bands.two_parts_on_a_round(10, 20, 1024, rp) == [(10, 15), (15, 20)]
    assert bands.two_parts_on_a_round(0, 135, 1920, 0) == [(0, 68), (68, 135)]
    assert bands.two_parts_on_a_round(5, 6, 64, rp) == [(5, 6)]
    assert bands.two_parts_on_a_round(5, 5, 64, rp) == []
    # chain programs (256 x 2 shape) on a tall band
    parts = bands.two_parts_on_a_round(0, 1024, 8192, 148 * 2 * 256)
    assert parts[0][1] == 55 * 148 * 2 * 256 // 8192 and parts[1][1] == 1024
