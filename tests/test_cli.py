"""The command-line front end (maray_b200/maray_cuda), shaped like the reference's examples/maray.rs:
`-i scene.maray -o out.png [-t textures...]`.  PNG input/output goes through the built-in codec."""
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from maray_b200 import expr as E
from maray_b200 import scenes
from oracle.oracle import OracleScene

from conftest import ROOT

CLI = os.path.join(ROOT, "maray_b200", "maray_cuda")


def test_cli_usage_and_no_cpu_fallback(tmp_path):
    assert os.path.exists(CLI), "build with make -C maray_b200/csrc"
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: maray_cuda -i" in r.stderr
    scene = tmp_path / "s.maray"
    scene.write_bytes(scenes.sdf(32, 16, 3))
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([CLI, "-c", "8", "-i", str(scene), "-o", str(tmp_path / "o.png")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "o.png").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("backend", ["nvrtc", "interp"])
def test_cli_renders_textured_scene_like_the_oracle(tmp_path, backend):
    """Textures in four PNG flavours (RGB, RGBA, palette, 16-bit grey) must load as image::open().to_rgb8()
    would, and the written PNG must decode to the oracle's bytes."""
    rgb = scenes.synthetic_textures(1, 32)[0]
    files, arrays = [], []
    for i, mode in enumerate(["RGB", "RGBA", "P", "I;16"]):
        path = str(tmp_path / f"t{i}.png")
        if mode == "RGB":
            Image.fromarray(rgb).save(path); arr = rgb
        elif mode == "RGBA":
            rgba = np.dstack([rgb[::-1], np.full(rgb.shape[:2], 77, np.uint8)])
            Image.fromarray(rgba, "RGBA").save(path); arr = rgb[::-1]
        elif mode == "P":
            pal = Image.fromarray(rgb).quantize(16)
            pal.save(path); arr = np.array(pal.convert("RGB"))
        else:
            g16 = (rgb[:, :, 0].astype(np.uint16) << 8) | 0x5A
            Image.fromarray(g16).save(path)
            # image 0.25.1 `to_rgb8()` rounds 16-bit samples: (v + 128) / 257 (examples/maray.rs:61)
            arr = np.repeat(((g16.astype(np.uint32) + 128) // 257).astype(np.uint8)[:, :, None], 3, axis=2)
        files.append(path); arrays.append(arr)
    x, y = E.x(), E.y()
    color = [E.app(E.channel(0, 0), x, y), E.add(E.app(E.channel(1, 1), x, y), E.app(E.channel(2, 2), y, x)),
             E.app(E.channel(3, 0), E.div(x, E.nat(2)), E.div(y, E.nat(2)))]
    scene_bytes = E.to_bytes([48, 40], color)
    scene = tmp_path / "s.maray"
    scene.write_bytes(scene_bytes)
    out = str(tmp_path / "out.png")
    r = subprocess.run([CLI, "-c", "8", "-i", str(scene), "-o", out, "-b", backend, "-t"] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.array(Image.open(out).convert("RGB"))
    want = OracleScene(scene_bytes, arrays).render()
    assert np.array_equal(got, want)


def test_png_codec_reads_what_pillow_writes(tmp_path):
    """The CLI's built-in PNG codec (csrc/png.cpp) against Pillow, on the CPU: RGB, RGBA, 8/4/2/1-bit
    palette, 8/4/2/1-bit and 16-bit grey must decode to what `convert("RGB")` gives (16-bit: rounded as
    the reference's image 0.25.1 `to_rgb8()` does, (v + 128) / 257 -- 0x01FF is 2, not 1), and what the
    writer writes must read back unchanged.  A header that declares an absurd size is an error, not a crash."""
    csrc = os.path.join(ROOT, "maray_b200", "csrc")
    exe = str(tmp_path / "pngtool")
    (tmp_path / "pngtool.cpp").write_text(r'''
#include "png.hpp"
#include <cstdio>
int main(int argc, char** argv) {
    std::vector<uint8_t> rgb; uint32_t w = 0, h = 0; std::string err;
    if (!maray::read_png_rgb8(argv[1], &w, &h, &rgb, &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (argc > 2 && !maray::write_png_rgb8(argv[2], w, h, rgb.data(), &err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::fwrite(rgb.data(), 1, rgb.size(), stdout);
    return 0;
}''')
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", csrc, "-o", exe, str(tmp_path / "pngtool.cpp"),
                           os.path.join(csrc, "png.cpp"), "-lz"])
    rgb = scenes.synthetic_textures(1, 37)[0][:29]          # odd sizes: rows do not end on byte boundaries
    grey = rgb[:, :, 0]
    cases = {"rgb": Image.fromarray(rgb), "rgba": Image.fromarray(np.dstack([rgb, grey]), "RGBA"),
             "grey8": Image.fromarray(grey), "grey16": Image.fromarray((grey.astype(np.uint16) << 8) | 0xFF),
             "bilevel": Image.fromarray(grey > 100)}
    for colours in (256, 16, 4, 2):
        cases[f"pal{colours}"] = Image.fromarray(rgb).quantize(colours)
    for name, img in cases.items():
        path = str(tmp_path / f"{name}.png")
        bits = {"pal16": 4, "pal4": 2, "pal2": 1}.get(name)
        img.save(path, **({"bits": bits} if bits else {}))
        back = str(tmp_path / f"{name}_back.png")
        r = subprocess.run([exe, path, back], capture_output=True)
        assert r.returncode == 0, (name, r.stderr)
        got = np.frombuffer(r.stdout, np.uint8).reshape(29, 37, 3)
        if name == "grey16":
            v16 = (grey.astype(np.uint32) << 8) | 0xFF
            want = np.repeat(((v16 + 128) // 257).astype(np.uint8)[:, :, None], 3, axis=2)
            assert (want[:, :, 0] != grey).any()         # rounding, not the high byte
        else:
            want = np.array(img.convert("RGB"))
        assert np.array_equal(got, want), name
        assert np.array_equal(np.array(Image.open(back).convert("RGB")), want), name
    # hostile IHDR: 2^31 x 2^31 pixels declared over a few bytes of data
    import struct, zlib
    def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    evil = tmp_path / "evil.png"
    evil.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 31, 1 << 31, 8, 2, 0, 0, 0)) +
                     chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b""))
    r = subprocess.run([exe, str(evil)], capture_output=True)
    assert r.returncode == 1 and b"too large" in r.stderr
    evil.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 20000, 15000, 8, 2, 0, 0, 0)) +
                     chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b""))
    r = subprocess.run([exe, str(evil)], capture_output=True)
    assert r.returncode == 1 and b"corrupt" in r.stderr
