// Minimal PNG reader/writer for the CLI (the reference uses the `image` crate: image::open(..).to_rgb8()
// for textures, RgbImage::save for the output -- reference examples/maray.rs:59-65, src/lib.rs:1207,1212).
// Reader: non-interlaced PNG, bit depth 8 or 16, colour types 0/2/3/4/6, converted to RGB8 as
// to_rgb8() does (alpha dropped, grey replicated, 16-bit taken from the high byte).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace maray {
bool write_png_rgb8(const std::string& path, uint32_t w, uint32_t h, const uint8_t* rgb, std::string* err);
bool read_png_rgb8(const std::string& path, uint32_t* w, uint32_t* h, std::vector<uint8_t>* rgb, std::string* err);
}  // namespace maray
