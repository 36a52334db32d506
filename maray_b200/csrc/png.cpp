#include "png.hpp"

#include <zlib.h>

#include <new>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace maray {
namespace {

void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
uint32_t get32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

void chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t len) {
    put32(out, uint32_t(len));
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    put32(out, uint32_t(crc32(0L, out.data() + start, uInt(len + 4))));
}

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

bool write_png_rgb8(const std::string& path, uint32_t w, uint32_t h, const uint8_t* rgb, std::string* err) {
    std::vector<uint8_t> raw;
    raw.reserve(size_t(h) * (size_t(w) * 3 + 1));
    for (uint32_t y = 0; y < h; y++) {
        raw.push_back(0);   // filter type 0 (None)
        raw.insert(raw.end(), rgb + size_t(y) * w * 3, rgb + size_t(y + 1) * w * 3);
    }
    uLongf clen = compressBound(uLong(raw.size()));
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), uLong(raw.size()), 6) != Z_OK) { if (err) *err = "zlib compress failed"; return false; }
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w); put32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGB
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", comp.data(), clen);
    chunk(out, "IEND", nullptr, 0);
    std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) { if (err) *err = "cannot open " + tmp + " for writing"; return false; }
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (std::fclose(f) == 0) && ok;
    if (ok) ok = std::rename(tmp.c_str(), path.c_str()) == 0;
    if (!ok && err) *err = "write to " + path + " failed";
    return ok;
}

bool read_png_rgb8(const std::string& path, uint32_t* w_out, uint32_t* h_out, std::vector<uint8_t>* rgb, std::string* err) {
    auto fail = [&](const std::string& m) { if (err) *err = path + ": " + m; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) return fail("not a PNG file");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    size_t pos = 8;
    while (pos + 12 <= buf.size()) {
        uint32_t len = get32(&buf[pos]);
        const uint8_t* type = &buf[pos + 4];
        if (pos + 12 + size_t(len) > buf.size()) return fail("truncated chunk");
        const uint8_t* data = &buf[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            w = get32(data); h = get32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + size_t(len);
    }
    if (!w || !h || ctype < 0) return fail("missing IHDR");
    if (interlace) return fail("interlaced PNG is not supported");
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) return fail("unsupported colour type");
    // PNG allows 1/2/4-bit samples for greyscale and palette images only, 16-bit for everything but palette
    const bool sub_byte = depth == 1 || depth == 2 || depth == 4;
    if (!(depth == 8 || (depth == 16 && ctype != 3) || (sub_byte && (ctype == 0 || ctype == 3))))
        return fail("unsupported bit depth for this colour type");
    // bytes per complete pixel as the filters see it (1 for sub-byte samples), bytes per row
    const size_t bpp = sub_byte ? 1 : size_t(channels) * (depth / 8);
    const size_t stride = sub_byte ? (size_t(w) * depth + 7) / 8 : size_t(w) * bpp;
    // A hostile IHDR must not drive the allocations: the inflated size is bounded (1 GiB of samples, which
    // also keeps every size below in range) and cannot exceed what deflate can expand the IDAT bytes to.
    const uint64_t raw_bytes = (uint64_t(stride) + 1) * h;
    if (raw_bytes > (uint64_t(1) << 30)) return fail("image too large (more than 1 GiB of samples)");
    if (raw_bytes > uint64_t(idat.size()) * 1032 + 1024) return fail("corrupt image data (IDAT too short for the declared size)");
    std::vector<uint8_t> raw;
    try {
        raw.resize(size_t(raw_bytes));
        rgb->assign(size_t(w) * h * 3, 0);
    } catch (const std::bad_alloc&) {
        return fail("out of memory");
    }
    uLongf rawlen = uLongf(raw.size());
    if (uncompress(raw.data(), &rawlen, idat.data(), uLong(idat.size())) != Z_OK || rawlen != raw.size()) return fail("corrupt image data");
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* line = &raw[(stride + 1) * y];
        int ft = line[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0, x = line[1 + i];
            switch (ft) {
            case 0: break;
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) / 2; break;
            case 4: x += paeth(a, b, c); break;
            default: return fail("bad filter type");
            }
            cur[i] = uint8_t(x);
        }
        uint8_t* dst = &(*rgb)[size_t(y) * w * 3];
        const size_t sample = depth / 8;
        // 16-bit samples (big-endian) are reduced the way the reference's image 0.25.1 `to_rgb8()` does it
        // (examples/maray.rs:61): round(v / 257) = (v + 128) / 257, not the high byte.
        auto s8 = [&](const uint8_t* q) -> uint8_t {
            return sample == 2 ? uint8_t(((unsigned(q[0]) << 8 | q[1]) + 128u) / 257u) : q[0];
        };
        for (uint32_t xx = 0; xx < w; xx++) {
            uint8_t packed;
            const uint8_t* px = &cur[xx * bpp];
            if (sub_byte) {
                // samples are packed most significant first; greyscale is scaled to 0..255 (v * 255 / max)
                const unsigned bit = unsigned(xx) * unsigned(depth);
                const unsigned v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
                packed = ctype == 0 ? uint8_t(v * 255u / ((1u << depth) - 1u)) : uint8_t(v);
                px = &packed;
            }
            switch (ctype) {
            case 0: case 4: dst[3 * xx] = dst[3 * xx + 1] = dst[3 * xx + 2] = s8(px); break;
            case 2: case 6: dst[3 * xx] = s8(px); dst[3 * xx + 1] = s8(px + sample); dst[3 * xx + 2] = s8(px + 2 * sample); break;
            case 3: {
                size_t idx = px[0];
                if (3 * idx + 2 >= plte.size()) return fail("palette index out of range");
                dst[3 * xx] = plte[3 * idx]; dst[3 * xx + 1] = plte[3 * idx + 1]; dst[3 * xx + 2] = plte[3 * idx + 2];
            } break;
            }
        }
        prev.swap(cur);
    }
    *w_out = w; *h_out = h;
    return true;
}

}  // namespace maray
