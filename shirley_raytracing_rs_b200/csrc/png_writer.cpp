// png_writer.cpp — b200rt_write_png: 8-bit RGB PNG through zlib.
// Replaces `dst.save_with_format(path, ImageFormat::Png)` (src/raytracer/image.rs:42, crate
// `image` -> `png`).  PNG is lossless, so any conforming encoder yields the same pixels.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/b200rt.h"

namespace {
void put_u32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    put_u32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4));
    put_u32(out, crc);
}
}  // namespace

// Encodes into memory; exposed for the host layer and the tests.
extern "C" int b200rt_encode_png(const uint8_t* rgb8, uint32_t width, uint32_t height, uint8_t** out_bytes, size_t* out_len) {
    if (!rgb8 || !out_bytes || !out_len || width == 0 || height == 0) return B200RT_EINVAL;
    std::vector<uint8_t> raw((size_t)height * ((size_t)width * 3 + 1));
    for (uint32_t y = 0; y < height; ++y) {
        uint8_t* row = raw.data() + (size_t)y * ((size_t)width * 3 + 1);
        row[0] = 0;   // filter: None
        memcpy(row + 1, rgb8 + (size_t)y * width * 3, (size_t)width * 3);
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) return B200RT_EIO;
    std::vector<uint8_t> png;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    png.insert(png.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_u32(ihdr, width); put_u32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit, truecolour
    chunk(png, "IHDR", ihdr.data(), ihdr.size());
    chunk(png, "IDAT", z.data(), bound);
    chunk(png, "IEND", nullptr, 0);
    uint8_t* buf = (uint8_t*)malloc(png.size());
    if (!buf) return B200RT_ENOMEM;
    memcpy(buf, png.data(), png.size());
    *out_bytes = buf; *out_len = png.size();
    return B200RT_OK;
}

extern "C" void b200rt_free(void* p) { free(p); }

extern "C" int b200rt_write_png(const char* path, const uint8_t* rgb8, uint32_t width, uint32_t height) {
    if (!path) return B200RT_EINVAL;
    uint8_t* bytes = nullptr; size_t len = 0;
    int rc = b200rt_encode_png(rgb8, width, height, &bytes, &len);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) { free(bytes); return B200RT_EIO; }
    size_t w = fwrite(bytes, 1, len, f);
    int c = fclose(f);
    free(bytes);
    return (w == len && c == 0) ? B200RT_OK : B200RT_EIO;
}
