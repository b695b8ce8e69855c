// png_decoder.cpp — PNG -> RGB8 for TextureLoader::ImagePath (image_texture.rs:23-26 `image::open`, then
// `to_rgb8()`-style texel reads at :44-55).  zlib does the inflate; this file does the container, the five
// scanline filters and the conversion to RGB8.  Supported: 8-bit greyscale, greyscale+alpha, RGB, RGBA and
// palette images (1/2/4/8-bit indices; 1/2/4-bit greyscale), non-interlaced.  16-bit and Adam7-interlaced
// files are rejected with an error.  Alpha is dropped (the reference reads pixel[0..3]).
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "raytracer.hpp"

namespace raytracer {
namespace scene {
namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

ImageData decode_png(const uint8_t* data, size_t size) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || memcmp(data, sig, 8) != 0) throw Error("png: bad signature");
    uint32_t W = 0, H = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    bool have_ihdr = false, have_iend = false;
    size_t pos = 8;
    while (pos + 12 <= size && !have_iend) {
        uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        if ((size_t)len > size - pos - 12) throw Error("png: truncated chunk");
        const uint8_t* body = data + pos + 8;
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len);
        if (crc != be32(body + len)) throw Error("png: chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) throw Error("png: bad IHDR");
            W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (W == 0 || H == 0 || W > 65536 || H > 65536) throw Error("png: unsupported dimensions");
            if (body[10] != 0 || body[11] != 0) throw Error("png: unknown compression or filter method");
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            if (len % 3 != 0 || len > 768) throw Error("png: bad PLTE");
            palette.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            have_iend = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || idat.empty()) throw Error("png: missing IHDR or IDAT");
    if (interlace != 0) throw Error("png: Adam7-interlaced files are not supported");
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;    // greyscale
        case 2: channels = 3; break;    // RGB
        case 3: channels = 1; break;    // palette
        case 4: channels = 2; break;    // greyscale + alpha
        case 6: channels = 4; break;    // RGBA
        default: throw Error("png: unknown colour type");
    }
    if (depth == 16) throw Error("png: 16-bit samples are not supported");
    bool sub_byte = depth < 8;
    if (!(depth == 8 || (sub_byte && (depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3)))) throw Error("png: unsupported bit depth for this colour type");
    if (ctype == 3 && palette.empty()) throw Error("png: palette image without PLTE");
    const size_t bpp = sub_byte ? 1 : (size_t)channels;                       // filter unit in bytes
    const size_t stride = ((size_t)W * channels * depth + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf out_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != raw.size()) throw Error("png: inflate failed or size mismatch");
    // undo the scanline filters in place (PNG spec 9.2)
    std::vector<uint8_t> zero(stride, 0);
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t* row = raw.data() + (size_t)y * (stride + 1);
        int ft = row[0];
        uint8_t* cur = row + 1;
        const uint8_t* up = y ? row - stride : zero.data();
        switch (ft) {
            case 0: break;
            case 1: for (size_t i = bpp; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
            case 2: for (size_t i = 0; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
            case 3: for (size_t i = 0; i < stride; ++i) { int a = i >= bpp ? cur[i - bpp] : 0; cur[i] = (uint8_t)(cur[i] + ((a + up[i]) >> 1)); } break;
            case 4: for (size_t i = 0; i < stride; ++i) { int a = i >= bpp ? cur[i - bpp] : 0, c = i >= bpp ? up[i - bpp] : 0; cur[i] = (uint8_t)(cur[i] + paeth(a, up[i], c)); } break;
            default: throw Error("png: unknown filter type");
        }
    }
    ImageData img; img.width = W; img.height = H; img.rgb.resize((size_t)W * H * 3);
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t* cur = raw.data() + (size_t)y * (stride + 1) + 1;
        uint8_t* o = &img.rgb[(size_t)y * W * 3];
        for (uint32_t x = 0; x < W; ++x, o += 3) {
            if (sub_byte) {
                size_t bit = (size_t)x * depth;
                int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                if (ctype == 3) {
                    if ((size_t)v * 3 + 2 >= palette.size()) throw Error("png: palette index out of range");
                    o[0] = palette[v * 3]; o[1] = palette[v * 3 + 1]; o[2] = palette[v * 3 + 2];
                } else { uint8_t g = (uint8_t)(v * 255 / ((1 << depth) - 1)); o[0] = o[1] = o[2] = g; }
            } else if (ctype == 3) {
                int v = cur[x];
                if ((size_t)v * 3 + 2 >= palette.size()) throw Error("png: palette index out of range");
                o[0] = palette[v * 3]; o[1] = palette[v * 3 + 1]; o[2] = palette[v * 3 + 2];
            } else if (channels <= 2) { o[0] = o[1] = o[2] = cur[(size_t)x * channels]; }
            else { const uint8_t* p = cur + (size_t)x * channels; o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; }
        }
    }
    return img;
}

// image::open: the format is sniffed from the file's first bytes.
ImageData load_image_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open image `" + path + "`");   // image_texture.rs:24 `image::open(path)?`
    std::vector<uint8_t> buf;
    uint8_t chunk[65536]; size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    fclose(f);
    if (buf.size() >= 2 && buf[0] == 0xFF && buf[1] == 0xD8) return decode_jpeg(buf.data(), buf.size());
    if (buf.size() >= 8 && buf[0] == 0x89 && buf[1] == 'P') return decode_png(buf.data(), buf.size());
    throw Error("image `" + path + "`: unsupported format (baseline/progressive JPEG and PNG are decoded here)");
}

}  // namespace scene
}  // namespace raytracer
