// png_decoder.cpp — PNG -> RGB8 for TextureLoader::ImagePath (image_texture.rs:23-26 `image::open`, then
// `to_rgb8()`-style texel reads at :44-55).  zlib does the inflate; this file does the container, the five
// scanline filters, Adam7 de-interlacing and the conversion to RGB8.  Supported: every colour type and bit depth
// of the PNG specification (greyscale 1/2/4/8/16, greyscale+alpha 8/16, RGB 8/16, RGBA 8/16, palette 1/2/4/8),
// interlaced or not.  16-bit samples become 8-bit the way the `image` crate's `get_pixel` on a DynamicImage does
// (u16 -> u8: (c + 128) / 257, image 0.24 color conversion).  Alpha is dropped (the reference reads pixel[0..3]);
// tRNS / gamma / colour-profile chunks are ignored like `image::open` + `get_pixel` ignore them.
#include <zlib.h>

#include <algorithm>
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
            if (W == 0 || H == 0 || W > 65536 || H > 65536 || (uint64_t)W * H > (1ull << 28)) throw Error("png: unsupported dimensions");
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
    if (interlace > 1) throw Error("png: unknown interlace method");
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;    // greyscale
        case 2: channels = 3; break;    // RGB
        case 3: channels = 1; break;    // palette
        case 4: channels = 2; break;    // greyscale + alpha
        case 6: channels = 4; break;    // RGBA
        default: throw Error("png: unknown colour type");
    }
    const bool sub_byte = depth < 8;
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok) throw Error("png: unsupported bit depth for this colour type");
    if (ctype == 3 && palette.empty()) throw Error("png: palette image without PLTE");
    const size_t bpp = sub_byte ? 1 : (size_t)channels * (depth / 8);         // filter unit in bytes
    auto row_bytes = [&](uint32_t w) { return ((size_t)w * channels * depth + 7) / 8; };
    // Adam7 (PNG spec 8.2): seven reduced images, each filtered on its own; pass p holds the pixels
    // (x0 + i * dx, y0 + j * dy).  A non-interlaced file is the single "pass" (0, 0, 1, 1).
    struct Pass { uint32_t x0, y0, dx, dy; };
    static const Pass ADAM7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass WHOLE = {0, 0, 1, 1};
    const Pass* passes = interlace ? ADAM7 : &WHOLE;
    const int n_passes = interlace ? 7 : 1;
    size_t total = 0;
    for (int k = 0; k < n_passes; ++k) {
        const Pass& ps = passes[k];
        uint32_t pw = W > ps.x0 ? (W - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = H > ps.y0 ? (H - ps.y0 + ps.dy - 1) / ps.dy : 0;
        if (pw && ph) total += (row_bytes(pw) + 1) * ph;
    }
    std::vector<uint8_t> raw(total);
    uLongf out_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != raw.size()) throw Error("png: inflate failed or size mismatch");
    ImageData img; img.width = W; img.height = H; img.rgb.resize((size_t)W * H * 3);
    auto to8 = [](const uint8_t* be16) { return (uint8_t)((((uint32_t)be16[0] << 8 | be16[1]) + 128u) / 257u); };   // image 0.24: u16 -> u8
    size_t off = 0;
    for (int k = 0; k < n_passes; ++k) {
        const Pass& ps = passes[k];
        const uint32_t pw = W > ps.x0 ? (W - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = H > ps.y0 ? (H - ps.y0 + ps.dy - 1) / ps.dy : 0;
        if (!pw || !ph) continue;
        const size_t stride = row_bytes(pw);
        std::vector<uint8_t> zero(stride, 0);
        for (uint32_t j = 0; j < ph; ++j) {
            // undo the scanline filter in place (PNG spec 9.2)
            uint8_t* row = raw.data() + off + (size_t)j * (stride + 1);
            const int ft = row[0];
            uint8_t* cur = row + 1;
            const uint8_t* up = j ? row - stride : zero.data();
            switch (ft) {
                case 0: break;
                case 1: for (size_t i = bpp; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
                case 2: for (size_t i = 0; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
                case 3: for (size_t i = 0; i < stride; ++i) { int a = i >= bpp ? cur[i - bpp] : 0; cur[i] = (uint8_t)(cur[i] + ((a + up[i]) >> 1)); } break;
                case 4: for (size_t i = 0; i < stride; ++i) { int a = i >= bpp ? cur[i - bpp] : 0, c = i >= bpp ? up[i - bpp] : 0; cur[i] = (uint8_t)(cur[i] + paeth(a, up[i], c)); } break;
                default: throw Error("png: unknown filter type");
            }
            const uint32_t y = ps.y0 + j * ps.dy;
            for (uint32_t i = 0; i < pw; ++i) {
                uint8_t* o = &img.rgb[((size_t)y * W + ps.x0 + (size_t)i * ps.dx) * 3];
                if (sub_byte) {
                    size_t bit = (size_t)i * depth;
                    int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                    if (ctype == 3) {
                        if ((size_t)v * 3 + 2 >= palette.size()) throw Error("png: palette index out of range");
                        o[0] = palette[v * 3]; o[1] = palette[v * 3 + 1]; o[2] = palette[v * 3 + 2];
                    } else { uint8_t g = (uint8_t)(v * 255 / ((1 << depth) - 1)); o[0] = o[1] = o[2] = g; }
                } else if (ctype == 3) {
                    int v = cur[i];
                    if ((size_t)v * 3 + 2 >= palette.size()) throw Error("png: palette index out of range");
                    o[0] = palette[v * 3]; o[1] = palette[v * 3 + 1]; o[2] = palette[v * 3 + 2];
                } else if (depth == 16) {
                    const uint8_t* p = cur + (size_t)i * channels * 2;
                    if (channels <= 2) o[0] = o[1] = o[2] = to8(p);
                    else { o[0] = to8(p); o[1] = to8(p + 2); o[2] = to8(p + 4); }
                } else if (channels <= 2) { o[0] = o[1] = o[2] = cur[(size_t)i * channels]; }
                else { const uint8_t* p = cur + (size_t)i * channels; o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; }
            }
        }
        off += (stride + 1) * ph;
    }
    return img;
}

// ---- the two other formats simple enough to carry: BMP and binary / ASCII PNM (image::open sniffs them as well) -------------
namespace {
uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace

// Windows BMP, uncompressed (BI_RGB): 24- and 32-bit true colour and 1/4/8-bit palette images, bottom-up or top-down.
ImageData decode_bmp(const uint8_t* d, size_t size) {
    if (size < 54 || d[0] != 'B' || d[1] != 'M') throw Error("bmp: bad signature");
    const uint32_t data_off = le32(d + 10), hdr = le32(d + 14);
    if (hdr < 40 || 14 + (size_t)hdr > size) throw Error("bmp: unsupported header");
    const int32_t w = (int32_t)le32(d + 18), hs = (int32_t)le32(d + 22);
    const uint16_t planes = le16(d + 26), bpp = le16(d + 28);
    const uint32_t comp = le32(d + 30);
    uint32_t n_colors = le32(d + 46);
    const bool top_down = hs < 0;
    const int64_t h = top_down ? -(int64_t)hs : hs;
    if (w <= 0 || h <= 0 || w > 65536 || h > 65536 || (uint64_t)w * (uint64_t)h > (1ull << 28) || planes != 1) throw Error("bmp: unsupported dimensions");
    if (!(comp == 0 || (comp == 3 && bpp == 32))) throw Error("bmp: compressed files are not supported");
    if (!(bpp == 1 || bpp == 4 || bpp == 8 || bpp == 24 || bpp == 32)) throw Error("bmp: unsupported bit depth");
    const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
    if ((size_t)data_off > size || stride * (size_t)h > size - data_off) throw Error("bmp: truncated pixel data");
    const uint8_t* pal = d + 14 + hdr;
    if (bpp <= 8) {
        if (n_colors == 0) n_colors = 1u << bpp;
        if (n_colors > 256 || 14 + (size_t)hdr + 4 * (size_t)n_colors > size) throw Error("bmp: bad palette");
    }
    ImageData img; img.width = (uint32_t)w; img.height = (uint32_t)h; img.rgb.resize((size_t)w * h * 3);
    for (int64_t y = 0; y < h; ++y) {
        const uint8_t* row = d + data_off + stride * (size_t)(top_down ? y : h - 1 - y);
        uint8_t* o = &img.rgb[(size_t)y * w * 3];
        for (int32_t x = 0; x < w; ++x, o += 3) {
            if (bpp >= 24) { const uint8_t* p = row + (size_t)x * (bpp / 8); o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; }
            else {
                size_t bit = (size_t)x * bpp;
                uint32_t v = (row[bit >> 3] >> (8 - bpp - (bit & 7))) & ((1u << bpp) - 1);
                if (v >= n_colors) throw Error("bmp: palette index out of range");
                o[0] = pal[4 * v + 2]; o[1] = pal[4 * v + 1]; o[2] = pal[4 * v];
            }
        }
    }
    return img;
}

// Netpbm: P5 / P6 (binary) and P2 / P3 (ASCII) grey and colour maps, maxval <= 65535 (samples are rescaled to 8 bits).
ImageData decode_pnm(const uint8_t* d, size_t size) {
    if (size < 3 || d[0] != 'P' || (d[1] != '2' && d[1] != '3' && d[1] != '5' && d[1] != '6')) throw Error("pnm: bad signature");
    const int kind = d[1] - '0';
    size_t pos = 2;
    auto number = [&]() -> uint32_t {
        for (;;) {                                           // white space and # comments
            while (pos < size && (d[pos] == ' ' || d[pos] == '\t' || d[pos] == '\r' || d[pos] == '\n')) ++pos;
            if (pos < size && d[pos] == '#') { while (pos < size && d[pos] != '\n') ++pos; } else break;
        }
        if (pos >= size || d[pos] < '0' || d[pos] > '9') throw Error("pnm: bad header");
        uint64_t v = 0;
        while (pos < size && d[pos] >= '0' && d[pos] <= '9') { v = v * 10 + (d[pos++] - '0'); if (v > 0xFFFFFFFFull) throw Error("pnm: bad header"); }
        return (uint32_t)v;
    };
    const uint32_t w = number(), h = number(), maxv = number();
    if (w == 0 || h == 0 || w > 65536 || h > 65536 || (uint64_t)w * h > (1ull << 28) || maxv == 0 || maxv > 65535) throw Error("pnm: unsupported dimensions or maxval");
    if (kind >= 5 && (uint64_t)w * h > size) throw Error("pnm: truncated pixel data");
    const int ch = (kind == 3 || kind == 6) ? 3 : 1;
    const size_t n = (size_t)w * h * ch;
    std::vector<uint32_t> v(n);
    if (kind >= 5) {
        ++pos;                                               // the single white-space byte after maxval
        const size_t bytes = maxv > 255 ? 2 : 1;
        if (pos > size || n * bytes > size - pos) throw Error("pnm: truncated pixel data");
        for (size_t i = 0; i < n; ++i) v[i] = bytes == 2 ? ((uint32_t)d[pos + 2 * i] << 8 | d[pos + 2 * i + 1]) : d[pos + i];
    } else for (size_t i = 0; i < n; ++i) v[i] = number();
    ImageData img; img.width = w; img.height = h; img.rgb.resize((size_t)w * h * 3);
    for (size_t p = 0; p < (size_t)w * h; ++p)
        for (int c = 0; c < 3; ++c) {
            uint32_t s = std::min(v[p * ch + (ch == 3 ? c : 0)], maxv);
            img.rgb[p * 3 + c] = (uint8_t)((s * 255u + maxv / 2) / maxv);
        }
    return img;
}

// image::open / image::load_from_memory: the format is sniffed from the first bytes.
ImageData decode_any(const uint8_t* d, size_t size) {
    if (size >= 2 && d[0] == 0xFF && d[1] == 0xD8) return decode_jpeg(d, size);
    if (size >= 8 && d[0] == 0x89 && d[1] == 'P') return decode_png(d, size);
    if (size >= 2 && d[0] == 'B' && d[1] == 'M') return decode_bmp(d, size);
    if (size >= 2 && d[0] == 'P' && d[1] >= '2' && d[1] <= '6' && d[1] != '4') return decode_pnm(d, size);
    throw Error("unsupported format (JPEG baseline/progressive, PNG, BMP and PNM are decoded here; GIF, TIFF, WebP, TGA, ... are not)");
}

// image::open: the format is sniffed from the file's first bytes.
ImageData load_image_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open image `" + path + "`");   // image_texture.rs:24 `image::open(path)?`
    std::vector<uint8_t> buf;
    uint8_t chunk[65536]; size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    fclose(f);
    try { return decode_any(buf.data(), buf.size()); }
    catch (const Error& e) { throw Error("image `" + path + "`: " + e.what()); }
}

}  // namespace scene
}  // namespace raytracer
