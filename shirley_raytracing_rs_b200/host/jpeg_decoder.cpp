// jpeg_decoder.cpp — baseline JPEG -> RGB8 for TextureLoader::{ImagePath, EarthBuiltin}.
//
// The reference decodes textures with `image::open` / `image::load_from_memory`
// (material/texture/image_texture.rs:23-31; crate image 0.24.3 -> jpeg-decoder 0.2.6, not
// vendored under /root/reference) and reads texels as RGB8 (image_texture.rs:44-55).  No JPEG
// library headers exist in this image, so this is a from-scratch decoder of the subset the
// texture path needs: sequential (SOF0/SOF1) and progressive (SOF2) DCT, 8-bit, Huffman, 1 or 3 components,
// any scan script, restart intervals, JFIF YCbCr or Adobe RGB.  Arithmetic follows the published libjpeg
// algorithms — the accurate integer inverse DCT ("islow", 13-bit constants), the 16-bit
// fixed-point YCbCr->RGB tables and the triangle-filter ("fancy") chroma upsampling for 2x1 and
// 2x2 subsampling — so that the bytes agree with libjpeg-based decoders (PIL, OpenCV), which
// tests/test_jpeg.py checks.  jpeg-decoder's own IDCT may differ by +-1 level on some texels;
// assets/earthmap.jpg is 4:4:4 baseline, where only the IDCT rounding is in play.
// Arithmetic-coded, lossless and hierarchical files are rejected with an error.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "raytracer.hpp"

namespace raytracer {
namespace scene {
namespace {

struct HuffTable {
    bool defined = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    // canonical decoding tables (ITU T.81 F.2.2.3)
    int mincode[17], maxcode[18], valptr[17];
    // 9-bit lookahead: (length << 8) | symbol, 0 = longer than 9 bits
    uint16_t look[512];
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        memset(look, 0, sizeof look);
        code = 0; k = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                int first = code << (9 - l);
                for (int f = 0; f < (1 << (9 - l)) && first + f < 512; ++f) look[first + f] = (uint16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
        defined = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;     // allocated blocks (whole MCUs)
    int width = 0, height = 0;          // downsampled size in samples: ceil(image * h / hmax)
    std::vector<uint8_t> plane;         // blocks_w*8 x blocks_h*8 samples
    std::vector<int16_t> coef;          // blocks_w x blocks_h blocks of 64 coefficients, natural order
    int dc_pred = 0;
};

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int n = 0;
    bool hit_marker = false;
    void fill() {
        while (n <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    int b2 = (p + 1 < end) ? p[1] : 0xD9;
                    if (b2 == 0x00) p += 2;
                    else { hit_marker = true; b = 0; }   // leave p at the marker, feed zeros
                } else ++p;
            }
            acc |= (uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; n -= k; }
    int get(int k) { if (k == 0) return 0; int v = peek(k); skip(k); return v; }
    void reset() { acc = 0; n = 0; hit_marker = false; }
};

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }   // T.81 F.2.2.1

int decode_symbol(BitReader& br, const HuffTable& h) {
    int idx = br.peek(9);
    uint16_t e = h.look[idx];
    if (e) { br.skip(e >> 8); return e & 0xff; }
    int code = br.peek(16);
    for (int l = 10; l <= 16; ++l) {
        int c = code >> (16 - l);
        if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) { br.skip(l); return h.vals[h.valptr[l] + c - h.mincode[l]]; }
    }
    throw Error("jpeg: bad Huffman code");
}

const uint8_t ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
inline uint8_t clamp8_64(int64_t x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// The accurate integer inverse DCT of libjpeg (jidctint.c, "islow"): Loeffler-Ligtenberg-
// Moschytz factorisation, 13-bit fixed-point constants, two intermediate fraction bits.
void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
    const int CONST_BITS = 13, PASS1_BITS = 2;
    const int64_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                  F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
    auto descale = [](int64_t x, int n) { return (x + (int64_t(1) << (n - 1))) >> n; };   // 64-bit: a corrupt stream's coefficients must not overflow (valid streams never reach 2^31)
    int64_t ws[64];
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 8; ++i) {
            int64_t in[8];
            if (pass == 0) for (int k = 0; k < 8; ++k) in[k] = (int64_t)coef[k * 8 + i] * (int64_t)q[k * 8 + i];
            else for (int k = 0; k < 8; ++k) in[k] = ws[i * 8 + k];
            int64_t z2 = in[2], z3 = in[6];
            int64_t z1 = (z2 + z3) * F_0_541;
            int64_t tmp2 = z1 + z3 * (-F_1_847);
            int64_t tmp3 = z1 + z2 * F_0_765;
            int64_t tmp0 = (in[0] + in[4]) * (1 << CONST_BITS);
            int64_t tmp1 = (in[0] - in[4]) * (1 << CONST_BITS);
            int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            int64_t z4 = tmp1 + tmp3;
            int64_t z5 = (z3 + z4) * F_1_175;
            tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
            z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            int64_t o[8] = {tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3};
            if (pass == 0) for (int k = 0; k < 8; ++k) ws[k * 8 + i] = descale(o[k], CONST_BITS - PASS1_BITS);
            else for (int k = 0; k < 8; ++k) out[i * stride + k] = clamp8_64(descale(o[k], CONST_BITS + PASS1_BITS + 3) + 128);
        }
    }
}

uint16_t rd16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

}  // namespace

namespace {

// One scan of entropy-coded data (T.81 F.2 sequential, G.1 progressive).  Coefficients accumulate in the
// components' `coef` arrays (natural order, 64 per block); the inverse DCT runs once after the last scan.
struct Scan {
    std::vector<Component*> comps;
    int Ss = 0, Se = 63, Ah = 0, Al = 0;
    bool progressive = false;
};

struct ScanDecoder {
    BitReader br;
    const HuffTable* dc; const HuffTable* ac;   // arrays of 4
    const Scan& sc;
    int eobrun = 0;
    ScanDecoder(const HuffTable* d, const HuffTable* a, const Scan& s) : dc(d), ac(a), sc(s) {}

    void block_sequential(Component& c, int16_t* blk) {                       // F.2.2
        int t = decode_symbol(br, dc[c.td]);
        if (t > 11) throw Error("jpeg: bad DC category");
        c.dc_pred += t ? extend(br.get(t), t) : 0;
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            int rs = decode_symbol(br, ac[c.ta]);
            int r = rs >> 4, s = rs & 15;
            if (s == 0) { if (r == 15) { k += 16; continue; } break; }       // ZRL / EOB
            k += r;
            if (k > 63) throw Error("jpeg: AC coefficient index out of range");
            blk[ZIGZAG[k]] = (int16_t)extend(br.get(s), s);
            ++k;
        }
    }
    void block_dc_first(Component& c, int16_t* blk) {                         // G.1.2.1
        int t = decode_symbol(br, dc[c.td]);
        if (t > 11) throw Error("jpeg: bad DC category");
        c.dc_pred += t ? extend(br.get(t), t) : 0;
        blk[0] = (int16_t)(c.dc_pred * (1 << sc.Al));
    }
    void block_dc_refine(int16_t* blk) { if (br.get(1)) blk[0] = (int16_t)(blk[0] | (1 << sc.Al)); }
    void block_ac_first(Component& c, int16_t* blk) {                         // G.1.2.2
        if (eobrun > 0) { --eobrun; return; }
        for (int k = sc.Ss; k <= sc.Se; ++k) {
            int rs = decode_symbol(br, ac[c.ta]);
            int r = rs >> 4, s = rs & 15;
            if (s) {
                k += r;
                if (k > 63) throw Error("jpeg: AC coefficient index out of range");
                blk[ZIGZAG[k]] = (int16_t)(extend(br.get(s), s) * (1 << sc.Al));
            } else if (r == 15) {
                k += 15;
            } else {
                eobrun = 1 << r;
                if (r) eobrun += br.get(r);
                --eobrun;
                break;
            }
        }
    }
    void block_ac_refine(Component& c, int16_t* blk) {                        // G.1.2.3
        const int p1 = 1 << sc.Al, m1 = -(1 << sc.Al);
        int k = sc.Ss;
        auto correct = [&](int16_t& coef) {
            if (br.get(1) && (coef & p1) == 0) coef = (int16_t)(coef + (coef >= 0 ? p1 : m1));
        };
        if (eobrun == 0) {
            for (; k <= sc.Se; ++k) {
                int rs = decode_symbol(br, ac[c.ta]);
                int r = rs >> 4, s = rs & 15;
                if (s) {
                    if (s != 1) throw Error("jpeg: bad refinement symbol");
                    s = br.get(1) ? p1 : m1;
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += br.get(r);
                    break;                                                    // the rest of the band is handled below
                }
                // skip r zero-history coefficients, correcting the non-zero ones passed on the way
                for (; k <= sc.Se; ++k) {
                    int16_t& coef = blk[ZIGZAG[k]];
                    if (coef != 0) correct(coef);
                    else if (--r < 0) break;
                }
                if (s && k <= sc.Se) blk[ZIGZAG[k]] = (int16_t)s;
            }
        }
        if (eobrun > 0) {
            for (; k <= sc.Se; ++k) { int16_t& coef = blk[ZIGZAG[k]]; if (coef != 0) correct(coef); }
            --eobrun;
        }
    }
    void block(Component& c, int16_t* blk) {
        if (!sc.progressive) block_sequential(c, blk);
        else if (sc.Ss == 0) { if (sc.Ah == 0) block_dc_first(c, blk); else block_dc_refine(blk); }
        else { if (sc.Ah == 0) block_ac_first(c, blk); else block_ac_refine(c, blk); }
    }
};

}  // namespace

ImageData decode_jpeg(const uint8_t* data, size_t size) {
    if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) throw Error("jpeg: missing SOI marker");
    uint16_t qt[4][64]; bool qt_def[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<Component> comps;
    int W = 0, H = 0, hmax = 1, vmax = 1, restart_interval = 0;
    bool adobe = false; int adobe_transform = -1; bool jfif = false;
    bool have_frame = false, progressive = false, done = false, any_scan = false;
    size_t pos = 2;
    int mcus_x = 0, mcus_y = 0;

    while (!done) {
        // next marker
        while (pos < size && data[pos] != 0xFF) ++pos;
        while (pos < size && data[pos] == 0xFF) ++pos;
        if (pos >= size) break;
        int m = data[pos++];
        if (m == 0xD9) break;                                             // EOI
        if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;   // stuffed byte left over from a scan / TEM / stray RSTn
        if (pos + 2 > size) throw Error("jpeg: truncated segment");
        size_t len = rd16(data + pos);
        if (len < 2 || pos + len > size) throw Error("jpeg: bad segment length");
        const uint8_t* seg = data + pos + 2; size_t n = len - 2;
        switch (m) {
            case 0xDB: {   // DQT
                size_t i = 0;
                while (i < n) {
                    int pq = seg[i] >> 4, tq = seg[i] & 15; ++i;
                    if (tq > 3) throw Error("jpeg: bad quantisation table id");
                    if (i + (pq ? 128 : 64) > n) throw Error("jpeg: truncated DQT");
                    for (int k = 0; k < 64; ++k) {
                        uint16_t v = pq ? rd16(seg + i + 2 * k) : seg[i + k];
                        qt[tq][ZIGZAG[k]] = v;                            // store in natural order
                    }
                    i += pq ? 128 : 64; qt_def[tq] = true;
                }
                break;
            }
            case 0xC4: {   // DHT
                size_t i = 0;
                while (i < n) {
                    if (i + 17 > n) throw Error("jpeg: truncated DHT");
                    int tc = seg[i] >> 4, th = seg[i] & 15; ++i;
                    if (tc > 1 || th > 3) throw Error("jpeg: bad Huffman table id");
                    HuffTable& t = tc ? ac[th] : dc[th];
                    int total = 0;
                    t.bits[0] = 0;
                    for (int l = 1; l <= 16; ++l) { t.bits[l] = seg[i + l - 1]; total += t.bits[l]; }
                    i += 16;
                    if (total > 256 || i + total > n) throw Error("jpeg: bad DHT counts");
                    {   // the code lengths must form a prefix code (Kraft): at most 2^l codes of length l remain at each
                        // length, else build()'s canonical codes run past their tables (libjpeg: "bogus Huffman table")
                        int code = 0;
                        for (int l = 1; l <= 16; ++l) {
                            code += t.bits[l];
                            if (code > (1 << l)) throw Error("jpeg: over-subscribed Huffman table");
                            code <<= 1;
                        }
                    }
                    memcpy(t.vals, seg + i, total); i += total;
                    t.build();
                }
                break;
            }
            case 0xC0: case 0xC1: case 0xC2: {   // SOF0 / SOF1 (sequential) / SOF2 (progressive), Huffman
                if (have_frame) throw Error("jpeg: more than one frame header");
                if (n < 6) throw Error("jpeg: truncated SOF");
                if (seg[0] != 8) throw Error("jpeg: only 8-bit samples are supported");
                progressive = m == 0xC2;
                H = rd16(seg + 1); W = rd16(seg + 3);
                int nc = seg[5];
                if (W == 0 || H == 0) throw Error("jpeg: empty image");
                if (nc != 1 && nc != 3) throw Error("jpeg: only 1- or 3-component images are supported");
                if (n < (size_t)(6 + 3 * nc)) throw Error("jpeg: truncated SOF");
                comps.resize(nc);
                for (int c = 0; c < nc; ++c) {
                    comps[c].id = seg[6 + 3 * c]; comps[c].h = seg[7 + 3 * c] >> 4; comps[c].v = seg[7 + 3 * c] & 15; comps[c].tq = seg[8 + 3 * c];
                    if (comps[c].h < 1 || comps[c].h > 4 || comps[c].v < 1 || comps[c].v > 4 || comps[c].tq > 3) throw Error("jpeg: bad component spec");
                    hmax = std::max(hmax, comps[c].h); vmax = std::max(vmax, comps[c].v);
                }
                if (nc == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }   // single-component scans are never interleaved
                mcus_x = (W + 8 * hmax - 1) / (8 * hmax); mcus_y = (H + 8 * vmax - 1) / (8 * vmax);
                for (auto& c : comps) {
                    c.blocks_w = mcus_x * c.h; c.blocks_h = mcus_y * c.v;
                    c.width = (W * c.h + hmax - 1) / hmax; c.height = (H * c.v + vmax - 1) / vmax;
                    c.plane.assign((size_t)c.blocks_w * 8 * c.blocks_h * 8, 0);
                    c.coef.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
                }
                have_frame = true;
                break;
            }
            case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                throw Error("jpeg: unsupported coding process (lossless / hierarchical / arithmetic)");
            case 0xDD: if (n < 2) throw Error("jpeg: truncated DRI"); restart_interval = rd16(seg); break;
            case 0xE0: if (n >= 5 && memcmp(seg, "JFIF", 5) == 0) jfif = true; break;
            case 0xEE: if (n >= 12 && memcmp(seg, "Adobe", 5) == 0) { adobe = true; adobe_transform = seg[11]; } break;
            case 0xDA: {   // SOS + entropy-coded data
                if (!have_frame) throw Error("jpeg: scan before frame header");
                if (n < 1) throw Error("jpeg: truncated SOS");
                int ns = seg[0];
                if (ns < 1 || ns > (int)comps.size() || n < (size_t)(1 + 2 * ns + 3)) throw Error("jpeg: bad SOS");
                Scan sc; sc.progressive = progressive;
                for (int s = 0; s < ns; ++s) {
                    int id = seg[1 + 2 * s]; Component* found = nullptr;
                    for (auto& c : comps) if (c.id == id) found = &c;
                    if (!found) throw Error("jpeg: scan names an unknown component");
                    found->td = seg[2 + 2 * s] >> 4; found->ta = seg[2 + 2 * s] & 15;
                    if (found->td > 3 || found->ta > 3) throw Error("jpeg: bad Huffman table selector");
                    sc.comps.push_back(found);
                }
                sc.Ss = seg[1 + 2 * ns]; sc.Se = seg[2 + 2 * ns]; sc.Ah = seg[3 + 2 * ns] >> 4; sc.Al = seg[3 + 2 * ns] & 15;
                if (!progressive) { sc.Ss = 0; sc.Se = 63; sc.Ah = sc.Al = 0; }
                if (sc.Ss > sc.Se || sc.Se > 63 || sc.Al > 13 || (progressive && sc.Ss > 0 && ns != 1) || (progressive && sc.Ss == 0 && sc.Se != 0))
                    throw Error("jpeg: bad progressive scan parameters");
                const bool need_dc = !progressive || (sc.Ss == 0 && sc.Ah == 0), need_ac = !progressive || sc.Ss > 0;
                for (Component* c : sc.comps) {
                    if ((need_dc && !dc[c->td].defined) || (need_ac && !ac[c->ta].defined)) throw Error("jpeg: scan uses an undefined Huffman table");
                    if (!qt_def[c->tq]) throw Error("jpeg: component uses an undefined quantisation table");
                    c->dc_pred = 0;
                }
                ScanDecoder dec(dc, ac, sc);
                dec.br.p = data + pos + len; dec.br.end = data + size;
                // an interleaved scan walks MCUs; a single-component scan walks that component's own blocks (A.2.3)
                const bool interleaved = ns > 1;
                Component& c0 = *sc.comps[0];
                const int units_x = interleaved ? mcus_x : (c0.width + 7) / 8, units_y = interleaved ? mcus_y : (c0.height + 7) / 8;
                int until_restart = restart_interval, next_rst = 0;
                for (int uy = 0; uy < units_y; ++uy) for (int ux = 0; ux < units_x; ++ux) {
                    if (restart_interval && until_restart == 0) {
                        dec.br.reset();
                        const uint8_t* p = dec.br.p;
                        while (p + 1 < dec.br.end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
                        if (p + 1 >= dec.br.end) throw Error("jpeg: missing restart marker");
                        if ((p[1] & 7) != next_rst) throw Error("jpeg: restart markers out of order");
                        next_rst = (next_rst + 1) & 7;
                        dec.br.p = p + 2;
                        for (Component* c : sc.comps) c->dc_pred = 0;
                        dec.eobrun = 0;
                        until_restart = restart_interval;
                    }
                    if (interleaved) {
                        for (Component* c : sc.comps) for (int by = 0; by < c->v; ++by) for (int bx = 0; bx < c->h; ++bx)
                            dec.block(*c, &c->coef[((size_t)(uy * c->v + by) * c->blocks_w + (ux * c->h + bx)) * 64]);
                    } else {
                        dec.block(c0, &c0.coef[((size_t)uy * c0.blocks_w + ux) * 64]);
                    }
                    if (restart_interval) --until_restart;
                }
                any_scan = true;
                pos = (size_t)(dec.br.p - data);      // the reader never runs past a marker: resume the marker search there
                if (!progressive && ns == (int)comps.size()) done = true;   // a fully interleaved baseline scan is the whole picture
                continue;                              // (pos already advanced past the segment and its entropy data)
            }
            default: break;   // APPn, COM, ...: skipped
        }
        pos += len;
    }
    if (!any_scan) throw Error("jpeg: no scan data found");
    for (auto& c : comps)
        for (int by = 0; by < c.blocks_h; ++by) for (int bx = 0; bx < c.blocks_w; ++bx)
            idct_islow(&c.coef[((size_t)by * c.blocks_w + bx) * 64], qt[c.tq], c.plane.data() + (size_t)by * 8 * c.blocks_w * 8 + bx * 8, c.blocks_w * 8);

    ImageData img; img.width = (uint32_t)W; img.height = (uint32_t)H; img.rgb.resize((size_t)W * H * 3);
    if (comps.size() == 1) {
        const Component& c = comps[0];
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            uint8_t v = c.plane[(size_t)y * c.blocks_w * 8 + x];
            uint8_t* o = &img.rgb[((size_t)y * W + x) * 3]; o[0] = o[1] = o[2] = v;
        }
        return img;
    }
    // upsample each component to full resolution
    std::vector<std::vector<uint8_t>> full(3);
    for (int ci = 0; ci < 3; ++ci) {
        const Component& c = comps[ci];
        const int sw = c.blocks_w * 8;
        std::vector<uint8_t>& out = full[ci];
        out.resize((size_t)W * H);
        auto src = [&](int x, int y) -> int {   // edge replication inside the component's real size
            x = x < 0 ? 0 : (x >= c.width ? c.width - 1 : x); y = y < 0 ? 0 : (y >= c.height ? c.height - 1 : y);
            return c.plane[(size_t)y * sw + x];
        };
        if (c.h == hmax && c.v == vmax) {
            for (int y = 0; y < H; ++y) memcpy(&out[(size_t)y * W], &c.plane[(size_t)y * sw], W);
        } else if (c.h * 2 == hmax && c.v == vmax) {
            // h2v1 triangle filter (libjpeg jdsample.c h2v1_fancy_upsample): 3/4 nearer + 1/4 farther, ordered-dither rounding
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                int i = x >> 1, v;
                if (c.width == 1) v = src(0, y);
                else if (x & 1) v = (i == c.width - 1) ? src(i, y) : (3 * src(i, y) + src(i + 1, y) + 2) >> 2;
                else v = (i == 0) ? src(0, y) : (3 * src(i, y) + src(i - 1, y) + 1) >> 2;
                out[(size_t)y * W + x] = (uint8_t)v;
            }
        } else if (c.h * 2 == hmax && c.v * 2 == vmax) {
            // h2v2 triangle filter (h2v2_fancy_upsample): vertical 3:1 column sums, then horizontal 3:1, /16
            for (int y = 0; y < H; ++y) {
                int j = y >> 1, jn = (y & 1) ? j + 1 : j - 1;     // nearer / farther source rows
                for (int x = 0; x < W; ++x) {
                    int i = x >> 1;
                    auto colsum = [&](int xi) { return 3 * src(xi, j) + src(xi, jn); };
                    int cur = colsum(i), v;
                    if (c.width == 1) v = (cur * 4 + 8) >> 4;
                    else if (x & 1) v = (i == c.width - 1) ? (cur * 4 + 7) >> 4 : (cur * 3 + colsum(i + 1) + 7) >> 4;
                    else v = (i == 0) ? (cur * 4 + 8) >> 4 : (cur * 3 + colsum(i - 1) + 8) >> 4;
                    out[(size_t)y * W + x] = (uint8_t)v;
                }
            }
        } else {
            // other ratios: sample replication (libjpeg int_upsample)
            if (hmax % c.h || vmax % c.v) throw Error("jpeg: fractional sampling ratios are not supported");
            int fx = hmax / c.h, fy = vmax / c.v;
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = (uint8_t)src(x / fx, y / fy);
        }
    }
    // colour: Adobe transform 0 = RGB; otherwise YCbCr (JFIF, Adobe transform 1, or unmarked with ids 1,2,3 / anything else)
    bool rgb_direct = adobe ? adobe_transform == 0 : (!jfif && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B');
    for (size_t i = 0, n = (size_t)W * H; i < n; ++i) {
        int y = full[0][i], cb = full[1][i], cr = full[2][i];
        uint8_t* o = &img.rgb[i * 3];
        if (rgb_direct) { o[0] = (uint8_t)y; o[1] = (uint8_t)cb; o[2] = (uint8_t)cr; continue; }
        // libjpeg jdcolor.c build_ycc_rgb_table: 16-bit fixed point, rounding folded into the tables
        int xb = cb - 128, xr = cr - 128;
        int r = y + ((91881 * xr + 32768) >> 16);
        int g = y + ((-22554 * xb + 32768 - 46802 * xr) >> 16);
        int b = y + ((116130 * xb + 32768) >> 16);
        o[0] = clamp8(r); o[1] = clamp8(g); o[2] = clamp8(b);
    }
    return img;
}

ImageData load_jpeg_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open image `" + path + "`");   // image_texture.rs:24 `image::open(path)?`
    std::vector<uint8_t> buf;
    uint8_t chunk[65536]; size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    fclose(f);
    return decode_jpeg(buf.data(), buf.size());
}

}  // namespace scene
}  // namespace raytracer
