// image_io.cpp -- see image_io.h. Host-only C++17, no dependencies: an inflate (RFC 1951), a PNG reader (RFC 2083) and an OpenEXR scanline reader
// (NONE / RLE / ZIPS / ZIP / PIZ; the PIZ Huffman coder and wavelet follow the published OpenEXR format, ImfHuf / ImfWav / ImfPizCompressor) small
// enough to audit. Checked against zlib-written PNGs and against OpenEXR's own decoding of every EXR file in the reference tree (tests/test_plugin_host.py).
#include "image_io.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/igb200.h"
#include "script_recognizer.h"

namespace igbh {

[[noreturn]] static void io_fail(const std::string& path, const std::string& what) { throw RecognizeError{"image '" + path + "': " + what}; }

// ------------------------------------------------------------------------------------------------ inflate (zlib stream)
namespace {
struct Bits {
    const uint8_t* p; size_t n, pos = 0; uint32_t buf = 0; int cnt = 0;
    bool need(int k) { while (cnt < k) { if (pos >= n) return false; buf |= (uint32_t)p[pos++] << cnt; cnt += 8; } return true; }
    uint32_t get(int k) { const uint32_t v = buf & ((1u << k) - 1u); buf >>= k; cnt -= k; return v; }
};
struct Huffman {
    uint16_t count[16]; uint16_t symbol[288];
    void build(const uint8_t* lengths, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; ++i) count[lengths[i]]++;
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
        count[0] = 0;
    }
    int decode(Bits& b) const {   // canonical code, one bit at a time (the files in question are small)
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; ++l) {
            if (!b.need(1)) return -1;
            code |= (int)b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
}  // namespace

static bool inflate_zlib(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    if (in.size() < 6 || (in[0] & 0x0F) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return false;
    Bits b{in.data() + 2, in.size() - 2};
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
        if (!b.need(3)) return false;
        const uint32_t last = b.get(1), type = b.get(2);
        if (type == 0) {   // stored
            b.get(b.cnt & 7);   // to the byte boundary
            if (!b.need(32)) return false;
            const uint32_t len = b.get(16), nlen = b.get(16);
            if ((len ^ 0xFFFFu) != nlen) return false;
            for (uint32_t i = 0; i < len; ++i) { if (!b.need(8)) return false; out.push_back((uint8_t)b.get(8)); }
        } else if (type == 1 || type == 2) {
            Huffman lit, dist;
            uint8_t lengths[320];
            if (type == 1) {
                for (int i = 0; i < 144; ++i) lengths[i] = 8;
                for (int i = 144; i < 256; ++i) lengths[i] = 9;
                for (int i = 256; i < 280; ++i) lengths[i] = 7;
                for (int i = 280; i < 288; ++i) lengths[i] = 8;
                lit.build(lengths, 288);
                for (int i = 0; i < 30; ++i) lengths[i] = 5;
                dist.build(lengths, 30);
            } else {
                if (!b.need(14)) return false;
                const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
                if (nlen > 286 || ndist > 30) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int i = 0; i < ncode; ++i) { if (!b.need(3)) return false; cl[order[i]] = (uint8_t)b.get(3); }
                Huffman lc; lc.build(cl, 19);
                int i = 0;
                while (i < nlen + ndist) {
                    const int sym = lc.decode(b);
                    if (sym < 0) return false;
                    if (sym < 16) lengths[i++] = (uint8_t)sym;
                    else {
                        int rep; uint8_t val = 0;
                        if (sym == 16) { if (i == 0 || !b.need(2)) return false; val = lengths[i - 1]; rep = 3 + (int)b.get(2); }
                        else if (sym == 17) { if (!b.need(3)) return false; rep = 3 + (int)b.get(3); }
                        else { if (!b.need(7)) return false; rep = 11 + (int)b.get(7); }
                        if (i + rep > nlen + ndist) return false;
                        while (rep--) lengths[i++] = val;
                    }
                }
                lit.build(lengths, nlen);
                dist.build(lengths + nlen, ndist);
            }
            for (;;) {
                const int sym = lit.decode(b);
                if (sym < 0) return false;
                if (sym < 256) { out.push_back((uint8_t)sym); continue; }
                if (sym == 256) break;
                const int ls = sym - 257;
                if (ls >= 29 || !b.need(len_extra[ls])) return false;
                const int len = len_base[ls] + (int)b.get(len_extra[ls]);
                const int ds = dist.decode(b);
                if (ds < 0 || ds >= 30 || !b.need(dist_extra[ds])) return false;
                const size_t d = dist_base[ds] + (size_t)b.get(dist_extra[ds]);
                if (d > out.size()) return false;
                for (int k = 0; k < len; ++k) out.push_back(out[out.size() - d]);
            }
        } else return false;
        if (last) return true;
    }
}

// ------------------------------------------------------------------------------------------------ sRGB bytes
const uint8_t* srgb_byte_to_linear_byte() {
    static uint8_t lut[256];
    static bool made = false;
    if (!made) {
        for (int c = 0; c < 256; ++c) {
            const float x = (float)c / 255.0f;
            const float lin = x <= 0.04045f ? x / 12.92f : std::pow((x + 0.055f) / 1.055f, 2.4f);
            lut[c] = (uint8_t)std::min<uint16_t>(255, (uint16_t)std::floor(lin * 255));
        }
        made = true;
    }
    return lut;
}

// ------------------------------------------------------------------------------------------------ PNG
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

DeviceImage load_packed_image(const std::string& path, bool already_linear) {
    std::vector<uint8_t> file;
    {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) io_fail(path, "cannot open the file");
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        file.resize(n > 0 ? (size_t)n : 0);
        const size_t got = file.empty() ? 0 : std::fread(file.data(), 1, file.size(), f);
        std::fclose(f);
        if (got != file.size()) io_fail(path, "short read");
    }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0)
        io_fail(path, "not a PNG file (this build decodes PNG only; other 8-bit formats and EXR / HDR go through the runtime's IG::Image)");
    uint32_t w = 0, h = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t n = be32(&file[pos]);
        const uint8_t* kind = &file[pos + 4];
        if (pos + 12 + (size_t)n > file.size()) io_fail(path, "truncated chunk");
        const uint8_t* body = &file[pos + 8];
        if (!std::memcmp(kind, "IHDR", 4) && n >= 13) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (!std::memcmp(kind, "IDAT", 4)) idat.insert(idat.end(), body, body + n);
        else if (!std::memcmp(kind, "PLTE", 4)) plte.assign(body, body + n);
        else if (!std::memcmp(kind, "tRNS", 4)) trns.assign(body, body + n);
        else if (!std::memcmp(kind, "IEND", 4)) break;
        pos += 12 + (size_t)n;
    }
    if (w == 0 || h == 0 || depth != 8 || interlace != 0 || !(ctype == 0 || ctype == 2 || ctype == 3 || ctype == 4 || ctype == 6))
        io_fail(path, "only 8-bit non-interlaced grey / grey+alpha / RGB / RGBA / palette PNG files are decoded here");
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (1 + (size_t)w * ch));
    if (!inflate_zlib(idat, raw) || raw.size() < (size_t)h * (1 + (size_t)w * ch)) io_fail(path, "corrupt image data");
    // ---- undo the scanline filters (RFC 2083 section 6)
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> px((size_t)h * stride);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(size_t)y * (1 + stride)];
        const int ft = line[0];
        uint8_t* cur = &px[(size_t)y * stride];
        const uint8_t* prev = y ? &px[(size_t)(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)ch) ? prev[i - ch] : 0;
            int v = line[1 + i];
            if (ft == 1) v += a;
            else if (ft == 2) v += b;
            else if (ft == 3) v += (a + b) >> 1;
            else if (ft == 4) { const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            else if (ft != 0) io_fail(path, "unknown scanline filter");
            cur[i] = (uint8_t)v;
        }
    }
    // ---- to what stb_image hands Image::loadAsPacked (Image.cpp:727-735): 1, 3 or 4 channels as they are, anything else re-read as RGBA
    const uint8_t* lut = already_linear ? nullptr : srgb_byte_to_linear_byte();
    auto col = [&](uint8_t v) { return lut ? lut[v] : v; };
    DeviceImage img;
    img.width = (int)w; img.height = (int)h;
    const bool mono = ctype == 0;
    img.format = mono ? IGB200_IMAGE_MONO8 : IGB200_IMAGE_RGBA8;
    img.bytes.resize((size_t)w * h * (mono ? 1 : 4));
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* src = &px[(size_t)(h - 1 - y) * stride];   // stbi_set_flip_vertically_on_load(1)
        uint8_t* dst = &img.bytes[(size_t)y * w * (mono ? 1 : 4)];
        for (uint32_t x = 0; x < w; ++x) {
            if (mono) { dst[x] = col(src[x]); continue; }
            uint8_t r, g, b, a = 255;
            if (ctype == 2) { r = src[3 * x]; g = src[3 * x + 1]; b = src[3 * x + 2]; }
            else if (ctype == 6) { r = src[4 * x]; g = src[4 * x + 1]; b = src[4 * x + 2]; a = src[4 * x + 3]; }
            else if (ctype == 4) { r = g = b = src[2 * x]; a = src[2 * x + 1]; }
            else {   // palette
                const size_t k = src[x];
                if (3 * k + 2 >= plte.size()) io_fail(path, "palette index out of range");
                r = plte[3 * k]; g = plte[3 * k + 1]; b = plte[3 * k + 2];
                if (k < trns.size()) a = trns[k];
            }
            dst[4 * x] = col(r); dst[4 * x + 1] = col(g); dst[4 * x + 2] = col(b); dst[4 * x + 3] = a;
        }
    }
    return img;
}

}  // namespace igbh

// ================================================================================================ OpenEXR (scanline, single part)
namespace igbh {
namespace {

float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s;
        else { int sh = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; ++sh; } bits = s | ((uint32_t)(113 - sh) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) bits = s | 0x7F800000u | (m << 13);
    else bits = s | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

// ---- PIZ: Huffman coder of 16-bit symbols (OpenEXR ImfHuf) ------------------------------------------------------
constexpr int HUF_ENCBITS = 16, HUF_DECBITS = 14, HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS, HUF_DECMASK = HUF_DECSIZE - 1;
struct HufDec { int len = 0, lit = 0; std::vector<int> p; };   // len != 0: short code (lit = symbol); else p = symbols with long codes sharing this prefix

struct BitReader {
    const uint8_t* p; const uint8_t* end; uint64_t c = 0; int lc = 0;
    bool fill() { if (p >= end) return false; c = (c << 8) | *p++; lc += 8; return true; }
    bool get(int n, uint64_t& out) { while (lc < n) if (!fill()) return false; lc -= n; out = (c >> lc) & ((1ull << n) - 1); return true; }
};

bool huf_unpack_enc_table(const uint8_t*& p, const uint8_t* end, int im, int iM, std::vector<uint64_t>& hcode) {
    BitReader br{p, end};
    for (; im <= iM; ++im) {
        uint64_t l;
        if (!br.get(6, l)) return false;
        hcode[im] = l;
        if (l == 63) {   // long run of zero-length codes
            uint64_t n;
            if (!br.get(8, n)) return false;
            int zerun = (int)n + 6;
            if (im + zerun > iM + 1) return false;
            while (zerun--) hcode[im++] = 0;
            --im;
        } else if (l >= 59) {   // short run
            int zerun = (int)l - 59 + 2;
            if (im + zerun > iM + 1) return false;
            while (zerun--) hcode[im++] = 0;
            --im;
        }
    }
    p = br.p;   // whole bytes consumed (the table ends on a byte boundary of the reader)
    return true;
}
void huf_canonical_code_table(std::vector<uint64_t>& hcode) {   // code lengths -> (code << 6) | length
    uint64_t n[59] = {0};
    for (int i = 0; i < HUF_ENCSIZE; ++i) n[hcode[i]] += 1;
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) { const uint64_t nc = (c + n[i]) >> 1; n[i] = c; c = nc; }
    for (int i = 0; i < HUF_ENCSIZE; ++i) { const int l = (int)hcode[i]; if (l > 0) hcode[i] = (uint64_t)l | (n[l]++ << 6); }
}
bool huf_build_dec_table(const std::vector<uint64_t>& hcode, int im, int iM, std::vector<HufDec>& hdec) {
    for (; im <= iM; ++im) {
        const uint64_t c = hcode[im] >> 6;
        const int l = (int)(hcode[im] & 63);
        if (c >> l) return false;
        if (l > HUF_DECBITS) {
            HufDec& pl = hdec[(size_t)(c >> (l - HUF_DECBITS))];
            if (pl.len) return false;
            pl.lit++;
            pl.p.push_back(im);
        } else if (l) {
            const size_t base = (size_t)(c << (HUF_DECBITS - l));
            for (uint64_t i = 0; i < (1ull << (HUF_DECBITS - l)); ++i) {
                HufDec& pl = hdec[base + (size_t)i];
                if (pl.len || !pl.p.empty()) return false;
                pl.len = l; pl.lit = im;
            }
        }
    }
    return true;
}
bool huf_decode(const std::vector<uint64_t>& hcode, const std::vector<HufDec>& hdec, const uint8_t* in, int ni /*bits*/, int rlc, int no, uint16_t* out) {
    uint16_t* outb = out; uint16_t* oe = out + no;
    const uint8_t* ie = in + (ni + 7) / 8;
    uint64_t c = 0; int lc = 0;
    auto emit = [&](int po) -> bool {   // one decoded symbol; the run-length symbol repeats the last value
        if (po == rlc) {
            if (lc < 8) { if (in >= ie) return false; c = (c << 8) | *in++; lc += 8; }
            lc -= 8;
            const int cs = (int)((c >> lc) & 0xFF);
            if (out + cs > oe || out == outb) return false;
            const uint16_t s = out[-1];
            for (int k = 0; k < cs; ++k) *out++ = s;
        } else { if (out >= oe) return false; *out++ = (uint16_t)po; }
        return true;
    };
    while (in < ie) {
        c = (c << 8) | *in++; lc += 8;
        while (lc >= HUF_DECBITS) {
            const HufDec& pl = hdec[(size_t)((c >> (lc - HUF_DECBITS)) & HUF_DECMASK)];
            if (pl.len) { lc -= pl.len; if (!emit(pl.lit)) return false; }
            else {
                if (pl.p.empty()) return false;
                size_t j = 0;
                for (; j < pl.p.size(); ++j) {
                    const int l = (int)(hcode[(size_t)pl.p[j]] & 63);
                    while (lc < l && in < ie) { c = (c << 8) | *in++; lc += 8; }
                    if (lc >= l && (hcode[(size_t)pl.p[j]] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1))) { lc -= l; if (!emit(pl.p[j])) return false; break; }
                }
                if (j == pl.p.size()) return false;
            }
        }
    }
    const int i = (8 - ni) & 7;   // the last byte's padding bits
    c >>= i; lc -= i;
    while (lc > 0) {
        const HufDec& pl = hdec[(size_t)((c << (HUF_DECBITS - lc)) & HUF_DECMASK)];
        if (!pl.len) return false;
        lc -= pl.len;
        if (lc < 0) return false;
        if (!emit(pl.lit)) return false;
    }
    return out == oe;
}
bool huf_uncompress(const uint8_t* data, int n, uint16_t* raw, int n_raw) {
    if (n == 0) return n_raw == 0;
    if (n < 20) return false;
    auto rd = [&](int o) { return (int)((uint32_t)data[o] | ((uint32_t)data[o + 1] << 8) | ((uint32_t)data[o + 2] << 16) | ((uint32_t)data[o + 3] << 24)); };
    const int im = rd(0), iM = rd(4), n_bits = rd(12);
    if (im < 0 || im >= HUF_ENCSIZE || iM < 0 || iM >= HUF_ENCSIZE) return false;
    const uint8_t* p = data + 20; const uint8_t* end = data + n;
    std::vector<uint64_t> hcode(HUF_ENCSIZE, 0);
    if (!huf_unpack_enc_table(p, end, im, iM, hcode)) return false;
    if (n_bits > 8 * (int)(end - p)) return false;
    huf_canonical_code_table(hcode);
    std::vector<HufDec> hdec(HUF_DECSIZE);
    if (!huf_build_dec_table(hcode, im, iM, hdec)) return false;
    return huf_decode(hcode, hdec, p, n_bits, iM, n_raw, raw);
}

// ---- PIZ: 2-D wavelet (OpenEXR ImfWav) -----------------------------------------------------------------------------------
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const short ls = (short)l, hs = (short)h;
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(short)ai; b = (uint16_t)(short)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xFFFF, aa = (d + bb - (1 << 15)) & 0xFFFF;
    b = (uint16_t)bb; a = (uint16_t)aa;
}
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1;
    while (p <= n) p <<= 1;
    p >>= 1;
    int p2 = p;
    p >>= 1;
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* ey = in + oy * (ny - p2);
        const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1; uint16_t* p10 = px + oy1; uint16_t* p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

struct Channel { std::string name; int type; int size; };   // type: 0 uint, 1 half, 2 float

bool piz_uncompress(const uint8_t* in, int n_in, const std::vector<Channel>& ch, int width, int lines, std::vector<uint8_t>& out) {
    if (n_in < 4) return false;
    const int min_nz = in[0] | (in[1] << 8), max_nz = in[2] | (in[3] << 8);
    std::vector<uint8_t> bitmap(8192, 0);
    const uint8_t* p = in + 4;
    if (min_nz <= max_nz) { if (max_nz >= 8192 || p + (max_nz - min_nz + 1) > in + n_in) return false; std::memcpy(&bitmap[(size_t)min_nz], p, (size_t)(max_nz - min_nz + 1)); p += max_nz - min_nz + 1; }
    std::vector<uint16_t> lut(65536, 0);
    int k = 0;
    for (int i = 0; i < 65536; ++i) if (i == 0 || (bitmap[(size_t)(i >> 3)] & (1 << (i & 7)))) lut[(size_t)k++] = (uint16_t)i;
    const uint16_t max_value = (uint16_t)(k - 1);
    if (p + 4 > in + n_in) return false;
    const int length = (int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    p += 4;
    if (length < 0 || p + length > in + n_in) return false;
    size_t total = 0;
    for (const Channel& c : ch) total += (size_t)width * lines * (c.size / 2);
    std::vector<uint16_t> tmp(total);
    if (!huf_uncompress(p, length, tmp.data(), (int)total)) return false;
    size_t off = 0;
    std::vector<size_t> start(ch.size());
    for (size_t c = 0; c < ch.size(); ++c) {
        start[c] = off;
        const int sz = ch[c].size / 2;
        for (int j = 0; j < sz; ++j) wav2_decode(tmp.data() + off + j, width, sz, lines, width * sz, max_value);
        off += (size_t)width * lines * sz;
    }
    for (uint16_t& v : tmp) v = lut[v];
    out.resize(total * 2);
    uint8_t* o = out.data();
    std::vector<size_t> pos = start;
    for (int y = 0; y < lines; ++y)
        for (size_t c = 0; c < ch.size(); ++c) {
            const size_t nw = (size_t)width * (ch[c].size / 2);
            std::memcpy(o, tmp.data() + pos[c], nw * 2);
            o += nw * 2; pos[c] += nw;
        }
    return true;
}

bool rle_uncompress(const uint8_t* in, int n_in, std::vector<uint8_t>& out, size_t expect) {
    out.clear();
    const uint8_t* end = in + n_in;
    while (in < end) {
        const int8_t c = (int8_t)*in++;
        if (c < 0) { const int n = -c; if (in + n > end) return false; out.insert(out.end(), in, in + n); in += n; }
        else { if (in >= end) return false; out.insert(out.end(), (size_t)c + 1, *in++); }
        if (out.size() > expect) return false;
    }
    return out.size() == expect;
}
void unpredict_and_interleave(std::vector<uint8_t>& buf) {   // the filter ZIP and RLE apply before compressing
    for (size_t i = 1; i < buf.size(); ++i) buf[i] = (uint8_t)(buf[i - 1] + buf[i] - 128);
    std::vector<uint8_t> out(buf.size());
    const size_t half = (buf.size() + 1) / 2;
    for (size_t i = 0, a = 0, b = half; i < buf.size();) { out[i++] = buf[a++]; if (i < buf.size()) out[i++] = buf[b++]; }
    buf.swap(out);
}

}  // namespace

// Radiance RGBE (.hdr), as stb_image decodes it for IG::Image (flat and run-length encoded scanlines, -Y +X orientation; value = mantissa * 2^(e - 136))
static FloatImage load_radiance_hdr(const std::string& path, const std::vector<uint8_t>& f) {
    size_t pos = 0;
    auto line = [&]() { std::string s; while (pos < f.size() && f[pos] != '\n') s += (char)f[pos++]; ++pos; return s; };
    bool format_ok = false;
    for (;;) {
        if (pos >= f.size()) io_fail(path, "truncated Radiance header");
        const std::string l = line();
        if (l.empty()) break;
        if (l == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    }
    if (!format_ok) io_fail(path, "only 32-bit_rle_rgbe Radiance files are decoded here");
    int h = 0, w = 0;
    if (std::sscanf(line().c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) io_fail(path, "only the -Y +X orientation of Radiance files is decoded here");
    FloatImage img;
    img.width = w; img.height = h;
    img.rgba.assign((size_t)w * h * 4, 1.0f);
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int y = 0; y < h; ++y) {
        if (pos + 4 > f.size()) io_fail(path, "truncated Radiance data");
        if (w >= 8 && w < 32768 && f[pos] == 2 && f[pos + 1] == 2 && !(f[pos + 2] & 0x80) && ((f[pos + 2] << 8) | f[pos + 3]) == w) {   // new run-length encoding: the four components one after another
            pos += 4;
            for (int k = 0; k < 4; ++k)
                for (int x = 0; x < w;) {
                    if (pos >= f.size()) io_fail(path, "truncated Radiance data");
                    int count = f[pos++];
                    if (count > 128) { count -= 128; if (pos >= f.size() || x + count > w) io_fail(path, "corrupt Radiance run"); const uint8_t v = f[pos++]; while (count--) scan[(size_t)(x++) * 4 + k] = v; }
                    else { if (count == 0 || pos + count > f.size() || x + count > w) io_fail(path, "corrupt Radiance run"); while (count--) scan[(size_t)(x++) * 4 + k] = f[pos++]; }
                }
        } else {   // flat
            if (pos + (size_t)w * 4 > f.size()) io_fail(path, "truncated Radiance data");
            std::memcpy(scan.data(), &f[pos], (size_t)w * 4);
            pos += (size_t)w * 4;
        }
        float* dst = &img.rgba[(size_t)(h - 1 - y) * w * 4];   // rows bottom-up
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &scan[(size_t)x * 4];
            if (p[3] == 0) { dst[4 * x] = dst[4 * x + 1] = dst[4 * x + 2] = 0; continue; }
            const float s1 = std::ldexp(1.0f, (int)p[3] - 136);
            dst[4 * x] = p[0] * s1; dst[4 * x + 1] = p[1] * s1; dst[4 * x + 2] = p[2] * s1;
        }
    }
    return img;
}

FloatImage load_float_image(const std::string& path) {
    std::vector<uint8_t> f;
    {
        FILE* fh = std::fopen(path.c_str(), "rb");
        if (!fh) io_fail(path, "cannot open the file");
        std::fseek(fh, 0, SEEK_END);
        const long n = std::ftell(fh);
        std::fseek(fh, 0, SEEK_SET);
        f.resize(n > 0 ? (size_t)n : 0);
        const size_t got = f.empty() ? 0 : std::fread(f.data(), 1, f.size(), fh);
        std::fclose(fh);
        if (got != f.size()) io_fail(path, "short read");
    }
    auto u32 = [&](size_t o) -> uint32_t { if (o + 4 > f.size()) io_fail(path, "truncated file"); return (uint32_t)f[o] | ((uint32_t)f[o + 1] << 8) | ((uint32_t)f[o + 2] << 16) | ((uint32_t)f[o + 3] << 24); };
    if (f.size() > 10 && (!std::memcmp(f.data(), "#?RADIANCE", 10) || !std::memcmp(f.data(), "#?RGBE", 6))) return load_radiance_hdr(path, f);
    if (f.size() < 8 || u32(0) != 20000630u) io_fail(path, "not an OpenEXR or Radiance HDR file (the float formats this build decodes)");
    const uint32_t version = u32(4);
    if ((version & 0xFF) != 2 || (version & 0x200) || (version & 0x800) || (version & 0x1000)) io_fail(path, "tiled, deep or multi-part OpenEXR files are not decoded here");
    // ---- header attributes
    std::vector<Channel> channels;
    int compression = -1, line_order = 0, xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    size_t pos = 8;
    auto cstr = [&]() { std::string s; while (pos < f.size() && f[pos]) s += (char)f[pos++]; if (pos >= f.size()) io_fail(path, "truncated header"); ++pos; return s; };
    for (;;) {
        const std::string name = cstr();
        if (name.empty()) break;
        const std::string type = cstr();
        const uint32_t size = u32(pos); pos += 4;
        if (pos + size > f.size()) io_fail(path, "truncated header");
        const size_t v = pos;
        if (name == "channels") {
            size_t q = v;
            while (q < v + size && f[q]) {
                Channel c;
                while (f[q]) c.name += (char)f[q++];
                ++q;
                c.type = (int)u32(q); q += 4 + 4;
                const int xs = (int)u32(q), ys = (int)u32(q + 4); q += 8;
                if (xs != 1 || ys != 1) io_fail(path, "sub-sampled channels are not decoded here");
                if (c.type < 0 || c.type > 2) io_fail(path, "unknown channel type");
                c.size = c.type == 1 ? 2 : 4;
                channels.push_back(c);
            }
        } else if (name == "compression") compression = f[v];
        else if (name == "lineOrder") line_order = f[v];
        else if (name == "dataWindow") { xmin = (int)u32(v); ymin = (int)u32(v + 4); xmax = (int)u32(v + 8); ymax = (int)u32(v + 12); }
        pos = v + size;
    }
    const int width = xmax - xmin + 1, height = ymax - ymin + 1;
    if (channels.empty() || width <= 0 || height <= 0) io_fail(path, "incomplete OpenEXR header");
    int lines_per_chunk;
    switch (compression) {
        case 0: case 1: case 2: lines_per_chunk = 1; break;    // NONE, RLE, ZIPS
        case 3: lines_per_chunk = 16; break;                    // ZIP
        case 4: lines_per_chunk = 32; break;                    // PIZ
        default: io_fail(path, "OpenEXR compression " + std::to_string(compression) + " (PXR24 / B44 / DWA) is not decoded here");
    }
    size_t line_bytes = 0;
    for (const Channel& c : channels) line_bytes += (size_t)width * c.size;
    const int n_chunks = (height + lines_per_chunk - 1) / lines_per_chunk;
    // which file channel feeds R, G, B, A (a single luminance channel feeds all three colours)
    int src[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < channels.size(); ++c) {
        const std::string& n = channels[c].name;
        if (n == "R") src[0] = (int)c; else if (n == "G") src[1] = (int)c; else if (n == "B") src[2] = (int)c; else if (n == "A") src[3] = (int)c;
        else if (n == "Y" && src[0] < 0) src[0] = src[1] = src[2] = (int)c;
    }
    if (src[0] < 0 || src[1] < 0 || src[2] < 0) io_fail(path, "no R, G, B (or Y) channels");
    (void)line_order;   // every chunk carries its own y
    FloatImage img;
    img.width = width; img.height = height;
    img.rgba.assign((size_t)width * height * 4, 1.0f);
    const size_t table = pos;
    for (int k = 0; k < n_chunks; ++k) {
        if (table + 8 * (size_t)k + 8 > f.size()) io_fail(path, "truncated offset table");
        uint64_t off = 0;
        for (int b = 7; b >= 0; --b) off = (off << 8) | f[table + 8 * (size_t)k + (size_t)b];
        if (off + 8 > f.size()) io_fail(path, "chunk offset outside the file");
        const int y0 = (int)u32((size_t)off) - ymin;
        const int n = (int)u32((size_t)off + 4);
        if (y0 < 0 || y0 >= height || n < 0 || off + 8 + (uint64_t)n > f.size()) io_fail(path, "corrupt chunk");
        const int lines = std::min(lines_per_chunk, height - y0);
        const size_t expect = line_bytes * (size_t)lines;
        const uint8_t* data = &f[(size_t)off + 8];
        std::vector<uint8_t> raw;
        if ((size_t)n == expect || compression == 0) raw.assign(data, data + n);   // stored as it is when compressing did not help
        else if (compression == 1) { if (!rle_uncompress(data, n, raw, expect)) io_fail(path, "corrupt RLE chunk"); unpredict_and_interleave(raw); }
        else if (compression == 2 || compression == 3) {
            std::vector<uint8_t> in(data, data + n);
            if (!inflate_zlib(in, raw) || raw.size() != expect) io_fail(path, "corrupt ZIP chunk");
            unpredict_and_interleave(raw);
        } else if (!piz_uncompress(data, n, channels, width, lines, raw) || raw.size() != expect) io_fail(path, "corrupt PIZ chunk");
        if (raw.size() != expect) io_fail(path, "chunk of unexpected size");
        for (int l = 0; l < lines; ++l) {
            const uint8_t* line = &raw[(size_t)l * line_bytes];
            const int y = y0 + l;
            float* dst = &img.rgba[(size_t)(height - 1 - y) * width * 4];   // Image::flipY: rows bottom-up
            size_t coff = 0;
            for (size_t c = 0; c < channels.size(); ++c) {
                const uint8_t* cp = line + coff;
                coff += (size_t)width * channels[c].size;
                for (int t = 0; t < 4; ++t) {
                    if (src[t] != (int)c) continue;
                    for (int x = 0; x < width; ++x) {
                        float v;
                        if (channels[c].type == 1) v = half_to_float((uint16_t)(cp[2 * x] | (cp[2 * x + 1] << 8)));
                        else if (channels[c].type == 2) std::memcpy(&v, cp + 4 * x, 4);
                        else { uint32_t u; std::memcpy(&u, cp + 4 * x, 4); v = (float)u; }
                        dst[4 * x + t] = v;
                    }
                }
            }
        }
    }
    return img;
}

}  // namespace igbh
