// image_io.cpp -- see image_io.h. Host-only C++17, no dependencies: an inflate (RFC 1951) and a PNG reader (RFC 2083) small enough to audit.
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
