// exr_codec.h — minimal OpenEXR scan-line reader/writer (host side, header-only).
//
// Why this exists: the reference loads/saves EXR through the OpenEXR library
// (reference src/bitmap.cpp:242-288, Imf::RgbaInputFile / RgbaOutputFile). That
// library is not available in this image, and the bundled cubemaps
// (data/env/forest, data/env/ocean) are PIZ-compressed HALF RGBA scan-line files,
// so this file re-states the published OpenEXR file layout and the PIZ
// (bitmap LUT + Haar wavelet + canonical Huffman), ZIP/ZIPS and RLE block codecs.
// It is validated against OpenCV's independent OpenEXR build in
// tests/test_exr_codec.py.
//
// Reader: single-part scan-line images, channels of type HALF/FLOAT/UINT,
//         compression NONE, RLE, ZIPS, ZIP, PIZ. Output: float RGBA, top-down.
// Writer: uncompressed HALF RGBA scan-line (alpha = 1), which every EXR reader
//         accepts (reference saveEXR writes HALF RGBA too, bitmap.cpp:270-288).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include <zlib.h>

namespace hxr {
namespace exr {

struct Image {
    int width = 0, height = 0;
    std::vector<float> rgba;  // 4 floats per pixel, row-major, top-down
};

// ---------------------------------------------------------------- half <-> float
inline float half_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h >> 15) << 31;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ff;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400));
            man &= 0x3ff;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

inline uint16_t float_to_half(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    uint16_t sign = (uint16_t)((x >> 16) & 0x8000);
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffff;
    if (((x >> 23) & 0xff) == 0xff) return sign | 0x7c00 | (man ? 0x200 : 0);  // inf / nan
    if (exp >= 31) return sign | 0x7c00;                                        // overflow -> inf
    if (exp <= 0) {
        if (exp < -10) return sign;  // underflow -> 0
        man |= 0x800000;
        int shift = 14 - exp;
        uint32_t r = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (r & 1))) r++;
        return sign | (uint16_t)r;
    }
    uint32_t r = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fff;
    if (rem > 0x1000 || (rem == 0x1000 && (r & 1))) r++;  // round to nearest even (may carry into exp)
    return sign | (uint16_t)r;
}

// ---------------------------------------------------------------- PIZ pieces
namespace piz {

const int HUF_ENCBITS = 16, HUF_DECBITS = 14;
const int HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS;
const int HUF_DECMASK = HUF_DECSIZE - 1;
const int SHORT_ZEROCODE_RUN = 59, LONG_ZEROCODE_RUN = 63;
const int SHORTEST_LONG_RUN = 2 + LONG_ZEROCODE_RUN - SHORT_ZEROCODE_RUN;

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t c = 0;
    int lc = 0;
    uint32_t get(int n)
    {
        while (lc < n) {
            if (p >= end) throw std::runtime_error("exr/piz: bitstream underrun");
            c = (c << 8) | *p++;
            lc += 8;
        }
        lc -= n;
        return (uint32_t)((c >> lc) & ((1u << n) - 1));
    }
};

// canonical code assignment: hcode[i] holds the code length on input,
// (code << 6 | length) on output.
inline void canonical_table(std::vector<uint64_t>& hcode)
{
    uint64_t n[59] = {0};
    for (int i = 0; i < HUF_ENCSIZE; i++) n[hcode[i]]++;
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) {
        uint64_t nc = (c + n[i]) >> 1;
        n[i] = c;
        c = nc;
    }
    for (int i = 0; i < HUF_ENCSIZE; i++) {
        int l = (int)hcode[i];
        if (l > 0) hcode[i] = (uint64_t)l | (n[l]++ << 6);
    }
}

struct DecEntry {
    int len = 0;   // short code length, 0 => long-code bucket
    int lit = 0;   // symbol (short) or count (long)
    std::vector<int> longs;
};

inline void huf_uncompress(const uint8_t* in, int nIn, uint16_t* out, int nOut)
{
    if (nIn == 0) {
        if (nOut != 0) throw std::runtime_error("exr/piz: empty huffman block");
        return;
    }
    if (nIn < 20) throw std::runtime_error("exr/piz: short huffman header");
    auto rd32 = [&](int off) { uint32_t v; memcpy(&v, in + off, 4); return v; };
    int im = (int)rd32(0), iM = (int)rd32(4);
    int nBits = (int)rd32(12);
    if (im < 0 || im >= HUF_ENCSIZE || iM < 0 || iM >= HUF_ENCSIZE) throw std::runtime_error("exr/piz: bad symbol range");
    const uint8_t* ptr = in + 20;
    const uint8_t* end = in + nIn;

    // unpack the 6-bit packed code-length table (with zero-run escapes)
    std::vector<uint64_t> hcode(HUF_ENCSIZE, 0);
    {
        BitReader br{ptr, end};
        for (int s = im; s <= iM; s++) {
            int l = (int)br.get(6);
            hcode[s] = l;
            if (l == LONG_ZEROCODE_RUN) {
                int zerun = (int)br.get(8) + SHORTEST_LONG_RUN;
                if (s + zerun > iM + 1) throw std::runtime_error("exr/piz: table overrun");
                while (zerun--) hcode[s++] = 0;
                s--;
            } else if (l >= SHORT_ZEROCODE_RUN) {
                int zerun = l - SHORT_ZEROCODE_RUN + 2;
                if (s + zerun > iM + 1) throw std::runtime_error("exr/piz: table overrun");
                while (zerun--) hcode[s++] = 0;
                s--;
            }
        }
        ptr = br.p;
    }
    canonical_table(hcode);
    if (nBits > 8 * (int)(end - ptr)) throw std::runtime_error("exr/piz: bad bit count");

    // decoding table: 14-bit primary lookup, overflow lists for longer codes
    std::vector<DecEntry> dec(HUF_DECSIZE);
    for (int s = im; s <= iM; s++) {
        uint64_t c = hcode[s] >> 6;
        int l = (int)(hcode[s] & 63);
        if (c >> l) throw std::runtime_error("exr/piz: invalid code");
        if (l > HUF_DECBITS) {
            DecEntry& e = dec[c >> (l - HUF_DECBITS)];
            if (e.len) throw std::runtime_error("exr/piz: code clash");
            e.lit++;
            e.longs.push_back(s);
        } else if (l) {
            size_t base = (size_t)(c << (HUF_DECBITS - l));
            for (size_t i = 0; i < ((size_t)1 << (HUF_DECBITS - l)); i++) {
                dec[base + i].len = l;
                dec[base + i].lit = s;
            }
        }
    }

    // decode
    const int rlc = iM;  // the encoder's run-length pseudo symbol
    uint64_t c = 0;
    int lc = 0;
    uint16_t* o = out;
    uint16_t* oe = out + nOut;
    const uint8_t* ie = ptr + (nBits + 7) / 8;
    auto emit = [&](int sym) {
        if (sym == rlc) {
            if (lc < 8) {
                if (ptr >= end) throw std::runtime_error("exr/piz: underrun");
                c = (c << 8) | *ptr++;
                lc += 8;
            }
            lc -= 8;
            int cs = (int)((c >> lc) & 0xff);
            if (o + cs > oe || o == out) throw std::runtime_error("exr/piz: bad run");
            uint16_t s = o[-1];
            while (cs-- > 0) *o++ = s;
        } else {
            if (o >= oe) throw std::runtime_error("exr/piz: too much data");
            *o++ = (uint16_t)sym;
        }
    };
    while (ptr < ie) {
        c = (c << 8) | *ptr++;
        lc += 8;
        while (lc >= HUF_DECBITS) {
            const DecEntry& e = dec[(c >> (lc - HUF_DECBITS)) & HUF_DECMASK];
            if (e.len) {
                lc -= e.len;
                emit(e.lit);
            } else {
                if (e.longs.empty()) throw std::runtime_error("exr/piz: invalid long code");
                size_t j = 0;
                for (; j < e.longs.size(); j++) {
                    int l = (int)(hcode[e.longs[j]] & 63);
                    while (lc < l && ptr < ie) {
                        c = (c << 8) | *ptr++;
                        lc += 8;
                    }
                    if (lc >= l && (hcode[e.longs[j]] >> 6) == ((c >> (lc - l)) & (((uint64_t)1 << l) - 1))) {
                        lc -= l;
                        emit(e.longs[j]);
                        break;
                    }
                }
                if (j == e.longs.size()) throw std::runtime_error("exr/piz: long code not found");
            }
        }
    }
    int i = (8 - nBits) & 7;
    c >>= i;
    lc -= i;
    while (lc > 0) {
        const DecEntry& e = dec[(c << (HUF_DECBITS - lc)) & HUF_DECMASK];
        if (!e.len) throw std::runtime_error("exr/piz: invalid tail code");
        lc -= e.len;
        emit(e.lit);
    }
    if (o != oe) throw std::runtime_error("exr/piz: not enough data");
}

// inverse of the 14-bit / 16-bit integer Haar lifting steps
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
    int16_t ls = (int16_t)l, hs = (int16_t)h;
    int hi = hs;
    int ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai;
    b = (uint16_t)(int16_t)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
    int m = l, d = h;
    int bb = (m - (d >> 1)) & 0xffff;
    int aa = (d + bb - (1 << 15)) & 0xffff;
    b = (uint16_t)bb;
    a = (uint16_t)aa;
}

inline void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx)
{
    bool w14 = mx < (1 << 14);
    int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1;
    p2 = p;
    p >>= 1;
    auto dec = [&](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
        if (w14) wdec14(l, h, a, b); else wdec16(l, h, a, b);
    };
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* ey = in + oy * (ny - p2);
        int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                uint16_t* p10 = px + oy1;
                uint16_t* p11 = p10 + ox1;
                dec(*px, *p10, i00, i10);
                dec(*p01, *p11, i01, i11);
                dec(i00, i01, *px, *p01);
                dec(i10, i11, *p10, *p11);
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                dec(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py;
            uint16_t* ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                dec(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

}  // namespace piz

// ---------------------------------------------------------------- reader
struct Channel {
    std::string name;
    int type;  // 0 uint, 1 half, 2 float
    int bytes() const { return type == 1 ? 2 : 4; }
};

inline void undo_predictor_and_interleave(std::vector<uint8_t>& buf)
{
    size_t n = buf.size();
    for (size_t i = 1; i < n; i++) buf[i] = (uint8_t)(buf[i - 1] + buf[i] - 128);
    std::vector<uint8_t> out(n);
    size_t half = (n + 1) / 2;
    for (size_t i = 0, a = 0, b = half; i < n;) {
        out[i++] = buf[a++];
        if (i < n) out[i++] = buf[b++];
    }
    buf.swap(out);
}

inline bool load(const char* filename, Image& img, std::string* err = nullptr)
{
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
    FILE* f = fopen(filename, "rb");
    if (!f) return fail("cannot open file");
    std::vector<uint8_t> data;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(sz > 0 ? sz : 0);
    if (sz > 0 && fread(data.data(), 1, sz, f) != (size_t)sz) { fclose(f); return fail("short read"); }
    fclose(f);
    try {
        size_t pos = 0;
        auto need = [&](size_t n) { if (pos + n > data.size()) throw std::runtime_error("exr: truncated file"); };
        auto rd32 = [&]() { need(4); int32_t v; memcpy(&v, &data[pos], 4); pos += 4; return v; };
        auto rdstr = [&]() { std::string s; for (;;) { need(1); char ch = (char)data[pos++]; if (!ch) break; s += ch; } return s; };
        if ((uint32_t)rd32() != 20000630u) throw std::runtime_error("exr: bad magic");
        uint32_t version = (uint32_t)rd32();
        if (version & 0x200) throw std::runtime_error("exr: tiled images unsupported");
        if (version & 0x1800) throw std::runtime_error("exr: multipart/deep unsupported");
        std::vector<Channel> channels;
        int compression = -1;
        int dw[4] = {0, 0, -1, -1};
        int lineOrder = 0;
        for (;;) {
            std::string name = rdstr();
            if (name.empty()) break;
            std::string type = rdstr();
            int size = rd32();
            need(size);
            size_t vpos = pos;
            if (name == "channels") {
                size_t p = vpos;
                while (p < vpos + size && data[p]) {
                    Channel c;
                    while (data[p]) c.name += (char)data[p++];
                    p++;
                    int32_t t; memcpy(&t, &data[p], 4);
                    c.type = t;
                    int32_t xs, ys; memcpy(&xs, &data[p + 8], 4); memcpy(&ys, &data[p + 12], 4);
                    if (xs != 1 || ys != 1) throw std::runtime_error("exr: subsampled channels unsupported");
                    p += 16;
                    channels.push_back(c);
                }
            } else if (name == "compression") {
                compression = data[vpos];
            } else if (name == "dataWindow") {
                memcpy(dw, &data[vpos], 16);
            } else if (name == "lineOrder") {
                lineOrder = data[vpos];
            }
            pos = vpos + size;
        }
        (void)lineOrder;  // chunks carry their own y; the offset table is not needed for a whole-file read
        int W = dw[2] - dw[0] + 1, H = dw[3] - dw[1] + 1;
        if (W <= 0 || H <= 0 || channels.empty()) throw std::runtime_error("exr: bad header");
        int linesPerBlock;
        switch (compression) {
            case 0: case 1: case 2: linesPerBlock = 1; break;
            case 3: linesPerBlock = 16; break;
            case 4: linesPerBlock = 32; break;
            default: throw std::runtime_error("exr: unsupported compression " + std::to_string(compression));
        }
        int nBlocks = (H + linesPerBlock - 1) / linesPerBlock;
        size_t bytesPerLine = 0;
        for (auto& c : channels) bytesPerLine += (size_t)c.bytes() * W;
        pos += (size_t)nBlocks * 8;  // skip the offset table
        img.width = W;
        img.height = H;
        img.rgba.assign((size_t)W * H * 4, 0.0f);
        for (size_t i = 0; i < (size_t)W * H; i++) img.rgba[i * 4 + 3] = 1.0f;
        int slot[4] = {-1, -1, -1, -1};
        for (size_t ci = 0; ci < channels.size(); ci++) {
            const std::string& n = channels[ci].name;
            if (n == "R") slot[0] = (int)ci; else if (n == "G") slot[1] = (int)ci;
            else if (n == "B") slot[2] = (int)ci; else if (n == "A") slot[3] = (int)ci;
            else if (n == "Y" && channels.size() <= 2) slot[0] = slot[1] = slot[2] = (int)ci;
        }
        for (int b = 0; b < nBlocks; b++) {
            int y0 = rd32();
            int csize = rd32();
            if (csize < 0) throw std::runtime_error("exr: bad chunk size");
            need(csize);
            const uint8_t* src = &data[pos];
            pos += csize;
            int row0 = y0 - dw[1];
            if (row0 < 0 || row0 >= H) throw std::runtime_error("exr: chunk outside data window");
            int nLines = std::min(linesPerBlock, H - row0);
            size_t rawSize = bytesPerLine * nLines;
            std::vector<uint8_t> raw(rawSize);
            if ((size_t)csize == rawSize || compression == 0) {
                if ((size_t)csize != rawSize) throw std::runtime_error("exr: raw chunk size mismatch");
                memcpy(raw.data(), src, rawSize);
            } else if (compression == 2 || compression == 3) {
                uLongf dl = (uLongf)rawSize;
                if (uncompress(raw.data(), &dl, src, (uLong)csize) != Z_OK || dl != rawSize)
                    throw std::runtime_error("exr: zlib error");
                undo_predictor_and_interleave(raw);
            } else if (compression == 1) {
                size_t o = 0;
                int i = 0;
                while (i < csize) {
                    int8_t cnt = (int8_t)src[i++];
                    if (cnt < 0) {
                        int n = -cnt;
                        if (i + n > csize || o + n > rawSize) throw std::runtime_error("exr: rle overrun");
                        memcpy(&raw[o], &src[i], n); o += n; i += n;
                    } else {
                        int n = cnt + 1;
                        if (i >= csize || o + n > rawSize) throw std::runtime_error("exr: rle overrun");
                        memset(&raw[o], src[i++], n); o += n;
                    }
                }
                if (o != rawSize) throw std::runtime_error("exr: rle size mismatch");
                undo_predictor_and_interleave(raw);
            } else {  // PIZ
                size_t nWords = rawSize / 2;
                std::vector<uint16_t> tmp(nWords);
                const uint8_t* p = src;
                const uint8_t* pe = src + csize;
                if (pe - p < 4) throw std::runtime_error("exr/piz: short block");
                uint16_t minNZ, maxNZ;
                memcpy(&minNZ, p, 2); memcpy(&maxNZ, p + 2, 2); p += 4;
                std::vector<uint8_t> bitmap(8192, 0);
                if (maxNZ >= 8192) throw std::runtime_error("exr/piz: bad bitmap range");
                if (minNZ <= maxNZ) {
                    size_t n = (size_t)maxNZ - minNZ + 1;
                    if ((size_t)(pe - p) < n) throw std::runtime_error("exr/piz: short bitmap");
                    memcpy(&bitmap[minNZ], p, n);
                    p += n;
                }
                std::vector<uint16_t> lut(65536, 0);
                int k = 0;
                for (int i = 0; i < 65536; i++)
                    if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
                uint16_t maxValue = (uint16_t)(k - 1);
                if (pe - p < 4) throw std::runtime_error("exr/piz: short block");
                int32_t hlen; memcpy(&hlen, p, 4); p += 4;
                if (hlen < 0 || hlen > pe - p) throw std::runtime_error("exr/piz: bad huffman length");
                piz::huf_uncompress(p, hlen, tmp.data(), (int)nWords);
                // per-channel planes inside the block: [channel][line][x][word]
                size_t off = 0;
                std::vector<size_t> chStart(channels.size());
                for (size_t ci = 0; ci < channels.size(); ci++) {
                    int wsz = channels[ci].bytes() / 2;
                    chStart[ci] = off;
                    for (int j = 0; j < wsz; j++)
                        piz::wav2_decode(&tmp[off + j], W, wsz, nLines, W * wsz, maxValue);
                    off += (size_t)W * nLines * wsz;
                }
                for (auto& v : tmp) v = lut[v];
                // back to the scan-line interleaved layout
                uint8_t* o = raw.data();
                std::vector<size_t> cur = chStart;
                for (int ly = 0; ly < nLines; ly++)
                    for (size_t ci = 0; ci < channels.size(); ci++) {
                        size_t n = (size_t)W * (channels[ci].bytes() / 2);
                        memcpy(o, &tmp[cur[ci]], n * 2);
                        o += n * 2;
                        cur[ci] += n;
                    }
            }
            // raw: for each line, for each channel (file order), W values
            const uint8_t* r = raw.data();
            for (int ly = 0; ly < nLines; ly++) {
                int y = row0 + ly;
                for (size_t ci = 0; ci < channels.size(); ci++) {
                    const Channel& c = channels[ci];
                    for (int s = 0; s < 4; s++) {
                        if (slot[s] != (int)ci) continue;
                        for (int x = 0; x < W; x++) {
                            float v;
                            if (c.type == 1) { uint16_t h; memcpy(&h, r + x * 2, 2); v = half_to_float(h); }
                            else if (c.type == 2) { memcpy(&v, r + x * 4, 4); }
                            else { uint32_t u; memcpy(&u, r + x * 4, 4); v = (float)u; }
                            img.rgba[((size_t)y * W + x) * 4 + s] = v;
                        }
                    }
                    r += (size_t)c.bytes() * W;
                }
            }
        }
        return true;
    } catch (std::exception& e) {
        img = Image();
        return fail(e.what());
    }
}

// ---------------------------------------------------------------- writer
// rows: when non-null, the already converted scan lines (per row: W halves of A, then B, G, R) are written as they are
inline bool save_half_impl(const char* filename, int W, int H, const float* rgb, int stride_floats, const uint16_t* rows)
{
    FILE* f = fopen(filename, "wb");
    if (!f) return false;
    std::vector<uint8_t> hdr;
    auto put = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; hdr.insert(hdr.end(), b, b + n); };
    auto put32 = [&](int32_t v) { put(&v, 4); };
    auto putstr = [&](const char* s) { put(s, strlen(s) + 1); };
    auto attr = [&](const char* name, const char* type, const void* v, int size) { putstr(name); putstr(type); put32(size); put(v, size); };
    put32(20000630);
    put32(2);
    {
        std::vector<uint8_t> ch;
        for (const char* n : {"A", "B", "G", "R"}) {
            ch.push_back((uint8_t)n[0]); ch.push_back(0);
            int32_t v[4] = {1, 0, 1, 1};  // HALF, pLinear+reserved, xSampling, ySampling
            const uint8_t* b = (const uint8_t*)v;
            ch.insert(ch.end(), b, b + 16);
        }
        ch.push_back(0);
        attr("channels", "chlist", ch.data(), (int)ch.size());
    }
    uint8_t comp = 0, lo = 0;
    attr("compression", "compression", &comp, 1);
    int32_t win[4] = {0, 0, W - 1, H - 1};
    attr("dataWindow", "box2i", win, 16);
    attr("displayWindow", "box2i", win, 16);
    attr("lineOrder", "lineOrder", &lo, 1);
    float one = 1.0f, centre[2] = {0, 0};
    attr("pixelAspectRatio", "float", &one, 4);
    attr("screenWindowCenter", "v2f", centre, 8);
    attr("screenWindowWidth", "float", &one, 4);
    hdr.push_back(0);
    size_t lineBytes = (size_t)W * 8;
    uint64_t base = hdr.size() + (uint64_t)H * 8;
    for (int y = 0; y < H; y++) {
        uint64_t off = base + (uint64_t)y * (8 + lineBytes);
        put(&off, 8);
    }
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    std::vector<uint16_t> line((size_t)W * 4);
    for (int y = 0; y < H && ok; y++) {
        const float* src = rows ? nullptr : rgb + (size_t)y * W * stride_floats;
        if (rows) memcpy(line.data(), rows + (size_t)y * W * 4, lineBytes);
        for (int x = 0; x < W && !rows; x++) {
            line[x] = float_to_half(1.0f);
            line[W + x] = float_to_half(src[x * stride_floats + 2]);
            line[2 * W + x] = float_to_half(src[x * stride_floats + 1]);
            line[3 * W + x] = float_to_half(src[x * stride_floats + 0]);
        }
        int32_t yy = y, n = (int32_t)lineBytes;
        ok = fwrite(&yy, 4, 1, f) == 1 && fwrite(&n, 4, 1, f) == 1 && fwrite(line.data(), 1, lineBytes, f) == lineBytes;
    }
    fclose(f);
    return ok;
}

inline bool save_half_rgba(const char* filename, int W, int H, const float* rgb, int stride_floats)
{
    return save_half_impl(filename, W, H, rgb, stride_floats, nullptr);
}
inline bool save_half_rows(const char* filename, int W, int H, const uint16_t* rows) { return save_half_impl(filename, W, H, nullptr, 0, rows); }

}  // namespace exr
}  // namespace hxr
