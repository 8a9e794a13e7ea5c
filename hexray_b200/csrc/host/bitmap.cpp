// bitmap.cpp — see bitmap.h for the reference lines each routine follows.
#include "bitmap.h"
#include <cstdio>
#include <cstdint>
#include <cctype>
#include <mutex>
#include "exr_codec.h"

namespace hxr {
namespace host {

static inline int nearestInt(float x) { return (int)std::floor(x + 0.5f); }

unsigned convertTo8bit_sRGB(float x)
{
    if (x <= 0) return 0;
    if (x >= 1) return 255;
    // the linear toe multiplies by 12.02 (not the standard 12.92) — kept for output parity
    if (x <= 0.0031308f) x = x * 12.02f;
    else x = (1.0f + 0.055f) * powf(x, 1.0f / 2.4f) - 0.055f;
    return (unsigned)nearestInt(x * 255.0f);
}

unsigned convertTo8bit_sRGB_cached(float x)
{
    static unsigned char lut[4097];
    static std::once_flag once;
    std::call_once(once, [] { for (int i = 0; i <= 4096; i++) lut[i] = (unsigned char)convertTo8bit_sRGB(i / 4096.0f); });
    if (x <= 0) return 0;
    if (x >= 1) return 255;
    return lut[int(x * 4096.0f)];
}

float decompress_sRGB(float x)
{
    if (x <= 0 || x >= 1) return x;
    if (x <= 0.04045f) return x / 12.92f;
    return powf((x + 0.055f) / 1.055f, 2.4f);
}

std::string extensionUpper(const char* fileName)
{
    std::string s(fileName);
    if (s.size() < 2) return "";
    size_t dot = s.rfind('.');
    if (dot == std::string::npos) return "";
    std::string e = s.substr(dot + 1);
    for (auto& c : e) c = (char)toupper((unsigned char)c);
    return e;
}

void Bitmap::freeMem()
{
    m_width = m_height = -1;
    std::vector<Color3>().swap(m_data);
}

void Bitmap::generateEmptyImage(int w, int h)
{
    freeMem();
    if (w <= 0 || h <= 0) return;
    m_width = w;
    m_height = h;
    m_data.assign((size_t)w * h, Color3(0, 0, 0));
}

Color3 Bitmap::getPixel(int x, int y) const
{
    if (m_data.empty() || x < 0 || x >= m_width || y < 0 || y >= m_height) return Color3(0, 0, 0);
    return m_data[x + (size_t)y * m_width];
}

void Bitmap::setPixel(int x, int y, const Color3& c)
{
    if (m_data.empty() || x < 0 || x >= m_width || y < 0 || y >= m_height) return;
    m_data[x + (size_t)y * m_width] = c;
}

namespace {
#pragma pack(push, 1)
struct BmpFileHeader {
    uint16_t magic;
    int32_t fileSize, reserved, dataOffset;
};
struct BmpInfoHeader {
    int32_t headerSize, width, height;
    uint16_t planes, bpp;
    int32_t compression, imageSize, ppmX, ppmY, colors, importantColors;
};
#pragma pack(pop)
struct FileCloser {
    FILE* f;
    ~FileCloser() { if (f) fclose(f); }
};
}  // namespace

bool Bitmap::loadBMP(const char* filename)
{
    freeMem();
    FILE* fp = fopen(filename, "rb");
    if (!fp) {
        printf("loadBMP: Can't open file: `%s'\n", filename);
        return false;
    }
    FileCloser closer{fp};
    BmpFileHeader fh;
    BmpInfoHeader ih;
    if (fread(&fh, sizeof fh, 1, fp) != 1) return false;
    if (fh.magic != 19778) {
        printf("loadBMP: `%s' is not a BMP file.\n", filename);
        return false;
    }
    if (fread(&ih, sizeof ih, 1, fp) != 1) return false;
    if (!(ih.bpp == 8 || ih.bpp == 24 || ih.bpp == 32)) {
        printf("loadBMP: Cannot handle file format at %d bpp.\n", ih.bpp);
        return false;
    }
    if (ih.planes != 1) {
        printf("loadBMP: cannot load multichannel .bmp!\n");
        return false;
    }
    Color3 palette[256];
    int paletteEntries = 0;
    if (ih.bpp <= 8) {
        paletteEntries = ih.colors ? ih.colors : (1 << ih.bpp);
        if (paletteEntries > 256) return false;
        for (int i = 0; i < paletteEntries; i++) {
            uint32_t e;
            if (fread(&e, 4, 1, fp) != 1) return false;
            // entry is 0x00RRGGBB: blue in the low byte
            palette[i] = Color3(((e >> 16) & 0xff) / 255.0f, ((e >> 8) & 0xff) / 255.0f, (e & 0xff) / 255.0f);
        }
    }
    fseek(fp, fh.dataOffset - (54 + paletteEntries * 4), SEEK_CUR);
    const int k = ih.bpp / 8;
    int rowsz = ih.width * k;
    if (rowsz % 4) rowsz = (rowsz / 4 + 1) * 4;
    std::vector<unsigned char> row(rowsz);
    generateEmptyImage(ih.width, ih.height);
    if (!isOK()) {
        printf("loadBMP: cannot allocate memory for bitmap! Check file integrity!\n");
        return false;
    }
    for (int j = ih.height - 1; j >= 0; j--) {  // stored bottom-up
        if (fread(row.data(), 1, rowsz, fp) == 0) {
            printf("loadBMP: short read while opening `%s', file is probably incomplete!\n", filename);
            freeMem();
            return false;
        }
        for (int i = 0; i < ih.width; i++) {
            if (ih.bpp > 8) setPixel(i, j, Color3(row[i * k + 2] / 255.0f, row[i * k + 1] / 255.0f, row[i * k] / 255.0f));
            else setPixel(i, j, palette[row[i * k]]);
        }
    }
    return true;
}

bool Bitmap::saveBMP(const char* filename) const
{
    FILE* fp = fopen(filename, "wb");
    if (!fp) return false;
    int rowsz = m_width * 3;
    if (rowsz % 4) rowsz += 4 - (rowsz % 4);
    BmpFileHeader fh = {19778, rowsz * m_height + 54, 0, 54};
    BmpInfoHeader ih = {40, m_width, m_height, 1, 24, 0, 0, 0, 0, 0, 0};
    fwrite(&fh, sizeof fh, 1, fp);
    fwrite(&ih, sizeof ih, 1, fp);
    std::vector<unsigned char> row(rowsz, 0);
    for (int y = m_height - 1; y >= 0; y--) {
        for (int x = 0; x < m_width; x++) {
            Color3 c = getPixel(x, y);
            row[x * 3 + 0] = (unsigned char)convertTo8bit_sRGB_cached(c.b);
            row[x * 3 + 1] = (unsigned char)convertTo8bit_sRGB_cached(c.g);
            row[x * 3 + 2] = (unsigned char)convertTo8bit_sRGB_cached(c.r);
        }
        fwrite(row.data(), rowsz, 1, fp);
    }
    fclose(fp);
    return true;
}

bool writeBmpFile(const char* filename, int width, int height, int rowsz, const unsigned char* rows)
{
    FILE* fp = fopen(filename, "wb");
    if (!fp) return false;
    BmpFileHeader fh = {19778, rowsz * height + 54, 0, 54};
    BmpInfoHeader ih = {40, width, height, 1, 24, 0, 0, 0, 0, 0, 0};
    bool ok = fwrite(&fh, sizeof fh, 1, fp) == 1 && fwrite(&ih, sizeof ih, 1, fp) == 1;
    ok = ok && fwrite(rows, (size_t)rowsz, (size_t)height, fp) == (size_t)height;
    fclose(fp);
    return ok;
}

bool Bitmap::loadEXR(const char* filename)
{
    exr::Image img;
    if (!exr::load(filename, img)) {
        m_width = m_height = 0;
        m_data.clear();
        return false;
    }
    m_width = img.width;
    m_height = img.height;
    m_data.resize((size_t)m_width * m_height);
    for (size_t i = 0; i < m_data.size(); i++) m_data[i] = Color3(img.rgba[i * 4], img.rgba[i * 4 + 1], img.rgba[i * 4 + 2]);
    return true;
}

bool Bitmap::saveEXR(const char* filename) const
{
    if (m_data.empty()) return false;
    return exr::save_half_rgba(filename, m_width, m_height, &m_data[0].r, 3);
}

bool Bitmap::loadImage(const char* filename)
{
    std::string e = extensionUpper(filename);
    if (e == "BMP") return loadBMP(filename);
    if (e == "EXR") return loadEXR(filename);
    return false;
}

bool Bitmap::saveImage(const char* filename) const
{
    std::string e = extensionUpper(filename);
    if (e == "BMP") return saveBMP(filename);
    if (e == "EXR") return saveEXR(filename);
    return false;
}

void Bitmap::differentiate()
{
    std::vector<Color3> out((size_t)m_width * m_height);
    for (int y = 0; y < m_height; y++)
        for (int x = 0; x < m_width; x++) {
            float me = getPixel(x, y).intensity();
            float dx = me - getPixel((x + 1) % m_width, y).intensity();
            float dy = me - getPixel(x, (y + 1) % m_height).intensity();
            out[x + (size_t)y * m_width] = Color3(dx, dy, 0);
        }
    m_data.swap(out);
}

void Bitmap::decompressGamma(float gamma)
{
    const bool srgb = fabsf(gamma - 2.2f) < 1e-6f;
    auto fix = [&](float& c) {
        if (c > 0) c = srgb ? decompress_sRGB(c) : powf(c, gamma);
    };
    for (auto& p : m_data) { fix(p.r); fix(p.g); fix(p.b); }
}

}  // namespace host
}  // namespace hxr
