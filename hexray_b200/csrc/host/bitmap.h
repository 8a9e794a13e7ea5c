// bitmap.h — host-side image container + BMP/EXR I/O.
// Same interface and behaviour as the reference's Bitmap (src/bitmap.h, src/bitmap.cpp):
//   loadBMP   8-bit palettised and 24/32-bit BGR(A), bottom-up rows     bitmap.cpp:127-200
//   saveBMP   24-bit through the 4097-entry sRGB LUT (12.02 quirk kept) bitmap.cpp:202-240,
//             color.h:36-47, sdl.cpp:404-419
//   loadEXR / saveEXR  via exr_codec.h instead of the OpenEXR library   bitmap.cpp:242-288
//   differentiate, decompressGamma                                      bitmap.cpp:304-338
#pragma once
#include <string>
#include <vector>
#include "math_types.h"

namespace hxr {
namespace host {

unsigned convertTo8bit_sRGB(float x);         // color.h:36-47
// header + an already converted pixel array (bottom-up BGR rows of rowsz bytes): what saveBMP writes (bitmap.cpp:202-240)
bool writeBmpFile(const char* filename, int width, int height, int rowsz, const unsigned char* rows);
unsigned convertTo8bit_sRGB_cached(float x);  // sdl.cpp:414-419
float decompress_sRGB(float x);               // color.h:49-57
std::string extensionUpper(const char* fileName);

class Bitmap {
    int m_width = -1, m_height = -1;
    std::vector<Color3> m_data;
public:
    void freeMem();
    int getWidth() const { return m_width; }
    int getHeight() const { return m_height; }
    bool isOK() const { return !m_data.empty(); }
    void generateEmptyImage(int width, int height);
    Color3 getPixel(int x, int y) const;
    void setPixel(int x, int y, const Color3& c);
    const std::vector<Color3>& data() const { return m_data; }

    bool loadBMP(const char* filename);
    bool saveBMP(const char* filename) const;
    bool loadEXR(const char* filename);
    bool saveEXR(const char* filename) const;
    bool loadImage(const char* filename);
    bool saveImage(const char* filename) const;

    void differentiate();
    void decompressGamma(float gamma);
};

}  // namespace host
}  // namespace hxr
