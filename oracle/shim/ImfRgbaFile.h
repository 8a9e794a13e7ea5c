// oracle/shim/ImfRgbaFile.h — TEST INFRASTRUCTURE ONLY (not product code).
// Stand-in for the OpenEXR RGBA interface used by reference src/bitmap.cpp:242-288
// (Imf::RgbaInputFile / RgbaOutputFile / Rgba, Imath::Box2i). Decoding is done by
// hexray_b200/csrc/host/exr_codec.h, which tests/test_exr_codec.py checks against
// OpenCV's own OpenEXR build, bit for bit, on every bundled .exr.
#pragma once
#include <string>
#include <vector>
#include <stdexcept>
#include "exr_codec.h"

namespace Iex { struct BaseExc : public std::runtime_error { using std::runtime_error::runtime_error; }; }
namespace Imath {
struct V2i { int x, y; };
struct Box2i { V2i min, max; };
}
namespace Imf {
// the reference only ever assigns floats to / reads floats from these fields;
// values coming out of a HALF file are exactly representable, and values going in
// are rounded to half by the writer, as OpenEXR's `half` type would.
struct Rgba { float r, g, b, a; };
enum RgbaChannels { WRITE_RGBA = 0xf };

class RgbaInputFile {
    hxr::exr::Image m_img;
    Rgba* m_base = nullptr;
    size_t m_xs = 1, m_ys = 0;
public:
    explicit RgbaInputFile(const char* fn)
    {
        std::string err;
        if (!hxr::exr::load(fn, m_img, &err)) throw Iex::BaseExc(err);
    }
    Imath::Box2i dataWindow() const { return Imath::Box2i{{0, 0}, {m_img.width - 1, m_img.height - 1}}; }
    void setFrameBuffer(Rgba* base, size_t xStride, size_t yStride) { m_base = base; m_xs = xStride; m_ys = yStride; }
    void readPixels(int y0, int y1)
    {
        for (int y = y0; y <= y1; y++)
            for (int x = 0; x < m_img.width; x++) {
                const float* p = &m_img.rgba[((size_t)y * m_img.width + x) * 4];
                m_base[x * m_xs + y * m_ys] = Rgba{p[0], p[1], p[2], p[3]};
            }
    }
};

class RgbaOutputFile {
    std::string m_fn;
    int m_w, m_h;
    const Rgba* m_base = nullptr;
    size_t m_xs = 1, m_ys = 0;
public:
    RgbaOutputFile(const char* fn, int w, int h, RgbaChannels) : m_fn(fn), m_w(w), m_h(h) {}
    void setFrameBuffer(const Rgba* base, size_t xStride, size_t yStride) { m_base = base; m_xs = xStride; m_ys = yStride; }
    void writePixels(int n)
    {
        std::vector<float> rgb((size_t)m_w * n * 3);
        for (int y = 0; y < n; y++)
            for (int x = 0; x < m_w; x++) {
                const Rgba& p = m_base[x * m_xs + y * m_ys];
                float* o = &rgb[((size_t)y * m_w + x) * 3];
                o[0] = p.r; o[1] = p.g; o[2] = p.b;
            }
        if (!hxr::exr::save_half_rgba(m_fn.c_str(), m_w, n, rgb.data(), 3)) throw Iex::BaseExc("cannot write " + m_fn);
    }
};
}  // namespace Imf
