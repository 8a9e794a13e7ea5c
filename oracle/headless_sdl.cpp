// oracle/headless_sdl.cpp — TEST INFRASTRUCTURE ONLY (not product code).
// Headless replacement for the reference's src/sdl.cpp: implements the surface of
// src/sdl.h that the render path touches, with the display calls as no-ops.
//   frameWidth/frameHeight  <- sdl.cpp:259-270 (here: the size given to initGraphics)
//   getBucketsList, Rect::clip <- sdl.cpp:272-294 (same zig-zag bucket order)
//   Color::init_sRGB_cache / convertTo8bit_sRGB_cached <- sdl.cpp:402-419
// The algorithms are restated from those descriptions, not copied.
#include <SDL.h>
#include <algorithm>
#include <chrono>
#include <vector>
#include "sdl.h"

bool isInteractive = false;
static int g_w = 0, g_h = 0;
static Uint8 g_keys[SDL_NUM_SCANCODES];

Uint32 SDL_GetTicks(void)
{
    using namespace std::chrono;
    static const steady_clock::time_point t0 = steady_clock::now();
    return (Uint32)duration_cast<milliseconds>(steady_clock::now() - t0).count();
}

bool initGraphics(int frameWidth, int frameHeight) { g_w = frameWidth; g_h = frameHeight; return true; }
void closeGraphics(void) {}
void displayVFB(Color[VFB_MAX_SIZE][VFB_MAX_SIZE]) {}
bool checkForUserExit(void) { return false; }
void getSDLInputs(const Uint8*& keystate, int& dx, int& dy, std::vector<SDL_Event>& events)
{
    keystate = g_keys; dx = dy = 0; events.clear();
}
int frameWidth(void) { return g_w; }
int frameHeight(void) { return g_h; }

void Rect::clip(int W, int H)
{
    if (x1 > W) x1 = W;
    if (y1 > H) y1 = H;
    w = x1 > x0 ? x1 - x0 : 0;
    h = y1 > y0 ? y1 - y0 : 0;
}

std::vector<Rect> getBucketsList(int bucketSize)
{
    // rows of buckets top to bottom; odd rows run right-to-left
    std::vector<Rect> out;
    const int nx = (g_w + bucketSize - 1) / bucketSize, ny = (g_h + bucketSize - 1) / bucketSize;
    for (int by = 0; by < ny; by++)
        for (int i = 0; i < nx; i++) {
            int bx = (by & 1) ? nx - 1 - i : i;
            Rect r(bx * bucketSize, by * bucketSize, (bx + 1) * bucketSize, (by + 1) * bucketSize);
            r.clip(g_w, g_h);
            out.push_back(r);
        }
    return out;
}

bool displayVFBRect(Rect, Color[VFB_MAX_SIZE][VFB_MAX_SIZE]) { return true; }
bool drawRect(Rect, const Color&) { return true; }
void showUpdatedFullscreen() {}
void showUpdated(Rect) {}
bool markRegion(Rect, const Color&) { return true; }
void markAApixels(bool[VFB_MAX_SIZE][VFB_MAX_SIZE]) {}
void uiMainLoop() {}

static unsigned char g_srgb_lut[4097];
void Color::init_sRGB_cache(void)
{
    for (int i = 0; i <= 4096; i++) g_srgb_lut[i] = (unsigned char)convertTo8bit_sRGB(i / 4096.0f);
}
unsigned convertTo8bit_sRGB_cached(float x)
{
    if (x <= 0) return 0;
    if (x >= 1) return 255;
    return g_srgb_lut[int(x * 4096.0f)];
}
