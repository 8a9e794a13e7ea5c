// oracle/shim/SDL.h — TEST INFRASTRUCTURE ONLY (not product code).
// A headless stand-in for <SDL.h>: declares only the handful of SDL names the
// reference's non-display translation units mention (main.cpp:391-524 gameloop /
// coarseRender, mesh.cpp:63-68, heightfield.cpp:256-258, sdl.h:59), so that the
// UNMODIFIED reference sources under /root/reference/src compile without SDL2.
#pragma once
#include <cstdint>
typedef uint32_t Uint32;
typedef uint8_t Uint8;
typedef int32_t Sint32;

Uint32 SDL_GetTicks(void);  // implemented in oracle/headless_sdl.cpp (steady_clock ms)

enum { SDL_KEYDOWN = 0x300, SDL_MOUSEMOTION = 0x400 };
enum { SDLK_r = 'r', SDLK_F5 = 0x4000003e };
enum {
    SDL_SCANCODE_RIGHT = 79, SDL_SCANCODE_LEFT = 80, SDL_SCANCODE_DOWN = 81, SDL_SCANCODE_UP = 82,
    SDL_SCANCODE_KP_2 = 90, SDL_SCANCODE_KP_4 = 92, SDL_SCANCODE_KP_6 = 94, SDL_SCANCODE_KP_8 = 96,
    SDL_NUM_SCANCODES = 512
};
struct SDL_Keysym { int scancode; int sym; unsigned short mod; };
struct SDL_KeyboardEvent { Uint32 type; SDL_Keysym keysym; };
struct SDL_MouseMotionEvent { Uint32 type; Sint32 x, y, xrel, yrel; };
union SDL_Event {
    Uint32 type;
    SDL_KeyboardEvent key;
    SDL_MouseMotionEvent motion;
};
