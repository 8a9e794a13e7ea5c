// oracle/shim/SDL_video.h — TEST INFRASTRUCTURE ONLY (intentionally empty).
#pragma once
