// oracle/shim/Iex.h — TEST INFRASTRUCTURE ONLY. Iex::BaseExc lives in ImfRgbaFile.h.
#pragma once
#include "ImfRgbaFile.h"
