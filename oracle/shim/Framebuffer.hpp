// case fix for reference src/framebuffer.cpp:1 (Windows file systems are case-insensitive)
#pragma once
#include "framebuffer.hpp"
