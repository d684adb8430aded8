// back-slash include path used by reference include/texture.hpp:4, src/framebuffer.cpp:2, src/camera.cpp:6
#pragma once
#include "tracy/Tracy.hpp"
