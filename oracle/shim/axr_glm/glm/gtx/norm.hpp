// axr mini-glm forwarding header (see ../glm.hpp)
#pragma once
#include "../glm.hpp"
