// back-slash include path used by reference include/math.hpp:13
#pragma once
#include <glm/glm.hpp>
