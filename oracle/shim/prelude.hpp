// Force-included (-include) ahead of every unmodified reference translation unit.
// The reference relies on MSVC's transitive includes (and on glm pulling in <immintrin.h>);
// this restores them for g++ without touching the reference sources. See SURVEY.md §8(c).
#pragma once
#include <functional>
#include <condition_variable>
#include <thread>
#include <vector>
#include <atomic>
#include <stdexcept>
#include <cstring>
#include <cstdint>
#include <memory>
#include <memory_resource>
#include <array>
#include <limits>
#include <iostream>
#include <future>
#include <string>
#include <sstream>
#include <algorithm>
#include <cmath>
#include <immintrin.h>
