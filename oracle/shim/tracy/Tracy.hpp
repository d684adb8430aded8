// Tracy profiler macros stubbed to no-ops (the reference fetches Tracy over the network,
// reference external/CMakeLists.txt:3-10; DISABLE_PROFILING=ON strips it anyway).
#pragma once
#ifndef ZoneScoped
#define ZoneScoped
#define ZoneScopedN(x)
#define FrameMark
#define TracyAlloc(p, s)
#define TracyFree(p)
#endif
