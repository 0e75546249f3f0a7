// NVTX ranges around the phases of a call (the reference logs phase times with myLog / clock(),
// src/util.cpp:149-155).  Header-only NVTX 3: without a profiler attached a range is a null-pointer check.
#ifndef STCSP_HOST_TRACE_RANGES_H
#define STCSP_HOST_TRACE_RANGES_H

#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define STCSP_HAVE_NVTX 1
#endif
#endif

namespace stcsp {

struct TraceRange {         // one nested range, closed when it goes out of scope
#ifdef STCSP_HAVE_NVTX
    explicit TraceRange(const char *name) { nvtxRangePushA(name); }
    ~TraceRange() { nvtxRangePop(); }
#else
    explicit TraceRange(const char *) {}
#endif
    TraceRange(const TraceRange &) = delete;
    TraceRange &operator=(const TraceRange &) = delete;
};

}  // namespace stcsp

#endif
