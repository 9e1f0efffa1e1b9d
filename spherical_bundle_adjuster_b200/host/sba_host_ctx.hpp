// sba_host_ctx.hpp -- process-wide libsba_b200 context used by the drop-in facade classes.
#pragma once
#include <stdexcept>
#include <string>

#include "sba_b200.h"

namespace sba_host {

// One context on CUDA device $SBA_DEVICE (default 0), created on first use.  The reference's classes
// carry no device state; the facade keeps that shape by sharing this context.  Not thread safe, like
// the reference (feature_matcher.hpp:44-48 keeps per-call state in members).
inline sba_ctx* ctx()
{
    static sba_ctx* c = nullptr;
    if (!c) {
        int dev = 0;
        if (const char* e = std::getenv("SBA_DEVICE")) dev = std::atoi(e);
        if (sba_ctx_create(dev, nullptr, &c) != SBA_OK) throw std::runtime_error(std::string("libsba_b200: ") + sba_last_error());
    }
    return c;
}

inline void check(int status)
{
    if (status != SBA_OK) throw std::runtime_error(std::string("libsba_b200: ") + sba_last_error());
}

}  // namespace sba_host
