// A/B and trace switches (SPK_NO_PDL, SPK_NO_ZIGZAG, SPK_STEM, SPK_*_TRACE, ... -- tools/README.md) are read from the
// environment ONLY in a debug build (`python -m sykepic_b200._build --debug` adds -DSPK_DEBUG_SWITCHES).  In the product
// build every switch is compiled out: one code path, no environment-dependent behaviour.
#pragma once

#include <cstdlib>

namespace spk {
#ifdef SPK_DEBUG_SWITCHES
inline const char* debug_env(const char* name) { return getenv(name); }
#else
inline const char* debug_env(const char*) { return nullptr; }
#endif
}  // namespace spk
