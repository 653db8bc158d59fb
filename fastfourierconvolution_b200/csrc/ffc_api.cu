// Library-level entry points: version and error text (include/ffc_b200.h).
#include "ffc_common.cuh"
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_ffc_error[512] = "";

void ffc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ffc_error, sizeof(g_ffc_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* ffc_last_error(void) { return g_ffc_error; }

extern "C" int ffc_version(void) { return 200; }   // 0.2.0 (ABI version: _C.ABI_VERSION must match)

// 1 when this shared object is the host emulation build used by tests/ (never shipped as product)
extern "C" int ffc_is_emulation(void) {
#ifdef FFC_EMU
    return 1;
#else
    return 0;
#endif
}

extern "C" size_t ffc_workspace_bytes(int batch, int max_channels) {
    if (batch < 0) batch = 0;
    if (max_channels < 1) max_channels = 1;
    // max over: 2*C doubles (BN, bias grad), B*C doubles (SE fwd), (3*B*C + B*hid) floats (SE bwd)
    return ((size_t)batch + 2) * (size_t)max_channels * 16 + 1024;
}

// Number of kernel launches issued by this library since load (monotonic; bench.py reports the
// difference over its timed region as "gpu_launches").
#ifdef FFC_EMU
static unsigned long long g_ffc_launches = 0;
void ffc_count_launch() { ++g_ffc_launches; }
extern "C" unsigned long long ffc_launch_count(void) { return g_ffc_launches; }
#else
#include <atomic>
static std::atomic<unsigned long long> g_ffc_launches{0};
void ffc_count_launch() { g_ffc_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long ffc_launch_count(void) { return g_ffc_launches.load(std::memory_order_relaxed); }
#endif
