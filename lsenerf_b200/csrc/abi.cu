// abi.cu -- error reporting, launch accounting and version of the C ABI.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char *what)
{
    count_launch();
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        return -5;
    }
    return 0;
}

}  // namespace lse

extern "C" int lse_abi_version(void) { return LSE_ABI_VERSION; }
extern "C" const char *lse_last_error(void) { return lse::g_err; }
extern "C" int64_t lse_launch_count(void) { return lse::g_launches.load(); }
