// abi.cu -- error reporting, launch accounting and version of the C ABI.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <vector>
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

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
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

static std::vector<det_setter_t> &det_setters()
{
    static std::vector<det_setter_t> v;
    return v;
}
void register_det_setter(det_setter_t fn) { det_setters().push_back(fn); }

__global__ void det_flush_kernel(float *__restrict__ grad, long long *__restrict__ shadow, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long q = shadow[i];
        if (q != 0) {
            grad[i] += (float)((double)q * (1.0 / 1125899906842624.0));
            shadow[i] = 0;
        }
    }
}

}  // namespace lse

extern "C" int lse_set_deterministic(int n_ranges, const void *const *bases, const int64_t *nbytes, void *const *shadows)
{
    LSE_CHECK_ARG(n_ranges >= 0 && n_ranges <= 4, "at most 4 gradient arenas");
    lse::DetCfg c{};
    c.n = n_ranges;
    for (int k = 0; k < n_ranges; ++k) {
        LSE_CHECK_ARG(bases && nbytes && shadows && bases[k] && shadows[k] && nbytes[k] > 0, "bad range");
        c.base[k] = (unsigned long long)bases[k];
        c.end[k] = c.base[k] + (unsigned long long)nbytes[k];
        c.shadow[k] = (unsigned long long)shadows[k];
    }
    cudaDeviceSynchronize();
    for (auto fn : lse::det_setters()) fn(c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { lse::set_error("lse_set_deterministic: %s", cudaGetErrorString(e)); return -5; }
    return 0;
}

extern "C" int lse_det_flush(float *grad, void *shadow, int64_t n, lse_stream_t stream)
{
    LSE_CHECK_ARG(grad && shadow, "missing pointer");
    if (n <= 0) return 0;
    lse::det_flush_kernel<<<lse::stride_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(grad, reinterpret_cast<long long *>(shadow), n);
    return lse::check_launch("lse_det_flush");
}

extern "C" int lse_abi_version(void) { return LSE_ABI_VERSION; }
extern "C" const char *lse_last_error(void) { return lse::g_err; }
extern "C" int64_t lse_launch_count(void) { return lse::g_launches.load(); }
