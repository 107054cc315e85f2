// Shared host/device helpers for libs2i: error reporting, launch accounting, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/s2i.h"

namespace s2i {

// error codes: S2I_OK / S2I_ERR_* macros from include/s2i.h

// Thread-local last-error message (s2i_last_error()).
int set_error(int code, const char* fmt, ...);
const char* last_error();

extern long g_launches;  // kernels launched by this library since load
// Bumped whenever a scratch buffer that kernels address directly (activation arena, LGP workspace, split-K / GroupNorm
// scratch) is (re)allocated: a captured CUDA graph is only valid for the generation it was captured in.
extern long g_alloc_gen;
inline void count_launch(int n = 1) { g_launches += n; }

// Opt-in per-launch device timing (bench.py's roofline figures): while a profile is open every launch site drops a
// CUDA event on the profiled stream; the gap to the previous event is attributed to the launch's tag.
extern bool g_prof_on;
void prof_mark(const char* tag, double flops, double bytes);
int prof_begin(cudaStream_t st);
void prof_set_peaks(double tflops, double gbs);
// Text report, one line per tag: "tag launches ms flops bytes roof_ms" (roof_ms = sum over the tag's launches of
// max(flops / peak_flops, bytes / peak_bw), see prof_set_peaks).  Returns the number of bytes written (or < 0).
int prof_end(char* buf, int cap);

#define S2I_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return ::s2i::set_error(S2I_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                                    cudaGetErrorString(e__));                                                \
    } while (0)

#define S2I_LAUNCH_CHECK()                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                                \
        if (e__ != cudaSuccess)                                                                              \
            return ::s2i::set_error(S2I_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,           \
                                    cudaGetErrorString(e__));                                                \
        ::s2i::count_launch();                                                                               \
        if (::s2i::g_prof_on) ::s2i::prof_mark(__func__, 0.0, 0.0);                                          \
    } while (0)

#define S2I_LAUNCH_CHECK_TAG(tag, flops, bytes)                                                              \
    do {                                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                                \
        if (e__ != cudaSuccess)                                                                              \
            return ::s2i::set_error(S2I_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__,           \
                                    cudaGetErrorString(e__));                                                \
        ::s2i::count_launch();                                                                               \
        if (::s2i::g_prof_on) ::s2i::prof_mark((tag), (flops), (bytes));                                     \
    } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of the sampling path is launched with programmatic stream serialization: its CTAs may become resident
// (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) while the previous kernel drains.
// Each kernel therefore executes pdl_wait() before it first touches memory written by earlier kernels, and
// pdl_launch() once it holds its own scarce resources (TMEM), so dependents never starve a still-starting primary.
// S2I_NO_PDL=1 in the environment turns the launch attribute off (the device-side instructions become no-ops).
extern bool g_pdl;
// True while the last operation enqueued by the library was one of its kernels: only then is the next kernel launched
// as a programmatic dependent (a copy / memset / graph launch in between gets an ordinary full dependency).
extern bool g_prev_kernel;
#define S2I_MEMOP(call)                \
    do {                               \
        ::s2i::g_prev_kernel = false;  \
        S2I_CUDA(call);                \
    } while (0)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && g_prev_kernel) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    g_prev_kernel = true;
}
// Same, as thread-block clusters of cluster_x CTAs along x (grid.x must be a multiple of cluster_x).
template <typename... KArgs, typename... Args>
inline void launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster_x, cudaStream_t st,
                                  Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && g_prev_kernel) ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    g_prev_kernel = true;
}
#define S2I_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::s2i::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#endif

#define S2I_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != 0) return rc__;    \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }
inline long round_up_l(long a, long b) { return ceil_div_l(a, b) * b; }

constexpr int kNumSMs = 148;

}  // namespace s2i
