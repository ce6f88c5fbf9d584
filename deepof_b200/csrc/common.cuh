// Shared helpers for the deepof_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <atomic>

#include "../../include/deepof_b200.h"

namespace dofb {

// ---- error plumbing (C ABI returns int, message kept thread-local) ---------
char *err_buf();
int set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define DOFB_CHECK_ARG(cond, ...)                      \
    do {                                               \
        if (!(cond)) return ::dofb::set_error(__VA_ARGS__); \
    } while (0)

#define DOFB_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return ::dofb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                     __FILE__, __LINE__);                                   \
    } while (0)

#define DOFB_LAUNCH_OK()                                                                   \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess)                                                            \
            return ::dofb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                                     __FILE__, __LINE__);                                  \
        ::dofb::count_launch();                                                            \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int num_sms();

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
// derivative of ELU expressed through its OUTPUT y: y>0 -> 1 else exp(x) = y+1
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }

}  // namespace dofb
