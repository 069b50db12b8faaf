// common.cuh -- host-side plumbing shared by the C-ABI translation units: error reporting,
// host/device pointer staging, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

#include "mcaz.h"

namespace mcaz {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<uint64_t> g_launches;  // kernels launched by this library (all engines)

#define MCAZ_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::mcaz::fail(MCAZ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

#define MCAZ_CHECK_LAUNCH()                                                                          \
    do {                                                                                             \
        ::mcaz::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess)                                                                       \
            return ::mcaz::fail(MCAZ_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

int require_device();  // MCAZ_OK or MCAZ_ENODEV (no CPU fallback)

inline bool is_device_pointer(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// An input that kernels can read: the caller's device pointer, or a staged copy of host data.
template <typename T>
struct In {
    const T* ptr = nullptr;
    T* owned = nullptr;
    int init(const T* src, size_t n, cudaStream_t st) {
        if (n == 0 || src == nullptr) { ptr = src; return MCAZ_OK; }
        if (is_device_pointer(src)) { ptr = src; return MCAZ_OK; }
        MCAZ_CUDA(cudaMalloc(&owned, n * sizeof(T)));
        MCAZ_CUDA(cudaMemcpyAsync(owned, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
        ptr = owned;
        return MCAZ_OK;
    }
    ~In() { if (owned) cudaFree(owned); }
};

// An output kernels can write: the caller's device pointer, or a device buffer copied back.
template <typename T>
struct Out {
    T* ptr = nullptr;
    T* owned = nullptr;
    T* host = nullptr;
    size_t count = 0;
    int init(T* dst, size_t n, cudaStream_t st, bool zero = false) {
        count = n;
        if (n == 0 || dst == nullptr) { ptr = dst; return MCAZ_OK; }
        if (is_device_pointer(dst)) { ptr = dst; }
        else {
            MCAZ_CUDA(cudaMalloc(&owned, n * sizeof(T)));
            ptr = owned;
            host = dst;
        }
        if (zero) MCAZ_CUDA(cudaMemsetAsync(ptr, 0, n * sizeof(T), st));
        return MCAZ_OK;
    }
    int finish(cudaStream_t st) {
        if (owned && host) MCAZ_CUDA(cudaMemcpyAsync(host, owned, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        return MCAZ_OK;
    }
    ~Out() { if (owned) cudaFree(owned); }
};

inline mc_rules rules_or_default(const mc_rules* r) {
    mc_rules d;
    mc_default_rules(&d);
    return r ? *r : d;
}

}  // namespace mcaz
