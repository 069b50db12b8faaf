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

// Grow-only device staging buffers, one per slot, reused across calls (cudaMalloc / cudaFree per
// call cost tens of milliseconds next to multi-GB tree arenas).
struct Scratch {
    static constexpr int SLOTS = 12;
    void* buf[SLOTS] = {};
    size_t cap[SLOTS] = {};
    int next = 0;
    void begin() { next = 0; }
    int take(size_t bytes, void** out) {
        if (next >= SLOTS) return fail(MCAZ_EINVAL, "staging slots exhausted");
        int s = next++;
        if (cap[s] < bytes) {
            if (buf[s]) cudaFree(buf[s]);
            buf[s] = nullptr; cap[s] = 0;
            size_t want = bytes + bytes / 4 + 256;
            MCAZ_CUDA(cudaMalloc(&buf[s], want));
            cap[s] = want;
        }
        *out = buf[s];
        return MCAZ_OK;
    }
    void release() {
        for (int s = 0; s < SLOTS; ++s) { if (buf[s]) cudaFree(buf[s]); buf[s] = nullptr; cap[s] = 0; }
    }
};
Scratch& thread_scratch();   // for the stateless mc_* entry points

// An input that kernels can read: the caller's device pointer, or a staged copy of host data.
template <typename T>
struct In {
    const T* ptr = nullptr;
    int init(const T* src, size_t n, cudaStream_t st, Scratch& sc) {
        if (n == 0 || src == nullptr) { ptr = src; return MCAZ_OK; }
        if (is_device_pointer(src)) { ptr = src; return MCAZ_OK; }
        void* d = nullptr;
        if (int rc = sc.take(n * sizeof(T), &d)) return rc;
        MCAZ_CUDA(cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
        ptr = static_cast<const T*>(d);
        return MCAZ_OK;
    }
};

// An output kernels can write: the caller's device pointer, or a device buffer copied back.
template <typename T>
struct Out {
    T* ptr = nullptr;
    T* host = nullptr;
    size_t count = 0;
    int init(T* dst, size_t n, cudaStream_t st, Scratch& sc, bool zero = false) {
        count = n;
        if (n == 0 || dst == nullptr) { ptr = dst; return MCAZ_OK; }
        if (is_device_pointer(dst)) { ptr = dst; }
        else {
            void* d = nullptr;
            if (int rc = sc.take(n * sizeof(T), &d)) return rc;
            ptr = static_cast<T*>(d);
            host = dst;
        }
        if (zero) MCAZ_CUDA(cudaMemsetAsync(ptr, 0, n * sizeof(T), st));
        return MCAZ_OK;
    }
    int finish(cudaStream_t st) {
        if (host) MCAZ_CUDA(cudaMemcpyAsync(host, ptr, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        return MCAZ_OK;
    }
};

inline mc_rules rules_or_default(const mc_rules* r) {
    mc_rules d;
    mc_default_rules(&d);
    return r ? *r : d;
}

}  // namespace mcaz
