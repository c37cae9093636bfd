// Shared internals of libvexb200.so: error reporting, tunables, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <atomic>
#include "../../include/vexb200.h"

namespace vexb {

// Thread-local "file:line: message" of the last failure, in the spirit of
// vexcl/backend/cuda/error.hpp:119-145 (which throws; the C ABI returns codes).
void set_error(const char *file, int line, const char *fmt, ...);
long param(const char *name, long dflt);
extern std::atomic<uint64_t> g_launches;

struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int sm_count(int dev);

inline size_t dtype_size(int dt) {
    switch (dt) {
        case VEXB_F64: case VEXB_I64: case VEXB_U64: return 8;
        case VEXB_F32: case VEXB_I32: case VEXB_U32: return 4;
    }
    return 0;
}

} // namespace vexb

#define VEXB_FAIL(code, ...) do { ::vexb::set_error(__FILE__, __LINE__, __VA_ARGS__); return (code); } while (0)

#define VEXB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    ::vexb::set_error(__FILE__, __LINE__, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    return VEXB_ERR_CUDA; } } while (0)

#define VEXB_CHECK(cond, ...) do { if (!(cond)) VEXB_FAIL(VEXB_ERR_INVALID, __VA_ARGS__); } while (0)

#define VEXB_TRY(expr) do { int s_ = (expr); if (s_ != VEXB_OK) return s_; } while (0)

#define VEXB_LAUNCHED() do { ::vexb::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    ::vexb::set_error(__FILE__, __LINE__, "kernel launch failed: %s", cudaGetErrorString(e_)); \
    return VEXB_ERR_CUDA; } } while (0)
