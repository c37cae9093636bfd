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

// Teardown at process exit.  The CUDA runtime registers its own exit handler at the first API call; a host program that
// keeps its vex::Context in a static (the reference's test fixture does) destroys queues, buffers and matrices AFTER that
// handler has run, and CUDA calls on a runtime that is being unloaded crash now and then.  note_cuda_started() registers
// an exit handler right after the first CUDA call -- it therefore runs BEFORE the runtime's -- and from then on every
// release entry point returns at once (the process is going away; the driver reclaims the memory).
void note_cuda_started();
bool process_exiting();

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

#define VEXB_RELEASE_GUARD() do { if (::vexb::process_exiting()) return VEXB_OK; } while (0)

#define VEXB_TRY(expr) do { int s_ = (expr); if (s_ != VEXB_OK) return s_; } while (0)

#define VEXB_LAUNCHED() do { ::vexb::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    ::vexb::set_error(__FILE__, __LINE__, "kernel launch failed: %s", cudaGetErrorString(e_)); \
    return VEXB_ERR_CUDA; } } while (0)
