// Lifecycle, device properties, streams/events, memory: the thin part of the
// C ABI that replaces vex::backend::{device,command_queue,device_vector,event}
// (vexcl/backend/cuda/context.hpp, device_vector.hpp, event.hpp).
#include "common.cuh"
#include <cstdarg>
#include <map>
#include <mutex>
#include <vector>

namespace vexb {

static thread_local char tl_error[1024] = "";

void set_error(const char *file, int line, const char *fmt, ...) {
    const char *base = strrchr(file, '/');
    int k = snprintf(tl_error, sizeof(tl_error), "%s:%d: ", base ? base + 1 : file, line);
    va_list ap; va_start(ap, fmt);
    vsnprintf(tl_error + k, sizeof(tl_error) - k, fmt, ap);
    va_end(ap);
}

std::atomic<uint64_t> g_launches{0};

static std::mutex g_mx;
static std::map<std::string, long> g_params;
static std::vector<int> g_sm_count;

long param(const char *name, long dflt) {
    std::lock_guard<std::mutex> lock(g_mx);
    auto it = g_params.find(name);
    return it == g_params.end() ? dflt : it->second;
}

int sm_count(int dev) {
    std::lock_guard<std::mutex> lock(g_mx);
    if (dev >= 0 && dev < (int)g_sm_count.size() && g_sm_count[dev] > 0) return g_sm_count[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0) { if ((int)g_sm_count.size() <= dev) g_sm_count.resize(dev + 1, 0); g_sm_count[dev] = n; }
    return n;
}

static std::atomic<bool> g_exiting{false};
static void mark_exiting() { g_exiting.store(true); }
bool process_exiting() { return g_exiting.load(std::memory_order_relaxed); }
void note_cuda_started() {
    static std::once_flag once;
    std::call_once(once, [] { atexit(mark_exiting); });
}

} // namespace vexb

using namespace vexb;

extern "C" {

int vexb_abi_version(void) { return VEXB_ABI_VERSION; }
const char *vexb_last_error(void) { return tl_error; }

int vexb_init(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) VEXB_FAIL(VEXB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    if (n <= 0) VEXB_FAIL(VEXB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    note_cuda_started();
    return VEXB_OK;
}

int vexb_shutdown(void) { return VEXB_OK; }

int vexb_device_count(int *n) {
    VEXB_CHECK(n, "n is NULL");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; VEXB_FAIL(VEXB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e)); }
    note_cuda_started();
    return VEXB_OK;
}

int vexb_device_props(int dev, vexb_devprops *p) {
    VEXB_CHECK(p, "p is NULL");
    cudaDeviceProp cp;
    VEXB_CUDA(cudaGetDeviceProperties(&cp, dev));
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "%s", cp.name);
    p->cc_major = cp.major; p->cc_minor = cp.minor;
    p->sm_count = cp.multiProcessorCount;
    p->max_threads_per_block = cp.maxThreadsPerBlock;
    p->warp_size = cp.warpSize;
    p->smem_per_block_optin = cp.sharedMemPerBlockOptin;
    p->total_mem = cp.totalGlobalMem;
    p->l2_bytes = (size_t)cp.l2CacheSize;
    return VEXB_OK;
}

int vexb_set_param(const char *name, long value) {
    VEXB_CHECK(name, "name is NULL");
    std::lock_guard<std::mutex> lock(g_mx);
    g_params[name] = value;
    return VEXB_OK;
}

int vexb_get_param(const char *name, long *value) {
    VEXB_CHECK(name && value, "NULL argument");
    std::lock_guard<std::mutex> lock(g_mx);
    auto it = g_params.find(name);
    if (it == g_params.end()) VEXB_FAIL(VEXB_ERR_INVALID, "unknown parameter '%s'", name);
    *value = it->second;
    return VEXB_OK;
}

int vexb_launch_count(uint64_t *n) { VEXB_CHECK(n, "n is NULL"); *n = g_launches.load(); return VEXB_OK; }

// ---- streams / events -------------------------------------------------------
int vexb_stream_create(int dev, void **stream) {
    VEXB_CHECK(stream, "stream is NULL");
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    cudaStream_t s; VEXB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    note_cuda_started();
    *stream = (void *)s; return VEXB_OK;
}
int vexb_stream_destroy(int dev, void *stream) {
    VEXB_RELEASE_GUARD();
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaStreamDestroy((cudaStream_t)stream)); return VEXB_OK;
}
int vexb_stream_sync(int dev, void *stream) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return VEXB_OK;
}
int vexb_device_sync(int dev) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaDeviceSynchronize()); return VEXB_OK;
}
int vexb_event_create(int dev, void **event) {
    VEXB_CHECK(event, "event is NULL");
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    cudaEvent_t e; VEXB_CUDA(cudaEventCreate(&e));
    *event = (void *)e; return VEXB_OK;
}
int vexb_event_destroy(int dev, void *event) {
    VEXB_RELEASE_GUARD();
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaEventDestroy((cudaEvent_t)event)); return VEXB_OK;
}
int vexb_event_record(int dev, void *event, void *stream) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream)); return VEXB_OK;
}
int vexb_event_sync(int dev, void *event) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaEventSynchronize((cudaEvent_t)event)); return VEXB_OK;
}
int vexb_stream_wait_event(int dev, void *stream, void *event) {
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0)); return VEXB_OK;
}
int vexb_event_elapsed_ms(void *start, void *stop, float *ms) {
    VEXB_CHECK(ms, "ms is NULL");
    VEXB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop)); return VEXB_OK;
}

// ---- memory -----------------------------------------------------------------
int vexb_malloc(int dev, size_t bytes, void **p) {
    VEXB_CHECK(p, "p is NULL");
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    *p = nullptr;
    if (bytes == 0) return VEXB_OK;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); VEXB_FAIL(VEXB_ERR_NOMEM, "cudaMalloc(%zu) out of memory", bytes); }
    VEXB_CUDA(e);
    note_cuda_started();
    return VEXB_OK;
}
int vexb_free(int dev, void *p) {
    if (!p) return VEXB_OK;
    VEXB_RELEASE_GUARD();
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaFree(p)); return VEXB_OK;
}
int vexb_host_alloc(size_t bytes, void **p) {
    VEXB_CHECK(p, "p is NULL");
    VEXB_CUDA(cudaMallocHost(p, bytes ? bytes : 1)); return VEXB_OK;
}
int vexb_host_free(void *p) { VEXB_RELEASE_GUARD(); if (p) VEXB_CUDA(cudaFreeHost(p)); return VEXB_OK; }

int vexb_h2d(int dev, void *dst, const void *src, size_t bytes, void *stream, int blocking) {
    if (!bytes) return VEXB_OK;
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (blocking) VEXB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return VEXB_OK;
}
int vexb_d2h(int dev, void *dst, const void *src, size_t bytes, void *stream, int blocking) {
    if (!bytes) return VEXB_OK;
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (blocking) VEXB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return VEXB_OK;
}
int vexb_d2d(int dev, void *dst, const void *src, size_t bytes, void *stream) {
    if (!bytes) return VEXB_OK;
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return VEXB_OK;
}
int vexb_memset(int dev, void *dst, int byte, size_t bytes, void *stream) {
    if (!bytes) return VEXB_OK;
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    VEXB_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t)stream));
    return VEXB_OK;
}

} // extern "C"
