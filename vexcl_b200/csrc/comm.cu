// NCCL plumbing over NVLink: replaces the two places where the reference moves
// data between devices through host memory --
//   * the halo round trip   device -> host rx[] -> device   (vexcl/spmat.hpp:149-176)
//   * the fold of reduction partials on the host            (vexcl/reductor.hpp:412-436)
// NCCL is dlopen()ed on first use so that the library (and its host-only entry
// points) load on machines without it.
#include "comm.hpp"
#include <dlfcn.h>
#include <mutex>
#include <vector>

namespace vexb {

NcclApi g_nccl;
static std::mutex g_nccl_mx;

int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mx);
    if (g_nccl.handle) return VEXB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) VEXB_FAIL(VEXB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define L(sym) do { *(void **)(&g_nccl.sym) = dlsym(h, #sym); if (!g_nccl.sym) { dlclose(h); \
        VEXB_FAIL(VEXB_ERR_NCCL, "symbol %s missing from NCCL", #sym); } } while (0)
    L(ncclGetUniqueId); L(ncclCommInitRank); L(ncclCommInitAll); L(ncclCommDestroy); L(ncclAllReduce);
    L(ncclSend); L(ncclRecv); L(ncclGroupStart); L(ncclGroupEnd); L(ncclGetErrorString);
#undef L
    g_nccl.handle = h;
    return VEXB_OK;
}

ncclDataType_t nccl_dtype(int dt) {
    switch (dt) {
        case VEXB_F64: return ncclDouble; case VEXB_F32: return ncclFloat;
        case VEXB_I32: return ncclInt32;  case VEXB_U32: return ncclUint32;
        case VEXB_I64: return ncclInt64;  default: return ncclUint64;
    }
}

} // namespace vexb

using namespace vexb;

#define VEXB_NCCL(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) { \
    ::vexb::set_error(__FILE__, __LINE__, "%s failed: %s", #expr, g_nccl.ncclGetErrorString(r_)); \
    return VEXB_ERR_NCCL; } } while (0)

extern "C" int vexb_comm_unique_id(void *id128) {
    VEXB_CHECK(id128, "id is NULL");
    VEXB_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == VEXB_UNIQUE_ID_BYTES, "ncclUniqueId size changed");
    ncclUniqueId id;
    VEXB_NCCL(g_nccl.ncclGetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return VEXB_OK;
}

extern "C" int vexb_comm_create_rank(int dev, int nranks, int rank, const void *id128, vexb_comm **comm) {
    VEXB_CHECK(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    VEXB_TRY(nccl_load());
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    auto *c = new vexb_comm();
    c->dev = dev; c->rank = rank; c->nranks = nranks;
    ncclResult_t r = g_nccl.ncclCommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; VEXB_FAIL(VEXB_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.ncclGetErrorString(r)); }
    *comm = c;
    return VEXB_OK;
}

extern "C" int vexb_comm_create_all(int ndev, const int *devs, vexb_comm **comms) {
    VEXB_CHECK(ndev >= 1 && devs && comms, "bad arguments");
    VEXB_TRY(nccl_load());
    std::vector<ncclComm_t> cs(ndev);
    VEXB_NCCL(g_nccl.ncclCommInitAll(cs.data(), ndev, devs));
    for (int k = 0; k < ndev; ++k) {
        auto *c = new vexb_comm();
        c->dev = devs[k]; c->rank = k; c->nranks = ndev; c->comm = cs[k];
        comms[k] = c;
    }
    return VEXB_OK;
}

extern "C" int vexb_comm_destroy(vexb_comm *comm) {
    if (!comm) return VEXB_OK;
    VEXB_RELEASE_GUARD();
    if (comm->comm && g_nccl.handle) { DeviceGuard g(comm->dev); g_nccl.ncclCommDestroy(comm->comm); }
    delete comm;
    return VEXB_OK;
}

extern "C" int vexb_comm_rank(const vexb_comm *comm, int *rank, int *nranks, int *dev) {
    VEXB_CHECK(comm, "comm is NULL");
    if (rank) *rank = comm->rank;
    if (nranks) *nranks = comm->nranks;
    if (dev) *dev = comm->dev;
    return VEXB_OK;
}

extern "C" int vexb_comm_allreduce(int nlocal, vexb_comm *const *comms, void *const *bufs, void *const *streams,
                                   int count, int dtype, int op) {
    VEXB_CHECK(nlocal >= 1 && comms && bufs && count >= 1, "bad arguments");
    VEXB_CHECK(dtype >= VEXB_F64 && dtype <= VEXB_U64, "bad dtype %d", dtype);
    VEXB_TRY(nccl_load());
    const size_t es = dtype_size(dtype);
    VEXB_NCCL(g_nccl.ncclGroupStart());
    for (int k = 0; k < nlocal; ++k) {
        cudaStream_t st = streams ? (cudaStream_t)streams[k] : nullptr;
        char *b = (char *)bufs[k];
        ncclResult_t r = ncclSuccess;
        switch (op) {
            case VEXB_SUM: case VEXB_SUM_KAHAN: r = g_nccl.ncclAllReduce(b, b, count, nccl_dtype(dtype), ncclSum, comms[k]->comm, st); break;
            case VEXB_MAX: r = g_nccl.ncclAllReduce(b, b, count, nccl_dtype(dtype), ncclMax, comms[k]->comm, st); break;
            case VEXB_MIN: r = g_nccl.ncclAllReduce(b, b, count, nccl_dtype(dtype), ncclMin, comms[k]->comm, st); break;
            case VEXB_MINMAX:   // count pairs of (min, max)
                for (int i = 0; i < count && r == ncclSuccess; ++i) {
                    r = g_nccl.ncclAllReduce(b + 2 * i * es, b + 2 * i * es, 1, nccl_dtype(dtype), ncclMin, comms[k]->comm, st);
                    if (r == ncclSuccess) r = g_nccl.ncclAllReduce(b + (2 * i + 1) * es, b + (2 * i + 1) * es, 1, nccl_dtype(dtype), ncclMax, comms[k]->comm, st);
                }
                break;
            default: g_nccl.ncclGroupEnd(); VEXB_FAIL(VEXB_ERR_INVALID, "bad reduce op %d", op);
        }
        if (r != ncclSuccess) { g_nccl.ncclGroupEnd(); VEXB_FAIL(VEXB_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.ncclGetErrorString(r)); }
    }
    VEXB_NCCL(g_nccl.ncclGroupEnd());
    return VEXB_OK;
}

extern "C" int vexb_comm_barrier(int nlocal, vexb_comm *const *comms, void *const *streams) {
    VEXB_CHECK(nlocal >= 1 && comms, "bad arguments");
    VEXB_TRY(nccl_load());
    std::vector<void *> bufs(nlocal);
    for (int k = 0; k < nlocal; ++k) {
        if (!comms[k]->scratch) { DeviceGuard g(comms[k]->dev); VEXB_CUDA(cudaMalloc(&comms[k]->scratch, 64)); VEXB_CUDA(cudaMemset(comms[k]->scratch, 0, 64)); }
        bufs[k] = comms[k]->scratch;
    }
    VEXB_TRY(vexb_comm_allreduce(nlocal, comms, bufs.data(), streams, 1, VEXB_I32, VEXB_SUM));
    for (int k = 0; k < nlocal; ++k) { DeviceGuard g(comms[k]->dev); VEXB_CUDA(cudaStreamSynchronize(streams ? (cudaStream_t)streams[k] : nullptr)); }
    return VEXB_OK;
}
