// Peer groups: mailboxes in device memory that every rank of the job can write over NVLink
// (include/vexb200.h, "Peer memory").  See peer.cuh for the protocol.
#include "peer.cuh"
#include <vector>
#include <mutex>

namespace vexb {

static const size_t kMailboxWords = 16 + 2 * VEXB_MAX_PEERS * 4;

template <class T> __device__ __forceinline__ unsigned long long bits_of(T v) { unsigned long long u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> __device__ __forceinline__ T of_bits(unsigned long long u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// In-place all-reduce of ONE value (two for MINMAX) per rank, for values that did not come out of
// vexb_reduce_all (identity of an empty slice, user buffers).
template <class T>
__global__ void peer_allreduce_kernel(PeerArgs pa, T *buf, int op) {
    __shared__ unsigned long long px[VEXB_MAX_PEERS], py[VEXB_MAX_PEERS];
    const unsigned long long v0 = bits_of<T>(buf[0]);
    const unsigned long long v1 = op == VEXB_MINMAX ? bits_of<T>(buf[1]) : 0ull;
    const bool arrived = peer_exchange(pa, v0, v1, px, py);
    if (threadIdx.x == 0 && !arrived) { buf[0] = peer_poison<T>(); if (op == VEXB_MINMAX) buf[1] = peer_poison<T>(); }
    if (threadIdx.x == 0 && arrived) {
        T a = of_bits<T>(px[0]), b = of_bits<T>(py[0]);
        for (int r = 1; r < pa.nranks; ++r) {
            const T x = of_bits<T>(px[r]), y = of_bits<T>(py[r]);
            switch (op) {
                case VEXB_MAX: a = a > x ? a : x; break;
                case VEXB_MIN: a = a < x ? a : x; break;
                case VEXB_MINMAX: a = a < x ? a : x; b = b > y ? b : y; break;
                default: a = a + x; break;
            }
        }
        buf[0] = a;
        if (op == VEXB_MINMAX) buf[1] = b;
    }
}

// One word per process, written by kernels of any device when a peer fails to arrive (pinned + mapped + portable:
// with unified addressing the host pointer is valid on every device).
unsigned long long *peer_fault_word() {
    static std::mutex mx;
    static unsigned long long *word = nullptr;
    std::lock_guard<std::mutex> lock(mx);
    if (!word) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memset(p, 0, 64);
        word = (unsigned long long *)p;
    }
    return word;
}

static int alloc_mailbox(vexb_peer *P) {
    P->fault_host = peer_fault_word();
    VEXB_CUDA(cudaMalloc((void **)&P->mailbox, kMailboxWords * 8));
    VEXB_CUDA(cudaMemset(P->mailbox, 0, kMailboxWords * 8));
    VEXB_CUDA(cudaDeviceSynchronize());
    P->peers[P->rank] = P->mailbox;
    return VEXB_OK;
}

} // namespace vexb

using namespace vexb;

extern "C" int vexb_peer_create(int dev, int rank, int nranks, vexb_peer **peer, void *handle64) {
    VEXB_CHECK(peer && handle64, "NULL argument");
    VEXB_CHECK(nranks >= 1 && nranks <= VEXB_MAX_PEERS && rank >= 0 && rank < nranks, "bad rank %d of %d (max %d)", rank, nranks, VEXB_MAX_PEERS);
    DeviceGuard g(dev); VEXB_CHECK(g.ok, "cannot select device %d", dev);
    auto *P = new vexb_peer();
    P->dev = dev; P->rank = rank; P->nranks = nranks;
    int st = alloc_mailbox(P);
    if (st != VEXB_OK) { delete P; return st; }
    static_assert(sizeof(cudaIpcMemHandle_t) == VEXB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size changed");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, P->mailbox);
    if (e != cudaSuccess) { cudaFree(P->mailbox); delete P; VEXB_FAIL(VEXB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, sizeof(h));
    *peer = P;
    return VEXB_OK;
}

extern "C" int vexb_peer_connect(vexb_peer *peer, const void *handles) {
    VEXB_CHECK(peer && handles, "NULL argument");
    DeviceGuard g(peer->dev); VEXB_CHECK(g.ok, "cannot select device %d", peer->dev);
    for (int p = 0; p < peer->nranks; ++p) {
        if (p == peer->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)p * VEXB_IPC_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        VEXB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        peer->peers[p] = (unsigned long long *)ptr;
        peer->ipc_opened[p] = true;
    }
    return VEXB_OK;
}

extern "C" int vexb_peer_create_all(int ndev, const int *devs, vexb_peer **peers) {
    VEXB_CHECK(ndev >= 1 && ndev <= VEXB_MAX_PEERS && devs && peers, "bad arguments");
    for (int a = 0; a < ndev; ++a) for (int b = a + 1; b < ndev; ++b)
        VEXB_CHECK(devs[a] != devs[b], "peer groups need distinct devices (device %d appears twice)", devs[a]);
    std::vector<vexb_peer *> ps(ndev, nullptr);
    for (int k = 0; k < ndev; ++k) {
        DeviceGuard g(devs[k]); VEXB_CHECK(g.ok, "cannot select device %d", devs[k]);
        ps[k] = new vexb_peer();
        ps[k]->dev = devs[k]; ps[k]->rank = k; ps[k]->nranks = ndev;
        int st = alloc_mailbox(ps[k]);
        if (st != VEXB_OK) return st;
        for (int j = 0; j < ndev; ++j) if (j != k) {
            int can = 0;
            VEXB_CUDA(cudaDeviceCanAccessPeer(&can, devs[k], devs[j]));
            if (!can) VEXB_FAIL(VEXB_ERR_UNSUPPORTED, "device %d cannot access device %d", devs[k], devs[j]);
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VEXB_CUDA(e);
            cudaGetLastError();
        }
    }
    for (int k = 0; k < ndev; ++k) { for (int j = 0; j < ndev; ++j) ps[k]->peers[j] = ps[j]->mailbox; peers[k] = ps[k]; }
    return VEXB_OK;
}

extern "C" int vexb_peer_destroy(vexb_peer *peer) {
    if (!peer) return VEXB_OK;
    VEXB_RELEASE_GUARD();
    DeviceGuard g(peer->dev);
    for (int p = 0; p < peer->nranks; ++p) if (peer->ipc_opened[p]) cudaIpcCloseMemHandle(peer->peers[p]);
    cudaFree(peer->mailbox);
    delete peer;
    return VEXB_OK;
}

extern "C" int vexb_peer_error(vexb_peer *peer, unsigned long long *epoch_of_timeout) {
    VEXB_CHECK(peer && epoch_of_timeout, "NULL argument");
    DeviceGuard g(peer->dev);
    VEXB_CUDA(cudaMemcpy(epoch_of_timeout, peer->mailbox + 1, 8, cudaMemcpyDeviceToHost));
    return VEXB_OK;
}

extern "C" int vexb_peer_fault(unsigned long long *epoch, int clear) {
    unsigned long long *w = peer_fault_word();
    const unsigned long long v = w ? *(volatile unsigned long long *)w : 0ull;
    if (epoch) *epoch = v;
    if (clear && w) *(volatile unsigned long long *)w = 0ull;
    return VEXB_OK;
}

extern "C" int vexb_peer_allreduce(vexb_peer *peer, void *stream, void *d_buf, int dtype, int op) {
    VEXB_CHECK(peer && d_buf, "NULL argument");
    VEXB_CHECK(dtype >= VEXB_F64 && dtype <= VEXB_U64, "bad dtype %d", dtype);
    VEXB_CHECK(op >= VEXB_SUM && op <= VEXB_MINMAX, "bad reduce op %d", op);
    if (peer->nranks <= 1) return VEXB_OK;
    DeviceGuard g(peer->dev); VEXB_CHECK(g.ok, "cannot select device %d", peer->dev);
    cudaStream_t st = (cudaStream_t)stream;
    const PeerArgs pa = peer->args();
    switch (dtype) {
        case VEXB_F64: peer_allreduce_kernel<double><<<1, 32, 0, st>>>(pa, (double *)d_buf, op); break;
        case VEXB_F32: peer_allreduce_kernel<float><<<1, 32, 0, st>>>(pa, (float *)d_buf, op); break;
        case VEXB_I32: peer_allreduce_kernel<int><<<1, 32, 0, st>>>(pa, (int *)d_buf, op); break;
        case VEXB_U32: peer_allreduce_kernel<unsigned><<<1, 32, 0, st>>>(pa, (unsigned *)d_buf, op); break;
        case VEXB_I64: peer_allreduce_kernel<long long><<<1, 32, 0, st>>>(pa, (long long *)d_buf, op); break;
        default:       peer_allreduce_kernel<unsigned long long><<<1, 32, 0, st>>>(pa, (unsigned long long *)d_buf, op); break;
    }
    VEXB_LAUNCHED();
    return VEXB_OK;
}
