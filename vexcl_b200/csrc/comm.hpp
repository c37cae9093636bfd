// NCCL function table (resolved with dlopen) and the communicator handle.
#pragma once
#include "common.cuh"
#include <nccl.h>

struct vexb_comm {
    int dev = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    void *scratch = nullptr;
};

namespace vexb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*ncclGetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*ncclCommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*ncclCommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*ncclCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*ncclAllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ncclSend)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ncclRecv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ncclGroupStart)() = nullptr;
    ncclResult_t (*ncclGroupEnd)() = nullptr;
    const char *(*ncclGetErrorString)(ncclResult_t) = nullptr;
};

extern NcclApi g_nccl;
int nccl_load();
ncclDataType_t nccl_dtype(int dt);

} // namespace vexb
