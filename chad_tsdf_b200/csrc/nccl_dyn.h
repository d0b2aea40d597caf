// NCCL bound at run time (dlopen "libnccl.so.2"): the single-GPU library has no link-time dependency on NCCL, and a process that
// has already loaded one (e.g. the copy bundled with PyTorch) shares it instead of mapping a second one. Only the handful of calls
// the Morton-range sharding needs: communicator set-up, grouped send / recv (the per-batch all-to-all of boundary runs, the
// per-submap gather of leaf chunks), all-gather and broadcast of a few words.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>  // types and enums only; every function is called through the table below

namespace chadgpu {

struct NcclApi {
    ncclResult_t (*GetVersion)(int*);
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommAbort)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};

// nullptr (and *why set) when no usable libnccl.so.2 can be loaded
const NcclApi* nccl_api(const char** why);

}  // namespace chadgpu
