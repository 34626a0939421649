/*
 * multi.h — NCCL plumbing of the multi-GPU merge (multi.cpp).  Not part of the ABI.
 *
 * libnccl is loaded at run time (the copy the process already holds — torch ships one — or the system's), so libb2q has
 * no link-time dependency on it and single-GPU use never touches it.
 */
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <string>

namespace b2q {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
  int (*GetVersion)(int*);
};

/* nullptr (and *why set) when no libnccl.so.2 can be loaded */
const NcclApi* nccl_api(std::string* why);

}  // namespace b2q

/* one rank of a communicator: a device, its NCCL handle and the stream the merge is enqueued on */
struct B2QComm {
  ncclComm_t comm = nullptr;
  int32_t rank = 0, nranks = 1, device = 0;
};
