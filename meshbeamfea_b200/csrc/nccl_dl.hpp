// NCCL entry points resolved at run time (dlopen), so libmbfea.so has no
// link-time dependency on a particular libnccl: single-GPU users never load it,
// and inside a torch process the copy torch already mapped is reused.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

namespace mbfea { namespace nccl {

struct UniqueId { char internal[128]; };  // == ncclUniqueId
typedef struct ncclComm *Comm;
enum DataType { kInt32 = 2, kInt64 = 4, kFloat64 = 8 };  // ncclInt32, ncclInt64, ncclFloat64
enum RedOp { kSum = 0, kMax = 2 };  // ncclSum, ncclMax

struct Api {
    int (*GetUniqueId)(UniqueId *);
    int (*CommInitRank)(Comm *, int, UniqueId, int);
    int (*CommDestroy)(Comm);
    const char *(*GetErrorString)(int);
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t);
    int (*Broadcast)(const void *, void *, size_t, int, int, Comm, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, Comm, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, Comm, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*GetVersion)(int *);
};

// nullptr + msg on failure
const Api *load(std::string &msg);

}}  // namespace mbfea::nccl
