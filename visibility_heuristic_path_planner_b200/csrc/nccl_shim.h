// nccl_shim.h -- the handful of NCCL entry points the strip-partitioned planner uses, resolved
// at run time with dlopen/dlsym: the library has no link-time dependency on NCCL (batches need
// no collective at all), and inside a torch process it binds to the libnccl.so.2 torch already
// loaded instead of a second copy.  Types follow nccl.h (2.18+: ncclCommSplit).
#ifndef VHP_NCCL_SHIM_H
#define VHP_NCCL_SHIM_H

#include <cuda_runtime.h>

#include <cstddef>
#include <string>

struct ncclComm;
typedef ncclComm *vhpNcclComm;
struct vhpNcclUniqueId { char internal[128]; };
enum { kNcclInt32 = 2, kNcclUint64 = 5, kNcclFloat64 = 8 }; // ncclDataType_t
enum { kNcclMax = 2 };                                      // ncclRedOp_t

struct VhpNccl {
  int (*GetVersion)(int *) = nullptr;
  int (*GetUniqueId)(vhpNcclUniqueId *) = nullptr;
  int (*CommInitRank)(vhpNcclComm *, int, vhpNcclUniqueId, int) = nullptr;
  int (*CommSplit)(vhpNcclComm, int, int, vhpNcclComm *, void *) = nullptr;
  int (*CommDestroy)(vhpNcclComm) = nullptr;
  int (*Send)(const void *, size_t, int, int, vhpNcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, vhpNcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, vhpNcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, vhpNcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string path; // what was opened
};

// nullptr (and *why) when no usable libnccl.so.2 is found.  env VHP_NCCL_LIB overrides the search.
const VhpNccl *vhp_nccl(std::string *why);

#endif
