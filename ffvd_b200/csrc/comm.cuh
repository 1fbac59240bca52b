// Multi-GPU collective behind the C ABI: one NCCL communicator per context, created from a 128-byte
// unique id that the caller ships between its ranks by any means (MPI, torch.distributed, a file).  The only data-path
// collective of the hot path is ONE all-reduce per evaluation of the packed shared-parameter gradients (and, for
// time-sharded evaluations, the scalar objective) -- x-bar never leaves its owner.
// NCCL is bound at run time (dlopen of libnccl.so.2; FFVD_NCCL_LIB overrides) so that single-GPU users need no NCCL and a
// process that already loaded one (PyTorch bundles its own) shares it.
#pragma once
#include <dlfcn.h>
#include <cuda_runtime.h>

namespace ffvd {

// the slice of nccl.h this library uses (ABI-stable since NCCL 2.0)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

inline const NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("FFVD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
      api.GetVersion = (int (*)(int*))dlsym(api.handle, "ncclGetVersion");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) api.handle = nullptr;
    }
  }
  if (!api.handle) {
    if (err) *err = "NCCL (libnccl.so.2) could not be loaded; set FFVD_NCCL_LIB to its path";
    return nullptr;
  }
  return &api;
}

// gather / scatter of up to 12 small tensors into / out of one contiguous buffer (a single launch each way)
struct PackList {
  double* ptr[12];
  long long off[13];     // prefix sums of the element counts; off[n] = total
  int n;
};
__global__ void pack_kernel(PackList pl, double* __restrict__ buf, int scatter) {
  const long long total = pl.off[pl.n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= pl.off[k + 1]) ++k;
    if (scatter) pl.ptr[k][i - pl.off[k]] = buf[i];
    else buf[i] = pl.ptr[k][i - pl.off[k]];
  }
}

}  // namespace ffvd
