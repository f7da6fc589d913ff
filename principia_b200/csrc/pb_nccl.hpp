// NCCL, resolved at run time: the library has no link-time dependency on
// libnccl, so that it loads on a box without it and, in a process that already
// carries an NCCL (the caller created the communicator with it; PyTorch bundles
// its own), uses THAT one instead of bringing a second copy.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

namespace pb {

// The part of nccl.h this library needs (NCCL 2.x ABI).
using NcclComm = struct ncclComm*;
constexpr int kNcclSuccess = 0;
constexpr int kNcclFloat64 = 8;  // ncclDouble.
constexpr int kNcclUniqueIdBytes = 128;
struct NcclUniqueId {
  char internal[kNcclUniqueIdBytes];
};

struct NcclApi {
  int (*all_gather)(void const* send, void* receive, std::size_t send_count,
                    int datatype, NcclComm comm, cudaStream_t stream) = nullptr;
  int (*comm_count)(NcclComm comm, int* count) = nullptr;
  int (*comm_user_rank)(NcclComm comm, int* rank) = nullptr;
  int (*get_unique_id)(NcclUniqueId* id) = nullptr;
  int (*comm_init_rank)(NcclComm* comm, int n_ranks, NcclUniqueId id,
                        int rank) = nullptr;
  int (*comm_destroy)(NcclComm comm) = nullptr;
  char const* (*get_error_string)(int result) = nullptr;
  bool loaded = false;
};

inline NcclApi const& Nccl() {
  static NcclApi const api = [] {
    NcclApi a;
    // The copy already in the process, whatever the scope it was loaded in.
    void* handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (handle == nullptr) {
      handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    }
    if (handle == nullptr) {
      handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    }
    if (handle == nullptr) {
      return a;
    }
    auto resolve = [&](char const* const name) { return dlsym(handle, name); };
    a.all_gather =
        reinterpret_cast<decltype(a.all_gather)>(resolve("ncclAllGather"));
    a.comm_count =
        reinterpret_cast<decltype(a.comm_count)>(resolve("ncclCommCount"));
    a.comm_user_rank = reinterpret_cast<decltype(a.comm_user_rank)>(
        resolve("ncclCommUserRank"));
    a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(
        resolve("ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(
        resolve("ncclCommInitRank"));
    a.comm_destroy =
        reinterpret_cast<decltype(a.comm_destroy)>(resolve("ncclCommDestroy"));
    a.get_error_string = reinterpret_cast<decltype(a.get_error_string)>(
        resolve("ncclGetErrorString"));
    a.loaded = a.all_gather != nullptr && a.comm_count != nullptr &&
               a.comm_user_rank != nullptr && a.get_unique_id != nullptr &&
               a.comm_init_rank != nullptr && a.comm_destroy != nullptr;
    return a;
  }();
  return api;
}

}  // namespace pb
