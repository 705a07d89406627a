// Run-time binding of the handful of NCCL entry points the sharded V-cycle
// uses (halo send/recv with the two neighbouring ranks, broadcast of the
// agglomerated level, one-double all-reduce).  libnccl.so.2 is dlopen'ed on
// first use so that single-GPU users of libamgb.so carry no NCCL dependency;
// inside a torch process the already-loaded (torch-bundled) NCCL is reused.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>

namespace amgb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  static NcclApi& get() {
    static NcclApi api = load();
    return api;
  }

 private:
  template <class F>
  static void bind(void* h, const char* name, F& fn) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    if (!fn) throw std::runtime_error(std::string("libnccl.so.2 lacks symbol ") + name);
  }
  static NcclApi load() {
    NcclApi a;
    a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.handle) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
    bind(a.handle, "ncclGetUniqueId", a.GetUniqueId);
    bind(a.handle, "ncclCommInitRank", a.CommInitRank);
    bind(a.handle, "ncclCommDestroy", a.CommDestroy);
    bind(a.handle, "ncclSend", a.Send);
    bind(a.handle, "ncclRecv", a.Recv);
    bind(a.handle, "ncclBroadcast", a.Broadcast);
    bind(a.handle, "ncclAllReduce", a.AllReduce);
    bind(a.handle, "ncclGroupStart", a.GroupStart);
    bind(a.handle, "ncclGroupEnd", a.GroupEnd);
    bind(a.handle, "ncclGetErrorString", a.GetErrorString);
    return a;
  }
};

}  // namespace amgb
