// halo.cpp -- transports for the stencil halo exchange (see halo.h).
#include "halo.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>

#include "engine_internal.h"

namespace scanner {
namespace internal {
namespace {

Result ok() {
  Result r;
  r.set_success(true);
  return r;
}

// ---- the slice of the NCCL C API this needs (nccl.h, stable since 2.7) ---------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[kHaloUniqueIdBytes];
};
enum { kNcclSuccess = 0, kNcclUint8 = 1 };  // ncclResult_t ncclSuccess, ncclDataType_t ncclUint8

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // SCN_NCCL_LIB first (torch's wheel keeps it under site-packages/nvidia/nccl/lib), then the loader path
    const char* env = getenv("SCN_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !n[0]) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("cannot load libnccl.so.2 (set SCN_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}

Result nccl_fail(const char* what, int rc) {
  Result r;
  NcclApi& a = nccl();
  RESULT_ERROR(&r, "%s: %s", what, a.GetErrorString ? a.GetErrorString(rc) : "NCCL error");
  return r;
}

class NcclTransport : public HaloTransport {
 public:
  NcclTransport(i32 gpu, i32 rank, i32 world) : gpu_(gpu), rank_(rank), world_(world) {}
  ~NcclTransport() override {
    ScopedDevice sd(gpu_);
    if (comm_) nccl().CommDestroy(comm_);
    if (stream_) cudaStreamDestroy(stream_);
  }
  Result init(const u8 id[kHaloUniqueIdBytes]) {
    Result r;
    NcclApi& a = nccl();
    if (!a.error.empty()) {
      RESULT_ERROR(&r, "%s", a.error.c_str());
      return r;
    }
    ScopedDevice sd(gpu_);
    if (cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) != cudaSuccess) {
      RESULT_ERROR(&r, "cannot create a stream on GPU %d", gpu_);
      return r;
    }
    NcclUniqueId uid;
    memcpy(uid.internal, id, kHaloUniqueIdBytes);
    const int rc = a.CommInitRank(&comm_, world_, uid, rank_);
    if (rc != kNcclSuccess) return nccl_fail("ncclCommInitRank", rc);
    return ok();
  }
  i32 rank() const override { return rank_; }
  i32 world() const override { return world_; }
  bool device_buffers() const override { return true; }
  i32 gpu_id() const override { return gpu_; }
  Result exchange(const std::vector<HaloXfer>& xfers) override {
    if (xfers.empty()) return ok();
    NcclApi& a = nccl();
    ScopedDevice sd(gpu_);
    int rc = a.GroupStart();
    if (rc != kNcclSuccess) return nccl_fail("ncclGroupStart", rc);
    for (const HaloXfer& x : xfers) {
      rc = x.send ? a.Send(x.buffer, x.bytes, kNcclUint8, x.peer, comm_, stream_)
                  : a.Recv(x.buffer, x.bytes, kNcclUint8, x.peer, comm_, stream_);
      if (rc != kNcclSuccess) {
        a.GroupEnd();
        return nccl_fail(x.send ? "ncclSend" : "ncclRecv", rc);
      }
    }
    rc = a.GroupEnd();
    if (rc != kNcclSuccess) return nccl_fail("ncclGroupEnd", rc);
    if (cudaStreamSynchronize(stream_) != cudaSuccess) {
      Result r;
      RESULT_ERROR(&r, "halo exchange failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
      return r;
    }
    return ok();
  }

 private:
  i32 gpu_, rank_, world_;
  ncclComm_t comm_ = nullptr;
  cudaStream_t stream_ = nullptr;
};

class CallbackTransport : public HaloTransport {
 public:
  CallbackTransport(i32 rank, i32 world, HaloExchangeFn fn, void* user) : rank_(rank), world_(world), fn_(fn), user_(user) {}
  i32 rank() const override { return rank_; }
  i32 world() const override { return world_; }
  bool device_buffers() const override { return false; }
  Result exchange(const std::vector<HaloXfer>& xfers) override {
    if (xfers.empty()) return ok();
    std::vector<int> peers, sends;
    std::vector<void*> bufs;
    std::vector<uint64_t> bytes;
    for (const HaloXfer& x : xfers) {
      peers.push_back(x.peer);
      bufs.push_back(x.buffer);
      bytes.push_back(x.bytes);
      sends.push_back(x.send ? 1 : 0);
    }
    const int rc = fn_(user_, (int)xfers.size(), peers.data(), bufs.data(), bytes.data(), sends.data());
    Result r = ok();
    if (rc != 0) RESULT_ERROR(&r, "halo exchange callback failed (%d)", rc);
    return r;
  }

 private:
  i32 rank_, world_;
  HaloExchangeFn fn_;
  void* user_;
};

}  // namespace

Result halo_nccl_unique_id(u8 out[kHaloUniqueIdBytes]) {
  Result r;
  NcclApi& a = nccl();
  if (!a.error.empty()) {
    RESULT_ERROR(&r, "%s", a.error.c_str());
    return r;
  }
  NcclUniqueId uid;
  const int rc = a.GetUniqueId(&uid);
  if (rc != kNcclSuccess) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out, uid.internal, kHaloUniqueIdBytes);
  return ok();
}

Result make_nccl_transport(i32 gpu_id, i32 rank, i32 world, const u8 id[kHaloUniqueIdBytes],
                           std::unique_ptr<HaloTransport>& out) {
  std::unique_ptr<NcclTransport> t(new NcclTransport(gpu_id, rank, world));
  Result r = t->init(id);
  if (r.success()) out = std::move(t);
  return r;
}

std::unique_ptr<HaloTransport> make_callback_transport(i32 rank, i32 world, HaloExchangeFn fn, void* user) {
  return std::unique_ptr<HaloTransport>(new CallbackTransport(rank, world, fn, user));
}

}  // namespace internal
}  // namespace scanner
