// halo.h -- the one exchange step of the data path: boundary frames of a clip whose rows are split
// into contiguous intervals across ranks (BASELINE configs[3]).
//
// The reference gives every task its stencil halo by loading and DECODING the extra rows again
// (derive_stencil_requirements adds row+s for every stencil offset, dag_analysis.cpp:1634-1657; the
// gather at evaluate_worker.cpp:1068-1089 then finds them in the task's own element cache).  Here a
// rank decodes only the rows of its own interval; rows a neighbouring rank's stencil reaches into
// are sent to it as decoded elements -- the packed NV12 surface when every consumer accepts it
// (3.1 MB per 1080p frame), RGB24 otherwise -- with ncclSend / ncclRecv over NVLink, all transfers of a
// run in ONE group before the pipeline instances start, so no instance stream ever waits for a peer.
// Per-frame ops need no exchange; this is the path's only collective.
//
// libnccl is resolved with dlopen (the library torch ships, or the system one): no link-time
// dependency, same as libnvcuvid.  CPU-only runs (tests: world_size-2 gloo) plug in a callback
// transport that moves host buffers with whatever the embedding process has (torch.distributed).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "scanner/util/common.h"

namespace scanner {
namespace internal {

struct HaloXfer {
  i32 peer;      // rank on the other side
  u8* buffer;    // device memory (NCCL transport) or host memory (callback transport)
  size_t bytes;
  bool send;
};

class HaloTransport {
 public:
  virtual ~HaloTransport() = default;
  virtual i32 rank() const = 0;
  virtual i32 world() const = 0;
  virtual bool device_buffers() const = 0;  // true: buffers are GPU memory of gpu_id()
  virtual i32 gpu_id() const { return -1; }
  // Performs every transfer; pairs of ranks list their mutual transfers in the same order.
  virtual Result exchange(const std::vector<HaloXfer>& xfers) = 0;
};

constexpr size_t kHaloUniqueIdBytes = 128;  // sizeof(ncclUniqueId)

// rank 0 calls this and hands the bytes to every rank (any side channel: torch.distributed, a file)
Result halo_nccl_unique_id(u8 out[kHaloUniqueIdBytes]);
// collective over all ranks of the job
Result make_nccl_transport(i32 gpu_id, i32 rank, i32 world, const u8 id[kHaloUniqueIdBytes],
                           std::unique_ptr<HaloTransport>& out);

// n transfers: peers[i], buffers[i], bytes[i], is_send[i]; returns 0 on success
using HaloExchangeFn = int (*)(void* user, int n, const int* peers, void* const* buffers, const uint64_t* bytes,
                               const int* is_send);
std::unique_ptr<HaloTransport> make_callback_transport(i32 rank, i32 world, HaloExchangeFn fn, void* user);

}  // namespace internal
}  // namespace scanner
