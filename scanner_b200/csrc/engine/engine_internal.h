// engine_internal.h -- small helpers shared by the engine translation units.
#pragma once
#include "scanner/util/common.h"

namespace scanner {
namespace internal {

// Bind a CUDA stream to (this thread, gpu): device_stream() returns it, so allocations, copies
// and kernels issued from this pipeline-instance thread are all ordered on one stream.
void set_thread_stream(int gpu_id, void* stream);

// Make externally owned memory (an input stream's frame storage) addressable by the refcounting
// API without copying it: add_buffer_ref / delete_buffer work on it, the memory is never freed.
void adopt_block(DeviceHandle device, u8* base, size_t size);
void disown_block(DeviceHandle device, u8* base);
// true when `buffer` lies in adopted memory (its lifetime is the input stream's, not the block's refcount)
bool block_is_external(DeviceHandle device, const u8* buffer);

struct ScopedDevice {
  explicit ScopedDevice(int id);
  ~ScopedDevice();
  int prev_ = -1;
};

}  // namespace internal
}  // namespace scanner
