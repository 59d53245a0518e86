// memory.cpp -- device-tagged, refcounted block allocator behind scanner/util/memory.h.
// (Replaces reference scanner/util/memory.cpp: System/Pool/Block allocators, :120-519, and the
// pinned-bounce memcpy path :863-1036.)
//
//  * GPU: cudaMallocAsync / cudaFreeAsync on the calling pipeline's stream from the device's
//    default memory pool with an unlimited release threshold -- a freed 400 MB frame block is
//    recycled by the next work packet without a driver call or a device-wide sync.
//  * CPU: pinned (cudaHostAlloc) blocks recycled through power-of-two free lists when CUDA is
//    usable, plain aligned malloc otherwise (CPU-only test runs).
//  * One table entry per BLOCK in an ordered map; any interior pointer resolves with one
//    upper_bound.  Refcounts are per block, as in the reference.
#include "scanner/util/memory.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <unordered_map>

#include "engine_internal.h"

namespace scanner {
namespace {

struct Block {
  size_t size;
  i32 refs;
  void* stream;  // GPU: stream the block was allocated on (frees are ordered on it too)
  bool pinned;
  bool external = false;  // adopted memory (an input stream's storage): refcounted, never freed here
};

struct DeviceTable {
  std::mutex mu;
  std::map<uintptr_t, Block> blocks;  // key = base address
  u64 current = 0, peak = 0;
};

std::mutex g_tables_mu;
std::map<std::pair<int, int>, DeviceTable*> g_tables;
MemoryPoolConfig g_config;
std::atomic<int> g_cuda_state{-1};  // -1 unknown, 0 no, 1 yes

DeviceTable& table_for(DeviceHandle d) {
  std::lock_guard<std::mutex> g(g_tables_mu);
  auto key = std::make_pair((int)(proto::DeviceType)d.type, d.is_gpu() ? d.id : 0);
  auto it = g_tables.find(key);
  if (it == g_tables.end()) it = g_tables.emplace(key, new DeviceTable()).first;
  return *it->second;
}

// ---- pinned host cache -------------------------------------------------------------------
std::mutex g_pin_mu;
std::unordered_map<size_t, std::vector<void*>> g_pin_free;  // rounded size -> free blocks

size_t round_pow2(size_t n) {
  size_t r = 256;
  while (r < n) r <<= 1;
  return r;
}

void* host_alloc(size_t size, bool& pinned) {
  pinned = false;
  if (g_config.pinned_cpu && cuda_available() && size >= 4096) {
    const size_t r = round_pow2(size);
    {
      std::lock_guard<std::mutex> g(g_pin_mu);
      auto& fl = g_pin_free[r];
      if (!fl.empty()) {
        void* p = fl.back();
        fl.pop_back();
        pinned = true;
        return p;
      }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, r, cudaHostAllocPortable) == cudaSuccess) {
      pinned = true;
      return p;
    }
    cudaGetLastError();
  }
  void* p = nullptr;
  if (posix_memalign(&p, 256, size ? size : 1) != 0)
    LOG(FATAL) << "host allocation of " << size << " B failed";
  return p;
}

void host_free(void* p, size_t size, bool pinned) {
  if (pinned) {
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pin_free[round_pow2(size)].push_back(p);
  } else {
    free(p);
  }
}

thread_local std::map<int, void*>* t_streams = nullptr;
std::mutex g_default_stream_mu;
std::map<int, void*> g_default_streams;

u8* alloc_block(DeviceHandle device, size_t size, i32 refs) {
  if (refs <= 0) refs = 1;
  Block b{size, refs, nullptr, false, false};
  void* p = nullptr;
  if (device.is_gpu()) {
    internal::ScopedDevice sd(device.id);
    b.stream = device_stream(device);
    cudaError_t e = cudaMallocAsync(&p, size ? size : 1, (cudaStream_t)b.stream);
    if (e != cudaSuccess)
      LOG(FATAL) << "GPU " << device.id << " allocation of " << size
                 << " B failed: " << cudaGetErrorString(e);
  } else {
    p = host_alloc(size, b.pinned);
  }
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  t.blocks[(uintptr_t)p] = b;
  t.current += size;
  if (t.current > t.peak) t.peak = t.current;
  return (u8*)p;
}

// find the block containing `buffer`; caller holds t.mu
std::map<uintptr_t, Block>::iterator find_block(DeviceTable& t, const u8* buffer) {
  auto it = t.blocks.upper_bound((uintptr_t)buffer);
  if (it == t.blocks.begin()) return t.blocks.end();
  --it;
  const uintptr_t base = it->first;
  const size_t span = it->second.size ? it->second.size : 1;
  if ((uintptr_t)buffer >= base + span) return t.blocks.end();
  return it;
}

}  // namespace

bool cuda_available() {
  int s = g_cuda_state.load();
  if (s < 0) {
    int n = 0;
    s = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0;
    if (!s) cudaGetLastError();
    g_cuda_state.store(s);
  }
  return s == 1;
}

void init_memory_allocators(MemoryPoolConfig config, std::vector<i32> gpu_device_ids) {
  g_config = config;
  if (!cuda_available()) return;
  for (i32 id : gpu_device_ids) {
    internal::ScopedDevice sd(id);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
      u64 thr = config.gpu_release_threshold;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
}

void destroy_memory_allocators() {
  std::lock_guard<std::mutex> g(g_pin_mu);
  for (auto& kv : g_pin_free)
    for (void* p : kv.second) cudaFreeHost(p);
  g_pin_free.clear();
}

void* device_stream(DeviceHandle device) {
  if (!device.is_gpu()) return nullptr;
  if (t_streams) {
    auto it = t_streams->find(device.id);
    if (it != t_streams->end()) return it->second;
  }
  std::lock_guard<std::mutex> g(g_default_stream_mu);
  auto it = g_default_streams.find(device.id);
  if (it != g_default_streams.end()) return it->second;
  internal::ScopedDevice sd(device.id);
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)
    LOG(FATAL) << "cannot create a stream on GPU " << device.id;
  g_default_streams[device.id] = s;
  return s;
}

namespace internal {
void set_thread_stream(int gpu_id, void* stream) {
  if (!t_streams) t_streams = new std::map<int, void*>();
  (*t_streams)[gpu_id] = stream;
}
ScopedDevice::ScopedDevice(int id) {
  if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
  if (prev_ != id)
    cudaSetDevice(id);
  else
    prev_ = -1;
}
ScopedDevice::~ScopedDevice() {
  if (prev_ >= 0) cudaSetDevice(prev_);
}
}  // namespace internal

void sync_device(DeviceHandle device) {
  if (!device.is_gpu()) return;
  internal::ScopedDevice sd(device.id);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)device_stream(device));
  if (e != cudaSuccess) LOG(FATAL) << "stream sync failed: " << cudaGetErrorString(e);
}

u8* new_buffer_(DeviceHandle device, size_t size, const char*, i32) {
  return alloc_block(device, size, 1);
}
u8* new_block_buffer_(DeviceHandle device, size_t size, i32 refs, const char*, i32) {
  return alloc_block(device, size, refs);
}
u8* new_block_buffer_sizes_(DeviceHandle device, const std::vector<size_t>& sizes, const char*,
                            i32) {
  size_t total = 0;
  for (size_t s : sizes) total += s;
  return alloc_block(device, total, (i32)sizes.size());
}
u8* new_block_buffer_size_(DeviceHandle device, size_t size, i32 copies, const char*, i32) {
  return alloc_block(device, size * (size_t)copies, copies);
}

void add_buffer_refs(DeviceHandle device, u8* buffer, i32 refs) {
  if (!buffer) return;
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  auto it = find_block(t, buffer);
  if (it == t.blocks.end())
    LOG(FATAL) << "add_buffer_ref: " << (void*)buffer << " is not a live buffer on " << device;
  it->second.refs += refs;
}
void add_buffer_ref(DeviceHandle device, u8* buffer) { add_buffer_refs(device, buffer, 1); }

void delete_buffer(DeviceHandle device, u8* buffer) {
  if (!buffer) return;
  DeviceTable& t = table_for(device);
  Block b;
  uintptr_t base;
  {
    std::lock_guard<std::mutex> g(t.mu);
    auto it = find_block(t, buffer);
    if (it == t.blocks.end())
      LOG(FATAL) << "delete_buffer: " << (void*)buffer << " is not a live buffer on " << device;
    if (--it->second.refs > 0) return;
    b = it->second;
    base = it->first;
    if (!b.external) t.current -= b.size;  // adopted memory was never counted
    t.blocks.erase(it);
  }
  if (b.external) return;
  if (device.is_gpu()) {
    internal::ScopedDevice sd(device.id);
    cudaFreeAsync((void*)base, (cudaStream_t)b.stream);
  } else {
    host_free((void*)base, b.size, b.pinned);
  }
}

namespace internal {
void adopt_block(DeviceHandle device, u8* base, size_t size) {
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  Block b{size, 1, nullptr, false, true};
  t.blocks[(uintptr_t)base] = b;
}
void disown_block(DeviceHandle device, u8* base) {
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  auto it = t.blocks.find((uintptr_t)base);
  if (it == t.blocks.end()) return;
  // drop the owner's reference; an entry somebody still references stays (orphaned, external: never
  // freed here) so that their delete_buffer finds it and erases it on the last reference
  if (--it->second.refs <= 0) t.blocks.erase(it);
}
bool block_is_external(DeviceHandle device, const u8* buffer) {
  if (!buffer) return false;
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  auto it = find_block(t, buffer);
  return it != t.blocks.end() && it->second.external;
}
}  // namespace internal

static void copy_async(u8* dst, DeviceHandle dd, const u8* src, DeviceHandle sd, size_t size) {
  if (size == 0) return;
  if (!dd.is_gpu() && !sd.is_gpu()) {
    memcpy(dst, src, size);
    return;
  }
  const DeviceHandle g = dd.is_gpu() ? dd : sd;
  internal::ScopedDevice scoped(g.id);
  cudaError_t e =
      cudaMemcpyAsync(dst, src, size, cudaMemcpyDefault, (cudaStream_t)device_stream(g));
  if (e != cudaSuccess) {
    // say what the driver thinks the two ranges are: the usual cause is a host range that is only
    // partly page-locked, or a device pointer that is no longer live
    auto describe = [](const void* p) {
      cudaPointerAttributes a;
      std::ostringstream os;
      os << p;
      if (cudaPointerGetAttributes(&a, p) == cudaSuccess)
        os << " (type " << (int)a.type << ", device " << a.device << ")";
      else
        cudaGetLastError();
      return os.str();
    };
    LOG(FATAL) << "memcpy " << sd << " -> " << dd << " of " << size << " B failed: " << cudaGetErrorString(e)
               << "; src " << describe(src) << " .. " << describe(src + size - 1) << ", dst " << describe(dst)
               << " .. " << describe(dst + size - 1) << ", stream " << device_stream(g) << " query="
               << cudaGetErrorString(cudaStreamQuery((cudaStream_t)device_stream(g)));
  }
}

void memcpy_buffer_async(u8* dst, DeviceHandle dd, const u8* src, DeviceHandle sd, size_t size) {
  copy_async(dst, dd, src, sd, size);
}

void memcpy_buffer(u8* dst, DeviceHandle dd, const u8* src, DeviceHandle sd, size_t size) {
  copy_async(dst, dd, src, sd, size);
  if (dd.is_gpu())
    sync_device(dd);
  else if (sd.is_gpu())
    sync_device(sd);
}

// true when both pointers lie inside the same live allocation of `device`
static bool same_block(DeviceHandle device, const u8* a, const u8* b) {
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  auto ia = find_block(t, a), ib = find_block(t, b);
  return ia != t.blocks.end() && ia == ib;
}

void memcpy_vec(std::vector<u8*>& dest_buffers, DeviceHandle dd, const std::vector<u8*>& src,
                DeviceHandle sd, const std::vector<size_t>& sizes) {
  // Coalesce runs that are contiguous on both sides into one DMA -- but only inside ONE
  // allocation on each side: two cudaMallocAsync blocks can be neighbours in the address space,
  // and the driver rejects a copy that spans them ("invalid argument", seen when single-frame
  // outputs of an unbatched kernel happened to be adjacent).
  size_t i = 0;
  while (i < src.size()) {
    size_t j = i, run = sizes[i];
    while (j + 1 < src.size() && src[j] + sizes[j] == src[j + 1] &&
           dest_buffers[j] + sizes[j] == dest_buffers[j + 1] && same_block(sd, src[i], src[j + 1]) &&
           same_block(dd, dest_buffers[i], dest_buffers[j + 1])) {
      ++j;
      run += sizes[j];
    }
    copy_async(dest_buffers[i], dd, src[i], sd, run);
    i = j + 1;
  }
  if (dd.is_gpu())
    sync_device(dd);
  else if (sd.is_gpu())
    sync_device(sd);
}

void copy_or_ref_buffers(std::vector<u8*>& dest_buffers, DeviceHandle dd,
                         const std::vector<u8*>& src_buffers, DeviceHandle sd,
                         const std::vector<size_t>& sizes) {
  dest_buffers.clear();
  if (dd.is_same_address_space(sd)) {
    for (u8* b : src_buffers) {
      add_buffer_ref(sd, b);
      dest_buffers.push_back(b);
    }
    return;
  }
  if (src_buffers.empty()) return;
  u8* block = new_block_buffer_sizes(dd, sizes);
  size_t off = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    dest_buffers.push_back(block + off);
    off += sizes[i];
  }
  memcpy_vec(dest_buffers, dd, src_buffers, sd, sizes);
}

u64 current_memory_allocated(DeviceHandle device) {
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  return t.current;
}
u64 max_memory_allocated(DeviceHandle device) {
  DeviceTable& t = table_for(device);
  std::lock_guard<std::mutex> g(t.mu);
  return t.peak;
}

}  // namespace scanner
