// Stand-in for storehouse/storage_backend.h (the reference's storage library, not in this image) while
// oracle/Makefile compiles the reference's H.264 index creator unmodified: the class only appends to a WriteFile.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace storehouse {
enum class StoreResult { Success = 0 };
class WriteFile {
 public:
  StoreResult append(size_t size, const uint8_t* data) {
    bytes.insert(bytes.end(), data, data + size);
    return StoreResult::Success;
  }
  std::vector<uint8_t> bytes;
};
class RandomReadFile {};
}  // namespace storehouse
