// Stand-in for the reference's scanner/util/storehouse.h: s_write as upstream (:26-32), appending to the stub WriteFile.
#pragma once
#include "scanner/util/common.h"
#include "storehouse/storage_backend.h"
namespace scanner {
inline void s_write(storehouse::WriteFile* file, const u8* buffer, size_t size) { file->append(size, buffer); }
}  // namespace scanner
