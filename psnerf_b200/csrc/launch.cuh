// Host-side launch helpers shared by the translation units.
#pragma once
#include "common.cuh"

namespace psn {

// Persistent kernels run one CTA per SM (148 on B200); cached per process.
int num_ctas();

// Bump allocator over the caller-provided workspace.
struct Workspace {
  char* base;
  size_t size, used;
  bool ok;
  Workspace(void* p, long long bytes) : base((char*)p), size(bytes > 0 ? (size_t)bytes : 0), used(0), ok(true) {}
  template <class T>
  T* take(size_t count) {
    const size_t off = (used + 255) / 256 * 256;
    const size_t end = off + count * sizeof(T);
    if (!base || end > size) {
      ok = false;
      used = end;
      return nullptr;
    }
    used = end;
    return reinterpret_cast<T*>(base + off);
  }
};

}  // namespace psn
