// devbuf.h -- growable device / pinned-host scratch buffers owned by a ctx.
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>

#include "cuda_compat.h"

namespace bsk {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define BSK_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      throw ::bsk::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  // contents are NOT preserved when the buffer grows
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    want = (want + 255) & ~(size_t)255;
    BSK_CUDA(cudaMalloc(&p, want));
    cap = want;
  }
  template <class T> T *get(size_t n) { reserve(n * sizeof(T)); return static_cast<T *>(p); }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuf {
  void *p = nullptr;
  size_t cap = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void reserve(size_t bytes, bool preserve = false, size_t used = 0) {
    if (bytes <= cap) return;
    size_t want = bytes + bytes / 4 + 4096;
    void *q = nullptr;
    BSK_CUDA(cudaHostAlloc(&q, want, cudaHostAllocMapped | cudaHostAllocPortable));  // kernels may write it (prim::copy_small)
    if (p) {
      if (preserve && used) memcpy(q, p, used);
      cudaFreeHost(p);
    }
    p = q;
    cap = want;
  }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace bsk
