// devbuf.h — owning device allocation with the few operations the library needs.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <vector>

namespace dtb {

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;  // elements allocated
  // Contents are NOT kept when the buffer has to grow.
  cudaError_t reserve(size_t want) {
    if (want <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (want == 0) want = 1;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  // Grows to at least `want` elements (geometrically) and keeps the first `keep` elements.
  cudaError_t grow(size_t want, size_t keep, cudaStream_t s) {
    if (want <= n && p) return cudaSuccess;
    size_t cap = std::max<size_t>(want, n + n / 2);
    if (cap == 0) cap = 1;
    T *q = nullptr;
    cudaError_t e = cudaMalloc((void **)&q, cap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) {
      e = cudaMemcpyAsync(q, p, std::min(keep, n) * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q;
    n = cap;
    return e;
  }
  cudaError_t upload(const std::vector<T> &v, cudaStream_t s) {
    cudaError_t e = reserve(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace dtb
