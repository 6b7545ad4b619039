// gpu_vector.h -- minimal device array with the interface of GPUMD's GPU_Vector<T>
// (src/utilities/gpu_vector.cuh:37-207: resize / size / data / copy_from_host / copy_to_host / fill),
// so that the adapter classes in this directory have the exact signatures of the reference's
// Potential / Ensemble virtuals.  Inside GPUMD itself the reference's own GPU_Vector is used.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define B2H_CHECK(call)                                                                       \
  do {                                                                                        \
    const cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "CUDA Error:\n    File: %s\n    Line: %d\n    Error text: %s\n", __FILE__, \
              __LINE__, cudaGetErrorString(e_));                                              \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)

template <typename T>
class GPU_Vector
{
public:
  GPU_Vector() = default;
  explicit GPU_Vector(size_t n) { resize(n); }
  GPU_Vector(const GPU_Vector&) = delete;
  GPU_Vector& operator=(const GPU_Vector&) = delete;
  GPU_Vector(GPU_Vector&& o) noexcept : data_(o.data_), size_(o.size_)
  {
    o.data_ = nullptr;
    o.size_ = 0;
  }
  ~GPU_Vector() { release(); }

  void resize(size_t n)
  {
    if (n == size_)
      return;
    release();
    if (n > 0)
      B2H_CHECK(cudaMalloc((void**)&data_, n * sizeof(T)));
    size_ = n;
  }
  void resize(size_t n, const T value)
  {
    resize(n);
    fill(value);
  }
  void fill(const T value)
  {
    std::vector<T> h(size_, value);
    copy_from_host(h.data());
  }
  void copy_from_host(const T* h) { copy_from_host(h, size_); }
  void copy_from_host(const T* h, size_t n)
  {
    if (n > 0)
      B2H_CHECK(cudaMemcpy(data_, h, n * sizeof(T), cudaMemcpyHostToDevice));
  }
  void copy_to_host(T* h) const { copy_to_host(h, size_); }
  void copy_to_host(T* h, size_t n) const
  {
    if (n > 0)
      B2H_CHECK(cudaMemcpy(h, data_, n * sizeof(T), cudaMemcpyDeviceToHost));
  }
  size_t size() const { return size_; }
  T* data() { return data_; }
  const T* data() const { return data_; }

private:
  void release()
  {
    if (data_)
      cudaFree(data_);
    data_ = nullptr;
    size_ = 0;
  }
  T* data_ = nullptr;
  size_t size_ = 0;
};
