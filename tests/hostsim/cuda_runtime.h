// TEST INFRASTRUCTURE ONLY: a host stand-in for the handful of CUDA constructs the simple kernels of
// pygcn_b200/csrc/batchnorm.cu use, so that the REAL kernel and launcher source can be compiled with g++ and executed on
// the CPU (tests/test_batchnorm_hostsim.py) while no GPU is at hand.  A "launch" runs the blocks one after another, each
// block as blockDim.x real threads; __syncthreads is a std::barrier over the block, __shfl_xor_sync an exchange through
// a per-warp buffer, __shared__ a function-local static (blocks never overlap).  Nothing under pygcn_b200/ includes this.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct hostsim_uint3 {
  unsigned x = 0, y = 0, z = 0;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline thread_local hostsim_uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

struct alignas(8) uint2 {
  unsigned x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <typename T>
inline T __ldg(const T* p) { return *p; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "hostsim"; }

// names common.cuh's launch_pdl template mentions (never instantiated here)
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { int id; cudaLaunchAttributeValue val; };
enum { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes;
  cudaStream_t stream;
  cudaLaunchAttribute* attrs;
  int numAttrs;
};
template <typename... A, typename... B>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t*, void (*)(A...), B&&...) { return cudaSuccess; }

struct hostsim_warp {
  std::barrier<> bar{32};
  double buf[32];
};
inline thread_local std::barrier<>* hostsim_block_barrier = nullptr;
inline thread_local hostsim_warp* hostsim_my_warp = nullptr;

inline void __syncthreads() { hostsim_block_barrier->arrive_and_wait(); }
inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  const int lane = threadIdx.x & 31;
  hostsim_my_warp->buf[lane] = v;
  hostsim_my_warp->bar.arrive_and_wait();
  const double r = hostsim_my_warp->buf[lane ^ lane_mask];
  hostsim_my_warp->bar.arrive_and_wait();
  return r;
}

// kernel<<<grid, block, smem, stream>>>(args) is rewritten by the test into hostsim_launch(grid, block, [&] { kernel(args); })
template <typename F>
inline void hostsim_launch(dim3 grid, dim3 block, F&& body) {
  for (unsigned b = 0; b < grid.x; ++b) {
    std::barrier<> bar((ptrdiff_t)block.x);
    std::vector<std::unique_ptr<hostsim_warp>> warps;
    for (unsigned w = 0; w < (block.x + 31) / 32; ++w) warps.emplace_back(new hostsim_warp());
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < block.x; ++t)
      threads.emplace_back([&, t] {
        threadIdx.x = t;
        blockIdx.x = b;
        blockDim = block;
        gridDim = grid;
        hostsim_block_barrier = &bar;
        hostsim_my_warp = warps[t / 32].get();
        body();
      });
    for (auto& th : threads) th.join();
  }
}
