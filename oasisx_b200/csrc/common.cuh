// common.cuh -- error handling, device buffers and reduction helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define B2_MAXK 3  // at most 3 velocity components solved together

struct B2Error : std::runtime_error {
  int code;
  B2Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define B2_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw B2Error(-100, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                              ":" + std::to_string(__LINE__) + ")");                           \
  } while (0)

#define B2_REQUIRE(cond, msg)                      \
  do {                                             \
    if (!(cond)) throw B2Error(-1, std::string(msg)); \
  } while (0)

template <typename T>
struct DBuf {  // owning device buffer
  T* p = nullptr;
  int64_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(int64_t count) {
    release();
    if (count > 0) B2_CUDA(cudaMalloc(&p, sizeof(T) * (size_t)count));
    n = count;
  }
  void zero(cudaStream_t s) {
    if (n) B2_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * (size_t)n, s));
  }
};

// streaming loads for the matrix stream (values, columns): read-only path, do not allocate in L1, so that
// the lines of the gathered vector are what stays resident there
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- warp / block reductions ------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid-wide sum of N per-thread values.  Every block publishes its partial sums;
// the block that takes the last ticket re-reads all partials in a fixed order and returns true
// (in every thread of that block) with `total` valid in thread 0.  `counter` is reset for the
// next launch.  Requires blockDim.x to be a multiple of 32 and <= 1024.
template <int N>
__device__ __forceinline__ bool grid_reduce(double (&v)[N], double* __restrict__ partials,
                                            unsigned* __restrict__ counter, double (&total)[N]) {
  __shared__ double sm[N][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = warp_sum(v[i]);
    if (lane == 0) sm[i][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = lane < nwarp ? sm[i][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) partials[(size_t)blockIdx.x * N + i] = s;
    }
    if (lane == 0) {
      __threadfence();
      unsigned t = atomicAdd(counter, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] += __ldcg(&partials[(size_t)b * N + i]);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = warp_sum(acc[i]);
    if (lane == 0) sm[i][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = lane < nwarp ? sm[i][lane] : 0.0;
      s = warp_sum(s);
      total[i] = s;
    }
    if (lane == 0) *counter = 0u;
  }
  return true;
}
