// cuda_shim.h -- TEST INFRASTRUCTURE ONLY (tests/hostsim).
//
// Just enough of the CUDA device vocabulary for g++ to compile zfit_b200/csrc/znll_device.cuh as host code, so that the
// per-event arithmetic of the kernels (every node's forward and reverse sweep, the fp64 exp / log / reciprocal, the
// Kahan accumulation) can be executed on a CPU and compared with the oracle in the `-m "not gpu"` suite.  Nothing in the
// product imports, links or executes this: the product path has no CPU implementation (zfit_b200/_znll.py raises
// without the CUDA library and a device).  The launch machinery (grid, shuffles, tickets, peer exchange) is only
// declared so that the templates parse; it is never instantiated here and stays covered by the `-m gpu` tests.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__

struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct HostsimDim3 {
  unsigned x, y, z;
};
static HostsimDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1}, blockDim = {1, 1, 1};

static inline int __double2hiint(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return (int)(uint32_t)(u >> 32);
}
static inline int __double2loint(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return (int)(uint32_t)u;
}
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x;
  std::memcpy(&x, &u, 8);
  return x;
}
static inline double __longlong_as_double(long long v) {
  double x;
  std::memcpy(&x, &v, 8);
  return x;
}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
template <class T>
static inline T __ldcg(const T* p) {
  return *p;
}
// MUFU.RCP64H: reciprocal seed good to about 20 bits with the low mantissa word cleared
static inline double hostsim_rcp_seed(double x) { return __hiloint2double(__double2hiint(1.0 / x), 0); }

// declared for parsing only (never called on the host)
double __shfl_down_sync(unsigned, double, int);
void __syncthreads();
void __threadfence();
void __threadfence_system();
unsigned atomicAdd(unsigned*, unsigned);
long long clock64();
size_t __cvta_generic_to_shared(const void*);
