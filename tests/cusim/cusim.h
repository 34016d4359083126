// cusim.h -- a tiny CUDA-on-CPU execution shim.  TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU.  To check kernel indexing and arithmetic
// before spending GPU minutes, tests/cusim/build_sim.py compiles the product's .cu
// sources with g++ (-x c++ -DB2A_SIM -include cusim.h) into tests/cusim/_build/libb2a_sim.so.
// Every CUDA thread of a block runs as a real host thread; __syncthreads / named
// barriers are std::barrier; warp shuffles go through a per-warp exchange slot;
// blocks of a launch run one after another.  Device pointers are host pointers.
//
// Nothing under audiotools_b200/ loads this library: it exists so that the *same kernel
// source* that ships can be executed here and compared with the oracle (tests/test_sim_*.py).
// Inline PTX paths are compiled out under B2A_SIM (plain C++ equivalents are used).
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(d, s, n);
  return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) {
  *n = 2;
  return 0;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return 0; }

namespace cusim {

struct BlockCtx {
  unsigned nthreads = 0;
  std::unique_ptr<std::barrier<>> all;
  std::mutex mu;
  std::map<std::pair<int, int>, std::unique_ptr<std::barrier<>>> named;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<std::array<unsigned long long, 32>> warp_slot;
  unsigned char* dyn_smem = nullptr;
};
inline BlockCtx*& ctx() { static BlockCtx* c = nullptr; return c; }
inline std::mutex& atomic_mu() { static std::mutex m; return m; }

struct TL { uint3 tid, bid; dim3 bdim, gdim; };
inline TL& tl() { static thread_local TL t; return t; }

inline void named_bar(int id, int n) {
  BlockCtx* c = ctx();
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> g(c->mu);
    auto& slot = c->named[{id, n}];
    if (!slot) slot.reset(new std::barrier<>(n));
    b = slot.get();
  }
  b->arrive_and_wait();
}

template <class Body>
void launch(dim3 grid, dim3 block, size_t smem_bytes, Body body) {
  unsigned nt = block.x * block.y * block.z;
  BlockCtx c;
  c.nthreads = nt;
  c.all.reset(new std::barrier<>(nt));
  unsigned nwarps = (nt + 31) / 32;
  for (unsigned w = 0; w < nwarps; ++w) {
    unsigned lanes = std::min(32u, nt - w * 32);
    c.warp_bar.emplace_back(new std::barrier<>(lanes));
  }
  c.warp_slot.resize(nwarps);
  std::vector<unsigned char> smem(smem_bytes + 1024);
  c.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 1023) & ~(uintptr_t)1023);
  ctx() = &c;
  std::vector<std::thread> th;
  th.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t] {
      TL& L = tl();
      L.bdim = block;
      L.gdim = grid;
      L.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            L.bid = uint3{bx, by, bz};
            body();
            c.all->arrive_and_wait();  // static __shared__ storage is reused by the next block
          }
    });
  }
  for (auto& x : th) x.join();
  ctx() = nullptr;
}

}  // namespace cusim

#define threadIdx (cusim::tl().tid)
#define blockIdx (cusim::tl().bid)
#define blockDim (cusim::tl().bdim)
#define gridDim (cusim::tl().gdim)
#define warpSize 32

static inline void __syncthreads() { cusim::ctx()->all->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  unsigned t = threadIdx.x + threadIdx.y * blockDim.x;
  cusim::ctx()->warp_bar[t / 32]->arrive_and_wait();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T> static inline T cusim_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl of <= 8 bytes");
  unsigned t = threadIdx.x + threadIdx.y * blockDim.x;
  unsigned w = t / 32, lane = t % 32;
  auto* c = cusim::ctx();
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  c->warp_slot[w][lane] = raw;
  c->warp_bar[w]->arrive_and_wait();
  unsigned long long r = c->warp_slot[w][(unsigned)src_lane & 31u];
  c->warp_bar[w]->arrive_and_wait();
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
static inline unsigned cusim_lane() { return (threadIdx.x + threadIdx.y * blockDim.x) % 32; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cusim_shfl(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return cusim_shfl(v, (int)(cusim_lane() ^ (unsigned)m));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int l = (int)cusim_lane();
  T r = cusim_shfl(v, l - (int)d < 0 ? l : l - (int)d);
  return r;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int l = (int)cusim_lane();
  T r = cusim_shfl(v, l + (int)d > 31 ? l : l + (int)d);
  return r;
}

static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned t = threadIdx.x + threadIdx.y * blockDim.x;
  unsigned w = t / 32, lane = t % 32;
  auto* c = cusim::ctx();
  c->warp_slot[w][lane] = pred ? 1ull : 0ull;
  c->warp_bar[w]->arrive_and_wait();
  unsigned nl = std::min(32u, c->nthreads - w * 32), m = 0;
  for (unsigned l = 0; l < nl; ++l) m |= (unsigned)c->warp_slot[w][l] << l;
  c->warp_bar[w]->arrive_and_wait();
  return m;
}
static inline int __all_sync(unsigned mask, int pred) {
  unsigned t = threadIdx.x + threadIdx.y * blockDim.x;
  unsigned nl = std::min(32u, cusim::ctx()->nthreads - (t / 32) * 32);
  unsigned full = nl == 32 ? 0xffffffffu : ((1u << nl) - 1u);
  return __ballot_sync(mask, pred) == full;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }

template <class T> static inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(cusim::atomic_mu());
  T old = *p;
  *p = old + v;
  return old;
}
template <class T> static inline T atomicExch(T* p, T v) {
  std::lock_guard<std::mutex> g(cusim::atomic_mu());
  T old = *p;
  *p = v;
  return old;
}
template <class T> static inline T atomicMax(T* p, T v) {
  std::lock_guard<std::mutex> g(cusim::atomic_mu());
  T old = *p;
  if (v > old) *p = v;
  return old;
}
template <class T> static inline T atomicCAS(T* p, T cmp, T v) {
  std::lock_guard<std::mutex> g(cusim::atomic_mu());
  T old = *p;
  if (old == cmp) *p = v;
  return old;
}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline void sincospif(float x, float* s, float* c) {
  *s = (float)sin(M_PI * (double)x);
  *c = (float)cos(M_PI * (double)x);
}
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned __float_as_uint(float v) { unsigned u; memcpy(&u, &v, 4); return u; }
static inline float __uint_as_float(unsigned u) { float v; memcpy(&v, &u, 4); return v; }
static inline float cospif(float x) { return (float)cos(M_PI * (double)x); }
static inline float sinpif(float x) { return (float)sin(M_PI * (double)x); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
using std::max;
using std::min;
