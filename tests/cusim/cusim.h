// cusim.h -- a tiny CUDA-on-CPU execution shim.  TEST INFRASTRUCTURE ONLY.
//
// The build container has nvcc but no GPU.  To check kernel indexing and arithmetic
// before spending GPU minutes, tests/cusim/build_sim.py compiles the product's .cu
// sources with g++ (-x c++ -DB2A_SIM -include cusim.h) into tests/cusim/_build/libb2a_sim.so.
// Every CUDA thread of a block runs as a FIBER (its own stack, cooperative scheduling on one
// worker thread per block in flight); __syncthreads / named barriers / __syncwarp are
// cooperative barriers; warp shuffles go through a per-warp exchange slot; the blocks of a
// launch are handed out in order to up to 8 worker threads.  Device pointers are host pointers.
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
#include <array>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static thread_local
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

// ---------------------------------------------------------------------------------------------
// Execution model.  The CUDA threads of a block are FIBERS (user-level contexts with their own stacks) that one OS
// worker thread schedules round-robin; __syncthreads / named barriers / __syncwarp / shuffles are cooperative
// barriers (a waiting fiber yields).  The blocks of a launch are handed out IN ORDER to up to 8 worker threads, so a
// block that spins on a flag of an earlier block (decoupled look-back) always finds its predecessor running.
// `__shared__` is `static thread_local`: one copy per worker, i.e. per block in flight.
// (The first version ran every CUDA thread as an OS thread with std::barrier: 256 threads on 8 cores spent four
//  fifths of the suite's time in futex calls.)
// ---------------------------------------------------------------------------------------------
struct TL { uint3 tid, bid; dim3 bdim, gdim; };
struct Bar { unsigned count = 0, gen = 0; };
struct Fiber { void* sp = nullptr; TL tl; bool done = false; };

struct Worker {
  unsigned nthreads = 0, cur = 0;
  std::vector<Fiber> fib;
  char* stacks = nullptr;
  size_t stack_bytes = 0;
  void* sched_sp = nullptr;
  Bar all;
  std::vector<Bar> warp_bar;
  std::map<std::pair<int, int>, Bar> named;
  std::vector<std::array<unsigned long long, 32>> warp_slot;
  unsigned char* dyn_smem = nullptr;
  void (*invoke)(void*) = nullptr;
  void* body = nullptr;
};
inline Worker*& ctx() { static thread_local Worker* w = nullptr; return w; }
inline std::mutex& atomic_mu() { static std::mutex m; return m; }
inline TL& tl() { Worker* w = ctx(); return w->fib[w->cur].tl; }

#if !defined(__x86_64__)
#error "tests/cusim: the fiber switch is written for x86-64 (System V ABI)"
#endif
extern "C" void cusim_switch(void** save_sp, void* load_sp);
// x86-64 System V: callee-saved registers on the old stack, swap stack pointers, restore, return into the new context
__asm__(
    ".text\n"
    ".weak cusim_switch\n"
    ".type cusim_switch,@function\n"
    "cusim_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size cusim_switch,.-cusim_switch\n");

// back to the worker's scheduler loop; resumed later at this point
inline void yield() {
  Worker* w = ctx();
  cusim_switch(&w->fib[w->cur].sp, w->sched_sp);
}
inline void bar_wait(Bar& b, unsigned n) {
  const unsigned gen = b.gen;
  if (++b.count >= n) { b.count = 0; ++b.gen; return; }
  while (b.gen == gen) yield();
}
inline void named_bar(int id, int n) { bar_wait(ctx()->named[{id, n}], (unsigned)n); }

inline void fiber_main() {
  Worker* w = ctx();
  w->invoke(w->body);
  w = ctx();
  w->fib[w->cur].done = true;
  cusim_switch(&w->fib[w->cur].sp, w->sched_sp);
  abort();  // a finished fiber is never resumed
}

inline void run_block(Worker* w, uint3 bid, dim3 block, dim3 grid) {
  const unsigned nt = w->nthreads;
  w->all = Bar();
  for (auto& b : w->warp_bar) b = Bar();
  w->named.clear();
  for (unsigned t = 0; t < nt; ++t) {
    Fiber& f = w->fib[t];
    f.done = false;
    f.tl.bdim = block; f.tl.gdim = grid; f.tl.bid = bid;
    f.tl.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    uintptr_t top = ((uintptr_t)(w->stacks + (size_t)(t + 1) * w->stack_bytes)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of fiber_main (keeps the ABI's entry alignment)
    *--sp = (void*)&fiber_main;      // `ret` of the first switch jumps here
    for (int r = 0; r < 6; ++r) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = sp;
  }
  // CUSIM_SHUFFLE=<seed>: visit the fibers in a fresh random order on every scheduling round instead of 0, 1, 2, ...
  // A missing barrier that the fixed order happens to satisfy (the producer always runs first) then shows up as a
  // wrong result: the poor man's racecheck (the suite is run under a few seeds before every GPU session).
  static const unsigned long long shuffle_seed = [] {
    const char* e = getenv("CUSIM_SHUFFLE");
    return e ? strtoull(e, nullptr, 10) * 2654435761ull + 88172645463325252ull : 0ull;
  }();
  std::vector<unsigned> order(nt);
  for (unsigned t = 0; t < nt; ++t) order[t] = t;
  unsigned long long rng = shuffle_seed ^ ((unsigned long long)bid.x * 0x9E3779B97F4A7C15ull + bid.y * 7919ull + bid.z);
  unsigned alive = nt;
  while (alive) {
    if (shuffle_seed) {
      for (unsigned t = nt - 1; t > 0; --t) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[t], order[(unsigned)(rng % (t + 1))]);
      }
    }
    for (unsigned k = 0; k < nt; ++k) {
      const unsigned t = order[k];
      if (w->fib[t].done) continue;
      w->cur = t;
      cusim_switch(&w->sched_sp, w->fib[t].sp);
      if (w->fib[t].done) --alive;
    }
  }
}

template <class Body>
void launch(dim3 grid, dim3 block, size_t smem_bytes, Body body) {
  const unsigned nt = block.x * block.y * block.z;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  if (nt == 0 || nblocks == 0) return;
  unsigned nw = std::thread::hardware_concurrency();
  if (nw == 0) nw = 4;
  if (nw > 8) nw = 8;
  if (nw > nblocks) nw = (unsigned)nblocks;
  std::atomic<unsigned long long> next{0};
  auto invoke = +[](void* b) { (*static_cast<Body*>(b))(); };
  auto work = [&]() {
    Worker w;
    w.nthreads = nt;
    w.fib.resize(nt);
    w.warp_bar.resize((nt + 31) / 32);
    w.warp_slot.resize((nt + 31) / 32);
    w.stack_bytes = 256 * 1024;
    const size_t total = (size_t)nt * w.stack_bytes + 64;
    w.stacks = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (w.stacks == (char*)MAP_FAILED) { fprintf(stderr, "cusim: mmap of fiber stacks failed\n"); abort(); }
    std::vector<unsigned char> smem(smem_bytes + 1024);
    w.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 1023) & ~(uintptr_t)1023);
    w.invoke = invoke;
    w.body = &body;
    Worker* prev = ctx();
    ctx() = &w;
    for (;;) {
      const unsigned long long b = next.fetch_add(1);
      if (b >= nblocks) break;
      const unsigned bx = (unsigned)(b % grid.x), by = (unsigned)((b / grid.x) % grid.y);
      const unsigned bz = (unsigned)(b / ((unsigned long long)grid.x * grid.y));
      run_block(&w, uint3{bx, by, bz}, block, grid);
    }
    ctx() = prev;
    munmap(w.stacks, total);
  };
  if (nw <= 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nw; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
}

}  // namespace cusim

#define threadIdx (cusim::tl().tid)
#define blockIdx (cusim::tl().bid)
#define blockDim (cusim::tl().bdim)
#define gridDim (cusim::tl().gdim)
#define warpSize 32

static inline void __syncthreads() { cusim::bar_wait(cusim::ctx()->all, cusim::ctx()->nthreads); }
static inline unsigned cusim_warp_lanes(unsigned w) { return std::min(32u, cusim::ctx()->nthreads - w * 32); }
static inline void cusim_warp_sync(unsigned w) { cusim::bar_wait(cusim::ctx()->warp_bar[w], cusim_warp_lanes(w)); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  unsigned t = threadIdx.x + threadIdx.y * blockDim.x;
  cusim_warp_sync(t / 32);
}
static inline void __nanosleep(unsigned) { cusim::yield(); }
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
  cusim_warp_sync(w);
  unsigned long long r = c->warp_slot[w][(unsigned)src_lane & 31u];
  cusim_warp_sync(w);
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
  cusim_warp_sync(w);
  unsigned nl = std::min(32u, c->nthreads - w * 32), m = 0;
  for (unsigned l = 0; l < nl; ++l) m |= (unsigned)c->warp_slot[w][l] << l;
  cusim_warp_sync(w);
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
