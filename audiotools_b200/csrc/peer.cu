// peer.cu -- one-sided exchange of per-item statistics between the GPUs of one node over NVLink peer memory.
//
// The hot path shards by batch item and has no data-path collective; the only exchange is the
// per-item loudness vector ([B/W] floats per rank) for whole-batch statistics.  An NCCL all-gather for those
// 256 bytes costs a rendezvous kernel that competes with the persistent spectral kernel for SM slots (measured:
// +50 us per 630 us step at 2 GPUs).  Here every rank instead STORES its vector straight into a small buffer
// of every peer (cudaIpc mapping, NVLink/NVSwitch P2P stores) and publishes a sequence number with a
// system-scope release; a reader waits on the flags in its OWN memory.  No rendezvous, no NCCL kernel, one
// tiny launch per step; slots are double buffered by sequence parity (a rank can be at most one step ahead of
// the slowest reader, see parallel.py).
//
// Buffer layout (floats / int32, per rank):  data[2][world][n_max] | flag[2][world] | status[1]
#include "b2a_common.h"

namespace b2a {
namespace peer {

constexpr int MAX_WORLD = 16;

struct Peers {
  float* buf[MAX_WORLD];
};

__host__ __device__ inline size_t flag_offset_floats(int world, int n_max) { return (size_t)2 * world * n_max; }

__device__ __forceinline__ void st_release_sys(int* p, int v) {
#ifdef B2A_SIM
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
#ifdef B2A_SIM
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}

// CTA p stores src[0..n) into slot (seq & 1, rank) of peer p's buffer, then publishes seq there.
__global__ void __launch_bounds__(256)
peer_put_kernel(const float* __restrict__ src, int n, const B2A_GRID_CONSTANT Peers peers, int world, int rank,
                int n_max, int seq) {
  float* base = peers.buf[blockIdx.x];
  float* dst = base + ((size_t)(seq & 1) * world + rank) * n_max;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    int* flags = reinterpret_cast<int*>(base + flag_offset_floats(world, n_max));
    st_release_sys(flags + (seq & 1) * world + rank, seq);
  }
}

// wait until every rank has published `seq` into THIS rank's buffer, then gather [world][n] -> out
__global__ void __launch_bounds__(256)
peer_collect_kernel(const float* __restrict__ local, int world, int n, int n_max, int seq, float* __restrict__ out,
                    long long max_spins) {
  const int* flags = reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max)) + (seq & 1) * world;
  int* status = const_cast<int*>(reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max))) + 2 * world;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    long long spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) - seq < 0) {  // sequence numbers only grow
      if (++spins > max_spins) { s_ok = 0; break; }           // a peer died: never hang the GPU
    }
  }
  __syncthreads();
  const float* data = local + (size_t)(seq & 1) * world * n_max;
  for (int i = threadIdx.x; i < world * n; i += blockDim.x) {
    const int r = i / n, k = i - r * n;
    out[i] = s_ok ? data[(size_t)r * n_max + k] : __int_as_float(0x7fc00000);
  }
  if (threadIdx.x == 0 && !s_ok) *status = seq;
}

// One launch per step: CTAs 0..world-1 put sequence `seq_put`, CTA `world` collects sequence `seq_col` (the
// previous step's, already published by every rank long ago) -- the steady state of a pipelined consumer.
__global__ void __launch_bounds__(256)
peer_exchange_kernel(const float* __restrict__ src, int n, const B2A_GRID_CONSTANT Peers peers, int world, int rank,
                     int n_max, int seq_put, const float* __restrict__ local, int n_col, int seq_col,
                     float* __restrict__ out, long long max_spins) {
  if ((int)blockIdx.x < world) {
    float* base = peers.buf[blockIdx.x];
    float* dst = base + ((size_t)(seq_put & 1) * world + rank) * n_max;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      int* flags = reinterpret_cast<int*>(base + flag_offset_floats(world, n_max));
      st_release_sys(flags + (seq_put & 1) * world + rank, seq_put);
    }
    return;
  }
  const int* flags = reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max)) + (seq_col & 1) * world;
  int* status = const_cast<int*>(reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max))) + 2 * world;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    long long spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) - seq_col < 0) {
      if (++spins > max_spins) { s_ok = 0; break; }
    }
  }
  __syncthreads();
  const float* data = local + (size_t)(seq_col & 1) * world * n_max;
  for (int i = threadIdx.x; i < world * n_col; i += blockDim.x) {
    const int r = i / n_col, k = i - r * n_col;
    out[i] = s_ok ? data[(size_t)r * n_max + k] : __int_as_float(0x7fc00000);
  }
  if (threadIdx.x == 0 && !s_ok) *status = seq_col;
}

}  // namespace peer
}  // namespace b2a

using namespace b2a::peer;

extern "C" size_t b2a_peer_buffer_bytes(int world, int n_max) {
  if (world < 1 || world > MAX_WORLD || n_max < 1) return 0;
  return (flag_offset_floats(world, n_max) + 2 * (size_t)world + 4) * 4;
}

extern "C" int b2a_peer_buffer_create(int world, int n_max, void** dev_ptr, unsigned char* handle_out /*[64]*/) {
  B2A_REQUIRE(dev_ptr && handle_out, B2A_E_INVALID, "peer: null pointer");
  const size_t bytes = b2a_peer_buffer_bytes(world, n_max);
  B2A_REQUIRE(bytes > 0, B2A_E_INVALID, "peer: world=%d (max %d) n_max=%d", world, MAX_WORLD, n_max);
#ifdef B2A_SIM
  *dev_ptr = calloc(1, bytes);
  memset(handle_out, 0, 64);
  memcpy(handle_out, dev_ptr, sizeof(void*));
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  B2A_CUDA_OK(cudaMalloc(&p, bytes));  // its own allocation: an IPC handle names a whole cudaMalloc block
  B2A_CUDA_OK(cudaMemset(p, 0, bytes));
  B2A_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  B2A_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
#endif
  return B2A_OK;
}

extern "C" int b2a_peer_buffer_open(const unsigned char* handle /*[64]*/, void** peer_ptr) {
  B2A_REQUIRE(handle && peer_ptr, B2A_E_INVALID, "peer: null pointer");
#ifdef B2A_SIM
  memcpy(peer_ptr, handle, sizeof(void*));
#else
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  B2A_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
#endif
  return B2A_OK;
}

extern "C" int b2a_peer_buffer_close(void* peer_ptr) {
#ifndef B2A_SIM
  if (peer_ptr) B2A_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
#endif
  return B2A_OK;
}

extern "C" int b2a_peer_buffer_destroy(void* dev_ptr) {
#ifdef B2A_SIM
  free(dev_ptr);
#else
  if (dev_ptr) B2A_CUDA_OK(cudaFree(dev_ptr));
#endif
  return B2A_OK;
}

extern "C" int b2a_peer_put_f32(const float* src, int n, void* const* peer_bufs_h, int world, int rank, int n_max,
                                int seq, void* stream) {
  B2A_REQUIRE(src && peer_bufs_h, B2A_E_INVALID, "peer_put: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world && n >= 1 && n <= n_max && seq >= 1,
              B2A_E_INVALID, "peer_put: bad argument");
  Peers peers;
  memset(&peers, 0, sizeof(peers));
  for (int i = 0; i < world; ++i) {
    B2A_REQUIRE(peer_bufs_h[i], B2A_E_INVALID, "peer_put: buffer of rank %d is not mapped", i);
    peers.buf[i] = (float*)peer_bufs_h[i];
  }
  B2A_LAUNCH(peer_put_kernel, dim3((unsigned)world), dim3(256), 0, stream, src, n, peers, world, rank, n_max, seq);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_peer_collect_f32(const void* local_buf, int world, int n, int n_max, int seq, float* out,
                                    void* stream) {
  B2A_REQUIRE(local_buf && out, B2A_E_INVALID, "peer_collect: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && n >= 1 && n <= n_max && seq >= 1, B2A_E_INVALID,
              "peer_collect: bad argument");
  const long long max_spins = 20000000LL;  // a few seconds of polling local memory, then give up (NaN + status)
  B2A_LAUNCH(peer_collect_kernel, dim3(1), dim3(256), 0, stream, (const float*)local_buf, world, n, n_max, seq, out,
             max_spins);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_peer_exchange_f32(const float* src, int n, void* const* peer_bufs_h, int world, int rank, int n_max,
                                     int seq_put, const void* local_buf, int n_collect, int seq_collect, float* out,
                                     void* stream) {
  B2A_REQUIRE(src && peer_bufs_h && local_buf && out, B2A_E_INVALID, "peer_exchange: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world && n >= 1 && n <= n_max &&
                  n_collect >= 1 && n_collect <= n_max && seq_put >= 2 && seq_collect == seq_put - 1,
              B2A_E_INVALID, "peer_exchange: bad argument (collects seq_put - 1)");
  Peers peers;
  memset(&peers, 0, sizeof(peers));
  for (int i = 0; i < world; ++i) {
    B2A_REQUIRE(peer_bufs_h[i], B2A_E_INVALID, "peer_exchange: buffer of rank %d is not mapped", i);
    peers.buf[i] = (float*)peer_bufs_h[i];
  }
  B2A_LAUNCH(peer_exchange_kernel, dim3((unsigned)world + 1), dim3(256), 0, stream, src, n, peers, world, rank, n_max,
             seq_put, (const float*)local_buf, n_collect, seq_collect, out, 20000000LL);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
