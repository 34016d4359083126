// peer.cu -- one-sided exchange of per-item statistics between the GPUs of one node over NVLink peer memory.
//
// The hot path shards by batch item and has no data-path collective; the only exchange is the
// per-item loudness vector ([B/W] floats per rank) for whole-batch statistics -- logging data.  A logging-only
// statistic must never be able to stall the data path, so nothing here makes one rank wait for another:
//
//   put      every rank STORES its vector straight into a small buffer of every peer (cudaIpc mapping, NVLink /
//            NVSwitch P2P stores) and publishes the sequence number with a system-scope release.  Never waits.
//   latest   reads, from the rank's OWN memory, the newest COMPLETE vector of every rank together with the sequence
//            number it carries (seqlock: flag, data, flag again).  Never waits: a rank that is behind simply shows
//            an older sequence number.
//   collect  the lock-step form (bounded spin until every rank has published exactly `seq`): used by tests, by the
//            bench's untimed validation against an NCCL all_gather, and by callers that want exact-step statistics;
//            it is issued on a side stream, never between two kernels of the data path.
//
// Slots rotate over NSLOT = 4 sequence numbers.  A writer invalidates a slot (flag = -seq) before it rewrites the
// data and publishes (flag = +seq) afterwards, so a reader can never accept a torn or half-overwritten vector
// whatever the skew between ranks (round 1's two-slot rotation silently returned a future vector to a slow rank).
//
// Buffer layout (floats / int32, per rank):  data[NSLOT][world][n_max] | flag[NSLOT][world] | status[4]
#include "b2a_common.h"

namespace b2a {
namespace peer {

constexpr int MAX_WORLD = 16;
constexpr int NSLOT = 4;

struct Peers {
  float* buf[MAX_WORLD];
};

__host__ __device__ inline size_t flag_offset_floats(int world, int n_max) { return (size_t)NSLOT * world * n_max; }

__device__ __forceinline__ void st_release_sys(int* p, int v) {
#ifdef B2A_SIM
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
#ifdef B2A_SIM
  cusim::yield();  // polled in spin loops
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}

// CTA p stores src[0..n) into slot (seq % NSLOT, rank) of peer p's buffer: invalidate, data, publish.
__global__ void __launch_bounds__(256)
peer_put_kernel(const float* __restrict__ src, int n, const B2A_GRID_CONSTANT Peers peers, int world, int rank,
                int n_max, int seq) {
  float* base = peers.buf[blockIdx.x];
  const int slot = seq % NSLOT;
  float* dst = base + ((size_t)slot * world + rank) * n_max;
  int* flag = reinterpret_cast<int*>(base + flag_offset_floats(world, n_max)) + slot * world + rank;
  if (threadIdx.x == 0) {
    st_release_sys(flag, -seq);  // readers reject the slot while it is being rewritten
    __threadfence_system();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(flag, seq);
}

// Seqlock read of rank r's slot `slot` in the local buffer into out[0..n): returns the sequence number the copy
// carries, or 0 when the slot is empty / being rewritten.  Called by one warp.
__device__ __forceinline__ int read_slot(const float* local, int world, int n, int n_max, int r, int slot, int want,
                                         float* out) {
  const int lane = threadIdx.x & 31;
  const int* flag = reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max)) + slot * world + r;
  int f1 = 0;
  if (lane == 0) f1 = ld_acquire_sys(flag);
  f1 = __shfl_sync(0xffffffffu, f1, 0);
  if (f1 <= 0 || (want > 0 && f1 != want)) return 0;
  const float* data = local + ((size_t)slot * world + r) * n_max;
  for (int k = lane; k < n; k += 32) out[k] = data[k];
  __threadfence_system();
  int f2 = 0;
  if (lane == 0) f2 = ld_acquire_sys(flag);
  f2 = __shfl_sync(0xffffffffu, f2, 0);
  return f2 == f1 ? f1 : 0;
}

// out[r][0..n) = the newest complete vector rank r has published here, seqs[r] = its sequence number (0: nothing
// yet -> NaNs).  One warp per rank (round-robin); never waits.
__global__ void __launch_bounds__(256)
peer_latest_kernel(const float* __restrict__ local, int world, int n, int n_max, float* __restrict__ out,
                   int* __restrict__ seqs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int r = warp; r < world; r += nwarp) {
    const int* flags = reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max));
    int got = 0;
    for (int attempt = 0; attempt < NSLOT && got == 0; ++attempt) {
      int best = 0, best_slot = 0;  // newest published slot right now
      for (int s = 0; s < NSLOT; ++s) {
        int f = 0;
        if (lane == 0) f = ld_acquire_sys(flags + s * world + r);
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f > best) { best = f; best_slot = s; }
      }
      if (best == 0) break;
      got = read_slot(local, world, n, n_max, r, best_slot, 0, out + (size_t)r * n);
    }
    if (got == 0)
      for (int k = lane; k < n; k += 32) out[(size_t)r * n + k] = __int_as_float(0x7fc00000);
    if (lane == 0) seqs[r] = got;
  }
}

// Lock-step gather: wait (bounded) until every rank has published exactly `seq` here, then out[world][n].  A rank
// that ran more than NSLOT - 1 steps ahead has overwritten the slot: its row is NaN and seqs[r] holds what was seen.
__global__ void __launch_bounds__(256)
peer_collect_kernel(const float* __restrict__ local, int world, int n, int n_max, int seq, float* __restrict__ out,
                    int* __restrict__ seqs, long long max_spins) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int slot = seq % NSLOT;
  int* status = const_cast<int*>(reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max))) + NSLOT * world;
  for (int r = warp; r < world; r += nwarp) {
    const int* flag = reinterpret_cast<const int*>(local + flag_offset_floats(world, n_max)) + slot * world + r;
    int got = 0, seen = 0;
    for (long long spins = 0; spins < max_spins; ++spins) {
      if (lane == 0) seen = ld_acquire_sys(flag);
      seen = __shfl_sync(0xffffffffu, seen, 0);
      const int a = seen < 0 ? -seen : seen;
      if (a > seq) break;  // overwritten by a later sequence number: gone
      if (seen == seq) {
        got = read_slot(local, world, n, n_max, r, slot, seq, out + (size_t)r * n);
        if (got == seq) break;
        got = 0;
      }
    }
    if (got != seq) {
      for (int k = lane; k < n; k += 32) out[(size_t)r * n + k] = __int_as_float(0x7fc00000);
      if (lane == 0) status[0] = seq;  // a peer died or lapped the slot: never hang the GPU
    }
    if (lane == 0 && seqs) seqs[r] = got == seq ? seq : -(seen < 0 ? -seen : seen);
  }
}

}  // namespace peer
}  // namespace b2a

using namespace b2a::peer;

extern "C" size_t b2a_peer_buffer_bytes(int world, int n_max) {
  if (world < 1 || world > MAX_WORLD || n_max < 1) return 0;
  return (flag_offset_floats(world, n_max) + (size_t)NSLOT * world + 4) * 4;
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
                                    int32_t* seqs_out, void* stream) {
  B2A_REQUIRE(local_buf && out, B2A_E_INVALID, "peer_collect: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && n >= 1 && n <= n_max && seq >= 1, B2A_E_INVALID,
              "peer_collect: bad argument");
  const long long max_spins = 20000000LL;  // a few seconds of polling local memory, then give up (NaN + status)
  B2A_LAUNCH(peer_collect_kernel, dim3(1), dim3(256), 0, stream, (const float*)local_buf, world, n, n_max, seq, out,
             seqs_out, max_spins);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_peer_latest_f32(const void* local_buf, int world, int n, int n_max, float* out, int32_t* seqs_out,
                                   void* stream) {
  B2A_REQUIRE(local_buf && out && seqs_out, B2A_E_INVALID, "peer_latest: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && n >= 1 && n <= n_max, B2A_E_INVALID, "peer_latest: bad argument");
  B2A_LAUNCH(peer_latest_kernel, dim3(1), dim3(256), 0, stream, (const float*)local_buf, world, n, n_max, out, seqs_out);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_peer_status(const void* local_buf, int world, int n_max, int32_t* status_out /*device [1]*/,
                               void* stream) {
  B2A_REQUIRE(local_buf && status_out, B2A_E_INVALID, "peer_status: null pointer");
  B2A_REQUIRE(world >= 1 && world <= MAX_WORLD && n_max >= 1, B2A_E_INVALID, "peer_status: bad argument");
  const int* st = reinterpret_cast<const int*>((const float*)local_buf + flag_offset_floats(world, n_max)) + NSLOT * world;
#ifdef B2A_SIM
  *status_out = *st;
#else
  B2A_CUDA_OK(cudaMemcpyAsync(status_out, st, sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
#endif
  return B2A_OK;
}
