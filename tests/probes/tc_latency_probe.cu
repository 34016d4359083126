// tc_latency_probe.cu -- how long does a chain of tcgen05.mma take when consecutive instructions accumulate into the
// SAME tensor-memory tile (dependent) versus round-robin over n_acc independent tiles?  M = 128, N = 16, K = 16 per
// instruction, fp16, A from TMEM or shared memory.  Prints cycles per MMA (clock64 from first issue to the commit's
// mbarrier arrival).  Hardware probe, not product code.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int N>
__global__ void __launch_bounds__(128) lat_kernel(long long* out, int n_mma, int n_acc, int a_from_tmem) {
  __shared__ __align__(128) __half sA[128 * 16];
  __shared__ __align__(128) __half sB[N * 16 * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 16; i += 128) sA[i] = __float2half(0.01f);
  for (int i = tid; i < N * 16 * 16; i += 128) sB[i] = __float2half(0.01f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    if (pred) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t da = make_desc((uint32_t)__cvta_generic_to_shared(sA), 128, 256);
      const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(sB);
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const int acc = i % n_acc;
        const uint64_t db = make_desc(b0 + (uint32_t)(i & 15) * (N * 32), 128, 256);
        const uint32_t d = tm + (uint32_t)(acc * N);
        if (a_from_tmem)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(tm + 448), "l"(db), "r"(idesc), "r"((uint32_t)(i >= n_acc)) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(i >= n_acc)) : "memory");
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
      const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
      asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
      t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
int main() {
  long long* d; CK(cudaMalloc(&d, 16));
  printf("{\"rows\": [");
  bool first = true;
  for (int N : {16, 32, 64}) for (int tmemA = 0; tmemA < 2; ++tmemA) for (int n_acc : {1, 2, 4, 8, 16}) {
    if (n_acc * N > 384) continue;
    const int n_mma = 192;
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      if (N == 16) lat_kernel<16><<<1, 128>>>(d, n_mma, n_acc, tmemA);
      else if (N == 32) lat_kernel<32><<<1, 128>>>(d, n_mma, n_acc, tmemA);
      else lat_kernel<64><<<1, 128>>>(d, n_mma, n_acc, tmemA);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("%s{\"N\": %d, \"a_tmem\": %d, \"n_acc\": %d, \"issue_cyc_per_mma\": %.1f, \"total_cyc_per_mma\": %.1f}", first ? "" : ", ", N, tmemA, n_acc, (double)h[0] / n_mma, (double)h[1] / n_mma);
    first = false;
  }
  printf("]}\n");
  return 0;
}
