// tc_probe.cu -- hardware probe for the tcgen05 building blocks of the tensor-core spectral kernel (csrc/spectral_tc.cu).
// Not product code: a standalone binary that answers, on a real B200, the questions the kernel design depends on:
//   T0  D[128xN] = A[128xK] B[NxK]^T, fp16 in / fp32 accumulate, A and B from shared memory (no-swizzle, K-major
//       canonical layout), accumulator read back with tcgen05.ld.32x32b  -> checks descriptor + idesc encodings
//   T1  same product with A taken from tensor memory (written with tcgen05.st)          -> TMEM A-operand layout
//   T2  B rows that OVERLAP in shared memory (row n starts 16 bytes after row n-1: LBO = 16 B) -> Toeplitz operand
//   T3  fp16 subnormal inputs: are they honoured or flushed by the tensor core?
//   T4  3-product split (x = h1 + h2, F = F1 + F2 in fp16, operands pre-scaled): achieved accuracy vs fp64
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; prints one JSON line.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 16, K = 32;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                     // D format f32
  d |= 0u << 7;                     // A format f16
  d |= 0u << 10;                    // B format f16
  d |= (uint32_t)(n >> 3) << 17;    // N
  d |= (uint32_t)(m >> 4) << 24;    // M
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(da), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b),
               "r"(parity) : "memory");
}

// mode 0: A smem; 1: A tmem; 2: B overlapping rows (B[n][k] = bseq[8 n + k])
__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ Ag, const __half* __restrict__ Bg,
                                                    float* __restrict__ Dg, int mode) {
  __shared__ __align__(128) __half sA[M * K];
  __shared__ __align__(128) __half sB[N * K + 512];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  // canonical no-swizzle K-major: core matrix = 8 rows x 16 bytes (8 halves), 128 B contiguous;
  // here: core (rg, kc) at ((rg * (K/8)) + kc) * 128 B  => LBO (next core along K) = 128 B, SBO (next 8 rows) = K/8 * 128 B
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    sA[((r / 8) * (K / 8) + k / 8) * 64 + (r % 8) * 8 + (k % 8)] = Ag[i];
  }
  if (mode == 2) {
    for (int i = tid; i < N * 8 + K; i += 128) sB[i] = Bg[i];  // linear sequence; row n = halves [8n, 8n + K)
  } else {
    for (int i = tid; i < N * K; i += 128) {
      const int r = i / K, k = i % K;
      sB[((r / 8) * (K / 8) + k / 8) * 64 + (r % 8) * 8 + (k % 8)] = Bg[i];
    }
  }
  if (tid == 0) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  const uint32_t d_tmem = tm;            // columns [0, 16): accumulator
  const uint32_t a_tmem = tm + 32;       // columns [32, 32 + K/2): A operand (two halves per 32-bit column)
  if (mode == 1) {
    // lane (row) r of the CTA = thread r: write K halves = K/2 packed words
    uint32_t w[K / 2];
    for (int j = 0; j < K / 2; ++j) {
      const __half lo = Ag[tid * K + 2 * j], hi = Ag[tid * K + 2 * j + 1];
      w[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    const uint32_t addr = a_tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
                 "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
    for (int s = 0; s < K / 16; ++s) {
      uint64_t db;
      if (mode == 2) db = make_desc(b0 + s * 32, 16, 128);                // overlapping rows: chunk(n, kc) = n + kc
      else db = make_desc(b0 + s * 256, 128, (K / 8) * 128);
      if (mode == 1) mma_ts(d_tmem, a_tmem + s * 8, db, idesc, s > 0);
      else mma_ss(d_tmem, make_desc(a0 + s * 256, 128, (K / 8) * 128), db, idesc, s > 0);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  const uint32_t addr = d_tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) Dg[tid * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64));
}

static double run(const std::vector<__half>& A, const std::vector<__half>& B, int mode, std::vector<float>& D) {
  __half *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, M * N * 4));
  probe_kernel<<<1, 128>>>(dA, dB, dD, mode);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  D.resize(M * N);
  CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  double err = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        const double b = (mode == 2) ? (double)__half2float(B[8 * n + k]) : (double)__half2float(B[n * K + k]);
        ref += (double)__half2float(A[r * K + k]) * b;
      }
      const double e = fabs(ref - (double)D[r * N + n]);
      if (!(e <= err)) err = e;  // NaN-propagating max
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err;
}

int main() {
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  std::vector<__half> A(M * K), B(N * K + 512);
  for (auto& a : A) a = __float2half(rnd());
  for (auto& b : B) b = __float2half(rnd());
  std::vector<float> D;
  const double e0 = run(A, B, 0, D);
  const double e1 = run(A, B, 1, D);
  const double e2 = run(A, B, 2, D);
  // T3: subnormals.  A = 2^-20 (fp16 subnormal: 2^-24 * 16), B = 1024  ->  each product 2^-10, sum over K = K * 2^-10
  std::vector<__half> As(M * K, __float2half(9.5367431640625e-07f)), Bs(N * K + 512, __float2half(1024.f));
  run(As, Bs, 0, D);
  const double sub_val = D[5 * N + 3], sub_expect = K * 9.5367431640625e-07 * 1024.0;
  // T4: 3-product split accuracy on a 128 x 16 x (K = 32) problem with fp32 data: x scaled to [512, 1024), F scaled by 64
  std::vector<float> Xf(N * K), Ff(M * K);
  for (auto& x : Xf) x = rnd() * 0.013f;
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) Ff[r * K + k] = (float)cos(2.0 * M_PI * r * k / 128.0);
  float mx = 0; for (auto x : Xf) mx = fmaxf(mx, fabsf(x));
  int e; frexpf(mx, &e); const float S = ldexpf(1.f, 10 - e);  // max |x S| in [512, 1024)
  std::vector<__half> X1(N * K + 512), X2(N * K + 512), F1(M * K), F2(M * K);
  for (int i = 0; i < N * K; ++i) { const float v = Xf[i] * S; X1[i] = __float2half(v); X2[i] = __float2half(v - __half2float(X1[i])); }
  for (int i = 0; i < M * K; ++i) { const float v = Ff[i] * 64.f; F1[i] = __float2half(v); F2[i] = __float2half(v - __half2float(F1[i])); }
  std::vector<float> D11, D12, D21;
  run(F1, X1, 0, D11); run(F1, X2, 0, D12); run(F2, X1, 0, D21);
  double split_err = 0, split_ref = 0;
  for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)Ff[r * K + k] * (double)Xf[n * K + k];
    const double got = ((double)D11[r * N + n] + (double)D12[r * N + n] + (double)D21[r * N + n]) / ((double)S * 64.0);
    split_err = fmax(split_err, fabs(got - ref)); split_ref = fmax(split_ref, fabs(ref));
  }
  printf("{\"t0_ss_max_abs_err\": %.3e, \"t1_ts_max_abs_err\": %.3e, \"t2_overlap_max_abs_err\": %.3e, "
         "\"t3_subnormal_got\": %.6e, \"t3_subnormal_expect\": %.6e, \"t4_split_rel_err\": %.3e}\n",
         e0, e1, e2, sub_val, sub_expect, split_err / split_ref);
  return 0;
}
