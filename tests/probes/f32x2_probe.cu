// f32x2_probe.cu -- issue rate and latency of the sm_100a packed FP32 instructions (FFMA2 / FADD2, PTX
// fma.rn.f32x2 / add.rn.f32x2) against scalar FFMA, plus a mix with shared-memory loads.  Decides whether packing the
// FFT butterflies of the issue-bound kernels pays.  Hardware probe, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// MODE 0: scalar FFMA, CH independent chains; 1: FFMA2; 2: FADD2; 3: FFMA2 with swapped/negated operand (LO_HI.NP);
// 4: alternating FFMA2 / FFMA; 5: FFMA2 + one LDS.64 per 4 math instructions
template <int MODE, int CH>
__global__ void __launch_bounds__(256) probe(float* out, long long* cyc, int iters, float a, float b) {
  __shared__ float2 sh[512];
  sh[threadIdx.x] = make_float2(a, b);
  sh[threadIdx.x + 256] = make_float2(b, a);
  __syncthreads();
  float s[CH];
  u64 v[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { s[i] = a + i; v[i] = pk(a + i, b - i); }
  const u64 m = pk(a, b), c = pk(b, a);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) s[i] = fma1(s[i], a, b);
      else if (MODE == 1) v[i] = fma2(v[i], m, c);
      else if (MODE == 2) v[i] = add2(v[i], c);
      else if (MODE == 3) { float2 f = upk(v[i]); v[i] = fma2(pk(f.y, -f.x), m, c); }
      else if (MODE == 4) { if (i & 1) s[i] = fma1(s[i], a, b); else v[i] = fma2(v[i], m, c); }
      else if (MODE == 5) {
        v[i] = fma2(v[i], m, c);
        if ((i & 3) == 3) {
          float2 f = sh[(threadIdx.x + it + i) & 511];
          v[i] = add2(v[i], pk(f.x, f.y));
        }
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float2 f = upk(v[i]); acc += s[i] + f.x + f.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE, int CH>
static int run(const char* name, int ctas_per_sm, float* out, long long* cyc, int n_sm) {
  const int iters = 4096;
  probe<MODE, CH><<<n_sm * ctas_per_sm, 256>>>(out, cyc, iters, 0.999f, 1e-3f);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe<MODE, CH><<<n_sm * ctas_per_sm, 256>>>(out, cyc, iters, 0.999f, 1e-3f);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  long long c = 0;
  CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  // warp instructions per SM sub-partition per cycle: (warps per partition) * CH * iters / cycles
  const double warps_pp = ctas_per_sm * 8 / 4.0;
  const double ipc = warps_pp * CH * iters / (double)c;
  printf("{\"probe\": \"%s\", \"chains\": %d, \"warps_per_scheduler\": %.0f, \"cycles\": %lld, \"math_instr_per_cycle_per_scheduler\": %.3f, \"cycles_per_instr_per_warp\": %.2f, \"ms\": %.4f}\n",
         name, CH, warps_pp, c, ipc, (double)c / iters / CH, ms);
  return 0;
}

int main() {
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 8 * 256 * 4 * 4)); CK(cudaMalloc(&cyc, 8));
  // latency: one warp per scheduler (half a CTA's warps idle is fine), one chain
  run<0, 1>("ffma latency (1 chain, 2 warps/sched)", 1, out, cyc, n_sm);
  run<1, 1>("ffma2 latency (1 chain, 2 warps/sched)", 1, out, cyc, n_sm);
  run<2, 1>("fadd2 latency (1 chain, 2 warps/sched)", 1, out, cyc, n_sm);
  // throughput: 8 chains, 4 warps per scheduler (2 CTAs of 256 per SM)
  run<0, 8>("ffma throughput", 2, out, cyc, n_sm);
  run<1, 8>("ffma2 throughput", 2, out, cyc, n_sm);
  run<2, 8>("fadd2 throughput", 2, out, cyc, n_sm);
  run<3, 8>("ffma2 swapped+negated operand", 2, out, cyc, n_sm);
  run<4, 8>("ffma2/ffma alternating", 2, out, cyc, n_sm);
  run<5, 8>("ffma2 + lds.64 per 4", 2, out, cyc, n_sm);
  return 0;
}
