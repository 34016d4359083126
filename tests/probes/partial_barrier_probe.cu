// Probe: does compute-sanitizer synccheck accept PTX named barriers over a SUBSET of a CTA's warps
// (bar.sync id, count) -- the construct kweight_energy_kernel uses (8 worker warps + 1 carry warp)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int* out) {
  __shared__ int s[288];
  const int tid = threadIdx.x, warp = tid >> 5;
  s[tid] = tid;
  __syncthreads();
  if (warp == 8) {  // "carry" warp: only joins barrier 2
    asm volatile("bar.sync 2, 288;" ::: "memory");
    return;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // workers only
  int v = s[(tid + 1) & 255];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  asm volatile("bar.sync 2, 288;" ::: "memory");
  out[tid] = v;
}
int main() {
  int* d;
  cudaMalloc(&d, 288 * 4);
  probe<<<2, 288>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
