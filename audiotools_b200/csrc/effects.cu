// effects.cu -- the element-wise / per-row-peak effects of EffectMixin on sm_100a: each is ONE pass
// over the waveform (HBM-bound: read x once, write y once) instead of the reference's chain of tensor temporaries.
//
//   b2a_row_absmax_f32   peak[row] = max |x|                      ref:audiotools/core/effects.py:194 (ensure_max_of_audio),
//                                                                  :155,:176 (apply_ir peak restore), :639 (alter_drr)
//   b2a_limit_peak_f32   y = x * (peak > max ? max / peak : 1)    ref :181-198
//   b2a_mix_f32          y = x + g[item] * other                  ref :27-64 (the normalize() multiply of `other` and the add)
//   b2a_quantize_f32     linear / mu-law quantisation             ref :463-523 (same float32 operation order, incl. the
//                                                                  `x - (x - q)` straight-through residual)
//   b2a_order_stats_f32  k-th smallest values of one row           ref :452-453 (torch.quantile's sorted gather), by
//                        (exact: 4-pass radix select)               radix selection instead of a full sort
//   b2a_clamp_items_f32  y = min(max(x, lo[item]), hi[item])       ref :459
#include "b2a_common.h"

namespace b2a {
namespace effects {

constexpr int TPB = 256;

__device__ __forceinline__ int64_t grid_threads() { return (int64_t)gridDim.x * blockDim.x; }

// peak must be zeroed by the caller (memset inside the entry point); |x| >= 0, so float bits order like ints
__global__ void __launch_bounds__(TPB) absmax_kernel(const float* __restrict__ x, int64_t T, int vec_ok,
                                                     float* __restrict__ peak) {
  const int row = blockIdx.y;
  const float* xr = x + (size_t)row * (size_t)T;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = grid_threads();
  float m = 0.f;
  if (vec_ok) {
    const int64_t n4 = T >> 2;
    for (int64_t i = gid; i < n4; i += nt) {
      const float4 v = ld_stream4(xr + 4 * i);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int64_t i = (n4 << 2) + gid; i < T; i += nt) m = fmaxf(m, fabsf(xr[i]));
  } else {
    for (int64_t i = gid; i < T; i += nt) m = fmaxf(m, fabsf(xr[i]));
  }
  m = warp_max(m);
  __shared__ float s[TPB / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < TPB / 32 ? s[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(peak + row), __float_as_int(m));
  }
}

// MODE 0: limit peak (scale = peak > lim ? lim / peak : 1)   MODE 1: clamp to [lo, hi] of the item
template <int MODE>
__global__ void __launch_bounds__(TPB) rows_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t T,
                                                   int vec_ok, const float* __restrict__ a, const float* __restrict__ b,
                                                   float lim) {
  const int row = blockIdx.y;
  const float* xr = x + (size_t)row * (size_t)T;
  float* yr = out + (size_t)row * (size_t)T;
  float p0, p1 = 0.f;
  if (MODE == 0) {
    const float pk = __ldg(a + row);
    p0 = pk > lim ? lim / pk : 1.0f;
  } else {
    p0 = __ldg(a + row); p1 = __ldg(b + row);
  }
  auto f = [&](float v) { return MODE == 0 ? v * p0 : fminf(fmaxf(v, p0), p1); };
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = grid_threads();
  if (vec_ok) {
    const int64_t n4 = T >> 2;
    for (int64_t i = gid; i < n4; i += nt) {
      float4 v = ld_stream4(xr + 4 * i);
      v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
      st_stream4(yr + 4 * i, v);
    }
    for (int64_t i = (n4 << 2) + gid; i < T; i += nt) yr[i] = f(xr[i]);
  } else {
    for (int64_t i = gid; i < T; i += nt) yr[i] = f(xr[i]);
  }
}

__global__ void __launch_bounds__(TPB) mix_kernel(const float* __restrict__ x, const float* __restrict__ other,
                                                  const float* __restrict__ gain, float* __restrict__ out,
                                                  int64_t per_item, int vec_ok) {
  const int b = blockIdx.y;
  const float g = gain ? __ldg(gain + b) : 1.0f;
  const float* xr = x + (size_t)b * per_item;
  const float* orow = other + (size_t)b * per_item;
  float* yr = out + (size_t)b * per_item;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = grid_threads();
  // the reference multiplies first (normalize) and adds afterwards: two roundings, not one fused multiply-add
  if (vec_ok) {
    const int64_t n4 = per_item >> 2;
    for (int64_t i = gid; i < n4; i += nt) {
      const float4 a = ld_stream4(xr + 4 * i), o = ld_stream4(orow + 4 * i);
      float4 v;
      v.x = __fadd_rn(a.x, __fmul_rn(o.x, g)); v.y = __fadd_rn(a.y, __fmul_rn(o.y, g));
      v.z = __fadd_rn(a.z, __fmul_rn(o.z, g)); v.w = __fadd_rn(a.w, __fmul_rn(o.w, g));
      st_stream4(yr + 4 * i, v);
    }
    for (int64_t i = (n4 << 2) + gid; i < per_item; i += nt) yr[i] = __fadd_rn(xr[i], __fmul_rn(orow[i], g));
  } else {
    for (int64_t i = gid; i < per_item; i += nt) yr[i] = __fadd_rn(xr[i], __fmul_rn(orow[i], g));
  }
}

// ref:audiotools/core/effects.py:481-491 (linear) and :509-523 (mu-law), operation by operation in float32
__device__ __forceinline__ float quant_linear(float a, float q) {
  float x = __fdiv_rn(__fadd_rn(a, 1.0f), 2.0f);
  x = floorf(__fmul_rn(x, q));
  x = __fdiv_rn(x, q);
  x = __fadd_rn(__fmul_rn(2.0f, x), -1.0f);
  const float residual = __fadd_rn(a, -x);
  return __fadd_rn(a, -residual);
}
__device__ __forceinline__ float quant_mulaw(float a, float mu, float l1p) {
  const float sg = (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f);
  float x = __fdiv_rn(__fmul_rn(sg, log1pf(__fmul_rn(mu, fabsf(a)))), l1p);
  x = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), mu), 0.5f);
  const float xi = (float)(long long)x;  // .to(torch.int64): truncation toward zero, then int / float
  x = __fadd_rn(__fmul_rn(__fdiv_rn(xi, mu), 2.0f), -1.0f);
  const float sx = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  x = __fdiv_rn(__fmul_rn(sx, __fadd_rn(expf(__fmul_rn(fabsf(x), l1p)), -1.0f)), mu);
  const float residual = __fadd_rn(a, -x);
  return __fadd_rn(a, -residual);
}

__global__ void __launch_bounds__(TPB) quantize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                       int64_t per_item, int vec_ok,
                                                       const float* __restrict__ channels, int mulaw) {
  const int b = blockIdx.y;
  const float q = __ldg(channels + b);
  const float mu = __fadd_rn(q, -1.0f), l1p = log1pf(mu);
  const float* xr = x + (size_t)b * per_item;
  float* yr = out + (size_t)b * per_item;
  auto f = [&](float v) { return mulaw ? quant_mulaw(v, mu, l1p) : quant_linear(v, q); };
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = grid_threads();
  if (vec_ok) {
    const int64_t n4 = per_item >> 2;
    for (int64_t i = gid; i < n4; i += nt) {
      float4 v = ld_stream4(xr + 4 * i);
      v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
      st_stream4(yr + 4 * i, v);
    }
    for (int64_t i = (n4 << 2) + gid; i < per_item; i += nt) yr[i] = f(xr[i]);
  } else {
    for (int64_t i = gid; i < per_item; i += nt) yr[i] = f(xr[i]);
  }
}

// ---- exact k-th smallest by 4-pass (8 bits each) radix selection: one CTA per requested order statistic
__device__ __forceinline__ unsigned ord_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_val(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr int ST = 1024;
__global__ void __launch_bounds__(ST) order_stat_kernel(const float* __restrict__ row, int64_t T,
                                                        const int64_t* __restrict__ ks, float* __restrict__ out) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix;
  __shared__ long long s_k;
  const int tid = threadIdx.x;
  long long k = ks[blockIdx.x];
  if (k < 0) k = 0;
  if (k > T - 1) k = T - 1;
  unsigned prefix = 0, mask = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += ST) hist[i] = 0;
    __syncthreads();
    for (int64_t i = tid; i < T; i += ST) {
      const unsigned key = ord_key(__ldg(row + i));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      long long c = 0;
      int bsel = 255;
      for (int bkt = 0; bkt < 256; ++bkt) {
        if (c + (long long)hist[bkt] > k) { bsel = bkt; break; }
        c += hist[bkt];
      }
      s_prefix = prefix | ((unsigned)bsel << shift);
      s_k = k - c;
    }
    __syncthreads();
    prefix = s_prefix;
    k = s_k;
    mask |= 0xffu << shift;
  }
  if (tid == 0) out[blockIdx.x] = ord_val(prefix);
}

static unsigned grid_x(int64_t work, int64_t rows) {
  const int64_t want = (work + TPB - 1) / TPB;
  int64_t cap = (int64_t)B2A_NUM_SMS * 8 / (rows > 0 ? rows : 1) + 1;  // ~8 resident CTAs per SM across the rows
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace effects
}  // namespace b2a

using namespace b2a::effects;

extern "C" int b2a_row_absmax_f32(const float* x, int64_t rows, int64_t T, float* peak, void* stream) {
  B2A_REQUIRE(x && peak, B2A_E_INVALID, "row_absmax: null pointer");
  B2A_REQUIRE(rows >= 1 && rows <= 65535 && T >= 1, B2A_E_INVALID, "row_absmax: bad shape");
  B2A_CUDA_OK(cudaMemsetAsync(peak, 0, (size_t)rows * sizeof(float), (cudaStream_t)stream));
  const int vec_ok = (((uintptr_t)x) % 16 == 0) && (T % 4 == 0);
  B2A_LAUNCH(absmax_kernel, dim3(grid_x(vec_ok ? T / 4 : T, rows), (unsigned)rows), dim3(TPB), 0, stream, x, T, vec_ok, peak);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_limit_peak_f32(const float* x, float* out, int64_t rows, int64_t T, const float* peak, float max_abs,
                                  void* stream) {
  B2A_REQUIRE(x && out && peak, B2A_E_INVALID, "limit_peak: null pointer");
  B2A_REQUIRE(rows >= 1 && rows <= 65535 && T >= 1, B2A_E_INVALID, "limit_peak: bad shape");
  const int vec_ok = (((uintptr_t)x | (uintptr_t)out) % 16 == 0) && (T % 4 == 0);
  B2A_LAUNCH(rows_kernel<0>, dim3(grid_x(vec_ok ? T / 4 : T, rows), (unsigned)rows), dim3(TPB), 0, stream, x, out, T, vec_ok,
             peak, (const float*)nullptr, max_abs);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_clamp_items_f32(const float* x, float* out, int64_t B, int64_t per_item, const float* lo,
                                   const float* hi, void* stream) {
  B2A_REQUIRE(x && out && lo && hi, B2A_E_INVALID, "clamp_items: null pointer");
  B2A_REQUIRE(B >= 1 && B <= 65535 && per_item >= 1, B2A_E_INVALID, "clamp_items: bad shape");
  const int vec_ok = (((uintptr_t)x | (uintptr_t)out) % 16 == 0) && (per_item % 4 == 0);
  B2A_LAUNCH(rows_kernel<1>, dim3(grid_x(vec_ok ? per_item / 4 : per_item, B), (unsigned)B), dim3(TPB), 0, stream, x, out,
             per_item, vec_ok, lo, hi, 0.f);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_mix_f32(const float* x, const float* other, const float* other_gain, float* out, int64_t B,
                           int64_t per_item, void* stream) {
  B2A_REQUIRE(x && other && out, B2A_E_INVALID, "mix: null pointer");
  B2A_REQUIRE(B >= 1 && B <= 65535 && per_item >= 1, B2A_E_INVALID, "mix: bad shape");
  const int vec_ok = (((uintptr_t)x | (uintptr_t)other | (uintptr_t)out) % 16 == 0) && (per_item % 4 == 0);
  B2A_LAUNCH(mix_kernel, dim3(grid_x(vec_ok ? per_item / 4 : per_item, B), (unsigned)B), dim3(TPB), 0, stream, x, other,
             other_gain, out, per_item, vec_ok);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_quantize_f32(const float* x, float* out, int64_t B, int64_t per_item, const float* channels, int mulaw,
                                void* stream) {
  B2A_REQUIRE(x && out && channels, B2A_E_INVALID, "quantize: null pointer");
  B2A_REQUIRE(B >= 1 && B <= 65535 && per_item >= 1, B2A_E_INVALID, "quantize: bad shape");
  const int vec_ok = (((uintptr_t)x | (uintptr_t)out) % 16 == 0) && (per_item % 4 == 0);
  B2A_LAUNCH(quantize_kernel, dim3(grid_x(vec_ok ? per_item / 4 : per_item, B), (unsigned)B), dim3(TPB), 0, stream, x, out,
             per_item, vec_ok, channels, mulaw ? 1 : 0);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

extern "C" int b2a_order_stats_f32(const float* row, int64_t T, const int64_t* k, int nk, float* out, void* stream) {
  B2A_REQUIRE(row && k && out, B2A_E_INVALID, "order_stats: null pointer");
  B2A_REQUIRE(T >= 1 && nk >= 1 && nk <= 65535, B2A_E_INVALID, "order_stats: bad shape");
  B2A_LAUNCH(order_stat_kernel, dim3((unsigned)nk), dim3(ST), 0, stream, row, T, k, out);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
