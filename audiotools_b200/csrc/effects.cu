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

namespace effects {
// ---------------------------------------------------------------------------------------------
// ImpulseResponseMixin.alter_drr (ref:audiotools/core/effects.py:540-647), one CTA per impulse-response row, one launch:
//   td = argmax(x), early = [td - t0, td + t0], window = the early region of the item's CHANNEL 0 (hann(1) == 1, so the
//   reference's window is that indicator), alpha from the quadratic of solve_alpha (:594-617, float32 like the
//   reference; with an indicator window b == 0), floored at max|late| / max|early|, out = alpha on (early AND window),
//   x elsewhere, then ensure_max_of_audio (:181-198).  The reference runs ~25 tensor passes for this.
// ---------------------------------------------------------------------------------------------
constexpr int DRR_T = 512;

struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {  // larger value, then the smaller index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ ArgMax block_argmax(const float* __restrict__ x, int T, ArgMax* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ArgMax m{-INFINITY, 0x7fffffff};
  for (int i = tid; i < T; i += DRR_T) { const float v = x[i]; if (v > m.v) { m.v = v; m.i = i; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax t{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
    m = better(m, t);
  }
  __syncthreads();
  if (lane == 0) sm[warp] = m;
  __syncthreads();
  ArgMax r = sm[0];
  for (int w = 1; w < DRR_T / 32; ++w) r = better(r, sm[w]);
  if (r.i == 0x7fffffff) r.i = 0;
  return r;
}
__device__ float block_sum_f(float v, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < DRR_T / 32; ++w) r += sm[w];
  return r;
}
__device__ float block_max_f(float v, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = sm[0];
  for (int w = 1; w < DRR_T / 32; ++w) r = fmaxf(r, sm[w]);
  return r;
}

__global__ void __launch_bounds__(DRR_T)
alter_drr_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int C, int t0, const float* __restrict__ drr,
                 float max_abs) {
  __shared__ ArgMax s_am[DRR_T / 32];
  __shared__ float s_f[DRR_T / 32];
  const int row = blockIdx.x, item = row / C, tid = threadIdx.x;
  const float* xr = x + (size_t)row * T;
  const float* x0 = x + (size_t)item * C * T;  // channel 0 of the item: its early region is the window
  const int td = block_argmax(xr, T, s_am).i;
  const int tw = (C == 1 || row == item * C) ? td : block_argmax(x0, T, s_am).i;
  float a = 0.f, ce = 0.f, lsq = 0.f, ml = 0.f, me = 0.f, mew = 0.f, meo = 0.f;
  for (int i = tid; i < T; i += DRR_T) {
    const float v = xr[i], av = fabsf(v);
    const bool e = (i >= td - t0) && (i <= td + t0), w = (i >= tw - t0) && (i <= tw + t0);
    if (e) {
      me = fmaxf(me, av);
      if (w) { a = fmaf(v, v, a); mew = fmaxf(mew, av); }
      else { ce = fmaf(v, v, ce); meo = fmaxf(meo, av); }
    } else {
      lsq = fmaf(v, v, lsq);
      ml = fmaxf(ml, av);
    }
  }
  a = block_sum_f(a, s_f); ce = block_sum_f(ce, s_f); lsq = block_sum_f(lsq, s_f);
  ml = block_max_f(ml, s_f); me = block_max_f(me, s_f); mew = block_max_f(mew, s_f); meo = block_max_f(meo, s_f);
  // solve_alpha with an indicator window: b = 0, c = sum_{early, outside the window} x^2 - 10^(drr/10) sum_late x^2
  const float c = ce - powf(10.0f, __ldg(drr + item) / 10.0f) * lsq;
  const float b = 0.f;
  const float expr = sqrtf(b * b - 4.0f * a * c);
  const float r1 = (-b - expr) / (2.0f * a), r2 = (-b + expr) / (2.0f * a);
  float alpha = (r1 != r1 || r2 != r2) ? NAN : fmaxf(r1, r2);  // torch.maximum propagates nan
  const float min_alpha = ml / me;
  alpha = (alpha != alpha || min_alpha != min_alpha) ? NAN : fmaxf(alpha, min_alpha);
  // ensure_max_of_audio on the altered response.  A non-finite alpha (a = 0: the row's early region does not meet
  // the window, e.g. a second channel whose direct path lies > 2.5 ms from channel 0's) makes the reference's
  // `alpha * window * early` NaN on the WHOLE row (NaN * 0), its peak NaN and its peak gain 1: same here.
  const bool finite = (alpha - alpha) == 0.0f;
  const float peak = fmaxf(fmaxf(fabsf(alpha) * mew, meo), ml);
  const float pg = (finite && peak > max_abs) ? max_abs / peak : 1.0f;
  float* o = out + (size_t)row * T;
  for (int i = tid; i < T; i += DRR_T) {
    const float v = xr[i];
    const bool e = (i >= td - t0) && (i <= td + t0), w = (i >= tw - t0) && (i <= tw + t0);
    const float wf = w ? 1.0f : 0.0f, ev = e ? v : 0.0f, lv = e ? 0.0f : v;
    o[i] = (alpha * wf * ev + (1.0f - wf) * ev + lv) * pg;  // the reference's expression, term by term (:642)
  }
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

extern "C" int b2a_alter_drr_f32(const float* ir, float* out, int64_t rows, int64_t T, int C, int t0, const float* drr,
                                 float max_abs, void* stream) {
  B2A_REQUIRE(ir && out && drr, B2A_E_INVALID, "alter_drr: null pointer");
  B2A_REQUIRE(rows >= 1 && T >= 1 && T < ((int64_t)1 << 31) && C >= 1 && rows % C == 0 && t0 >= 0, B2A_E_INVALID,
              "alter_drr: bad shape");
  B2A_REQUIRE(out != ir, B2A_E_INVALID, "alter_drr: out must not alias ir");
  B2A_LAUNCH(alter_drr_kernel, dim3((unsigned)rows), dim3(DRR_T), 0, stream, ir, out, (int)T, C, t0, drr, max_abs);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
