// pitch.cu -- length-preserving pitch shift of [rows, T] waveforms on sm_100a.
//
// Replaces EffectMixin.pitch_shift (ref:audiotools/core/effects.py:247-277), which moves the batch to the
// CPU and runs libsox `pitch -q <cents>` + `rate` row by row.  SoX's pitch effect is WSOLA time-scale
// modification followed by a rate change; NO numeric output of it is pinned anywhere in the reference
// (its tests only check batch[0] == single, ref:tests/core/test_effects.py:156-181), so this file defines
// the same construction natively and parity is stated through properties (tests/test_gpu_parity.py):
// exact length, pitch ratio 2^(n/12), batch == per-item, determinism.
//
//   r = 2^(semitones/12).
//   1. WSOLA search   frames of W samples, synthesis hop Hs = W/2, nominal analysis hop Hs/r; frame j starts at
//                     p_j = a_j + d_j, d_j in [-D, D) maximising the correlation (every 2nd sample, W/4 terms)
//                     of x[a_j + d ...] with the natural continuation x[p_{j-1} + Hs ...] of the previous frame.
//   2. overlap-add    s[u] = hann(u - J Hs) x[p_J + u - J Hs] + (1 - hann(u - J Hs)) x[p_{J-1} + u - (J-1) Hs],
//                     J = floor(u / Hs)   (periodic Hann at 50 % overlap sums to one) -> x stretched by r.
//   3. rate change    y[n] = sum_k w_k s[floor(n r) + k - half + 1]: windowed sinc, cutoff c = 0.95 min(1, 1/r),
//                     8 zero crossings each side, Hann window, weights normalised to sum 1 -> length T again.
//
// Kernel 1  wsola_search_kernel  one CTA (512 threads) per row.  Frames are sequential (each depends on the
//           previous choice); inside a frame the 2D x (W/4) correlation table is a register-tiled FIR: the
//           candidate window is staged in shared memory split by sample parity (the stride-2 correlation
//           then reads unit-stride streams), a thread owns 8 candidates x a slice of the taps with a sliding
//           float4 window (6 shared loads per 32 FMAs), slices are summed through shared memory and the
//           arg-max is a shuffle tree.  The NEXT frame's window only depends on the nominal positions, so it
//           streams in with cp.async underneath the current frame's arithmetic.
// Kernel 2  wsola_ola_kernel     one thread per stretched sample (coalesced reads of x, write s once).
// Kernel 3  rate_kernel          one thread per output; the 2*half weights are generated in registers with
//           two Chebyshev recurrences (sin(pi c t) and the window's cos) instead of a phase table: a table
//           indexed by each lane's own fractional phase costs ~32 cache lines per load.
#include "b2a_common.h"

namespace b2a {
namespace pitch {

constexpr int ST = 512;  // search threads per CTA

struct Geo {
  int W, Hs, D, Lc;  // frame, synthesis hop, search radius, correlation taps (every 2nd sample)
  int J;             // frames
  int half;          // interpolation taps each side
  int H;             // left halo of the stretched row (== half)
  int rcap;          // capacity (samples) of one staged search region
  int64_t SL;        // floats per stretched row (halo + samples, multiple of 4)
  double r;          // pitch ratio = stretch factor
  float c, pic;      // cutoff, pi*c
  float cb, sb;      // cos, sin of pi*c        (sinc numerator recurrence)
  float cw, sw;      // cos, sin of pi/half     (window recurrence)
  float inv_half;
  float inv_Hs;       // 1 / Hs (the Hann argument is 2 t / W = t / Hs)
  int log2Hs;
};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
#ifdef B2A_SIM
  *smem_dst = *gmem_src;
#else
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifndef B2A_SIM
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

// nominal analysis position of frame j (double arithmetic, once per call; every row shares the table)
__global__ void nominal_kernel(int* __restrict__ nom, Geo g) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.J) nom[j] = (int)floor((double)j * (double)g.Hs / g.r + 0.5);
}

// prefer the larger correlation; ties go to the smaller |offset|, then the smaller offset (deterministic)
__device__ __forceinline__ bool better(float o, int od, float v, int d) {
  return (o > v) || (o == v && (abs(od) < abs(d) || (abs(od) == abs(d) && od < d)));
}

// Region of frame j: covers its candidate window [a_j - D, a_j + D + span) and every possible continuation
// of frame j-1, [a_{j-1} - D + Hs, a_{j-1} + D + Hs + span).  lo is chosen so that a_j - D - lo is a multiple
// of 8: the parity streams of the candidate window then start float4-aligned.
__device__ __forceinline__ void region_of(int aj, int ap, const Geo& g, int& lo, int& rn) {
  const int span = 2 * g.Lc;
  const int mn = min(aj - g.D, ap - g.D + g.Hs);
  const int mx = max(aj + g.D + span, ap + g.D + g.Hs + span);
  lo = aj - g.D - 8 * ((aj - g.D - mn + 7) >> 3);
  rn = min((mx - lo + 7) & ~7, g.rcap);
}

__global__ void __launch_bounds__(ST)
wsola_search_kernel(const float* __restrict__ x, int T, Geo g, const int* __restrict__ nom /*[J]*/,
                    int* __restrict__ pos /*[rows, J]*/) {
  B2A_DYN_SMEM(smem);
  // layout (floats): [2 buffers][E: RH][O: RH] | part[8*ST] | tfb[Lc] ; RH = rcap/2 + 16 rounded to 16 mod 32
  const int RH = (((g.rcap >> 1) + 16 + 31) & ~31) + 16;
  float* reg = reinterpret_cast<float*>(smem);
  float* part = reg + 4 * RH;
  float* tfb = part + 8 * ST;
  __shared__ float wv[ST / 32];
  __shared__ int wd[ST / 32];
  __shared__ int s_prev[2];
  const float* xr = x + (size_t)blockIdx.x * (size_t)T;
  int* pr = pos + (size_t)blockIdx.x * g.J;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int span = 2 * g.Lc;
  const int ne = g.D;       // candidates per parity
  const int G = ne >> 3;    // 8-candidate groups per parity
  const int og = tid % (2 * G), kq = tid / (2 * G);
  const int parity = og / G, gi = og - parity * G;
  const int KS = min(ST / (2 * G), g.Lc >> 2);  // tap slices
  const int TS = g.Lc / KS;                     // taps per slice (multiple of 4)
  const bool active = kq < KS;
  if (tid == 0) { pr[0] = 0; s_prev[0] = 0; }
  if (g.J > 1) {
    int lo, rn;
    region_of(__ldg(nom + 1), __ldg(nom), g, lo, rn);
    float* E = reg + 2 * RH;  // buffer 1
    for (int i = tid; i < rn; i += ST) {
      const int u = lo + i;
      float* dst = E + ((i & 1) ? RH : 0) + (i >> 1);
      if (u >= 0 && u < T) cp_async4(dst, xr + u); else *dst = 0.f;
    }
  }
  for (int j = 1; j < g.J; ++j) {
    cp_async_wait_all();
    __syncthreads();  // region j landed; s_prev[(j-1)&1] visible; buffer (j+1)&1 and part[] are free again
    const int prev = s_prev[(j - 1) & 1];
    const int a = __ldg(nom + j);
    const int cont = prev + g.Hs;  // natural continuation of frame j-1
    int lo, rn;
    region_of(a, __ldg(nom + j - 1), g, lo, rn);
    const float* E = reg + (j & 1) * 2 * RH;
    const float* O = E + RH;
    if (j + 1 < g.J) {  // stream the next region in underneath this frame's correlations
      int lo2, rn2;
      region_of(__ldg(nom + j + 1), a, g, lo2, rn2);
      float* E2 = reg + ((j + 1) & 1) * 2 * RH;
      for (int i = tid; i < rn2; i += ST) {
        const int u = lo2 + i;
        float* dst = E2 + ((i & 1) ? RH : 0) + (i >> 1);
        if (u >= 0 && u < T) cp_async4(dst, xr + u); else *dst = 0.f;
      }
    }
    int best = min(max(a, 0), max(T - g.W, 0));
    if (cont + span <= T && a - g.D >= 0 && a + g.D + span <= T) {  // uniform over the CTA
      const int s = a - g.D - lo;  // multiple of 8, >= 0
      const int c = cont - lo;
      const float* Tp;
      if (c >= 0 && c + span <= rn && s + 2 * g.D + span <= rn) {
        Tp = ((c & 1) ? O : E) + (c >> 1);
      } else {  // continuation outside the staged region (only after a clamped frame): fetch it
        for (int i = tid; i < g.Lc; i += ST) tfb[i] = __ldg(xr + cont + 2 * i);
        __syncthreads();
        Tp = tfb;
      }
      if (active) {
        const float* S = (parity ? O : E) + (s >> 1) + TS * kq;
        const float* tq = Tp + TS * kq;
        const int b0 = 4 * gi, b1 = b0 + (ne >> 1);
        float4 wa0 = *reinterpret_cast<const float4*>(S + b0);
        float4 wa1 = *reinterpret_cast<const float4*>(S + b1);
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll 2
        for (int i = 0; i < TS; i += 4) {
          const float4 wb0 = *reinterpret_cast<const float4*>(S + b0 + i + 4);
          const float4 wb1 = *reinterpret_cast<const float4*>(S + b1 + i + 4);
          const float t0 = tq[i], t1 = tq[i + 1], t2 = tq[i + 2], t3 = tq[i + 3];
          acc[0] = fmaf(t0, wa0.x, fmaf(t1, wa0.y, fmaf(t2, wa0.z, fmaf(t3, wa0.w, acc[0]))));
          acc[1] = fmaf(t0, wa0.y, fmaf(t1, wa0.z, fmaf(t2, wa0.w, fmaf(t3, wb0.x, acc[1]))));
          acc[2] = fmaf(t0, wa0.z, fmaf(t1, wa0.w, fmaf(t2, wb0.x, fmaf(t3, wb0.y, acc[2]))));
          acc[3] = fmaf(t0, wa0.w, fmaf(t1, wb0.x, fmaf(t2, wb0.y, fmaf(t3, wb0.z, acc[3]))));
          acc[4] = fmaf(t0, wa1.x, fmaf(t1, wa1.y, fmaf(t2, wa1.z, fmaf(t3, wa1.w, acc[4]))));
          acc[5] = fmaf(t0, wa1.y, fmaf(t1, wa1.z, fmaf(t2, wa1.w, fmaf(t3, wb1.x, acc[5]))));
          acc[6] = fmaf(t0, wa1.z, fmaf(t1, wa1.w, fmaf(t2, wb1.x, fmaf(t3, wb1.y, acc[6]))));
          acc[7] = fmaf(t0, wa1.w, fmaf(t1, wb1.x, fmaf(t2, wb1.y, fmaf(t3, wb1.z, acc[7]))));
          wa0 = wb0;
          wa1 = wb1;
        }
        float* pq = part + kq * (2 * ne) + parity * ne;
        *reinterpret_cast<float4*>(pq + b0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(pq + b1) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      __syncthreads();
      float bv = -3.4e38f;
      int bd = 0;
      if (tid < 2 * ne) {
        float v = 0.f;
        for (int q = 0; q < KS; ++q) v += part[q * (2 * ne) + tid];
        const int par = tid >= ne, e = tid - par * ne;
        bv = v;
        bd = 2 * e + par - g.D;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int od = __shfl_xor_sync(0xffffffffu, bd, o);
        if (better(ov, od, bv, bd)) { bv = ov; bd = od; }
      }
      if (lane == 0) { wv[warp] = bv; wd[warp] = bd; }
      __syncthreads();
      if (warp == 0) {
        bv = lane < ST / 32 ? wv[lane] : -3.4e38f;
        bd = lane < ST / 32 ? wd[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int od = __shfl_xor_sync(0xffffffffu, bd, o);
          if (better(ov, od, bv, bd)) { bv = ov; bd = od; }
        }
        best = a + bd;
      }
    }
    if (tid == 0) { pr[j] = best; s_prev[j & 1] = best; }
  }
}

// stretched row: sbuf[row][H + u] = s[u]; the H-sample halo in front is zero
__global__ void __launch_bounds__(256)
wsola_ola_kernel(const float* __restrict__ x, const int* __restrict__ pos, float* __restrict__ sbuf, int T, Geo g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.SL) return;
  const int row = blockIdx.y;
  const int u = i - g.H;
  float v = 0.f;
  if (u >= 0) {
    const float* xr = x + (size_t)row * (size_t)T;
    const int* pr = pos + (size_t)row * g.J;
    const int J0 = u >> g.log2Hs, t0 = u - (J0 << g.log2Hs);
    const float h0 = 0.5f - 0.5f * cospif((float)t0 * g.inv_Hs);
    if (J0 < g.J) {
      const int idx = __ldg(pr + J0) + t0;
      if (idx >= 0 && idx < T) v = h0 * __ldg(xr + idx);
    }
    if (J0 >= 1 && J0 - 1 < g.J) {
      const int idx = __ldg(pr + J0 - 1) + t0 + g.Hs;
      if (idx >= 0 && idx < T) v = fmaf(1.0f - h0, __ldg(xr + idx), v);
    }
  }
  sbuf[(size_t)row * (size_t)g.SL + i] = v;
}

__global__ void __launch_bounds__(256)
rate_kernel(const float* __restrict__ sbuf, float* __restrict__ y, int T, Geo g) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= T) return;
  const int row = blockIdx.y;
  const double P = (double)n * g.r;  // read position in the stretched signal
  const int ip = (int)P;
  const float f = (float)(P - (double)ip);
  // tap k reads s[ip + k - half + 1] = sbuf[H + ...] = sbuf[ip + k + 1]; its distance to P is t_k = t0 + k
  const float* sp = sbuf + (size_t)row * (size_t)g.SL + ip + 1;
  const float t0 = (float)(1 - g.half) - f;
  float s0, c0, ws0, wc0;
  sincospif(g.c * t0, &s0, &c0);            // sin, cos(pi c t0)
  sincospif(t0 * g.inv_half, &ws0, &wc0);   // sin, cos(pi t0 / half)
  float sm = s0 * g.cb - c0 * g.sb;         // sin(pi c (t0 - 1))
  float cm = wc0 * g.cw + ws0 * g.sw;       // cos(pi (t0 - 1) / half)
  float sk = s0, ck = wc0;
  const float two_cb = 2.0f * g.cb, two_cw = 2.0f * g.cw;
  float acc = 0.f, wsum = 0.f;
  const int half = g.half;
  // taps 0 .. half-2 and half+1 .. 2 half-1 have |t| >= 1: sin(pi c t)/t straight from the recurrence
#define B2A_RATE_STEP()                                          \
  {                                                              \
    const float sn = fmaf(two_cb, sk, -sm), cn = fmaf(two_cw, ck, -cm); \
    sm = sk; sk = sn; cm = ck; ck = cn;                          \
  }
#pragma unroll 4
  for (int k = 0; k < half - 1; ++k) {
    const float w = fmaf(0.5f, ck, 0.5f) * __fdividef(sk, t0 + (float)k);
    wsum += w;
    acc = fmaf(w, __ldg(sp + k), acc);
    B2A_RATE_STEP();
  }
  // the two taps around the read position (t = -f and 1 - f): where |t| is small the recurrence's absolute
  // error would be amplified by 1/t, so use the series of sin(z)/z there
#pragma unroll
  for (int k2 = 0; k2 < 2; ++k2) {
    const int k = half - 1 + k2;
    const float t = (float)k2 - f;
    const float z2 = (g.pic * t) * (g.pic * t);
    const float sinc = fabsf(t) < 0.1f ? g.pic * fmaf(z2, fmaf(z2, 1.0f / 120.0f, -1.0f / 6.0f), 1.0f) : __fdividef(sk, t);
    const float w = fmaf(0.5f, ck, 0.5f) * sinc;
    wsum += w;
    acc = fmaf(w, __ldg(sp + k), acc);
    B2A_RATE_STEP();
  }
#pragma unroll 4
  for (int k = half + 1; k < 2 * half; ++k) {
    const float w = fmaf(0.5f, ck, 0.5f) * __fdividef(sk, t0 + (float)k);
    wsum += w;
    acc = fmaf(w, __ldg(sp + k), acc);
    B2A_RATE_STEP();
  }
#undef B2A_RATE_STEP
  y[(size_t)row * (size_t)T + n] = __fdividef(acc, wsum);
}

static int geometry(int64_t rows, int64_t T, int sr, float semitones, Geo* g) {
  (void)rows;
  const double r = pow(2.0, (double)semitones / 12.0);
  int W = 1;
  const double target = 0.046 * sr;
  while (W * 2 <= target * 1.4142135623730951) W *= 2;  // nearest power of two (in log scale)
  if (W < 64) W = 64;
  if (W > 2048) W = 2048;
  memset(g, 0, sizeof(*g));
  g->W = W; g->Hs = W / 2; g->D = W / 8; g->Lc = W / 4; g->r = r;
  g->J = (int)((double)T * r / g->Hs) + 2;
  const double c = 0.95 * (r > 1.0 ? 1.0 / r : 1.0);
  g->half = (int)ceil(8.0 / c);
  g->H = g->half;
  const int64_t Ls = (int64_t)ceil((double)T * r) + g->half + 2;
  g->SL = (g->H + Ls + 3) / 4 * 4;
  const int drift = (int)ceil(fabs((double)g->Hs / r - (double)g->Hs)) + 1;
  g->rcap = (2 * g->D + 2 * g->Lc + drift + 16 + 63) / 64 * 64;
  const double PI = 3.14159265358979323846;
  g->c = (float)c; g->pic = (float)(PI * c);
  g->cb = (float)cos(PI * c); g->sb = (float)sin(PI * c);
  g->cw = (float)cos(PI / g->half); g->sw = (float)sin(PI / g->half);
  g->inv_half = (float)(1.0 / g->half);
  g->inv_Hs = 1.0f / (float)g->Hs;
  g->log2Hs = 0;
  while ((1 << g->log2Hs) < g->Hs) ++g->log2Hs;
  return 0;
}

static size_t pos_bytes(int64_t rows, const Geo& g) { return ((size_t)(rows + 1) * g.J * 4 + 255) / 256 * 256; }  // + nominal[J]

}  // namespace pitch
}  // namespace b2a

using namespace b2a::pitch;

extern "C" size_t b2a_pitch_shift_workspace_bytes(int64_t rows, int64_t T, int sr, float semitones) {
  if (rows < 1 || T < 1 || sr < 1 || !(fabsf(semitones) <= 24.f)) return 0;
  Geo g;
  geometry(rows, T, sr, semitones, &g);
  return pos_bytes(rows, g) + (size_t)rows * (size_t)g.SL * 4;
}

extern "C" int b2a_pitch_shift_f32(const float* x, int64_t rows, int64_t T, int sr, float semitones, float* out,
                                   void* ws, size_t ws_bytes, void* stream) {
  B2A_REQUIRE(x && out && ws, B2A_E_INVALID, "pitch_shift: null pointer");
  B2A_REQUIRE(rows >= 1 && T >= 1 && sr >= 1, B2A_E_INVALID, "pitch_shift: bad argument");
  B2A_REQUIRE(fabsf(semitones) <= 24.f, B2A_E_UNSUPPORTED, "pitch_shift: |semitones| > 24");
  B2A_REQUIRE(rows * T < ((int64_t)1 << 40) && T < ((int64_t)1 << 28) && rows <= 65535, B2A_E_UNSUPPORTED,
              "pitch_shift: too large");
  B2A_REQUIRE(((uintptr_t)ws & 15) == 0, B2A_E_INVALID, "pitch_shift: workspace must be 16-byte aligned");
  Geo g;
  geometry(rows, T, sr, semitones, &g);
  const size_t pb = pos_bytes(rows, g);
  B2A_REQUIRE(ws_bytes >= pb + (size_t)rows * (size_t)g.SL * 4, B2A_E_INVALID, "pitch_shift: workspace too small");
  int* pos = (int*)ws;
  int* nom = pos + (size_t)rows * g.J;
  float* sbuf = (float*)((char*)ws + pb);
  const int RH = (((g.rcap >> 1) + 16 + 31) & ~31) + 16;
  const size_t smem = (size_t)(4 * RH + 8 * ST + g.Lc) * 4;
  B2A_CUDA_OK(cudaFuncSetAttribute(wsola_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2A_LAUNCH(nominal_kernel, dim3((unsigned)((g.J + 255) / 256)), dim3(256), 0, stream, nom, g);
  B2A_LAUNCH(wsola_search_kernel, dim3((unsigned)rows), dim3(ST), smem, stream, x, (int)T, g, (const int*)nom, pos);
  B2A_LAUNCH(wsola_ola_kernel, dim3((unsigned)((g.SL + 255) / 256), (unsigned)rows), dim3(256), 0, stream, x,
             (const int*)pos, sbuf, (int)T, g);
  B2A_LAUNCH(rate_kernel, dim3((unsigned)((T + 255) / 256), (unsigned)rows), dim3(256), 0, stream, (const float*)sbuf,
             out, (int)T, g);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
