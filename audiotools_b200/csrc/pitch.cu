// pitch.cu -- length-preserving pitch shift of [rows, T] waveforms on sm_100a.
//
// Replaces EffectMixin.pitch_shift (ref:audiotools/core/effects.py:247-277), which moves the batch to the
// CPU and runs libsox `pitch -q <cents>` + `rate` row by row.  SoX's pitch effect is WSOLA time-scale
// modification followed by a rate change; NO numeric output of it is pinned anywhere in the reference
// (its tests only check batch[0] == single, ref:tests/core/test_effects.py:156-181), so this kernel defines
// the same construction natively and parity is stated through properties (tests/test_gpu_parity.py):
// exact length, pitch ratio 2^(n/12), batch == per-item, determinism.
//
//   r = 2^(semitones/12).  Stretch x by r with WSOLA (Hann frames of W samples, synthesis hop W/2,
//   analysis hop W/(2r), best of 2*D+1 offsets by cross-correlation with the natural continuation of the
//   previous frame), then read the stretched signal r times faster:
//     y[n] = sum_{j in {J-1, J}} hann(u - j*Hs) * x~(p_j + u - j*Hs),   u = n*r,  J = floor(u / Hs)
//   where p_j is frame j's chosen start and x~ is band-limited interpolation of x (windowed sinc, cutoff
//   0.95*min(1, 1/r), 8 zero crossings, 256 tabulated phases, linearly interpolated).  The stretched signal
//   is never materialised: the render kernel evaluates the overlap-add directly from x.
//
// Kernel 1  wsola_search_kernel  one CTA per row; frames are sequential (each depends on the previous
//           choice), the 2*D+1 candidate correlations of a frame are parallel across the CTA.
// Kernel 2  pitch_render_kernel  one thread per output sample (HBM: read x ~r times through L1/L2, write y once).
#include "b2a_common.h"

namespace b2a {
namespace pitch {

constexpr int ST = 256;  // search threads per CTA

struct Geo {
  int W, Hs, D, Lc;  // frame, synthesis hop, search radius, correlation length (samples, decimated by 2)
  int J;             // frames
  float r;           // pitch ratio = stretch factor
};

__global__ void __launch_bounds__(ST)
wsola_search_kernel(const float* __restrict__ x, int T, Geo g, int* __restrict__ pos /*[rows, J]*/) {
  __shared__ float tmpl[2048];      // natural continuation, Lc*2 <= 2048 samples
  __shared__ float sv[ST];
  __shared__ int si[ST];
  __shared__ int s_prev;
  const float* xr = x + (size_t)blockIdx.x * (size_t)T;
  int* pr = pos + (size_t)blockIdx.x * g.J;
  const int tid = threadIdx.x;
  if (tid == 0) { pr[0] = 0; s_prev = 0; }
  __syncthreads();
  const int span = 2 * g.Lc;  // samples covered by a correlation (stride 2)
  for (int j = 1; j < g.J; ++j) {
    const int prev = s_prev;
    __syncthreads();  // everyone has read s_prev before thread 0 overwrites it below
    const int a = (int)floorf((float)j * (float)g.Hs / g.r + 0.5f);  // nominal analysis position
    const int cont = prev + g.Hs;                                      // natural continuation of frame j-1
    int best = min(max(a, 0), max(T - g.W, 0));
    if (cont + span <= T && a - g.D >= 0 && a + g.D + span <= T) {
      for (int i = tid; i < span; i += ST) tmpl[i] = __ldg(xr + cont + i);
      __syncthreads();
      float bv = -3.4e38f;
      int bd = 0;
      for (int d = tid - g.D; d <= g.D; d += ST) {
        const float* c = xr + a + d;
        float acc = 0.f;
        for (int i = 0; i < span; i += 2) acc = fmaf(tmpl[i], __ldg(c + i), acc);
        if (acc > bv) { bv = acc; bd = d; }
      }
      sv[tid] = bv;
      si[tid] = bd;
      __syncthreads();
      for (int s = ST / 2; s > 0; s >>= 1) {
        if (tid < s) {
          const float o = sv[tid + s];
          const int od = si[tid + s];
          // prefer the larger correlation; ties go to the smaller |offset| then the smaller offset (deterministic)
          const bool better = (o > sv[tid]) || (o == sv[tid] && (abs(od) < abs(si[tid]) || (abs(od) == abs(si[tid]) && od < si[tid])));
          if (better) { sv[tid] = o; si[tid] = od; }
        }
        __syncthreads();
      }
      best = a + si[0];
    }
    if (tid == 0) { pr[j] = best; s_prev = best; }
    __syncthreads();
  }
}

// table: [Q+1][NT] windowed-sinc weights for fractional phase q/Q; tap k reads x[floor(pos) + k - NT/2 + 1]
__global__ void __launch_bounds__(256)
pitch_render_kernel(const float* __restrict__ x, float* __restrict__ y, const int* __restrict__ pos, int T, Geo g,
                    const float* __restrict__ table, int Q, int NT, int64_t total) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int row = (int)(gid / T), n = (int)(gid - (int64_t)row * T);
  const float* xr = x + (size_t)row * (size_t)T;
  const int* pr = pos + (size_t)row * g.J;
  const double u = (double)n * (double)g.r;  // position in the (virtual) stretched signal
  const int j1 = (int)(u / g.Hs);
  float acc = 0.f;
#pragma unroll
  for (int dj = 0; dj < 2; ++dj) {
    const int j = j1 - 1 + dj;
    if (j < 0 || j >= g.J) continue;
    const double t = u - (double)j * g.Hs;  // offset inside frame j, [0, W)
    if (t < 0.0 || t >= (double)g.W) continue;
    const float wv = 0.5f - 0.5f * cospif(2.0f * (float)(t / g.W));
    const double p = (double)__ldg(pr + j) + t;  // read position in x
    const int ip = (int)floor(p);
    const float fq = (float)(p - ip) * Q;
    const int q = min((int)fq, Q - 1);
    const float fr = fq - q;
    const float* t0 = table + (size_t)q * NT;
    const float* t1 = t0 + NT;
    float s = 0.f;
    const int base = ip - NT / 2 + 1;
    for (int k = 0; k < NT; ++k) {
      const int idx = base + k;
      if (idx >= 0 && idx < T) {
        const float w0 = __ldg(t0 + k);
        s = fmaf(fmaf(fr, __ldg(t1 + k) - w0, w0), __ldg(xr + idx), s);
      }
    }
    acc = fmaf(wv, s, acc);
  }
  y[gid] = acc;
}

static int geometry(int64_t T, int sr, float semitones, Geo* g) {
  const double r = pow(2.0, (double)semitones / 12.0);
  int W = 1;
  const double target = 0.046 * sr;
  while (W * 2 <= target * 1.4142135623730951) W *= 2;  // nearest power of two (in log scale)
  if (W < 64) W = 64;
  if (W > 2048) W = 2048;
  g->W = W; g->Hs = W / 2; g->D = W / 8; g->Lc = W / 4; g->r = (float)r;
  g->J = (int)((double)T * r / g->Hs) + 2;
  return 0;
}

}  // namespace pitch
}  // namespace b2a

using namespace b2a::pitch;

extern "C" size_t b2a_pitch_shift_workspace_bytes(int64_t rows, int64_t T, int sr, float semitones) {
  if (rows < 1 || T < 1 || sr < 1) return 0;
  Geo g;
  geometry(T, sr, semitones, &g);
  return (size_t)rows * g.J * 4 + 256;
}

extern "C" int b2a_pitch_shift_f32(const float* x, int64_t rows, int64_t T, int sr, float semitones,
                                   const float* table, int Q, int NT, float* out, void* ws, size_t ws_bytes,
                                   void* stream) {
  B2A_REQUIRE(x && out && ws && table, B2A_E_INVALID, "pitch_shift: null pointer");
  B2A_REQUIRE(rows >= 1 && T >= 1 && sr >= 1 && Q >= 1 && NT >= 2, B2A_E_INVALID, "pitch_shift: bad argument");
  B2A_REQUIRE(fabsf(semitones) <= 24.f, B2A_E_UNSUPPORTED, "pitch_shift: |semitones| > 24");
  B2A_REQUIRE(rows * T < ((int64_t)1 << 40) && T < ((int64_t)1 << 30), B2A_E_UNSUPPORTED, "pitch_shift: too large");
  Geo g;
  geometry(T, sr, semitones, &g);
  B2A_REQUIRE(ws_bytes >= (size_t)rows * g.J * 4, B2A_E_INVALID, "pitch_shift: workspace too small");
  int* pos = (int*)ws;
  B2A_LAUNCH(wsola_search_kernel, dim3((unsigned)rows), dim3(ST), 0, stream, x, (int)T, g, pos);
  const int64_t total = rows * T;
  B2A_LAUNCH(pitch_render_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, out,
             (const int*)pos, (int)T, g, table, Q, NT, total);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
