// specmask.cu -- SpecAugment-style band masks on a complex STFT, in place, on sm_100a.
//
// Replaces the mask construction of DSPMixin.mask_frequencies / mask_timesteps
// (ref:audiotools/core/dsp.py:217-306): the reference takes |X| and angle(X), repeats the bin axis to the full
// [B, 1, F, N] shape, builds the boolean mask, masked_fills magnitude and phase and recombines
// mag * exp(1j * phase) -- about a dozen passes over the spectrogram.  Cells outside the band are unchanged by
// that round trip (up to its polar/rectangular rounding), so this kernel touches ONLY the masked cells: it
// evaluates  lo[item] <= axis_val < hi[item]  in float32 exactly as the reference does (axis_val = the
// reference's own torch.linspace values, passed in) and stores the constant fill = val * exp(1j * val).
// No loads of the spectrogram at all; bytes written = masked fraction x 8 B.
#include "b2a_common.h"

namespace b2a {
namespace specmask {

__global__ void __launch_bounds__(256)
band_mask_kernel(float2* __restrict__ spec, long long total, int F, int N, const float* __restrict__ axis_vals,
                 const float* __restrict__ lo, const float* __restrict__ hi, int rows_per_item, int axis, float2 fill) {
  const long long FN = (long long)F * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / FN;
    const int rem = (int)(idx - row * FN);
    const int f = rem / N, n = rem - f * N;
    const float v = __ldg(axis_vals + (axis == 0 ? f : n));
    const int item = (int)(row / rows_per_item);
    if (__ldg(lo + item) <= v && v < __ldg(hi + item)) spec[idx] = fill;
  }
}

}  // namespace specmask
}  // namespace b2a

extern "C" int b2a_spec_band_mask_f32(float* spec, int64_t rows, int F, int N, const float* axis_vals, const float* lo,
                                      const float* hi, int rows_per_item, int axis, float fill_re, float fill_im,
                                      void* stream) {
  using namespace b2a::specmask;
  B2A_REQUIRE(spec && axis_vals && lo && hi, B2A_E_INVALID, "spec_band_mask: null pointer");
  B2A_REQUIRE(rows >= 1 && F >= 1 && N >= 1 && rows_per_item >= 1 && (axis == 0 || axis == 1), B2A_E_INVALID,
              "spec_band_mask: bad argument");
  B2A_REQUIRE(((uintptr_t)spec & 7) == 0, B2A_E_INVALID, "spec_band_mask: spectra must be 8-byte aligned");
  const long long total = (long long)rows * F * N;
  B2A_REQUIRE((long long)F * N < ((long long)1 << 31), B2A_E_UNSUPPORTED, "spec_band_mask: F*N too large");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)B2A_NUM_SMS * 32;
  if (blocks > cap) blocks = cap;
  B2A_LAUNCH(band_mask_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, reinterpret_cast<float2*>(spec), total, F, N,
             axis_vals, lo, hi, rows_per_item, axis, make_float2(fill_re, fill_im));
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
