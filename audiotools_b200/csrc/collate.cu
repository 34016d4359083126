// collate.cu -- the caller side of the hot path on the device: ragged signals -> one padded /
// truncated [N, C, T_out] batch in ONE launch.
//
// Replaces the per-signal zero_pad / truncate_samples + torch.cat of AudioSignal.batch
// (ref:audiotools/core/audio_signal.py:380-470), which util.collate calls for every AudioSignal entry of a list of
// samples (ref:audiotools/core/util.py:426-479), and the excerpt gathering of the loudness-screened
// AudioSignal.salient_excerpt (ref :227-286): N candidate windows of one long signal are N (pointer, length) pairs.
// HBM-bound: every output sample is written once, every kept input sample is read once (16 B vectors when the source
// row, the destination row and T_out allow it).
#include "b2a_common.h"

namespace b2a {
namespace collate {

constexpr int TPB = 256;

// item i = C rows of src_len[i] samples at src[i] (row r at src[i] + r * src_stride[i]); out[i, r, t] = the sample
// t + src_off[i] of that row when 0 <= t + src_off[i] < src_len[i] and t < keep[i], else 0
__global__ void __launch_bounds__(TPB) pack_rows_kernel(const float* const* __restrict__ src,
                                                        const int64_t* __restrict__ src_len,
                                                        const int64_t* __restrict__ src_stride,
                                                        const int64_t* __restrict__ src_off, int C, int64_t T_out,
                                                        float* __restrict__ out) {
  const int row = blockIdx.y;  // item * C + channel
  const int item = row / C, ch = row - item * C;
  const float* s = src[item] + (size_t)ch * (size_t)src_stride[item];
  const int64_t len = src_len[item], off = src_off ? src_off[item] : 0;
  float* o = out + (size_t)row * (size_t)T_out;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)(s + off)) | ((uintptr_t)o)) % 16 == 0) && (T_out % 4 == 0) && off >= 0;
  if (vec) {
    const int64_t n4 = T_out >> 2;
    for (int64_t i = gid; i < n4; i += nt) {
      const int64_t t = 4 * i;
      float4 v;
      if (t + off + 3 < len) {
        v = ld_stream4(s + off + t);
      } else {
        v.x = (t + off < len) ? s[off + t] : 0.f;
        v.y = (t + off + 1 < len) ? s[off + t + 1] : 0.f;
        v.z = (t + off + 2 < len) ? s[off + t + 2] : 0.f;
        v.w = 0.f;
      }
      st_stream4(o + t, v);
    }
  } else {
    for (int64_t t = gid; t < T_out; t += nt) {
      const int64_t u = t + off;
      o[t] = (u >= 0 && u < len) ? __ldg(s + u) : 0.f;
    }
  }
}

}  // namespace collate
}  // namespace b2a

extern "C" int b2a_pack_rows_f32(const float* const* src_ptrs, const int64_t* src_len, const int64_t* src_stride,
                                 const int64_t* src_off, int64_t n_items, int C, int64_t T_out, float* out,
                                 void* stream) {
  B2A_REQUIRE(src_ptrs && src_len && src_stride && out, B2A_E_INVALID, "pack_rows: null pointer");
  B2A_REQUIRE(n_items >= 1 && C >= 1 && T_out >= 1 && n_items * C <= 65535, B2A_E_INVALID, "pack_rows: bad shape");
  const int64_t rows = n_items * C;
  const int64_t want = (T_out / 4 + b2a::collate::TPB - 1) / b2a::collate::TPB + 1;
  int64_t cap = (int64_t)B2A_NUM_SMS * 8 / rows + 1;
  const unsigned gx = (unsigned)(want < cap ? want : cap);
  B2A_LAUNCH(b2a::collate::pack_rows_kernel, dim3(gx, (unsigned)rows), dim3(b2a::collate::TPB), 0, stream, src_ptrs, src_len,
             src_stride, src_off, C, T_out, out);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}
