// spectral_tc.cu -- the fused framing -> window -> real DFT -> |.| -> banded mel -> post-op kernel with the DFT's
// first (radix-128) stage on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).
//
// Same contract as spectral_warp_kernel<10,0> (spectral.cu): replaces torch.stft + abs + mel matmul + log of
// ref:audiotools/core/audio_signal.py:1195-1202,1355,1367 and ref:audiotools/metrics/spectral.py:187-190 for
// window_length 2048, and (optionally) the x*gain of EffectMixin.normalize (ref:audiotools/core/effects.py:219).
//
// Factorisation of the 2048-point real DFT, n = 16 a + b (a < 128, b < 16), k = c + 128 d (c < 128, d < 16):
//   X[c + 128 d] = sum_b W16^{bd} . W2048^{bc} . Y_b[c],   Y_b[c] = sum_a xw[16 a + b] W128^{ac}
// * Y_b[c] for c = 0..63 (the other half follows from xw being real) is ONE GEMM per group b on the tensor cores:
//   D[128 x 16 frames] = F[128 x 128] . XW_b[128 x 16 frames], rows r = 2c + {0: cos, 1: -sin}; row 1 (Im Y[0] = 0)
//   carries c = 64 instead ((-1)^a).  Operands are fp16 with an exact two-term split of BOTH sides
//   (x = h1 + h2, F = F1 + F2; products h1 F1 + h2 F1 + h1 F2, fp32 accumulation in TMEM): measured relative error
//   1.2e-7 (tests/probes/tc_probe.cu), i.e. fp32 quality; the samples of a tile are pre-scaled by a power of two so
//   that max |x| lands in [512, 1024) and F by 64, which keeps both correction terms out of the fp16 subnormals.
//   F lives in TENSOR MEMORY (A operand of tcgen05.mma "ts" form, written once per CTA with tcgen05.st); the
//   windowed frames are the B operand in shared memory (canonical no-swizzle K-major layout).
// * the second stage -- twiddle by W2048^{bc} and a 16-point complex DFT over b -- runs in registers, one thread per
//   (frame, c): lanes 2c and 2c+1 hold Re / Im of the same Y and swap one of two frames with a shuffle so that each
//   processes one whole frame.  The odd lane ends up with (Im, Re) = i conj(z): its DFT is i conj(X[-d]), whose
//   MAGNITUDE is that of bin -d -- so it conjugates its twiddles and mirrors its store index instead of un-swapping.
//   Outputs d >= 8 are the mirrored bins 2048 - k.  c = 0 / c = 64 (real inputs) pack the two frames of the pair
//   into one complex DFT and separate them with the usual even/odd split.
// * |X| of the 16 frames of a tile goes to shared memory (on top of the dead B operand), the banded FP32 mel
//   projection + post-op + coalesced tile store are those of spectral.cu.
//
// Per 16-frame tile: 1 TMA-staged span (19 bulk copies, padded per 512 samples -> conflict-free strided reads),
// 48 tcgen05.mma (M 128, N 128 = 8 groups x 16 frames, K 16).
#ifndef B2A_SIM
#include <cuda_fp16.h>
#endif

#include "b2a_common.h"
#include "fft_warp.cuh"
#include "spectral_internal.h"

#ifndef B2A_SIM
#include <cstdlib>
#endif

namespace b2a {
namespace spectral {

namespace tc {

constexpr int NFFT = 2048;
constexpr int FR = 16;        // frames per tile = N of the MMA
constexpr int NG = 16;        // groups b
constexpr int KA = 128;       // a per group = K of a group's GEMM
constexpr int THREADS = 512;
constexpr int NWARP = THREADS / 32;
constexpr int XBS = 1156;     // floats per |X| slot (1025 + band over-read slack; 1156 % 32 == 4: frames interleave)
constexpr int BLK = 512;      // span padding granule (samples)
constexpr int BLKP = BLK + 4; // padded granule
constexpr int B_PART = NG * 4096;  // bytes of one fp16 part (hi or lo) of the B operand
constexpr int TM_COLS = 512;
constexpr int TM_D = 0, TM_F1 = 256, TM_F2 = 320;

__host__ __device__ __forceinline__ int pad_idx(int i) { return i + 4 * (i >> 9); }

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM wrappers.  Under the CPU simulator (tests/cusim) tensor memory is a per-block array and the MMA
// a plain loop over the same shared-memory bytes, so that the kernel's indexing is checked before any GPU time.
// ---------------------------------------------------------------------------------------------
#ifdef B2A_SIM
static thread_local uint32_t g_tmem[128][TM_COLS];  // per simulator worker = per block in flight
static inline float h2f(uint16_t h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
static inline uint16_t f2h(float f) { _Float16 v = (_Float16)f; uint16_t h; memcpy(&h, &v, 2); return h; }
#endif

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
#ifdef B2A_SIM
  return (uint32_t)f2h(lo) | ((uint32_t)f2h(hi) << 16);
#else
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
#endif
}
__device__ __forceinline__ float f16lo_to_f32(uint32_t w) {
#ifdef B2A_SIM
  return h2f((uint16_t)(w & 0xffff));
#else
  return __half2float(__ushort_as_half((unsigned short)(w & 0xffff)));
#endif
}
__device__ __forceinline__ float f16hi_to_f32(uint32_t w) {
#ifdef B2A_SIM
  return h2f((uint16_t)(w >> 16));
#else
  return __half2float(__ushort_as_half((unsigned short)(w >> 16)));
#endif
}
// two-term fp16 split of a pair: hi = rn(v), lo = rn(v - hi)  (v - hi is exact in fp32)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2(v0, v1);
  lo = pack_f16x2(v0 - f16lo_to_f32(hi), v1 - f16hi_to_f32(hi));
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
#ifdef B2A_SIM
  *slot = 0;
#else
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(slot)), "r"(TM_COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_free(uint32_t base) {
#ifndef B2A_SIM
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(TM_COLS) : "memory");
#else
  (void)base;
#endif
}
__device__ __forceinline__ void tc_fence_before() {
#ifndef B2A_SIM
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_fence_after() {
#ifndef B2A_SIM
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_async_smem() {
#ifndef B2A_SIM
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_st16(uint32_t base, int lane_row, int col, const uint32_t (&w)[16]) {
#ifdef B2A_SIM
  (void)base;
  for (int j = 0; j < 16; ++j) g_tmem[lane_row][col + j] = w[j];
#else
  const uint32_t addr = base + (uint32_t)col + ((uint32_t)(lane_row & ~31) << 16);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
               "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
               "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
#endif
}
__device__ __forceinline__ void tmem_st_wait() {
#ifndef B2A_SIM
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_ld2(uint32_t base, int lane_row, int col, float& v0, float& v1) {
#ifdef B2A_SIM
  (void)base;
  memcpy(&v0, &g_tmem[lane_row][col], 4);
  memcpy(&v1, &g_tmem[lane_row][col + 1], 4);
#else
  const uint32_t addr = base + (uint32_t)col + ((uint32_t)(lane_row & ~31) << 16);
  uint32_t a, b;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
  v0 = __uint_as_float(a);
  v1 = __uint_as_float(b);
#endif
}
// tcgen05.ld is asynchronous: the registers it names may only be read after tcgen05.wait::ld.  The values are
// threaded THROUGH the wait ("+f") so that no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
#ifndef B2A_SIM
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                 "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
#else
  (void)v;
#endif
}

#ifndef B2A_SIM
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); layout type 0 = no swizzle
  return d;
}
#endif
// D[128 x NN] (+)= A[128 x 16] (TMEM columns a_col .. a_col + 8, two halves per column) . B[NN x 16]^T (shared
// memory, canonical no-swizzle K-major: 8-row x 16-byte core matrices, LBO = 128 B between the two K chunks, SBO =
// 2048 B between consecutive 8-row groups); fp16 in, fp32 accumulate.  B row n' = 16 (group) + frame: the groups'
// operand regions are 4096 B apart = two row groups, so ONE instruction with NN = 16 * (number of groups) multiplies
// the same F slice with every group's frames and lands in D columns 16 b + n -- measured (tests/probes/
// tc_latency_probe.cu): a tcgen05.mma costs ~60 cycles to issue whatever its N, so 48 wide instructions per tile
// replace 384 narrow ones.  One thread issues it.
template <int NN>
__device__ __forceinline__ void mma_ts(uint32_t base, int d_col, int a_col, const unsigned char* b_smem, int accumulate) {
#ifdef B2A_SIM
  (void)base;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < NN; ++n) {
      float acc = 0.f;
      if (accumulate) memcpy(&acc, &g_tmem[m][d_col + n], 4);
      for (int k = 0; k < 16; ++k) {
        const uint32_t aw = g_tmem[m][a_col + (k >> 1)];
        const float a = h2f((uint16_t)((k & 1) ? (aw >> 16) : (aw & 0xffff)));
        uint16_t bh;
        memcpy(&bh, b_smem + (n >> 3) * 2048 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2, 2);
        acc += a * h2f(bh);
      }
      memcpy(&g_tmem[m][d_col + n], &acc, 4);
    }
#else
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32
  const uint64_t db = smem_desc((uint32_t)__cvta_generic_to_shared(b_smem), 128, 2048);
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(base + (uint32_t)d_col),
               "r"(base + (uint32_t)a_col), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
#endif
}
// one lane of a converged warp (warp-uniform control flow around it: the tcgen05.mma it guards is issued once, with
// no per-lane serialisation loop)
__device__ __forceinline__ bool elect_one() {
#ifdef B2A_SIM
  return (threadIdx.x & 31) == 0;
#else
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
#endif
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
#ifndef B2A_SIM
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)) : "memory");
#else
  (void)bar;
#endif
}
// arm `bar` with the byte count of the whole span, then one bulk copy per 512-sample granule (padded destination)
__device__ __forceinline__ void tma_span(float* sp, const float* src, int span, unsigned long long* bar) {
#ifdef B2A_SIM
  for (int i = 0; i < span; ++i) sp[pad_idx(i)] = src[i];
  (void)bar;
#else
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)span * 4u) : "memory");
  for (int i = 0; i < span; i += BLK) {
    const unsigned bytes = (unsigned)min(BLK, span - i) * 4u;
    const unsigned d = (unsigned)__cvta_generic_to_shared(sp + pad_idx(i));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(src + i), "r"(bytes), "r"(b) : "memory");
  }
#endif
}

struct Smem {
  int off_b, off_span, off_win, off_tw, off_mel, off_mpk, off_mseg, off_red, total;
};
static Smem smem_layout(const Params& p) {
  Smem s;
  auto al = [](int v) { return (v + 127) & ~127; };
  int o = 0;
  s.off_b = o; o = al(o + 2 * B_PART);                       // B operand (hi, lo); later the |X| slots
  const int nblk = (p.span + BLK - 1) / BLK;
  s.off_span = o; o = al(o + nblk * BLKP * 4);
  s.off_win = o; o = al(o + NFFT * 4);
  s.off_tw = o; o = al(o + NG * 128 * 8);
  s.off_mel = o; o = al(o + p.n_mels * (FR + 1) * 4);
  s.off_mpk = o; o = al(o + p.mel_packed_len * 4);
  s.off_mseg = o; o = al(o + p.n_mels * 16);
  s.off_red = o; o = al(o + 256);
  s.total = o;
  return s;
}

struct KParams {
  Params p;
  Smem s;
};

__global__ void __launch_bounds__(THREADS, 1) spectral_tc_kernel(const B2A_GRID_CONSTANT KParams kp) {
  const Params& p = kp.p;
  B2A_DYN_SMEM(smem);
  unsigned char* bop = smem + kp.s.off_b;
  float* xs = reinterpret_cast<float*>(smem + kp.s.off_b);  // |X| slots alias the B operand (dead after the MMAs)
  float* sp = reinterpret_cast<float*>(smem + kp.s.off_span);
  float* wsc = reinterpret_cast<float*>(smem + kp.s.off_win);
  float2* tw = reinterpret_cast<float2*>(smem + kp.s.off_tw);  // [b][row]
  float* melt = reinterpret_cast<float*>(smem + kp.s.off_mel);
  float* mpk = reinterpret_cast<float*>(smem + kp.s.off_mpk);
  int4* mseg = reinterpret_cast<int4*>(smem + kp.s.off_mseg);
  float* red = reinterpret_cast<float*>(smem + kp.s.off_red);

  __shared__ __align__(8) unsigned long long s_bar_tma, s_bar_mma;
  __shared__ uint32_t s_tmem;
  __shared__ int s_clamp;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = warp & 3;             // TMEM lane quarter this warp may access
  const int row = 32 * q + lane;      // GEMM row = TMEM lane: r = 2c + ri
  const int c = row >> 1, ri = row & 1;
  const int hop = p.hop, F = NFFT / 2 + 1;
  const int total_tiles = p.rows * p.n_tiles;

  if (tid == 0) {
    mbar_init(&s_bar_tma, 1);
    mbar_init(&s_bar_mma, 1);
  }
  if (warp == 0) tmem_alloc(&s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmb = s_tmem;

  // ---- first tile's span in flight while the tables are built
  int t = blockIdx.x;
  bool by_tma = false;
  auto stage = [&](int tt) -> bool {
    const int rw = tt / p.n_tiles, tile = tt - rw * p.n_tiles;
    const int ws = (tile * FR + p.drop_edge) * hop + p.origin;
    const float* xr = p.x + (size_t)rw * (size_t)p.T;
    const bool interior = (ws >= 0) && (ws + p.span <= p.T);
    if (interior && ((((uintptr_t)(xr + ws)) & 15) == 0) && ((p.span & 3) == 0)) {
      if (tid == 0) tma_span(sp, xr + ws, p.span, &s_bar_tma);
      return true;
    }
    for (int i = tid; i < p.span; i += THREADS) {
      const int u = src_index(ws + i, p.T, p.pad, p.right_pad, p.pad_mode, p.center);
      sp[pad_idx(i)] = (u >= 0) ? __ldg(xr + u) : 0.f;
    }
    return false;
  };
  if (t < total_tiles) by_tma = stage(t);
  unsigned par_tma = 0, par_mma = 0;

  // ---- DFT-128 matrix rows into tensor memory: F[r][a] = 64 cos(2 pi a c / 128) (ri = 0), -64 sin(..) (ri = 1);
  //      row 1 = 64 (-1)^a (the c = 64 component).  Warp (q, w/4) writes columns a in [32 (w/4), +32).
  {
    const int a0 = 32 * (warp >> 2);
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = a0 + 2 * j + e;
        float sn, cs;
        sincospif((float)((a * c) & 127) * (1.0f / 64.0f), &sn, &cs);
        v[e] = 64.0f * (ri ? -sn : cs);
        if (row == 1) v[e] = (a & 1) ? -64.0f : 64.0f;
      }
      split2(v[0], v[1], h[j], l[j]);
    }
    tmem_st16(tmb, row, TM_F1 + a0 / 2, h);
    tmem_st16(tmb, row, TM_F2 + a0 / 2, l);
    tmem_st_wait();
  }
  // ---- second-stage twiddles W2048^{b c} (conjugated for the odd lane, which works on i conj(z));
  //      row 1 (c = 64 packed pair): W2048^{64 b}, not conjugated
  for (int i = tid; i < NG * 128; i += THREADS) {
    const int b = i >> 7, r = i & 127;
    const int cc = (r == 1) ? 64 : (r >> 1);
    float sn, cs;
    sincospif((float)(b * cc) * (1.0f / 1024.0f), &sn, &cs);
    tw[i] = make_float2(cs, ((r & 1) && r != 1) ? sn : -sn);
  }
  // ---- banded mel weights (same packing as spectral_warp_kernel: rows grouped by (w = m % 8, step i))
  const bool packed = p.mel_packed_len > 0;
  if (packed) {
    for (int m = tid; m < p.n_mels; m += THREADS) {
      const int lo4 = __ldg(p.mel_lo + m) & ~3;
      int n4 = (((__ldg(p.mel_hi + m) + 3) & ~3) - lo4) >> 2;
      mseg[m] = make_int4(0, lo4, n4 < 0 ? 0 : n4, 0);
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0, reach = 0;
      for (int w = 0; w < 8; ++w)
        for (int i = 0; 4 * (w + 8 * i) < p.n_mels; ++i) {
          int mx = 0;
          for (int j = 0; j < 4; ++j) { const int m = 4 * (w + 8 * i) + j; if (m < p.n_mels) mx = max(mx, mseg[m].z); }
          mx = (mx + 1) & ~1;
          for (int j = 0; j < 4; ++j) {
            const int m = 4 * (w + 8 * i) + j;
            if (m < p.n_mels) { mseg[m].x = run; mseg[m].w = mx; run += mx; reach = max(reach, mseg[m].y + 4 * mx); }
          }
        }
      s_clamp = reach > XBS;
    }
    __syncthreads();
    for (int m = warp; m < p.n_mels; m += NWARP) {
      const int4 sg = mseg[m];
      const float* wrow = p.mel_fb + (size_t)m * F;
      for (int i = lane; i < 4 * sg.w; i += 32) {
        const int k = sg.y + i;
        mpk[4 * sg.x + i] = (i < 4 * sg.z && k < F) ? __ldg(wrow + k) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

#pragma unroll 1
  for (; t < total_tiles; t += gridDim.x) {
    const int rw = t / p.n_tiles, tile = t - rw * p.n_tiles;
    const int n0 = tile * FR;
    const int ws = (n0 + p.drop_edge) * hop + p.origin;
    const float g = p.gain ? __ldg(p.gain + rw / p.rows_per_gain) : 1.0f;
    if (by_tma) { mbar_wait(&s_bar_tma, par_tma); par_tma ^= 1u; }
    __syncthreads();

    // ---- (1) power-of-two scale of the tile: max |x| S in [512, 1024)
    float mx = 0.f;
    for (int j = tid; j < (p.span >> 2); j += THREADS) {
      const float4 v = *reinterpret_cast<const float4*>(sp + pad_idx(4 * j));
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int i = (p.span & ~3) + tid; i < p.span; i += THREADS) mx = fmaxf(mx, fabsf(sp[pad_idx(i)]));
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) mx = fmaxf(mx, red[w]);
    float S = 1.0f;
    {
      const int E = (int)((__float_as_uint(mx) >> 23) & 0xff);  // mx = m 2^(E-126), m in [0.5, 1)
      if (E >= 27 && E <= 230) S = __uint_as_float((uint32_t)(263 - E) << 23);  // 2^(10 - (E - 126))
    }
    const float inv_scale = 1.0f / (S * 64.0f);
    // ---- (2) window x S
    for (int i = tid * 4; i < NFFT; i += THREADS * 4) {
      float4 w = __ldg(reinterpret_cast<const float4*>(p.window + i));
      w.x *= S; w.y *= S; w.z *= S; w.w *= S;
      *reinterpret_cast<float4*>(wsc + i) = w;
    }
    __syncthreads();

    // ---- (3) + (4): windowed frames -> fp16 (hi, lo) B operand, in two halves of 8 groups; the MMAs of a half are
    //      issued (warp 0, one elected lane) as soon as the half is in shared memory, so the tensor core works on groups
    //      0-7 while the CUDA cores convert groups 8-15.  Unit u (64 per half) = (nh, ac, gq'): frames 8 nh + (lane & 7),
    //      a in [8 ac, +8), groups b = 4 (2 half + gq') + e.  Sample of (n, a, b) = span[n hop + 16 a + b].
    const bool blk_aligned = (hop & (BLK - 1)) == 0;  // every frame starts on a padding granule
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      {
        const int u = 4 * warp + (lane >> 3);
        const int nh = u >> 5, ac = (u >> 1) & 15, gq = 2 * half + (u & 1);
        const int n = 8 * nh + (lane & 7);
        const int o0 = 128 * ac + 4 * gq;
        const int i0 = n * hop + o0;
        // all 8 loads of a unit fall into one 512-sample granule when the frames start on granule boundaries
        const float* sbase = sp + (blk_aligned ? pad_idx(i0) : 0);
        uint32_t hw[4][4], lw[4][4];  // [e = b - 4 gq][pair of a]
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          float4 pr[2];
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int j = 2 * jp + e2;
            float4 xv;
            if (blk_aligned) {
              xv = *reinterpret_cast<const float4*>(sbase + 16 * j);
            } else if ((hop & 3) == 0) {
              xv = *reinterpret_cast<const float4*>(sp + pad_idx(i0 + 16 * j));
            } else {
              const int i = i0 + 16 * j;
              xv = make_float4(sp[pad_idx(i)], sp[pad_idx(i + 1)], sp[pad_idx(i + 2)], sp[pad_idx(i + 3)]);
            }
            const float4 wv = *reinterpret_cast<const float4*>(wsc + o0 + 16 * j);
            pr[e2] = make_float4(xv.x * wv.x, xv.y * wv.y, xv.z * wv.z, xv.w * wv.w);
          }
          split2(pr[0].x, pr[1].x, hw[0][jp], lw[0][jp]);
          split2(pr[0].y, pr[1].y, hw[1][jp], lw[1][jp]);
          split2(pr[0].z, pr[1].z, hw[2][jp], lw[2][jp]);
          split2(pr[0].w, pr[1].w, hw[3][jp], lw[3][jp]);
        }
        unsigned char* dst = bop + (4 * gq) * 4096 + nh * 2048 + ac * 128 + (lane & 7) * 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          *reinterpret_cast<int4*>(dst + e * 4096) = make_int4((int)hw[e][0], (int)hw[e][1], (int)hw[e][2], (int)hw[e][3]);
          *reinterpret_cast<int4*>(dst + B_PART + e * 4096) = make_int4((int)lw[e][0], (int)lw[e][1], (int)lw[e][2], (int)lw[e][3]);
        }
      }
      fence_async_smem();  // generic-proxy writes of the operand -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        // 3 products x 8 k-steps, each ONE instruction over the 8 groups of the half (N = 128 = 8 groups x 16 frames)
        tc_fence_after();
        if (elect_one()) {
          const unsigned char* bh0 = bop + (8 * half) * 4096;
#pragma unroll
          for (int prod = 0; prod < 3; ++prod) {
            const int a_col = (prod == 2) ? TM_F2 : TM_F1;
            const unsigned char* bp = bh0 + (prod == 1 ? B_PART : 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_ts<8 * FR>(tmb, TM_D + FR * 8 * half, a_col + 8 * ks, bp + ks * 256, (prod | ks) != 0);
          }
          if (half == 1) mma_commit(&s_bar_mma);
        }
        __syncwarp();
      }
    }
    // everybody but the issuing warp streams the scaled waveform out while the tensor core works
    if (p.y_out && warp != 0) {  // y = g x for the samples this tile owns ([n0 hop, (n0 + FR) hop), the last tile up to T)
      const int wt = tid - 32, WT = THREADS - 32;
      const int own_lo = n0 * hop;
      const int own_hi = (tile == p.n_tiles - 1) ? p.T : min(p.T, (n0 + FR) * hop);
      float* yr = p.y_out + (size_t)rw * (size_t)p.T;
      const int lo = max(own_lo, ws), hi = min(own_hi, ws + p.span);
      const bool vec = (((lo - ws) & 3) == 0) && ((((uintptr_t)(yr + lo)) & 15) == 0);
      if (vec) {
        const int n4 = (hi - lo) >> 2;
        for (int i = wt; i < n4; i += WT) {
          float4 v = *reinterpret_cast<const float4*>(sp + pad_idx(lo - ws + 4 * i));
          v.x *= g; v.y *= g; v.z *= g; v.w *= g;
          st_stream4(yr + lo + 4 * i, v);
        }
        for (int w = lo + 4 * n4 + wt; w < hi; w += WT) yr[w] = sp[pad_idx(w - ws)] * g;
      } else {
        for (int w = lo + wt; w < hi; w += WT) yr[w] = sp[pad_idx(w - ws)] * g;
      }
      const float* xr = p.x + (size_t)rw * (size_t)p.T;
      for (int w = max(own_lo, ws + p.span) + wt; w < own_hi; w += WT) yr[w] = __ldg(xr + w) * g;
    }
    __syncthreads();  // the span is dead: the next tile's samples stream in underneath the MMAs and the epilogue
    {
      const int tn = t + gridDim.x;
      by_tma = (tn < total_tiles) ? stage(tn) : false;
    }
    mbar_wait(&s_bar_mma, par_mma);
    par_mma ^= 1u;
    tc_fence_after();

    // ---- (5) second stage in registers: thread (row, frame pair) -> one frame's 16 bins k = +-c + 128 d
#pragma unroll 1
    for (int pp = (warp >> 2); pp < FR / 2; pp += NWARP / 4) {
      float v0[NG], v1[NG];
#pragma unroll
      for (int b = 0; b < NG; ++b) tmem_ld2(tmb, row, TM_D + FR * b + 2 * pp, v0[b], v1[b]);
      tmem_ld_wait(v0);
      tmem_ld_wait(v1);
      float2 z[NG];
#pragma unroll
      for (int b = 0; b < NG; ++b) {
        const float mine = ri ? v1[b] : v0[b];
        const float other = ri ? v0[b] : v1[b];
        const float recv = __shfl_xor_sync(0xffffffffu, other, 1);
        z[b] = make_float2(mine, recv);
        if (c == 0) z[b].y = other;  // c = 0 / 64: both frames of the pair in one complex DFT (lanes 0, 1 of 4 warps)
      }
#pragma unroll
      for (int b = 1; b < NG; ++b) z[b] = cmul(z[b], tw[b * 128 + row]);
      float2 o[NG];
      DFT<NG, 1>::run(z, o);
      const int f = 2 * pp + ri;  // even lane: frame 2 pp, odd lane: frame 2 pp + 1
      if (c != 0) {
        float* xf = xs + f * XBS;
        float* P1 = ri ? xf - c : xf + c;  // d = 1..7  -> +-c + 128 d
        float* P2 = ri ? xf + c : xf - c;  // d = 9..15 -> -+c + 128 (16 - d)
        xf[c] = fast_sqrt(fmaf(o[0].x, o[0].x, o[0].y * o[0].y));
#pragma unroll
        for (int d = 1; d < 8; ++d) P1[128 * d] = fast_sqrt(fmaf(o[d].x, o[d].x, o[d].y * o[d].y));
        xf[1024 - c] = fast_sqrt(fmaf(o[8].x, o[8].x, o[8].y * o[8].y));
#pragma unroll
        for (int d = 9; d < 16; ++d) P2[128 * (16 - d)] = fast_sqrt(fmaf(o[d].x, o[d].x, o[d].y * o[d].y));
      } else {
        // Y = DFT(u + i v) of two real-input problems u, v (the two frames of the pair):
        //   row 0 (c = 0):  U[d] = (Y[d] + conj Y[16-d]) / 2, V[d] = (Y[d] - conj Y[16-d]) / 2i  -> bins 128 d, d = 0..8,
        //                   u = frame 2 pp, v = frame 2 pp + 1
        //   row 1 (c = 64): partner index 15 - d, bins 64 + 128 d, d = 0..7, u = frame 2 pp + 1 (mine), v = frame 2 pp
        float* xu = xs + (2 * pp + ri) * XBS;
        float* xv = xs + (2 * pp + 1 - ri) * XBS;
        const int koff = ri ? 64 : 0;
#pragma unroll
        for (int d = 0; d < 9; ++d) {  // (static register indices: the partner is selected, not indexed)
          const float2 yd = o[d];
          const float2 ya = o[(16 - d) & 15], yb = o[(15 - d) & 15];
          const float2 yn = ri ? yb : ya;
          const float2 U = make_float2(0.5f * (yd.x + yn.x), 0.5f * (yd.y - yn.y));
          const float2 V = make_float2(0.5f * (yd.y + yn.y), 0.5f * (yn.x - yd.x));
          if (d < 8 || !ri) {
            xu[koff + 128 * d] = sqrtf(fmaf(U.x, U.x, U.y * U.y));
            xv[koff + 128 * d] = sqrtf(fmaf(V.x, V.x, V.y * V.y));
          }
        }
      }
    }
    if (tid < FR) {  // band rows read 4 wide: keep the three floats behind bin 1024 finite (x 0 weight)
      float* xf = xs + tid * XBS;
      xf[1025] = 0.f; xf[1026] = 0.f; xf[1027] = 0.f;
    }
    tc_fence_before();
    __syncthreads();

    // ---- (6) banded mel projection + post-op: warps 0-7 take frames 0-7, warps 8-15 frames 8-15
    {
      const int fl = lane & 7, jq = lane >> 3, w8 = warp & 7;
      const int f = 8 * (warp >> 3) + fl;
      const float lscale = p.post_power * 0.30102999566398120f;
      const float ga = fabsf(g) * inv_scale;  // |X| is stored in the tile's scaled units: undo S and the 64 of F here
      const float* xf = xs + f * XBS;
      if (packed) {
        const float4* mpk4 = reinterpret_cast<const float4*>(mpk);
        const int lim = XBS - 4;
        for (int mm = 4 * w8 + jq; mm < p.n_mels; mm += 32) {
          const int4 sg = mseg[mm];
          const float4* w4 = mpk4 + sg.x;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          if (!s_clamp) {
            const float4* v4 = reinterpret_cast<const float4*>(xf + sg.y);
            for (int it = 0; it < sg.w; it += 2) {
              const float4 wa = w4[it], wb = w4[it + 1], va = v4[it], vb = v4[it + 1];
              a0 = fmaf(wa.x, va.x, a0); a1 = fmaf(wa.y, va.y, a1);
              a2 = fmaf(wb.x, vb.x, a2); a3 = fmaf(wb.y, vb.y, a3);
              a0 = fmaf(wa.z, va.z, a0); a1 = fmaf(wa.w, va.w, a1);
              a2 = fmaf(wb.z, vb.z, a2); a3 = fmaf(wb.w, vb.w, a3);
            }
          } else {
            for (int it = 0; it < sg.w; ++it) {
              const float4 w = w4[it];
              const float4 v = *reinterpret_cast<const float4*>(xf + min(sg.y + 4 * it, lim));
              a0 = fmaf(w.x, v.x, a0); a1 = fmaf(w.y, v.y, a1);
              a0 = fmaf(w.z, v.z, a0); a1 = fmaf(w.w, v.w, a1);
            }
          }
          float acc = ((a0 + a1) + (a2 + a3)) * ga;
          if (p.post == B2A_POST_LOG10) acc = lscale * fast_log2(fmaxf(acc, p.post_eps));
          else if (p.post == B2A_POST_LN) acc = logf(acc + p.post_eps);
          melt[mm * (FR + 1) + f] = acc;
        }
      } else {
        for (int mm = 4 * w8 + jq; mm < p.n_mels; mm += 32) {
          const int lo = __ldg(p.mel_lo + mm), hi = __ldg(p.mel_hi + mm);
          const float* wrow = p.mel_fb + (size_t)mm * F;
          float acc = 0.f;
          for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(wrow + k), xf[k], acc);
          acc *= ga;
          if (p.post == B2A_POST_LOG10) acc = lscale * fast_log2(fmaxf(acc, p.post_eps));
          else if (p.post == B2A_POST_LN) acc = logf(acc + p.post_eps);
          melt[mm * (FR + 1) + f] = acc;
        }
      }
    }
    __syncthreads();
    {
      const int nf = min(FR, p.n_frames - n0);
      float* o = p.mel_out + (size_t)rw * p.n_mels * p.n_frames + n0;
      for (int i = tid; i < p.n_mels * FR; i += THREADS) {
        const int m = i / FR, f = i - m * FR;
        if (f < nf) o[(size_t)m * p.n_frames + f] = melt[m * (FR + 1) + f];
      }
    }
    // (the barrier at the top of the next iteration orders these reads of melt / xs before they are rewritten)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tmb);
}

}  // namespace tc

static int g_tc_enabled = -1;
static int tc_enabled() {
  if (g_tc_enabled < 0) {
    // Opt-in: measured on B200 (profiles/README.md, round 2) this kernel runs 64 x 2ch x 10 s in 0.62 ms against 0.45 ms
    // for the FP32 warp kernel, so the FP32 kernel stays the default; B2A_SPECTRAL_TC=1 or b2a_spectral_tc_enable(1)
    // selects the tensor-core path.
    const char* e = getenv("B2A_SPECTRAL_TC");
    g_tc_enabled = (e && e[0] == '1') ? 1 : 0;
  }
  return g_tc_enabled;
}

static int tc_num_sms() {
  int n = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = B2A_NUM_SMS;
  return n;
}

bool tc_supported(const Params& p) {
  if (!tc_enabled()) return false;
  if (p.n_fft != tc::NFFT || !p.mel_out || p.stft_out) return false;
  if (p.hop < 1 || p.hop > 512) return false;
  if (!p.center || p.row_origin) return false;
  Params q = p;
  q.span = (tc::FR - 1) * q.hop + q.n_fft;
  tc::Smem s = tc::smem_layout(q);
  if (s.total > 227 * 1024 - 64) {
    q.mel_packed_len = 0;  // band table from global
    s = tc::smem_layout(q);
    if (s.total > 227 * 1024 - 64) return false;
  }
  return true;
}

int launch_tc(Params& p, void* stream) {
  p.span = (tc::FR - 1) * p.hop + p.n_fft;
  p.n_tiles = (p.n_frames + tc::FR - 1) / tc::FR;
  tc::KParams kp;
  kp.s = tc::smem_layout(p);
  if (kp.s.total > 227 * 1024 - 64) {
    p.mel_packed_len = 0;
    kp.s = tc::smem_layout(p);
  }
  B2A_REQUIRE(kp.s.total <= 227 * 1024 - 64, B2A_E_UNSUPPORTED, "spectral_tc: %d bytes of shared memory", kp.s.total);
  p.smem_bytes = kp.s.total;
  kp.p = p;
  const int64_t total = (int64_t)p.rows * p.n_tiles;
  B2A_REQUIRE(total < (int64_t)2147483647, B2A_E_UNSUPPORTED, "spectral_tc: too many tiles");
  B2A_CUDA_OK(cudaFuncSetAttribute(tc::spectral_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.s.total));
  const int64_t cap = tc_num_sms();
  const unsigned grid = (unsigned)(total < cap ? total : cap);
  B2A_LAUNCH(tc::spectral_tc_kernel, dim3(grid), dim3(tc::THREADS), (size_t)kp.s.total, stream, kp);
  B2A_CUDA_OK(cudaGetLastError());
  return B2A_OK;
}

}  // namespace spectral
}  // namespace b2a

extern "C" int b2a_spectral_uses_tensor_cores(int n_fft, int hop, int want_mel, int want_stft) {
  b2a::spectral::Params p;
  memset(&p, 0, sizeof(p));
  p.n_fft = n_fft; p.hop = hop; p.center = 1; p.n_mels = 128; p.mel_packed_len = 0;
  p.mel_out = want_mel ? reinterpret_cast<float*>(8) : nullptr;
  p.stft_out = want_stft ? reinterpret_cast<float2*>(8) : nullptr;
  return b2a::spectral::tc_supported(p) ? 1 : 0;
}

extern "C" int b2a_spectral_tc_enable(int on) {
  const int prev = b2a::spectral::tc_enabled();
  b2a::spectral::g_tc_enabled = on ? 1 : 0;
  return prev;
}
