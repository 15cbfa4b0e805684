// Mel-spectrogram conditioner (model/condition.py:7-19): ReflectionPad1d((n_fft/2 - hop/2, n_fft/2 + hop/2)) ->
// windowed STFT (center=False) -> |.|^power -> mel filterbank -> log(. + eps), fused in ONE kernel that writes the
// (B, n_mels, frames) tensor the conditioning upsampler (upsample.cu) consumes.  Nothing but the audio is read from
// HBM and nothing but the log-mel frames is written: the padded signal, the (B, n_fft/2+1, frames) complex STFT and the
// power spectrogram of the reference's torchaudio pipeline never exist.
//
// One CTA per (batch, frame).  The n_fft real samples are packed as n_fft/2 complex numbers z[n] = x[2n] + i x[2n+1]
// (window applied, reflected indices resolved while loading), transformed by an in-shared-memory radix-2 FFT of half
// the length, and split back into the n_fft/2+1 one-sided bins.  A thread per mel band then runs over that band's
// support [lo, hi) of the triangular filter.  The real workloads are small (24 x 63 training frames, 863 frames per
// 10 s utterance), so the kernel is latency bound and a CTA per frame -- the widest decomposition -- is the fast one.
#include "common.cuh"

namespace cmwg {

__device__ __forceinline__ int reflect_index(int i, int T) {
  // ReflectionPad1d semantics (edge sample not repeated); valid for pads < T
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

template <int NFFT>
__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ x, long long x_bstride, int T,
                                                      const float* __restrict__ window,
                                                      const float* __restrict__ fbt, const int* __restrict__ fb_lo,
                                                      const int* __restrict__ fb_hi, int n_mels, int hop,
                                                      int pad_left, int frames, int power_is_one, float eps,
                                                      int take_log, float* __restrict__ out) {
  constexpr int M = NFFT / 2;  // complex FFT length
  constexpr int LOGM = (M == 64) ? 6 : (M == 128) ? 7 : (M == 256) ? 8 : (M == 512) ? 9 : (M == 1024) ? 10 : 11;
  constexpr int NT = 256;
  // one pad slot per 16 entries: the bit-reversed scatter (stride M/2) and the strided twiddle reads of the late
  // stages (stride NFFT >> (s+1)) would otherwise hit a single bank 32 ways
  constexpr int MP = M + M / 16;
#define CMWG_PADI(i) ((i) + ((i) >> 4))
  __shared__ float2 z[MP];
  __shared__ float2 tw[MP];      // tw[PAD(k)] = exp(-2 pi i k / NFFT), k < NFFT/2
  __shared__ float pw[M + 1];    // |X[k]|^power, k <= NFFT/2

  const int b = blockIdx.x / frames, fr = blockIdx.x % frames;
  const float* xb = x + (long long)b * x_bstride;
  const int start = fr * hop - pad_left;
  int my_lo = 0, my_hi = 0;
  if (threadIdx.x < n_mels) {
    my_lo = fb_lo[threadIdx.x];
    my_hi = fb_hi[threadIdx.x];
  }

  for (int k = threadIdx.x; k < M; k += NT) {
    float s, c;
    sincospif(-2.f * (float)k / (float)NFFT, &s, &c);
    tw[CMWG_PADI(k)] = make_float2(c, s);
  }
  // bit-reversed scatter of the packed, windowed frame
  for (int n = threadIdx.x; n < M; n += NT) {
    int i0 = reflect_index(start + 2 * n, T), i1 = reflect_index(start + 2 * n + 1, T);
    float a = xb[i0] * window[2 * n], c = xb[i1] * window[2 * n + 1];
    int r = (int)(__brev((unsigned)n) >> (32 - LOGM));
    z[CMWG_PADI(r)] = make_float2(a, c);
  }
  __syncthreads();
  // radix-2 decimation-in-time stages on M points; twiddle exp(-2 pi i pos / (2 half)) = tw[pos * (NFFT / (2 half))]
#pragma unroll 1
  for (int s = 0; s < LOGM; ++s) {
    const int half = 1 << s;
    const int tstep = NFFT >> (s + 1);
    for (int j = threadIdx.x; j < M / 2; j += NT) {
      int pos = j & (half - 1);
      int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      float2 w = tw[CMWG_PADI(pos * tstep)];
      i0 = CMWG_PADI(i0);
      i1 = CMWG_PADI(i1);
      float2 u = z[i0], v = z[i1];
      float2 t = make_float2(fmaf(v.x, w.x, -v.y * w.y), fmaf(v.x, w.y, v.y * w.x));
      z[i0] = make_float2(u.x + t.x, u.y + t.y);
      z[i1] = make_float2(u.x - t.x, u.y - t.y);
    }
    __syncthreads();
  }
  // split: X[k] = (Z[k] + conj Z[M-k]) / 2 - (i/2) exp(-2 pi i k / NFFT) (Z[k] - conj Z[M-k]),  k = 0..M  (Z[M] = Z[0])
  for (int k = threadIdx.x; k <= M; k += NT) {
    float2 a = z[CMWG_PADI(k & (M - 1))], c = z[CMWG_PADI((M - k) & (M - 1))];
    float er = 0.5f * (a.x + c.x), ei = 0.5f * (a.y - c.y);   // even part
    float dr = 0.5f * (a.x - c.x), di = 0.5f * (a.y + c.y);   // (Z[k] - conj Z[M-k]) / 2
    float2 w = (k < M) ? tw[CMWG_PADI(k)] : make_float2(-1.f, 0.f);
    // -i * w * d
    float pr = w.x * dr - w.y * di, pi = w.x * di + w.y * dr;
    float xr = er + pi, xi = ei - pr;
    float p = fmaf(xr, xr, xi * xi);
    pw[k] = power_is_one ? sqrtf(p) : p;
  }
  __syncthreads();
  // a thread per mel band: its [lo, hi) bounds were fetched before the FFT, and the loop's loads are independent of
  // each other, so the band sums cost one L2 round trip instead of one per band
  for (int m = threadIdx.x, i = 0; m < n_mels; m += NT, ++i) {
    int lo = i == 0 ? my_lo : fb_lo[m], hi = i == 0 ? my_hi : fb_hi[m];
    const float* frow = fbt + (long long)m * (M + 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int f = lo;
    for (; f + 4 <= hi; f += 4) {
      a0 = fmaf(pw[f], frow[f], a0);
      a1 = fmaf(pw[f + 1], frow[f + 1], a1);
      a2 = fmaf(pw[f + 2], frow[f + 2], a2);
      a3 = fmaf(pw[f + 3], frow[f + 3], a3);
    }
    for (; f < hi; ++f) a0 = fmaf(pw[f], frow[f], a0);
    float v = (a0 + a1) + (a2 + a3) + eps;
    out[((long long)b * n_mels + m) * frames + fr] = take_log ? logf(v) : v;
  }
#undef CMWG_PADI
}

}  // namespace cmwg

using namespace cmwg;

extern "C" int cmwg_melspec_frames(int T, int n_fft, int hop) {
  // reflect pad adds n_fft samples in total; center=False framing
  if (T <= 0 || hop <= 0) return 0;
  return T / hop + 1;
}

extern "C" int cmwg_melspec_fwd(const float* x, long long x_bstride, int B, int T, const float* window,
                                const float* fbt, const int* fb_lo, const int* fb_hi, int n_fft, int hop, int n_mels,
                                int power_is_one, float eps, int take_log, float* out, void* stream) {
  CMWG_REQUIRE(B >= 0 && T >= 0 && n_mels > 0 && hop > 0, "melspec: bad sizes B=%d T=%d n_mels=%d hop=%d", B, T,
               n_mels, hop);
  if (B == 0 || T == 0) return CMWG_OK;
  const int pad_left = n_fft / 2 - hop / 2, pad_right = n_fft / 2 + hop / 2;
  CMWG_REQUIRE(pad_right < T, "melspec: reflection padding (%d, %d) needs more than %d samples, got T=%d", pad_left,
               pad_right, pad_right, T);
  const int frames = (T + pad_left + pad_right - n_fft) / hop + 1;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)(B * frames)), block(256);
#define CMWG_MEL_CASE(N)                                                                                        \
  case N:                                                                                                       \
    melspec_kernel<N><<<grid, block, 0, st>>>(x, x_bstride, T, window, fbt, fb_lo, fb_hi, n_mels, hop, pad_left, \
                                              frames, power_is_one, eps, take_log, out);                       \
    break;
  switch (n_fft) {
    CMWG_MEL_CASE(128)
    CMWG_MEL_CASE(256)
    CMWG_MEL_CASE(512)
    CMWG_MEL_CASE(1024)
    CMWG_MEL_CASE(2048)
    CMWG_MEL_CASE(4096)
    default:
      set_error("melspec: n_fft must be a power of two in [128, 4096], got %d", n_fft);
      return CMWG_ERR_UNSUPPORTED;
  }
#undef CMWG_MEL_CASE
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}
