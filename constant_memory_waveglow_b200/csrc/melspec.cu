// Mel-spectrogram conditioner (model/condition.py:7-19): ReflectionPad1d((n_fft/2 - hop/2, n_fft/2 + hop/2)) ->
// windowed STFT (center=False) -> |.|^power -> mel filterbank -> log(. + eps), fused in ONE kernel that writes the
// (B, n_mels, frames) tensor the conditioning upsampler (upsample.cu) consumes.  Nothing but the audio is read from
// HBM and nothing but the log-mel frames is written: the padded signal, the (B, n_fft/2+1, frames) complex STFT and the
// power spectrogram of the reference's torchaudio pipeline never exist.
//
// One CTA per (batch, frame).  The n_fft real samples are packed as n_fft/2 complex numbers z[n] = x[2n] + i x[2n+1]
// (window applied, reflected indices resolved while loading), transformed by an in-shared-memory radix-2 FFT of half
// the length, and split back into the n_fft/2+1 one-sided bins.  A warp per mel band then runs over that band's
// support [lo, hi) of the triangular filter.
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
                                                      const float* __restrict__ fb, const int* __restrict__ fb_lo,
                                                      const int* __restrict__ fb_hi, int n_mels, int hop,
                                                      int pad_left, int frames, int power_is_one, float eps,
                                                      int take_log, float* __restrict__ out) {
  constexpr int M = NFFT / 2;  // complex FFT length
  constexpr int LOGM = (M == 64) ? 6 : (M == 128) ? 7 : (M == 256) ? 8 : (M == 512) ? 9 : (M == 1024) ? 10 : 11;
  constexpr int NT = 256;
  __shared__ float2 z[M];
  __shared__ float2 tw[M];       // tw[k] = exp(-2 pi i k / NFFT), k < NFFT/2
  __shared__ float pw[M + 1];    // |X[k]|^power, k <= NFFT/2

  const int b = blockIdx.x / frames, fr = blockIdx.x % frames;
  const float* xb = x + (long long)b * x_bstride;
  const int start = fr * hop - pad_left;

  for (int k = threadIdx.x; k < M; k += NT) {
    float s, c;
    sincospif(-2.f * (float)k / (float)NFFT, &s, &c);
    tw[k] = make_float2(c, s);
  }
  // bit-reversed scatter of the packed, windowed frame
  for (int n = threadIdx.x; n < M; n += NT) {
    int i0 = reflect_index(start + 2 * n, T), i1 = reflect_index(start + 2 * n + 1, T);
    float a = xb[i0] * window[2 * n], c = xb[i1] * window[2 * n + 1];
    int r = (int)(__brev((unsigned)n) >> (32 - LOGM));
    z[r] = make_float2(a, c);
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
      float2 w = tw[pos * tstep];
      float2 u = z[i0], v = z[i1];
      float2 t = make_float2(fmaf(v.x, w.x, -v.y * w.y), fmaf(v.x, w.y, v.y * w.x));
      z[i0] = make_float2(u.x + t.x, u.y + t.y);
      z[i1] = make_float2(u.x - t.x, u.y - t.y);
    }
    __syncthreads();
  }
  // split: X[k] = (Z[k] + conj Z[M-k]) / 2 - (i/2) exp(-2 pi i k / NFFT) (Z[k] - conj Z[M-k]),  k = 0..M  (Z[M] = Z[0])
  for (int k = threadIdx.x; k <= M; k += NT) {
    float2 a = z[k & (M - 1)], c = z[(M - k) & (M - 1)];
    float er = 0.5f * (a.x + c.x), ei = 0.5f * (a.y - c.y);   // even part
    float dr = 0.5f * (a.x - c.x), di = 0.5f * (a.y + c.y);   // (Z[k] - conj Z[M-k]) / 2
    float2 w = (k < M) ? tw[k] : make_float2(-1.f, 0.f);
    // -i * w * d
    float pr = w.x * dr - w.y * di, pi = w.x * di + w.y * dr;
    float xr = er + pi, xi = ei - pr;
    float p = fmaf(xr, xr, xi * xi);
    pw[k] = power_is_one ? sqrtf(p) : p;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < n_mels; m += NT / 32) {
    int lo = fb_lo[m], hi = fb_hi[m];
    float acc = 0.f;
    for (int f = lo + lane; f < hi; f += 32) acc = fmaf(pw[f], fb[(long long)f * n_mels + m], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + eps;
      out[((long long)b * n_mels + m) * frames + fr] = take_log ? logf(v) : v;
    }
  }
}

}  // namespace cmwg

using namespace cmwg;

extern "C" int cmwg_melspec_frames(int T, int n_fft, int hop) {
  // reflect pad adds n_fft samples in total; center=False framing
  if (T <= 0 || hop <= 0) return 0;
  return T / hop + 1;
}

extern "C" int cmwg_melspec_fwd(const float* x, long long x_bstride, int B, int T, const float* window,
                                const float* fb, const int* fb_lo, const int* fb_hi, int n_fft, int hop, int n_mels,
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
    melspec_kernel<N><<<grid, block, 0, st>>>(x, x_bstride, T, window, fb, fb_lo, fb_hi, n_mels, hop, pad_left, \
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
