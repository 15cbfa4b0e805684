// WSRGlow conditioning front end (model/wsrglow.py:37-50, `_get_cond`) as ONE kernel.
//
//   c (B, Tc) low-rate signal, clipped to [-1, 1] IN PLACE like the reference's c.clip_(-1, 1) (:38)
//   rows [0, 8E)            mu-law code embedding:  cond[b, j*E + e, f] = emb[code(c[b, 8f + j])][e]         (:39)
//   rows [8E, 8E + 9)       STFT magnitudes:        |STFT(reflect_pad(c, 4), n_fft 16, hop 8, hann, center=False)[k, f]|  (:40-47)
//   rows [8E + 9, 8E+9+9P)  phase-code embedding:   cond[b, 8E + 9 + k*P + e, f] = aemb[idx(angle[k, f])][e]   (:48-49)
// with E = 400, P = 50, 256 mu-law codes, 120 phase codes and F = Tc / 8 frames.  The output is the (B, 3659, F) fp32 NCL
// tensor the reference builds with Embedding + view + transpose + stft + cat: 90 MB at the VCTK training shape, so the kernel is
// bound by that write.  A CTA owns 32 frames of one batch item: the 264 samples it needs sit in shared memory, every table
// row is read 128 contiguous bytes at a time (lane = embedding column) and transposed through shared memory so that the
// output is written 128 contiguous bytes at a time (lane = frame).
//
// The 16-point windowed DFT is evaluated directly (9 bins x 16 taps per frame).  DC and Nyquist bins of a real signal are
// exactly real: a real FFT returns imag = +0 there, and atan2(+0, re < 0) = +pi selects the LAST phase code, so their
// imaginary parts are forced to +0 instead of the rounding noise a sine sum would leave.
#include "common.cuh"

namespace cmwg {

constexpr int WSR_FT = 32;          // frames per CTA
constexpr int WSR_NFFT = 16, WSR_HOP = 8, WSR_BINS = 9;

__device__ __forceinline__ int mu_law_code(float x, int channels) {
  // torchaudio.functional.mu_law_encoding: sign(x) * log1p(mu |x|) / log1p(mu), then ((x_mu + 1) / 2 * mu + 0.5) -> int64
  const float mu = (float)(channels - 1);
  const float sgn = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
  const float x_mu = sgn * log1pf(mu * fabsf(x)) / log1pf(mu);
  return (int)((x_mu + 1.f) / 2.f * mu + 0.5f);
}

__global__ void __launch_bounds__(256) wsrglow_cond_kernel(float* __restrict__ c, int Tc, const float* __restrict__ emb, int E,
                                                           int n_codes, const float* __restrict__ aemb, int P, int n_phase,
                                                           const float* __restrict__ window, float* __restrict__ out,
                                                           int* __restrict__ codes_out, int* __restrict__ phase_out) {
  __shared__ float xs[WSR_FT * WSR_HOP + WSR_NFFT];       // reflect-padded samples of this tile
  __shared__ int code[WSR_FT * WSR_HOP];
  __shared__ float mag[WSR_BINS][WSR_FT];
  __shared__ int pidx[WSR_BINS][WSR_FT];
  __shared__ float tile[32][33];
  __shared__ float win[WSR_NFFT];
  const int F = Tc / WSR_HOP;
  const int b = blockIdx.y, f0 = blockIdx.x * WSR_FT;
  const int nf = min(WSR_FT, F - f0);
  float* cb = c + (long long)b * Tc;
  const int rows_total = 8 * E + WSR_BINS + WSR_BINS * P;
  float* ob = out + (long long)b * rows_total * F;
  if (threadIdx.x < WSR_NFFT) win[threadIdx.x] = window[threadIdx.x];
  // padded index p (0 .. Tc + 7) -> sample: reflect pad of 4 on both sides (F.pad(..., (4, 4), mode='reflect'), :42)
  for (int i = threadIdx.x; i < nf * WSR_HOP + WSR_HOP; i += 256) {
    int p = f0 * WSR_HOP + i;            // padded coordinate
    int s = p - 4;
    if (s < 0) s = -s;
    if (s >= Tc) s = 2 * (Tc - 1) - s;
    xs[i] = fminf(fmaxf(cb[s], -1.f), 1.f);
  }
  __syncthreads();
  // clip in place (each sample is owned by exactly one tile) + mu-law codes of the tile's own samples
  for (int i = threadIdx.x; i < nf * WSR_HOP; i += 256) {
    const int s = f0 * WSR_HOP + i;
    const float v = fminf(fmaxf(cb[s], -1.f), 1.f);
    const int cd = min(max(mu_law_code(v, n_codes), 0), n_codes - 1);
    code[i] = cd;
    if (codes_out) codes_out[(long long)b * Tc + s] = cd;
  }
  // windowed DFT: one (bin, frame) per thread
  for (int i = threadIdx.x; i < WSR_BINS * nf; i += 256) {
    const int k = i / nf, f = i - k * nf;
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int n = 0; n < WSR_NFFT; ++n) {
      const float v = xs[f * WSR_HOP + n] * win[n];
      float sn, cs;
      sincospif((float)(2 * k * n) / (float)WSR_NFFT, &sn, &cs);   // exact at multiples of pi/8
      re = fmaf(v, cs, re);
      im = fmaf(-v, sn, im);
    }
    if (k == 0 || k == WSR_NFFT / 2) im = 0.f;
    mag[k][f] = sqrtf(re * re + im * im);
    const float ang = atan2f(im, re);
    // AngleEmbedding.forward (:15-17): ((angle / pi + 1) * 0.5 * (embed_num - 1)).long()
    int id = (int)((ang / 3.14159265358979323846f + 1.f) * 0.5f * (float)(n_phase - 1));
    id = min(max(id, 0), n_phase - 1);
    pidx[k][f] = id;
    if (phase_out) phase_out[((long long)b * WSR_BINS + k) * F + f0 + f] = id;
  }
  __syncthreads();
  // make the clipped samples visible in global memory only now: neighbouring tiles read the 4-sample halo through
  // the same clamp, so the order of these writes against their reads does not matter
  for (int i = threadIdx.x; i < nf * WSR_HOP; i += 256) {
    const int s = f0 * WSR_HOP + i;
    cb[s] = fminf(fmaxf(cb[s], -1.f), 1.f);
  }
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;   // 32 x 8
  // ---- mu-law embedding rows: for sample slot j and a 32-column slice of the table
  const int etiles = (E + 31) / 32;
  for (int jt = 0; jt < 8 * etiles; ++jt) {
    const int j = jt / etiles, e0 = (jt - j * etiles) * 32;
    for (int f = wy; f < WSR_FT; f += 8) {
      float v = 0.f;
      if (f < nf && e0 + lane < E) v = emb[(long long)code[f * WSR_HOP + j] * E + e0 + lane];
      tile[f][lane] = v;
    }
    __syncthreads();
    for (int ey = wy; ey < 32; ey += 8) {
      const int e = e0 + ey;
      if (e < E && lane < nf) ob[(long long)(j * E + e) * F + f0 + lane] = tile[lane][ey];
    }
    __syncthreads();
  }
  // ---- magnitudes
  for (int k = wy; k < WSR_BINS; k += 8)
    if (lane < nf) ob[(long long)(8 * E + k) * F + f0 + lane] = mag[k][lane];
  // ---- phase-code embedding rows (24 KB table: gathered directly)
  for (int r = wy; r < WSR_BINS * P; r += 8) {
    const int k = r / P, e = r - k * P;
    if (lane < nf) ob[(long long)(8 * E + WSR_BINS + r) * F + f0 + lane] = aemb[(long long)pidx[k][lane] * P + e];
  }
}

}  // namespace cmwg

using namespace cmwg;

extern "C" int cmwg_wsrglow_cond(float* c, int B, int Tc, const float* emb, int E, int n_codes, const float* aemb, int P,
                                 int n_phase, const float* window, float* out, int* codes, int* phase_codes, void* stream) {
  CMWG_REQUIRE(c && emb && aemb && window && out, "cmwg_wsrglow_cond: null argument");
  CMWG_REQUIRE(Tc > 0 && Tc % WSR_HOP == 0 && Tc >= 8, "cmwg_wsrglow_cond: low-rate length %d must be a positive multiple of 8", Tc);
  CMWG_REQUIRE(E > 0 && P > 0 && n_codes > 1 && n_phase > 1, "cmwg_wsrglow_cond: bad table sizes");
  if (B == 0) return CMWG_OK;
  const int F = Tc / WSR_HOP;
  dim3 grid(ceil_div(F, WSR_FT), B);
  wsrglow_cond_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, Tc, emb, E, n_codes, aemb, P, n_phase, window, out, codes,
                                                              phase_codes);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}
