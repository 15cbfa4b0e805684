// WaveFlow glue around the 2-D WN (model/waveflow.py:154-265): the row-shifted affine transform with the
// height flip between flows (forward, inverse on a line window, backward) and the conditioning upsampler
// (replication pad -> weight-normed DENSE ConvTranspose1d -> LeakyReLU).  All fp32, HBM / latency bound,
// coalesced along the time axis, deterministic reductions.
#include "common.cuh"

namespace cmwg {

// ------------------------------------------------------------------------------------------------
// affine transform of image lines.  With y = cat(x0, xout) in UNFLIPPED line order (y[0] = x[0]):
//   forward (:203-206)   y[j] = x[j] * exp(log_s[j-1]) + t[j-1]          j >= 1
//   inverse (:253)       y[j] = (z[j] - t[j-1]) / exp(log_s[j-1])        j >= 1
// `in_flip` / `out_flip` read / write line H-1-j instead of j: the flip(2) of :211 and :230 (new image =
// cat(xout.flip(2), x0) = the full height flip of y).  lst = (B, 2, (H-1)*W): log_s then t, line j-1 at
// offset (j-1)*W.  Only lines [j0, j0+nj) are processed (the row-recurrent inverse generates one line per call).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) waveflow_affine_kernel(const float* __restrict__ in, int in_flip,
                                                              const float* __restrict__ lst, float* __restrict__ out,
                                                              int out_flip, int H, int W, int j0, int inverse) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.z, j = j0 + blockIdx.y;
  const int w = blockIdx.x * 256 + threadIdx.x;
  if (w >= W) return;
  const long long img = (long long)b * H * W;
  const float xv = in[img + (long long)(in_flip ? H - 1 - j : j) * W + w];
  float r = xv;
  if (j > 0) {
    const long long lo = (long long)b * 2 * (H - 1) * W + (long long)(j - 1) * W + w;
    const float ls = lst[lo], t = lst[lo + (long long)(H - 1) * W];
    r = inverse ? (xv - t) / expf(ls) : fmaf(xv, expf(ls), t);
  }
  out[img + (long long)(out_flip ? H - 1 - j : j) * W + w] = r;
}

// backward of the forward transform: dy[j] = dout[out_flip ? H-1-j : j];
//   dx[0] = dy[0];  dx[j] = dy[j] * s;  dlog_s[j-1] = dy[j] * x[j] * s + dlogdet[b];  dt[j-1] = dy[j]
__global__ void __launch_bounds__(256) waveflow_affine_bwd_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ lst,
                                                                  const float* __restrict__ dout, int out_flip,
                                                                  const float* __restrict__ dlogdet,
                                                                  float* __restrict__ dx, float* __restrict__ dlst,
                                                                  int H, int W) {
  const int b = blockIdx.z, j = blockIdx.y;
  const int w = blockIdx.x * 256 + threadIdx.x;
  if (w >= W) return;
  const long long img = (long long)b * H * W;
  const float dy = dout[img + (long long)(out_flip ? H - 1 - j : j) * W + w];
  if (j == 0) {
    dx[img + w] = dy;
    return;
  }
  const long long lo = (long long)b * 2 * (H - 1) * W + (long long)(j - 1) * W + w;
  const float s = expf(lst[lo]);
  const float xv = x[img + (long long)j * W + w];
  dx[img + (long long)j * W + w] = dy * s;
  dlst[lo] = fmaf(dy * xv, s, dlogdet ? dlogdet[b] : 0.f);
  dlst[lo + (long long)(H - 1) * W] = dy;
}

// ------------------------------------------------------------------------------------------------
// conditioning upsampler (:169-175): hp = replication-pad(h, (0, rpad)); y = leaky(bias + convT(hp, w_eff));
// weight norm over dim 0 of the (in, out, K) transposed-conv weight, i.e. per INPUT channel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dense_weff_kernel(const float* __restrict__ g, const float* __restrict__ v, int L,
                                                         float* __restrict__ weff, float* __restrict__ inv_norm) {
  __shared__ float red[4];
  const int i = blockIdx.x;
  const float* vi = v + (long long)i * L;
  float scale = 1.f;
  if (g) {
    float ss = 0.f;
    for (int l = threadIdx.x; l < L; l += 128) ss = fmaf(vi[l], vi[l], ss);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    const float norm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    scale = g[i] / norm;
    if (threadIdx.x == 0) inv_norm[i] = 1.f / norm;
  }
  for (int l = threadIdx.x; l < L; l += 128) weff[(long long)i * L + l] = vi[l] * scale;
}

__global__ void __launch_bounds__(128) upsample_dense_fwd_kernel(const float* __restrict__ h,
                                                                 const float* __restrict__ weff,
                                                                 const float* __restrict__ bias, int C, int F, int Fp,
                                                                 int K, int stride, int pad, int Tout, float slope,
                                                                 float* __restrict__ y) {
  const int b = blockIdx.z, o = blockIdx.y;
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= Tout) return;
  float acc = bias ? bias[o] : 0.f;
  const int tp = t + pad;
  for (int k = tp % stride; k < K; k += stride) {
    const int f = (tp - k) / stride;
    if (tp - k < 0 || f >= Fp) continue;
    const int fs = f < F ? f : F - 1;  // replication pad on the right
    const float* hb = h + (long long)b * C * F + fs;
    const float* wk = weff + (long long)o * K + k;
    for (int i = 0; i < C; ++i) acc = fmaf(wk[(long long)i * C * K], hb[(long long)i * F], acc);
  }
  y[((long long)b * C + o) * Tout + t] = acc > 0.f ? acc : acc * slope;
}

// one CTA per INPUT channel i: thread (o, k) accumulates d w_eff[i][o][k] over (b, f) in a fixed order, then the
// weight-norm backward of that channel (utils.py:14-16 -> torch weight_norm dim 0) is finished in place
__global__ void __launch_bounds__(256) upsample_dense_bwd_w_kernel(const float* __restrict__ h,
                                                                   const float* __restrict__ g,
                                                                   const float* __restrict__ v,
                                                                   const float* __restrict__ inv_norm,
                                                                   const float* __restrict__ y,
                                                                   const float* __restrict__ dy, int B, int C, int F,
                                                                   int Fp, int K, int stride, int pad, int Tout,
                                                                   float slope, float* __restrict__ dg,
                                                                   float* __restrict__ dv) {
  extern __shared__ float sm[];  // [B * Fp] padded input row of channel i, then 8 floats of reduction scratch
  float* hs = sm;
  float* red = sm + B * Fp;
  const int i = blockIdx.x;
  const int L = C * K;
  for (int idx = threadIdx.x; idx < B * Fp; idx += 256) {
    const int b = idx / Fp, f = idx % Fp;
    hs[idx] = h[((long long)b * C + i) * F + (f < F ? f : F - 1)];
  }
  __syncthreads();
  float dot = 0.f;
  for (int l = threadIdx.x; l < L; l += 256) {
    const int o = l / K, k = l % K;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* yb = y + ((long long)b * C + o) * Tout;
      const float* dyb = dy + ((long long)b * C + o) * Tout;
      for (int f = 0; f < Fp; ++f) {
        const int t = f * stride - pad + k;
        if (t < 0 || t >= Tout) continue;
        const float d = dyb[t] * (yb[t] > 0.f ? 1.f : slope);
        acc = fmaf(hs[b * Fp + f], d, acc);
      }
    }
    dv[(long long)i * L + l] = acc;  // d w_eff for now
    if (g) dot = fmaf(acc, v[(long long)i * L + l], dot);
  }
  if (!g) return;
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  float tot = 0.f;
  for (int wi = 0; wi < 8; ++wi) tot += red[wi];
  const float inv = inv_norm[i];
  if (threadIdx.x == 0 && dg) dg[i] = tot * inv;
  const float gs = g[i] * inv, kk = tot * inv * inv;
  for (int l = threadIdx.x; l < L; l += 256) {
    const long long q = (long long)i * L + l;
    dv[q] = gs * (dv[q] - v[q] * kk);  // same thread wrote dv[q] above
  }
}

__global__ void __launch_bounds__(256) upsample_dense_bwd_bias_kernel(const float* __restrict__ y,
                                                                      const float* __restrict__ dy, int B, int C,
                                                                      int Tout, float slope, float* __restrict__ dbias) {
  __shared__ float red[8];
  const int o = blockIdx.x;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const long long base = ((long long)b * C + o) * Tout;
    for (int t = threadIdx.x; t < Tout; t += 256) acc += dy[base + t] * (y[base + t] > 0.f ? 1.f : slope);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int wi = 0; wi < 8; ++wi) tot += red[wi];
    dbias[o] = tot;
  }
}

}  // namespace cmwg

using namespace cmwg;

extern "C" {

int cmwg_waveflow_affine(const float* in, int in_flip, const float* lst, float* out, int out_flip, int B, int H, int W,
                         int line_begin, int line_count, int inverse, void* stream) {
  CMWG_REQUIRE(in && out && (lst || (line_begin == 0 && line_count <= 1)), "cmwg_waveflow_affine: null argument");
  CMWG_REQUIRE(H >= 1 && line_begin >= 0 && line_count >= 0 && line_begin + line_count <= H,
               "cmwg_waveflow_affine: lines [%d, %d) outside [0, %d)", line_begin, line_begin + line_count, H);
  if (B == 0 || W == 0 || line_count == 0) return CMWG_OK;
  dim3 grid(ceil_div(W, 256), line_count, B);
  CMWG_CHECK_CUDA(launch_pdl(waveflow_affine_kernel, grid, dim3(256), 0, (cudaStream_t)stream, in, in_flip, lst, out,
                             out_flip, H, W, line_begin, inverse));
  CMWG_COUNT_LAUNCH();
  return CMWG_OK;
}

int cmwg_waveflow_affine_bwd(const float* x, const float* lst, const float* dout, int out_flip, const float* dlogdet,
                             float* dx, float* dlst, int B, int H, int W, void* stream) {
  CMWG_REQUIRE(x && lst && dout && dx && dlst, "cmwg_waveflow_affine_bwd: null argument");
  if (B == 0 || W == 0 || H == 0) return CMWG_OK;
  dim3 grid(ceil_div(W, 256), H, B);
  waveflow_affine_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, lst, dout, out_flip, dlogdet, dx, dlst, H, W);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

size_t cmwg_upsample_dense_workspace(int C, int K) { return ((size_t)C * C * K + C) * sizeof(float) + 256; }

int cmwg_upsample_dense_fwd(const float* h, const float* g, const float* v, const float* bias, int B, int C, int F,
                            int K, int stride, int pad, int rpad, float slope, float* y, void* workspace,
                            void* stream) {
  CMWG_REQUIRE(h && v && y && workspace, "cmwg_upsample_dense_fwd: null argument");
  CMWG_REQUIRE(C >= 1 && F >= 1 && K >= 1 && stride >= 1 && pad >= 0 && rpad >= 0, "cmwg_upsample_dense_fwd: bad dims");
  const int Fp = F + rpad, Tout = (Fp - 1) * stride - 2 * pad + K;
  CMWG_REQUIRE(Tout >= 1, "cmwg_upsample_dense_fwd: empty output");
  if (B == 0) return CMWG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* weff = reinterpret_cast<float*>(workspace);
  float* inv_norm = weff + (size_t)C * C * K;
  dense_weff_kernel<<<C, 128, 0, st>>>(g, v, C * K, weff, inv_norm);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  upsample_dense_fwd_kernel<<<dim3(ceil_div(Tout, 128), C, B), 128, 0, st>>>(h, weff, bias, C, F, Fp, K, stride, pad,
                                                                             Tout, slope, y);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

int cmwg_upsample_dense_bwd(const float* h, const float* g, const float* v, const float* y, const float* dy, int B,
                            int C, int F, int K, int stride, int pad, int rpad, float slope, float* dg, float* dv,
                            float* dbias, void* workspace, void* stream) {
  CMWG_REQUIRE(h && v && y && dy && dv && workspace, "cmwg_upsample_dense_bwd: null argument");
  const int Fp = F + rpad, Tout = (Fp - 1) * stride - 2 * pad + K;
  if (B == 0) return CMWG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* weff = reinterpret_cast<float*>(workspace);
  float* inv_norm = weff + (size_t)C * C * K;
  if (g) {  // 1 / ||v_i||
    dense_weff_kernel<<<C, 128, 0, st>>>(g, v, C * K, weff, inv_norm);
    CMWG_COUNT_LAUNCH();
    CMWG_LAUNCH_CHECK();
  }
  size_t smem = ((size_t)B * Fp + 8) * sizeof(float);
  CMWG_REQUIRE(smem <= 200 * 1024, "cmwg_upsample_dense_bwd: batch x frames = %d too large for one CTA", B * Fp);
  if (smem > 48 * 1024)
    CMWG_CHECK_CUDA(cudaFuncSetAttribute(upsample_dense_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  upsample_dense_bwd_w_kernel<<<C, 256, smem, st>>>(h, g, v, inv_norm, y, dy, B, C, F, Fp, K, stride, pad, Tout, slope,
                                                    dg, dv);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  if (dbias) {
    upsample_dense_bwd_bias_kernel<<<C, 256, 0, st>>>(y, dy, B, C, Tout, slope, dbias);
    CMWG_COUNT_LAUNCH();
    CMWG_LAUNCH_CHECK();
  }
  return CMWG_OK;
}

}  // extern "C"
