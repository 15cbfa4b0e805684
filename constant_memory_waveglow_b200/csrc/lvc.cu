// Location-variable convolution + gate of MelGlow's WN_LVC transform (model/melglow.py:52-92, NonCausalLayerLVC.forward):
//
//   z[b, oc, t]  = sum_{ic, k} w[b, s(t), oc, ic, k] * x[b, ic, t + (k - c) * dil]         s(t) = t / span, zero padding
//   g[b, c, t]   = tanh(z[b, c, t]) * sigmoid(z[b, Cd + c, t])                               (fused_gate, model/waveglow.py:13-15)
//
// Every conditioning frame s owns `span` consecutive samples and its OWN kernel w[b, s] (2Cd x Cr x radix floats, predicted
// from the mel frame), so this is one small GEMM per (batch item, frame) whose "weights" are activations: 0.88 MFLOP against
// 55 KB of kernel per frame and layer at the LJ config (Cd = Cr = 48, radix 3, span 32) -- 16 FLOP per byte, bound by
// reading the predicted kernels from HBM.  CUDA cores (FFMA) are the right unit: one CTA per (frame, batch item) keeps the
// frame's kernel and the haloed input tile in shared memory; a warp owns the 32 samples of ... one group of gate channels, so
// the kernel reads are warp-uniform broadcasts and the input reads are consecutive.
//
// Backward (the autograd of the same two lines):
//   pass A  recompute z, a = tanh, b = sigmoid;  dz_t = dg * b * (1 - a^2),  dz_s = dg * a * b * (1 - b)   -> dz (B, 2Cd, T)
//           dw[b, s, oc, ic, k] = sum_{t in frame s} dz[b, oc, t] * x[b, ic, t + (k - c) * dil]
//   pass B  dx[b, ic, u] = sum_k sum_oc w[b, s(u - (k - c) dil), oc, ic, k] * dz[b, oc, u - (k - c) dil]
//           (a tap's source sample may lie in a neighbouring frame: its kernel is that frame's)
// All sums run in a fixed order: deterministic.
#include "common.cuh"

namespace cmwg {

constexpr int LVC_THREADS = 256;

struct LvcDims {
  int B, T, frames, span;   // T = frames * span
  int Cd, Cr, R, dil;
};

__device__ __forceinline__ size_t lvc_w_elems(const LvcDims& d) { return (size_t)2 * d.Cd * d.Cr * d.R; }

// shared-memory layout: ws [2Cd][Cr*R] (as in global), xs [Cr][span + (R-1)*dil]
template <bool BWD>
__global__ void __launch_bounds__(LVC_THREADS) lvc_gate_kernel(const LvcDims d, const float* __restrict__ x,
                                                               const float* __restrict__ w, float* __restrict__ g,
                                                               const float* __restrict__ dg, float* __restrict__ dz,
                                                               float* __restrict__ dw) {
  extern __shared__ float sm[];
  const int s = blockIdx.x, b = blockIdx.y;
  const int CR = d.Cr * d.R, O2 = 2 * d.Cd;
  const int halo = (d.R - 1) / 2 * d.dil, XW = d.span + 2 * halo;
  float* ws = sm;                      // [O2][CR]
  float* xs = ws + (size_t)O2 * CR;    // [Cr][XW]
  float* dzs = xs + (size_t)d.Cr * XW; // [O2][span]   (backward only)
  const float* wb = w + ((size_t)b * d.frames + s) * lvc_w_elems(d);
  for (int i = threadIdx.x; i < O2 * CR; i += LVC_THREADS) ws[i] = wb[i];
  const int t0 = s * d.span;
  for (int i = threadIdx.x; i < d.Cr * XW; i += LVC_THREADS) {
    const int ic = i / XW, j = i - ic * XW;
    const int t = t0 - halo + j;
    xs[i] = (t >= 0 && t < d.T) ? x[((size_t)b * d.Cr + ic) * d.T + t] : 0.f;
  }
  __syncthreads();
  // (sample, gate channel) pairs; a warp's 32 consecutive indices are 32 consecutive samples of one channel when span >= 32
  for (int idx = threadIdx.x; idx < d.Cd * d.span; idx += LVC_THREADS) {
    const int c = idx / d.span, j = idx - c * d.span;
    const float* wt = ws + (size_t)c * CR;
    const float* wsg = ws + (size_t)(d.Cd + c) * CR;
    float zt = 0.f, zs = 0.f;
    for (int ic = 0; ic < d.Cr; ++ic) {
      const float* xr = xs + (size_t)ic * XW + j;
      for (int k = 0; k < d.R; ++k) {
        const float xv = xr[k * d.dil];
        zt = fmaf(wt[ic * d.R + k], xv, zt);
        zs = fmaf(wsg[ic * d.R + k], xv, zs);
      }
    }
    const float a = tanhf(zt), sg = 1.f / (1.f + expf(-zs));
    const size_t o = ((size_t)b * d.Cd + c) * d.T + t0 + j;
    if (!BWD) {
      g[o] = a * sg;
    } else {
      const float gd = dg[o];
      const float dt = gd * sg * (1.f - a * a), ds = gd * a * sg * (1.f - sg);
      dzs[c * d.span + j] = dt;
      dzs[(d.Cd + c) * d.span + j] = ds;
      dz[((size_t)b * O2 + c) * d.T + t0 + j] = dt;
      dz[((size_t)b * O2 + d.Cd + c) * d.T + t0 + j] = ds;
    }
  }
  if (!BWD) return;
  __syncthreads();
  // kernel gradient of this frame: dw[oc][ic][k] = sum_j dz[oc][j] * x[ic][j + k*dil]
  float* dwb = dw + ((size_t)b * d.frames + s) * lvc_w_elems(d);
  for (int i = threadIdx.x; i < O2 * CR; i += LVC_THREADS) {
    const int oc = i / CR, r = i - oc * CR;
    const int ic = r / d.R, k = r - ic * d.R;
    const float* dr = dzs + (size_t)oc * d.span;
    const float* xr = xs + (size_t)ic * XW + k * d.dil;
    float acc = 0.f;
    for (int j = 0; j < d.span; ++j) acc = fmaf(dr[j], xr[j], acc);
    dwb[i] = acc;
  }
}

// pass B: dx for the samples of frame s.  Tap k of output sample u reads dz at v = u - (k - c) * dil, which belongs to frame
// v / span: per tap at most two frames contribute, their tap-k kernel slices [O2][Cr] are staged one after the other.
__global__ void __launch_bounds__(LVC_THREADS) lvc_dx_kernel(const LvcDims d, const float* __restrict__ w,
                                                             const float* __restrict__ dz, float* __restrict__ dx) {
  extern __shared__ float sm[];
  const int s = blockIdx.x, b = blockIdx.y;
  const int CR = d.Cr * d.R, O2 = 2 * d.Cd;
  float* wk = sm;                          // [O2][Cr]   tap-k slice of one frame's kernel
  float* dzs = wk + (size_t)O2 * d.Cr;     // [O2][span] dz at the shifted samples (zero where another frame / outside)
  float* acc = dzs + (size_t)O2 * d.span;  // [Cr][span]
  const int t0 = s * d.span;
  for (int i = threadIdx.x; i < d.Cr * d.span; i += LVC_THREADS) acc[i] = 0.f;
  const int c = (d.R - 1) / 2;
  for (int k = 0; k < d.R; ++k) {
    const int shift = (k - c) * d.dil;                 // v = u - shift
    const int v_lo = t0 - shift, v_hi = t0 + d.span - 1 - shift;
    const int f_lo = v_lo >= 0 ? v_lo / d.span : -1, f_hi = v_hi >= 0 ? v_hi / d.span : -1;
    for (int f = f_lo; f <= f_hi; ++f) {
      if (f < 0 || f >= d.frames) continue;
      __syncthreads();
      const float* wf = w + ((size_t)b * d.frames + f) * lvc_w_elems(d);
      for (int i = threadIdx.x; i < O2 * d.Cr; i += LVC_THREADS) wk[i] = wf[(size_t)i * d.R + k];
      for (int i = threadIdx.x; i < O2 * d.span; i += LVC_THREADS) {
        const int oc = i / d.span, j = i - oc * d.span;
        const int v = t0 + j - shift;
        dzs[i] = (v >= 0 && v < d.T && v / d.span == f) ? dz[((size_t)b * O2 + oc) * d.T + v] : 0.f;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < d.Cr * d.span; i += LVC_THREADS) {
        const int ic = i / d.span, j = i - ic * d.span;
        float a = acc[i];
        for (int oc = 0; oc < O2; ++oc) a = fmaf(wk[oc * d.Cr + ic], dzs[oc * d.span + j], a);
        acc[i] = a;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.Cr * d.span; i += LVC_THREADS) {
    const int ic = i / d.span, j = i - ic * d.span;
    dx[((size_t)b * d.Cr + ic) * d.T + t0 + j] = acc[i];
  }
}

static int lvc_dims(LvcDims* d, int B, int T, int frames, int Cd, int Cr, int R, int dil) {
  CMWG_REQUIRE(B >= 0 && T > 0 && frames > 0 && T % frames == 0, "lvc: T=%d is not a multiple of frames=%d", T, frames);
  CMWG_REQUIRE(Cd > 0 && Cr > 0 && R >= 1 && (R & 1) && dil >= 1, "lvc: bad channel / radix / dilation");
  d->B = B; d->T = T; d->frames = frames; d->span = T / frames; d->Cd = Cd; d->Cr = Cr; d->R = R; d->dil = dil;
  return CMWG_OK;
}

}  // namespace cmwg

using namespace cmwg;

extern "C" {

int cmwg_lvc_gate_forward(const float* x, const float* w, int B, int T, int frames, int Cd, int Cr, int radix, int dilation,
                          float* g, void* stream) {
  LvcDims d;
  CMWG_PROPAGATE(lvc_dims(&d, B, T, frames, Cd, Cr, radix, dilation));
  CMWG_REQUIRE(x && w && g, "cmwg_lvc_gate_forward: null argument");
  if (B == 0) return CMWG_OK;
  const size_t smem = ((size_t)2 * Cd * Cr * radix + (size_t)Cr * (d.span + (radix - 1) * dilation)) * sizeof(float);
  CMWG_REQUIRE(smem <= 227 * 1024, "cmwg_lvc_gate_forward: frame kernel + input tile (%zu bytes) exceed shared memory", smem);
  CMWG_CHECK_CUDA(cudaFuncSetAttribute(lvc_gate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lvc_gate_kernel<false><<<dim3(frames, B), LVC_THREADS, smem, (cudaStream_t)stream>>>(d, x, w, g, nullptr, nullptr, nullptr);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

int cmwg_lvc_gate_backward(const float* x, const float* w, const float* dg, int B, int T, int frames, int Cd, int Cr,
                           int radix, int dilation, float* dz, float* dx, float* dw, void* stream) {
  LvcDims d;
  CMWG_PROPAGATE(lvc_dims(&d, B, T, frames, Cd, Cr, radix, dilation));
  CMWG_REQUIRE(x && w && dg && dz && dx && dw, "cmwg_lvc_gate_backward: null argument");
  if (B == 0) return CMWG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_a = ((size_t)2 * Cd * Cr * radix + (size_t)Cr * (d.span + (radix - 1) * dilation) + (size_t)2 * Cd * d.span) *
                        sizeof(float);
  CMWG_REQUIRE(smem_a <= 227 * 1024, "cmwg_lvc_gate_backward: %zu bytes of shared memory needed", smem_a);
  CMWG_CHECK_CUDA(cudaFuncSetAttribute(lvc_gate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  lvc_gate_kernel<true><<<dim3(frames, B), LVC_THREADS, smem_a, st>>>(d, x, w, nullptr, dg, dz, dw);
  CMWG_COUNT_LAUNCH();
  const size_t smem_b = ((size_t)2 * Cd * Cr + (size_t)2 * Cd * d.span + (size_t)Cr * d.span) * sizeof(float);
  CMWG_CHECK_CUDA(cudaFuncSetAttribute(lvc_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  lvc_dx_kernel<<<dim3(frames, B), LVC_THREADS, smem_b, st>>>(d, w, dz, dx);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

}  // extern "C"
