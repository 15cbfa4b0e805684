// Conditioning upsampler (model/waveglow.py:126-130, 210-212): depthwise ConvTranspose1d with weight
// norm (dim 0 -> one norm per channel over the K taps) and bias.
//   y[b,c,to] = bias[c] + sum_f h[b,c,f] * w[c, to + pad - f*stride]        (0 <= tap < K)
#include "common.cuh"

namespace cmwg {

__device__ __forceinline__ float block_sum128(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return r;
}

// one block per (b, c) row
__global__ void __launch_bounds__(128) upsample_fwd_kernel(const float* __restrict__ h, const float* __restrict__ g,
                                                           const float* __restrict__ v,
                                                           const float* __restrict__ bias, int C, int F, int K,
                                                           int stride, int pad, int Tout, float* __restrict__ y,
                                                           int stage_h) {
  extern __shared__ float sm[];
  float* w = sm;       // [K]
  __shared__ float red[4];
  int b = blockIdx.x / C, c = blockIdx.x % C;
  float ss = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    float vv = v[(long long)c * K + k];
    w[k] = vv;
    ss = fmaf(vv, vv, ss);
  }
  // the row of frames is staged in shared memory when it fits (stage_h); a longer row (full-utterance super-resolution:
  // F = T / 16 frames with upsample factor 1) is read in place -- consecutive outputs touch consecutive frames, L1 serves them
  const float* hs = h + ((long long)b * C + c) * F;
  if (stage_h) {
    float* hsm = sm + K;  // [F]
    for (int f = threadIdx.x; f < F; f += 128) hsm[f] = hs[f];
    hs = hsm;
  }
  ss = block_sum128(ss, red);
  float scale = g ? g[c] / sqrtf(ss) : 1.f;
  float bv = bias ? bias[c] : 0.f;
  for (int to = threadIdx.x; to < Tout; to += 128) {
    int num = to + pad;
    int f_hi = min(num / stride, F - 1);
    int f_lo = max((num - K + stride) / stride, 0);  // ceil((num-K+1)/stride) for num-K+1 >= 0, else 0
    if (num - K + 1 <= 0) f_lo = 0;
    float acc = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) {
      int k = num - f * stride;
      if (k >= 0 && k < K) acc = fmaf(hs[f], w[k] * scale, acc);
    }
    y[((long long)b * C + c) * Tout + to] = acc + bv;
  }
}

// one block per channel: dw_eff[k] = sum_{b,f} h[b,c,f] dy[b,c,f*stride-pad+k]; then weight-norm backward
__global__ void __launch_bounds__(128) upsample_bwd_kernel(const float* __restrict__ h, const float* __restrict__ g,
                                                           const float* __restrict__ v,
                                                           const float* __restrict__ dy, long long dy_bs,
                                                           long long dy_cs, int B, int C, int F, int K, int stride,
                                                           int pad, int Tvalid, float* __restrict__ dg,
                                                           float* __restrict__ dv, float* __restrict__ dbias) {
  __shared__ float red[4];
  int c = blockIdx.x;
  // bias gradient
  float sb = 0.f;
  for (int b = 0; b < B; ++b)
    for (int t = threadIdx.x; t < Tvalid; t += 128) sb += dy[b * dy_bs + c * dy_cs + t];
  sb = block_sum128(sb, red);
  if (threadIdx.x == 0 && dbias) dbias[c] = sb;
  // per-tap effective-weight gradient (taps strided over threads; K may exceed 128)
  float ss = 0.f, dot = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    float vv = v[(long long)c * K + k];
    float acc = 0.f;
    for (int b = 0; b < B; ++b)
      for (int f = 0; f < F; ++f) {
        int to = f * stride - pad + k;
        if (to >= 0 && to < Tvalid) acc = fmaf(h[((long long)b * C + c) * F + f], dy[b * dy_bs + c * dy_cs + to], acc);
      }
    ss = fmaf(vv, vv, ss);
    dot = fmaf(acc, vv, dot);
    // stash dw_eff in dv for the second pass
    dv[(long long)c * K + k] = acc;
  }
  ss = block_sum128(ss, red);
  dot = block_sum128(dot, red);
  if (g == nullptr) return;  // no weight norm: dv already holds d(weight)
  float inv = 1.f / sqrtf(ss);
  if (threadIdx.x == 0 && dg) dg[c] = dot * inv;
  float gs = g[c] * inv, kk = dot * inv * inv;
  for (int k = threadIdx.x; k < K; k += 128) {
    float dwk = dv[(long long)c * K + k];
    dv[(long long)c * K + k] = gs * (dwk - v[(long long)c * K + k] * kk);
  }
}

// Two-pass form of the same gradient for batches: pass 1, one CTA per (channel, batch group), writes partial dw_eff[k] and
// partial dbias to the workspace; pass 2, one CTA per channel, folds the groups in a fixed order and applies the weight-norm
// backward.  (The one-pass kernel above walks B * F products per tap in ONE thread: 195 us at the LJ training shape.)
constexpr int UPS_GROUPS = 8;
__global__ void __launch_bounds__(128) upsample_bwd_partial_kernel(const float* __restrict__ h, const float* __restrict__ dy,
                                                                   long long dy_bs, long long dy_cs, int B, int C, int F,
                                                                   int K, int stride, int pad, int Tvalid, int groups,
                                                                   float* __restrict__ ws) {
  __shared__ float red[4];
  const int c = blockIdx.x, gi = blockIdx.y;
  const int b0 = (int)((long long)B * gi / groups), b1 = (int)((long long)B * (gi + 1) / groups);
  float* out = ws + ((long long)gi * C + c) * (K + 1);
  float sb = 0.f;
  for (int b = b0; b < b1; ++b)
    for (int t = threadIdx.x; t < Tvalid; t += 128) sb += dy[b * dy_bs + c * dy_cs + t];
  sb = block_sum128(sb, red);
  if (threadIdx.x == 0) out[K] = sb;
  for (int k = threadIdx.x; k < K; k += 128) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      const float* hr = h + ((long long)b * C + c) * F;
      const float* dr = dy + b * dy_bs + c * dy_cs;
      for (int f = 0; f < F; ++f) {
        const int to = f * stride - pad + k;
        if (to >= 0 && to < Tvalid) acc = fmaf(hr[f], dr[to], acc);
      }
    }
    out[k] = acc;
  }
}

__global__ void __launch_bounds__(128) upsample_bwd_final_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                                                 const float* __restrict__ ws, int C, int K, int groups,
                                                                 float* __restrict__ dg, float* __restrict__ dv,
                                                                 float* __restrict__ dbias) {
  __shared__ float red[4];
  const int c = blockIdx.x;
  if (threadIdx.x == 0 && dbias) {
    float sb = 0.f;
    for (int gi = 0; gi < groups; ++gi) sb += ws[((long long)gi * C + c) * (K + 1) + K];
    dbias[c] = sb;
  }
  float ss = 0.f, dot = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    float acc = 0.f;
    for (int gi = 0; gi < groups; ++gi) acc += ws[((long long)gi * C + c) * (K + 1) + k];
    const float vv = v[(long long)c * K + k];
    ss = fmaf(vv, vv, ss);
    dot = fmaf(acc, vv, dot);
    dv[(long long)c * K + k] = acc;
  }
  ss = block_sum128(ss, red);
  dot = block_sum128(dot, red);
  if (g == nullptr) return;
  const float inv = 1.f / sqrtf(ss);
  if (threadIdx.x == 0 && dg) dg[c] = dot * inv;
  const float gs = g[c] * inv, kk = dot * inv * inv;
  for (int k = threadIdx.x; k < K; k += 128) {
    const float dwk = dv[(long long)c * K + k];
    dv[(long long)c * K + k] = gs * (dwk - v[(long long)c * K + k] * kk);
  }
}

// input gradient of the transposed conv: dh[b,c,f] = sum_k w_eff[c,k] dy[b,c,f*stride-pad+k]; one block per (b, c) row
__global__ void __launch_bounds__(128) upsample_bwd_input_kernel(const float* __restrict__ g,
                                                                 const float* __restrict__ v,
                                                                 const float* __restrict__ dy, long long dy_bs,
                                                                 long long dy_cs, int C, int F, int K, int stride,
                                                                 int pad, int Tvalid, float* __restrict__ dh) {
  extern __shared__ float sm[];
  float* w = sm;  // [K]
  __shared__ float red[4];
  int b = blockIdx.x / C, c = blockIdx.x % C;
  float ss = 0.f;
  for (int k = threadIdx.x; k < K; k += 128) {
    float vv = v[(long long)c * K + k];
    w[k] = vv;
    ss = fmaf(vv, vv, ss);
  }
  ss = block_sum128(ss, red);
  float scale = g ? g[c] / sqrtf(ss) : 1.f;
  const float* dyr = dy + b * dy_bs + c * dy_cs;
  for (int f = threadIdx.x; f < F; f += 128) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      int to = f * stride - pad + k;
      if (to >= 0 && to < Tvalid) acc = fmaf(w[k], dyr[to], acc);
    }
    dh[((long long)b * C + c) * F + f] = acc * scale;
  }
}

}  // namespace cmwg

using namespace cmwg;

extern "C" {

int cmwg_upsample_fwd(const float* h, const float* g, const float* v, const float* bias, int B, int C, int F, int K,
                      int stride, int pad, float* y, void* stream) {
  int Tout = (F - 1) * stride - 2 * pad + K;
  CMWG_REQUIRE(Tout > 0 && stride > 0 && K > 0, "cmwg_upsample_fwd: bad geometry F=%d K=%d stride=%d pad=%d", F, K,
               stride, pad);
  if (B == 0 || C == 0) return CMWG_OK;
  size_t smem = (size_t)(K + F) * sizeof(float);
  const int stage_h = smem <= 48 * 1024 ? 1 : 0;
  if (!stage_h) smem = (size_t)K * sizeof(float);
  CMWG_REQUIRE(smem <= 48 * 1024, "cmwg_upsample_fwd: K=%d taps exceed the shared-memory staging buffer", K);
  upsample_fwd_kernel<<<B * C, 128, smem, (cudaStream_t)stream>>>(h, g, v, bias, C, F, K, stride, pad, Tout, y, stage_h);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

size_t cmwg_upsample_bwd_workspace(int B, int C, int K) {
  (void)B;
  return (size_t)UPS_GROUPS * C * (K + 1) * sizeof(float);
}

int cmwg_upsample_bwd(const float* h, const float* g, const float* v, const float* dy, long long dy_bstride,
                      long long dy_cstride, int B, int C, int F, int K, int stride, int pad, int Tvalid, float* dg,
                      float* dv, float* dbias, void* workspace, void* stream) {
  CMWG_REQUIRE(dv != nullptr, "cmwg_upsample_bwd: dv must not be NULL");
  if (C == 0) return CMWG_OK;
  if (workspace != nullptr && B >= 2) {
    const int groups = B < UPS_GROUPS ? B : UPS_GROUPS;
    float* ws = reinterpret_cast<float*>(workspace);
    upsample_bwd_partial_kernel<<<dim3(C, groups), 128, 0, (cudaStream_t)stream>>>(h, dy, dy_bstride, dy_cstride, B, C, F, K,
                                                                                   stride, pad, Tvalid, groups, ws);
    CMWG_COUNT_LAUNCH();
    upsample_bwd_final_kernel<<<C, 128, 0, (cudaStream_t)stream>>>(g, v, ws, C, K, groups, dg, dv, dbias);
    CMWG_COUNT_LAUNCH();
    CMWG_LAUNCH_CHECK();
    return CMWG_OK;
  }
  upsample_bwd_kernel<<<C, 128, 0, (cudaStream_t)stream>>>(h, g, v, dy, dy_bstride, dy_cstride, B, C, F, K, stride, pad,
                                                           Tvalid, dg, dv, dbias);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

int cmwg_upsample_bwd_input(const float* g, const float* v, const float* dy, long long dy_bstride,
                            long long dy_cstride, int B, int C, int F, int K, int stride, int pad, int Tvalid,
                            float* dh, void* stream) {
  CMWG_REQUIRE(v && dy && dh, "cmwg_upsample_bwd_input: null argument");
  if (B == 0 || C == 0) return CMWG_OK;
  upsample_bwd_input_kernel<<<B * C, 128, (size_t)K * sizeof(float), (cudaStream_t)stream>>>(
      g, v, dy, dy_bstride, dy_cstride, C, F, K, stride, pad, Tvalid, dh);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

}  // extern "C"
