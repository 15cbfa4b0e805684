// Epilogue functors of the tcgen05 engine.  One call covers ONE accumulator row (= one (batch, time)
// column of the reference's NCL tensors) and 32 consecutive epilogue columns; inputs and outputs are
// 32-wide register fragments that the engine moves through swizzled shared-memory staging and TMA.
//
//   kPaired    the tile's columns [0, BN/2) and [BN/2, BN) are partner pre-activations (tanh / sigmoid)
//   kOut       output streams (each a slab the engine TMA-stores 32 rows x 32 columns at a time)
//   kOutF32    outputs are fp32 (else 16-bit operands, two per register)
//   kOutBufs   staging buffers per output stream (2 = the store of chunk i overlaps chunk i+1)
//   kIn        16-bit input streams, TMA-loaded one chunk ahead
//   out_col(i, c0) / in_col(i, c0)   channel coordinate of stream i for epilogue column c0
#pragma once
#include "epilogues.cuh"

namespace cmwg {

// ---- gate: g = tanh(pre_t) * sigmoid(pre_s)  (model/waveglow.py:13-15,42-44) ---------------------
// SAVE additionally stores tanh(pre_t) and sigmoid(pre_s) for the backward pass.
template <bool SAVE>
struct GateTcEpi {
  static constexpr bool kPaired = true, kOutF32 = false;
  static constexpr int kOut = SAVE ? 3 : 1, kIn = 0, kOutBufs = 2;
  const float* bias;  // nullptr or [2][Cd]
  int Cd, f16;
  __device__ __forceinline__ int out_col(int, int c0) const { return c0; }
  __device__ __forceinline__ int in_col(int, int c0) const { return c0; }
  __device__ __forceinline__ void compute(int ch0, const float (&lo)[32], const float (&hi)[32],
                                          uint32_t (&o)[kOut][16]) const {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a[2], b[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float pt = lo[j + u], ps = hi[j + u];
        if (bias) { pt += __ldg(bias + ch0 + j + u); ps += __ldg(bias + Cd + ch0 + j + u); }
        if (f16) {
          a[u] = tanh_ex2(pt);
          b[u] = sigmoid_ex2(ps);
        } else {
          a[u] = tanh_f<true>(pt);
          b[u] = sigmoid_f<true>(ps);
        }
      }
      o[0][j >> 1] = pack2(a[0] * b[0], a[1] * b[1], f16);
      if constexpr (SAVE) {
        o[1][j >> 1] = pack2(a[0], a[1], f16);
        o[2][j >> 1] = pack2(b[0], b[1], f16);
      }
    }
  }
};

// ---- (hi, lo) split store with optional (hi, lo) addend ------------------------------------------
// The tcgen05 pipeline keeps the residual stream (and its gradient) as a PAIR of 16-bit slabs,
// x = hi + lo with hi = rn16(x), lo = rn16(x - hi) (precision of the pair: 2^-17 relative); hi doubles
// as the GEMM operand of the next dilated conv.  ADD: out = acc + in_hi + in_lo, the residual add of
// model/waveglow.py:46 (forward) or the upstream residual gradient (backward), summed in fp32.
template <bool ADD>
struct SplitTcEpi {
  static constexpr bool kPaired = false, kOutF32 = false;
  static constexpr int kOut = 2, kIn = ADD ? 2 : 0, kOutBufs = 1;
  const float* bias;
  int f16;
  __device__ __forceinline__ int out_col(int, int c0) const { return c0; }
  __device__ __forceinline__ int in_col(int, int c0) const { return c0; }
  __device__ __forceinline__ void split(int col0, float (&x)[32], uint32_t (&o)[2][16]) const {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float x0 = x[j], x1 = x[j + 1];
      if (bias) { x0 += __ldg(bias + col0 + j); x1 += __ldg(bias + col0 + j + 1); }
      uint32_t h = pack2(x0, x1, f16);
      float h0, h1;
      unpack2(h, f16, h0, h1);
      o[0][j >> 1] = h;
      o[1][j >> 1] = pack2(x0 - h0, x1 - h1, f16);
    }
  }
  __device__ __forceinline__ void compute(int col0, const float (&v)[32], uint32_t (&o)[2][16]) const {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = v[j];
    split(col0, x, o);
  }
  __device__ __forceinline__ void compute(int col0, const float (&v)[32], const uint32_t (&in)[2][16],
                                          uint32_t (&o)[2][16]) const {
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float h0, h1, l0, l1;
      unpack2(in[0][j >> 1], f16, h0, h1);
      unpack2(in[1][j >> 1], f16, l0, l1);
      x[j] = v[j] + (h0 + l0);
      x[j + 1] = v[j + 1] + (h1 + l1);
    }
    split(col0, x, o);
  }
};

// ---- gate backward: dpre = dg * d(tanh * sigmoid) ------------------------------------------------
// inputs: saved tanh / sigmoid values; outputs: the tanh-half and sigmoid-half gradients, columns
// [0, Cd) and [Cd, 2Cd) of the dpre slab (two tensor maps, one per column window).
struct GateBwdTcEpi {
  static constexpr bool kPaired = false, kOutF32 = false;
  static constexpr int kOut = 2, kIn = 2, kOutBufs = 1;
  int Cd, f16;
  __device__ __forceinline__ int out_col(int, int c0) const { return c0; }
  __device__ __forceinline__ int in_col(int, int c0) const { return c0; }
  __device__ __forceinline__ void compute(int, const float (&v)[32], const uint32_t (&in)[2][16],
                                          uint32_t (&o)[2][16]) const {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a0, a1, b0, b1;
      unpack2(in[0][j >> 1], f16, a0, a1);
      unpack2(in[1][j >> 1], f16, b0, b1);
      o[0][j >> 1] = pack2(v[j] * b0 * (1.f - a0 * a0), v[j + 1] * b1 * (1.f - a1 * a1), f16);
      o[1][j >> 1] = pack2(v[j] * a0 * b0 * (1.f - b0), v[j + 1] * a1 * b1 * (1.f - b1), f16);
    }
  }
};

// ---- plain fp32 store (skip sum, conditioning gradient, self tests) -------------------------------
struct StoreTcEpi {
  static constexpr bool kPaired = false, kOutF32 = true;
  static constexpr int kOut = 1, kIn = 0, kOutBufs = 2;
  const float* bias;  // nullptr or [N]
  __device__ __forceinline__ int out_col(int, int c0) const { return c0; }
  __device__ __forceinline__ int in_col(int, int c0) const { return c0; }
  __device__ __forceinline__ void compute(int col0, const float (&v)[32], uint32_t (&o)[1][32]) const {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[0][j] = __float_as_uint(v[j] + (bias ? __ldg(bias + col0 + j) : 0.f));
  }
};

}  // namespace cmwg
