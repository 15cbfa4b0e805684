// CUDA-core kernels around the WN GEMMs: weight norm (forward / backward), operand packing,
// conditioning layout change, the tiny-K `start` conv and tiny-N `end` conv (forward / backward),
// fixed-order reductions of block partials.
#pragma once
#include "common.cuh"
#include "wn_layout.cuh"

namespace cmwg {

constexpr int ROWS_PER_BLOCK = 32;  // start/end conv kernels: rows (time steps) per CTA

// ------------------------------------------------------------------------------------------------
// deterministic block reduction (blockDim.x = 128)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_128(float v, float* red /*[4]*/) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------------
// weight norm forward for ALL convs of one WN in one launch  (utils.py:14-16 -> torch weight_norm)
// ------------------------------------------------------------------------------------------------
struct WeffEntry {
  const float* g;  // nullptr: no weight norm (w = v)
  const float* v;
  float* w;        // effective weight out, same layout as v
  float* inv_norm; // [O] or nullptr
  int O, L;        // out channels, elements per out channel
  int row_begin;   // prefix sum of O
};
struct WeffTable {
  WeffEntry e[2 * CMWG_MAX_DEPTH + 3];
  int n;
};

static __global__ void __launch_bounds__(128) weight_eff_kernel(const WeffTable tb) {
  __shared__ float red[4];
  int row = blockIdx.x;
  int ci = 0;
  while (ci + 1 < tb.n && row >= tb.e[ci + 1].row_begin) ++ci;
  const WeffEntry& e = tb.e[ci];
  int o = row - e.row_begin;
  const float* v = e.v + (long long)o * e.L;
  float* w = e.w + (long long)o * e.L;
  if (e.g == nullptr) {
    for (int l = threadIdx.x; l < e.L; l += 128) w[l] = v[l];
    return;
  }
  float ss = 0.f;
  for (int l = threadIdx.x; l < e.L; l += 128) ss = fmaf(v[l], v[l], ss);
  ss = block_sum_128(ss, red);
  float norm = sqrtf(ss);
  float scale = e.g[o] / norm;
  for (int l = threadIdx.x; l < e.L; l += 128) w[l] = v[l] * scale;
  if (threadIdx.x == 0 && e.inv_norm) e.inv_norm[o] = 1.f / norm;
}

// weight norm backward: (dw_eff, v, g, 1/||v||) -> (dg, dv); without weight norm dv = dw_eff
static __global__ void __launch_bounds__(128) weight_norm_bwd_kernel(const float* __restrict__ dw,
                                                                     const float* __restrict__ v,
                                                                     const float* __restrict__ g,
                                                                     const float* __restrict__ inv_norm, int L,
                                                                     float* __restrict__ dg, float* __restrict__ dv) {
  __shared__ float red[4];
  int o = blockIdx.x;
  const float* dwo = dw + (long long)o * L;
  if (g == nullptr) {
    if (dv)
      for (int l = threadIdx.x; l < L; l += 128) dv[(long long)o * L + l] = dwo[l];
    return;
  }
  const float* vo = v + (long long)o * L;
  float dot = 0.f;
  for (int l = threadIdx.x; l < L; l += 128) dot = fmaf(dwo[l], vo[l], dot);
  dot = block_sum_128(dot, red);
  float inv = inv_norm[o];
  if (threadIdx.x == 0 && dg) dg[o] = dot * inv;
  if (dv) {
    float gs = g[o] * inv;
    float k = dot * inv * inv;
    for (int l = threadIdx.x; l < L; l += 128) dv[(long long)o * L + l] = gs * (dwo[l] - vo[l] * k);
  }
}

// ------------------------------------------------------------------------------------------------
// operand packing: one launch packs the five GEMM matrices of every layer
// ------------------------------------------------------------------------------------------------
struct PackParams {
  WnDims d;
  const float* wV;
  const float* wW[CMWG_MAX_DEPTH];
  const float* wWo[CMWG_MAX_DEPTH];
  void* PA[CMWG_MAX_DEPTH];
  void* PB[CMWG_MAX_DEPTH];
  void* Q1[CMWG_MAX_DEPTH];
  void* Q2[CMWG_MAX_DEPTH];
  void* QV;  // shared [auxp][ldQV]
  void* PS;  // [Cs][ldPS] (tc)
  int is_fp16;
};

template <typename OpT>
static __global__ void __launch_bounds__(256) pack_operands_kernel(const PackParams p) {
  const WnDims& d = p.d;
  const int i = blockIdx.y;     // layer
  const int kind = blockIdx.z;  // 0 PA, 1 PB, 2 Q1, 3 Q2, 4 QV, 5 PS
  const int nb = d.nb(i), k1 = d.k1(i), cr_eff = d.cr_eff(i);
  long long size;
  switch (kind) {
    case 0: size = (long long)d.npadA * d.KA; break;
    case 1: size = (long long)nb * d.ldPB; break;
    case 2: size = (long long)d.Cd * k1; break;
    case 3: size = (long long)d.Cr * d.ldQ2; break;
    case 4: size = (long long)d.auxp * d.Cd2p; break;
    default: size = d.tc ? (long long)d.Cs * d.Cdp : 0; break;
  }
  const float* wW = p.wW[i];
  const float* wWo = p.wWo[i];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < size;
       idx += (long long)gridDim.x * blockDim.x) {
    float val = 0.f;
    OpT* dst;
    long long didx = idx;  // destination element index (differs from idx for the K-concatenated matrices)
    if (kind == 0) {
      int n = (int)(idx / d.KA), k = (int)(idx % d.KA);
      int tile = n / d.bn_gate, r = n % d.bn_gate;
      int half = r / d.G, ch = tile * d.G + (r % d.G);
      if (ch < d.Cd) {
        int oc = half * d.Cd + ch;
        if (k < d.R * d.Crp) {
          int tap = k / d.Crp, ic = k % d.Crp;
          if (ic < d.Cr) val = wW[((long long)oc * d.Cr + ic) * d.R + tap];
        } else {
          int kk = k - d.R * d.Crp;
          if (kk < d.aux) val = p.wV[((long long)i * 2 * d.Cd + oc) * d.aux + kk];
        }
      }
      dst = reinterpret_cast<OpT*>(p.PA[i]);
    } else if (kind == 1) {
      int n = (int)(idx / d.ldPB), k = (int)(idx % d.ldPB);
      if (k < d.Cd) val = wWo[(long long)n * d.Cd + k];
      dst = reinterpret_cast<OpT*>(p.PB[i]);
    } else if (kind == 2) {
      int n = (int)(idx / k1), k = (int)(idx % k1);
      int ro = -1;
      if (cr_eff > 0) {
        if (k < d.Crp) { if (k < d.Cr) ro = k; }
        else { int kk = k - d.Crp; if (kk < d.Cs) ro = d.Cr + kk; }
      } else {
        if (k < d.Cs) ro = k;
      }
      if (ro >= 0) val = wWo[(long long)ro * d.Cd + n];
      dst = reinterpret_cast<OpT*>(p.Q1[i]);
    } else if (kind == 3) {
      int ld = d.ldQ2;
      int n = (int)(idx / ld), k = (int)(idx % ld);
      int tap = k / d.Cd2p, oc = k % d.Cd2p;
      if (oc < 2 * d.Cd) val = wW[((long long)oc * d.Cr + n) * d.R + tap];
      dst = reinterpret_cast<OpT*>(p.Q2[i]);
    } else if (kind == 4) {
      int n = (int)(idx / d.Cd2p), k = (int)(idx % d.Cd2p);
      if (n < d.aux && k < 2 * d.Cd) val = p.wV[((long long)i * 2 * d.Cd + k) * d.aux + n];
      dst = reinterpret_cast<OpT*>(p.QV);
      didx = (long long)n * d.ldQV + (long long)i * d.Cd2p + k;
    } else {
      int n = (int)(idx / d.Cdp), k = (int)(idx % d.Cdp);
      if (k < d.Cd) val = wWo[((long long)cr_eff + n) * d.Cd + k];
      dst = reinterpret_cast<OpT*>(p.PS);
      didx = (long long)n * d.ldPS + (long long)i * d.Cdp + k;
    }
    OpTraits<OpT>::store(dst + didx, val, p.is_fp16);
  }
}

struct BiasPackParams {
  WnDims d;
  const float* bV;
  const float* bStart;
  const float* bEnd;
  const float* bW[CMWG_MAX_DEPTH];
  const float* bWo[CMWG_MAX_DEPTH];
  float* biasA[CMWG_MAX_DEPTH];
  float* biasB[CMWG_MAX_DEPTH];
  float* biasStart;
  float* biasEnd;
  float* biasS;
};
static __global__ void __launch_bounds__(256) pack_bias_kernel(const BiasPackParams p) {
  const WnDims& d = p.d;
  int i = blockIdx.y;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * d.Cd; idx += gridDim.x * blockDim.x)
    p.biasA[i][idx] = p.bW[i][idx] + p.bV[i * 2 * d.Cd + idx];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d.nb(i); idx += gridDim.x * blockDim.x)
    p.biasB[i][idx] = p.bWo[i][idx];
  if (i == 0) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d.Cr; idx += gridDim.x * blockDim.x)
      p.biasStart[idx] = p.bStart[idx];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * d.cin; idx += gridDim.x * blockDim.x)
      p.biasEnd[idx] = p.bEnd[idx];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d.Cs; idx += gridDim.x * blockDim.x) {
      float sacc = 0.f;
      for (int l = 0; l < d.depth; ++l) sacc += p.bWo[l][d.cr_eff(l) + idx];
      p.biasS[idx] = sacc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// conditioning: NCL fp32 (strided) <-> slab
// ------------------------------------------------------------------------------------------------
template <typename OpT>
static __global__ void __launch_bounds__(256) cond_pack_kernel(const float* __restrict__ y, long long bs, long long cs,
                                                               long long ts, int aux, int auxp, int T,
                                                               OpT* __restrict__ ycl, int is_fp16) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < aux && t < T) ? y[b * bs + c * cs + t * ts] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int t = t0 + j, c = c0 + tx;
    if (t < T && c < auxp) OpTraits<OpT>::store(ycl + ((long long)b * T + t) * auxp + c, tile[tx][j], is_fp16);
  }
}

static __global__ void __launch_bounds__(256) cond_unpack_grad_kernel(const float* __restrict__ dycl, int aux, int auxp,
                                                                      int T, float* __restrict__ dy) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    int t = t0 + j, c = c0 + tx;
    tile[j][tx] = (t < T && c < auxp) ? dycl[((long long)b * T + t) * auxp + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, t = t0 + tx;
    if (c < aux && t < T) dy[((long long)b * aux + c) * T + t] = tile[tx][j];
  }
}

// ------------------------------------------------------------------------------------------------
// start conv (cin -> Cr, kernel 1): NCL input, slab output            model/waveglow.py:74,99
// ------------------------------------------------------------------------------------------------
template <typename OpT>
static __global__ void __launch_bounds__(256) start_fwd_kernel(const float* __restrict__ x, long long x_bs,
                                                               const float* __restrict__ ws,
                                                               const float* __restrict__ bias, int cin, int Cr,
                                                               int T, int blocks_per_batch,
                                                               float* __restrict__ h32, OpT* __restrict__ hop,
                                                               OpT* __restrict__ hlo, int is_fp16) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [cin][32]
  float* wsm = sm + cin * ROWS_PER_BLOCK;  // [Cr][cin]
  int b = blockIdx.x / blocks_per_batch;
  int t0 = (blockIdx.x % blocks_per_batch) * ROWS_PER_BLOCK;
  for (int idx = threadIdx.x; idx < cin * ROWS_PER_BLOCK; idx += 256) {
    int i = idx / ROWS_PER_BLOCK, r = idx % ROWS_PER_BLOCK;
    int t = t0 + r;
    xs[idx] = (t < T) ? x[b * x_bs + (long long)i * T + t] : 0.f;
  }
  for (int idx = threadIdx.x; idx < Cr * cin; idx += 256) wsm[idx] = ws[idx];
  __syncthreads();
  for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * Cr; idx += 256) {
    int r = idx / Cr, o = idx % Cr;
    int t = t0 + r;
    if (t >= T) continue;
    float acc = bias ? bias[o] : 0.f;
    for (int i = 0; i < cin; ++i) acc = fmaf(wsm[o * cin + i], xs[i * ROWS_PER_BLOCK + r], acc);
    long long off = ((long long)b * T + t) * Cr + o;
    if (h32) h32[off] = acc;
    if (hop) OpTraits<OpT>::store(hop + off, acc, is_fp16);
    if (hlo) {  // tc: residual stream as hi + lo
      float hi = OpTraits<OpT>::load(hop + off, is_fp16);
      OpTraits<OpT>::store(hlo + off, acc - hi, is_fp16);
    }
  }
}

// start conv backward: dx[:, :cin] += Ws^T dh0 ; block partials of dWs (and dbias)
static __global__ void __launch_bounds__(256) start_bwd_kernel(const float* __restrict__ dh0,
                                                               const uint16_t* __restrict__ dh_hi,
                                                               const uint16_t* __restrict__ dh_lo,
                                                               const float* __restrict__ x, long long x_bs,
                                                               const float* __restrict__ ws, int cin, int Cr, int T,
                                                               int blocks_per_batch, float* __restrict__ dx,
                                                               long long dx_bs, float* __restrict__ partial_w,
                                                               float* __restrict__ partial_b) {
  extern __shared__ float sm[];
  const int LD = Cr + 1;
  float* dhs = sm;                          // [32][Cr+1]
  float* xs = dhs + ROWS_PER_BLOCK * LD;    // [cin][32]
  int b = blockIdx.x / blocks_per_batch;
  int t0 = (blockIdx.x % blocks_per_batch) * ROWS_PER_BLOCK;
  for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * Cr; idx += 256) {
    int r = idx / Cr, o = idx % Cr;
    int t = t0 + r;
    float val = 0.f;
    if (t < T) {
      long long off = ((long long)b * T + t) * Cr + o;
      // tc engine: dh_0 arrives as a (hi, lo) bf16 pair
      val = dh0 ? dh0[off] : (op16_to_f32(dh_hi[off], 0) + op16_to_f32(dh_lo[off], 0));
    }
    dhs[r * LD + o] = val;
  }
  for (int idx = threadIdx.x; idx < cin * ROWS_PER_BLOCK; idx += 256) {
    int i = idx / ROWS_PER_BLOCK, r = idx % ROWS_PER_BLOCK;
    int t = t0 + r;
    xs[idx] = (t < T) ? x[b * x_bs + (long long)i * T + t] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < cin * ROWS_PER_BLOCK; idx += 256) {
    int i = idx / ROWS_PER_BLOCK, r = idx % ROWS_PER_BLOCK;
    int t = t0 + r;
    if (t >= T) continue;
    float acc = 0.f;
    for (int o = 0; o < Cr; ++o) acc = fmaf(ws[o * cin + i], dhs[r * LD + o], acc);
    dx[b * dx_bs + (long long)i * T + t] += acc;
  }
  float* pw = partial_w + (long long)blockIdx.x * Cr * cin;
  for (int idx = threadIdx.x; idx < Cr * cin; idx += 256) {
    int o = idx / cin, i = idx % cin;
    float acc = 0.f;
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) acc = fmaf(dhs[r * LD + o], xs[i * ROWS_PER_BLOCK + r], acc);
    pw[idx] = acc;
  }
  if (partial_b) {
    float* pb = partial_b + (long long)blockIdx.x * Cr;
    for (int o = threadIdx.x; o < Cr; o += 256) {
      float acc = 0.f;
      for (int r = 0; r < ROWS_PER_BLOCK; ++r) acc += dhs[r * LD + o];
      pb[o] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// end conv (Cs -> 2cin, kernel 1): slab fp32 input, NCL output          model/waveglow.py:92,105
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) end_fwd_kernel(const float* __restrict__ skip,
                                                             const float* __restrict__ we,
                                                             const float* __restrict__ bias, int cout, int Cs, int T,
                                                             int blocks_per_batch, float* __restrict__ lst) {
  extern __shared__ float sm[];
  const int LD = Cs + 1;
  float* sk = sm;  // [32][Cs+1]
  int b = blockIdx.x / blocks_per_batch;
  int t0 = (blockIdx.x % blocks_per_batch) * ROWS_PER_BLOCK;
  for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * Cs; idx += 256) {
    int r = idx / Cs, k = idx % Cs;
    int t = t0 + r;
    sk[r * LD + k] = (t < T) ? skip[((long long)b * T + t) * Cs + k] : 0.f;
  }
  __syncthreads();
  int r = threadIdx.x & 31, og = threadIdx.x >> 5;
  int t = t0 + r;
  for (int oc = og; oc < cout; oc += 8) {
    float acc = bias ? bias[oc] : 0.f;
    const float* w = we + (long long)oc * Cs;
    for (int k = 0; k < Cs; ++k) acc = fmaf(w[k], sk[r * LD + k], acc);
    if (t < T) lst[((long long)b * cout + oc) * T + t] = acc;
  }
}

template <typename OpT>
static __global__ void __launch_bounds__(256) end_bwd_dskip_kernel(const float* __restrict__ dlst,
                                                                   const float* __restrict__ we, int cout, int Cs,
                                                                   int T, int blocks_per_batch,
                                                                   OpT* __restrict__ dskip, int is_fp16) {
  extern __shared__ float sm[];
  float* dl = sm;  // [cout][32]
  int b = blockIdx.x / blocks_per_batch;
  int t0 = (blockIdx.x % blocks_per_batch) * ROWS_PER_BLOCK;
  for (int idx = threadIdx.x; idx < cout * ROWS_PER_BLOCK; idx += 256) {
    int oc = idx / ROWS_PER_BLOCK, r = idx % ROWS_PER_BLOCK;
    int t = t0 + r;
    dl[idx] = (t < T) ? dlst[((long long)b * cout + oc) * T + t] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * Cs; idx += 256) {
    int r = idx / Cs, k = idx % Cs;
    int t = t0 + r;
    if (t >= T) continue;
    float acc = 0.f;
    for (int oc = 0; oc < cout; ++oc) acc = fmaf(we[(long long)oc * Cs + k], dl[oc * ROWS_PER_BLOCK + r], acc);
    OpTraits<OpT>::store(dskip + ((long long)b * T + t) * Cs + k, acc, is_fp16);
  }
}

static __global__ void __launch_bounds__(256) end_bwd_dw_kernel(const float* __restrict__ dlst,
                                                                const float* __restrict__ skip, int cout, int Cs,
                                                                int T, int blocks_per_batch,
                                                                float* __restrict__ partial_w,
                                                                float* __restrict__ partial_b) {
  extern __shared__ float sm[];
  float* dl = sm;                           // [cout][32]
  float* sk = sm + cout * ROWS_PER_BLOCK;   // [32][Cs]
  int b = blockIdx.x / blocks_per_batch;
  int t0 = (blockIdx.x % blocks_per_batch) * ROWS_PER_BLOCK;
  for (int idx = threadIdx.x; idx < cout * ROWS_PER_BLOCK; idx += 256) {
    int oc = idx / ROWS_PER_BLOCK, r = idx % ROWS_PER_BLOCK;
    int t = t0 + r;
    dl[idx] = (t < T) ? dlst[((long long)b * cout + oc) * T + t] : 0.f;
  }
  for (int idx = threadIdx.x; idx < ROWS_PER_BLOCK * Cs; idx += 256) {
    int r = idx / Cs, k = idx % Cs;
    int t = t0 + r;
    sk[idx] = (t < T) ? skip[((long long)b * T + t) * Cs + k] : 0.f;
  }
  __syncthreads();
  float* pw = partial_w + (long long)blockIdx.x * cout * Cs;
  for (int idx = threadIdx.x; idx < cout * Cs; idx += 256) {
    int oc = idx / Cs, k = idx % Cs;
    float acc = 0.f;
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) acc = fmaf(dl[oc * ROWS_PER_BLOCK + r], sk[r * Cs + k], acc);
    pw[idx] = acc;
  }
  if (partial_b) {
    float* pb = partial_b + (long long)blockIdx.x * cout;
    for (int oc = threadIdx.x; oc < cout; oc += 256) {
      float acc = 0.f;
      for (int r = 0; r < ROWS_PER_BLOCK; ++r) acc += dl[oc * ROWS_PER_BLOCK + r];
      pb[oc] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
// out[p] = sum_j partial[j][p], fixed order, fp64 accumulator
static __global__ void __launch_bounds__(128) reduce_blocks_kernel(const float* __restrict__ partial, int nblocks,
                                                                   int P, float* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double s = 0.0;
  for (int j = 0; j < nblocks; ++j) s += (double)partial[(long long)j * P + p];
  out[p] = (float)s;
}

// split-K partials [splits][M][N] of several weight-gradient problems -> strided destinations
struct WgReduceEntry {
  const float* partial;
  int M, N, n_valid;     // stored dims, valid columns
  float* out;
  long long sm, sn, off; // out[off + m*sm + n*sn]
};
struct WgReduceTable {
  WgReduceEntry e[TC_MAX_WG_REDUCE];
  int n, splits;
};

static __global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgReduceTable tb) {
  const WgReduceEntry& e = tb.e[blockIdx.y];
  const int nv4 = (e.n_valid + 3) >> 2;  // stored N is a multiple of 4, so the float4 loads stay in bounds
  long long total = (long long)e.M * nv4;
  const long long stride = (long long)e.M * e.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int m = (int)(idx / nv4), n = (int)(idx % nv4) * 4;
    const float* p = e.partial + (long long)m * e.N + n;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < tb.splits; ++j) {  // fixed order: deterministic
      float4 v = *reinterpret_cast<const float4*>(p + j * stride);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = e.out + e.off + m * e.sm + n * e.sn;
    o[0] = s.x;
    if (n + 1 < e.n_valid) o[e.sn] = s.y;
    if (n + 2 < e.n_valid) o[2 * e.sn] = s.z;
    if (n + 3 < e.n_valid) o[3 * e.sn] = s.w;
  }
}

// two-level fixed-order reduction of block partials: [nblocks][P] -> [ceil(nblocks/64)][P]
static __global__ void __launch_bounds__(128) reduce_blocks_stage_kernel(const float* __restrict__ partial, int nblocks,
                                                                         int P, float* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int j0 = blockIdx.y * 64, j1 = min(j0 + 64, nblocks);
  float s = 0.f;
  for (int j = j0; j < j1; ++j) s += partial[(long long)j * P + p];
  out[(long long)blockIdx.y * P + p] = s;
}

// column sums of a slab (bias gradients): block partials over 32 rows
template <typename OpT>
static __global__ void __launch_bounds__(256) colsum_partial_kernel(const OpT* __restrict__ a, int ld, int C,
                                                                    long long rows, float* __restrict__ partial,
                                                                    int is_fp16) {
  long long r0 = (long long)blockIdx.x * ROWS_PER_BLOCK;
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (int r = 0; r < ROWS_PER_BLOCK; ++r) {
      long long row = r0 + r;
      if (row < rows) acc += OpTraits<OpT>::load(a + row * ld + c, is_fp16);
    }
    partial[(long long)blockIdx.x * C + c] = acc;
  }
}

}  // namespace cmwg
