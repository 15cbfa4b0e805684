// Dimension bookkeeping and HBM buffer layouts of the WN pipeline (host side).
//
// Slab layout: every activation between WN layers is stored channel-last, [B*T rows][C channels],
// so that (a) a dilated-conv tap is a pure ROW SHIFT of the GEMM A operand (TMA coordinate offset,
// zero fill outside [0,T) gives the 'same' padding for free), (b) the K dimension of every GEMM is
// contiguous (K-major operands, 128B-swizzle friendly), (c) weight-gradient GEMMs read the very same
// buffers as MN-major operands with K = time.
#pragma once
#include "common.cuh"

namespace cmwg {

constexpr int MAX_SEG = 16;  // K segments of one GEMM: radix taps + conditioning / one per layer
constexpr int TC_MAX_WG_REDUCE = 48;  // weight-gradient problems reduced per launch

struct WnDims {
  int cin, aux, Cd, Cr, Cs, depth, bias, prec;
  int R;        // taps of W in weight-layout order: radix (1-D) or radix*radix (2-D, tap = kh*radix + kw)
  int radix;
  int H;        // lines per batch item (1: 1-D WN)
  int hdil[CMWG_MAX_DEPTH];
  // time / line offset of tap s of layer i
  __host__ __device__ int tap_dt(int i, int s) const {
    int kw = (H > 1) ? s % radix : s;
    return (kw - (radix - 1) / 2) * (1 << i);
  }
  __host__ __device__ int tap_dh(int i, int s) const {
    return (H > 1) ? (s / radix - (radix - 1)) * hdil[i] : 0;  // causal in height: rows h-2hd, h-hd, h
  }
  bool tc;      // tcgen05 engine (16-bit operands) vs FFMA engine (fp32 operands)
  int opsize;   // bytes per operand element
  int kb;       // K granule every GEMM segment is padded to: 64 (tc) / 16 (ff)
  int auxp;     // padded conditioning channels
  int Crp, Cdp, Csp, Cd2p;  // channel counts rounded up to kb (Cd2p: 2*Cd)
  int bn_gate;  // N tile of the gate GEMM; tanh/sigmoid partner channels are G = bn_gate/2 apart
  int G;
  int npadA;    // padded rows of the gate GEMM weight matrix
  int KA;       // K of the gate GEMM: R*Crp + auxp
  int ldPB;     // K of the res(/skip) GEMM: Cdp
  int ldQ2;     // K of the dx GEMM: R*Cd2p
  int ldQV;     // K of the conditioning-gradient GEMM: depth*Cd2p (all layers concatenated)
  int ldPS;     // K of the skip GEMM (tc): depth*Cdp
  __host__ __device__ int nb(int i) const { return (i < depth - 1) ? Cr + Cs : Cs; }     // rows of W_o[i]
  __host__ __device__ int cr_eff(int i) const { return (i < depth - 1) ? Cr : 0; }       // residual rows of W_o[i]
  __host__ __device__ int k1(int i) const { return (i < depth - 1) ? Crp + Csp : Csp; }  // K of the dgate GEMM
};

inline bool wn_tc_shapes_ok(const cmwg_wn_config& c) {
  return c.dil_channels % 64 == 0 && c.res_channels % 64 == 0 && c.skip_channels % 64 == 0 &&
         c.radix <= 7 && c.dil_channels >= 64;
}

inline int make_dims(const cmwg_wn_config* c, WnDims* d) {
  CMWG_REQUIRE(c != nullptr, "null cmwg_wn_config");
  CMWG_REQUIRE(c->depth >= 1 && c->depth <= CMWG_MAX_DEPTH, "WN depth %d out of range [1,%d]", c->depth,
               CMWG_MAX_DEPTH);
  CMWG_REQUIRE(c->radix >= 1 && (c->radix % 2) == 1 && c->radix <= 7, "WN radix %d unsupported (odd, <= 7)",
               c->radix);
  CMWG_REQUIRE(c->in_channels >= 1 && c->in_channels <= 64, "WN in_channels %d out of range [1,64]",
               c->in_channels);
  CMWG_REQUIRE(c->aux_channels >= 1, "WN aux_channels %d invalid", c->aux_channels);
  CMWG_REQUIRE(c->dil_channels % 4 == 0 && c->res_channels % 4 == 0 && c->skip_channels % 4 == 0 &&
                   c->dil_channels > 0 && c->res_channels > 0 && c->skip_channels > 0,
               "WN channel counts (dil %d, res %d, skip %d) must be positive multiples of 4", c->dil_channels,
               c->res_channels, c->skip_channels);
  CMWG_REQUIRE(c->precision == CMWG_PREC_FP32 || c->precision == CMWG_PREC_BF16 || c->precision == CMWG_PREC_FP16,
               "unknown precision %d", c->precision);
  d->cin = c->in_channels; d->aux = c->aux_channels; d->Cd = c->dil_channels; d->Cr = c->res_channels;
  d->Cs = c->skip_channels; d->depth = c->depth; d->radix = c->radix; d->bias = c->has_bias ? 1 : 0;
  d->prec = c->precision;
  d->H = c->height > 1 ? c->height : 1;
  d->R = d->H > 1 ? c->radix * c->radix : c->radix;
  CMWG_REQUIRE(d->H == 1 || c->radix == 3, "2-D WN: radix %d unsupported (3 x 3 only)", c->radix);
  CMWG_REQUIRE(d->R + 1 <= MAX_SEG, "too many taps");
  for (int i = 0; i < CMWG_MAX_DEPTH; ++i) {
    d->hdil[i] = (d->H > 1 && i < c->depth) ? c->h_dilation[i] : 0;
    CMWG_REQUIRE(d->hdil[i] >= 0 && d->hdil[i] < 4096, "2-D WN: bad height dilation");
  }
  d->tc = c->precision != CMWG_PREC_FP32;
  if (d->tc && !wn_tc_shapes_ok(*c)) {
    set_error("tensor-core precision requested but WN channels (dil %d, res %d, skip %d) are not multiples of 64",
              c->dil_channels, c->res_channels, c->skip_channels);
    return CMWG_ERR_UNSUPPORTED;
  }
  d->opsize = d->tc ? 2 : 4;
  d->kb = d->tc ? 64 : 16;
  d->auxp = round_up(d->aux, d->kb);
  d->Crp = round_up(d->Cr, d->kb);
  d->Cdp = round_up(d->Cd, d->kb);
  d->Csp = round_up(d->Cs, d->kb);
  d->Cd2p = round_up(2 * d->Cd, d->kb);
  d->bn_gate = (d->tc && d->Cd % 128 == 0) ? 256 : 128;
  d->G = d->bn_gate / 2;
  d->npadA = ceil_div(d->Cd, d->G) * d->bn_gate;
  d->KA = d->R * d->Crp + d->auxp;
  d->ldPB = d->Cdp;
  d->ldQ2 = d->R * d->Cd2p;
  d->ldQV = d->depth * d->Cd2p;
  d->ldPS = d->depth * d->Cdp;
  return CMWG_OK;
}

// ---- packed weights ----------------------------------------------------------------------------
struct PackedLayout {
  // fp32, natural PyTorch layouts (effective weights after weight norm) + 1/||v|| per out channel
  size_t wV, wStart, wEnd, wW[CMWG_MAX_DEPTH], wWo[CMWG_MAX_DEPTH];
  size_t nV, nStart, nW[CMWG_MAX_DEPTH], nWo[CMWG_MAX_DEPTH];
  // fp32 bias vectors in engine order (only when has_bias)
  size_t biasA[CMWG_MAX_DEPTH];  // [2][Cd]: tanh-half bias, sigmoid-half bias (W bias + V bias)
  size_t biasB[CMWG_MAX_DEPTH];  // [nb(i)]
  size_t biasStart, biasEnd;
  size_t biasS;               // [Cs] sum over layers of the skip rows' biases (tc skip GEMM)
  size_t PS;                  // skip GEMM      [Cs][depth*Cdp]      = skip rows of every W_o, K-concatenated (tc)
  // GEMM operand matrices (operand element type), row-major [N][K]
  size_t PA[CMWG_MAX_DEPTH];  // gate GEMM      [npadA][KA]
  size_t PB[CMWG_MAX_DEPTH];  // res/skip GEMM  [nb(i)][ldPB]
  size_t Q1[CMWG_MAX_DEPTH];  // dgate GEMM     [Cd][k1(i)]          = W_o^T
  size_t Q2[CMWG_MAX_DEPTH];  // dx GEMM        [Cr][ldQ2]           = W^T per tap
  size_t QV[CMWG_MAX_DEPTH];  // dy GEMM        [auxp][ldQV] shared; QV[i] points at column i*Cd2p
  // layer 0 with the start conv folded in (fold0_ok): h_0 = W_start x_a has K = in_channels, so layer 0's dilated conv is
  // a K = taps * in_channels GEMM in disguise.  The taps of x_a ride in the padding columns of the conditioning slab.
  size_t PA0f;                // gate GEMM of layer 0   [npadA][auxp]: V_0 | W_0,tap W_start per tap | 0
  size_t PB0f;                // residual GEMM of layer 0 [Cr][Cdp + kb]: W_res,0 | W_start under the centre tap's columns
  // dgate GEMM with the `end` conv folded in (foldend_shapes_ok): dskip = W_end^T d(log_s, t) has K = 2 in_channels, so the
  // skip half of the dgate GEMM is a K = 2 in_channels GEMM in disguise:  dg_i = W_res,i^T dh_{i+1} + (W_end W_skip,i)^T d(log_s, t)
  size_t Q1f[CMWG_MAX_DEPTH]; // [Cd][k1f(i)] = W_res,i^T | (W_end W_skip,i)^T in the first 2 in_channels columns of one k-block
  size_t total;
};
inline int k1f(const WnDims& d, int i) { return (i < d.depth - 1 ? d.Crp : 0) + d.kb; }
inline bool foldend_shapes_ok(const WnDims& d) { return d.tc && d.H == 1 && !d.bias && 2 * d.cin <= 16; }

// Layer 0 can run without the start conv when the taps of x_a fit behind the conditioning channels in the LAST k-block of the
// padded conditioning slab (1-D WN on the tcgen05 engine, no bias, at least one residual layer).
inline bool fold0_shapes_ok(const WnDims& d) {
  return d.tc && d.H == 1 && !d.bias && d.depth >= 2 && d.aux - (d.auxp - d.kb) + d.R * d.cin <= d.kb;
}

inline PackedLayout make_packed_layout(const WnDims& d) {
  PackedLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.wV = take((size_t)2 * d.Cd * d.depth * d.aux * 4);
  L.wStart = take((size_t)d.Cr * d.cin * 4);
  L.wEnd = take((size_t)(2 * d.cin > 16 ? 2 * d.cin : 16) * d.Cs * 4);  // >= 16 rows, tail zero (fused end conv epilogue)
  L.nV = take((size_t)2 * d.Cd * d.depth * 4);
  L.nStart = take((size_t)d.Cr * 4);
  L.biasStart = take((size_t)d.Cr * 4);
  L.biasEnd = take((size_t)2 * d.cin * 4);
  L.biasS = take((size_t)d.Cs * 4);
  L.PS = take((size_t)d.Cs * d.ldPS * d.opsize);
  size_t qv_base = take((size_t)d.auxp * d.ldQV * d.opsize);
  for (int i = 0; i < d.depth; ++i) {
    L.wW[i] = take((size_t)2 * d.Cd * d.Cr * d.R * 4);
    L.wWo[i] = take((size_t)d.nb(i) * d.Cd * 4);
    L.nW[i] = take((size_t)2 * d.Cd * 4);
    L.nWo[i] = take((size_t)d.nb(i) * 4);
    L.biasA[i] = take((size_t)2 * d.Cd * 4);
    L.biasB[i] = take((size_t)d.nb(i) * 4);
    L.PA[i] = take((size_t)d.npadA * d.KA * d.opsize);
    L.PB[i] = take((size_t)d.nb(i) * d.ldPB * d.opsize);
    L.Q1[i] = take((size_t)d.Cd * d.k1(i) * d.opsize);
    L.Q2[i] = take((size_t)d.Cr * d.ldQ2 * d.opsize);
    L.QV[i] = qv_base + (size_t)i * d.Cd2p * d.opsize;
  }
  L.PA0f = take((size_t)d.npadA * d.auxp * d.opsize);
  L.PB0f = take((size_t)d.Cr * (d.Cdp + d.kb) * d.opsize);
  for (int i = 0; i < d.depth; ++i) L.Q1f[i] = take((size_t)d.Cd * k1f(d, i) * d.opsize);
  L.total = off;
  return L;
}

// ---- forward workspace / saved activations -----------------------------------------------------
struct FwdLayout {
  // workspace
  size_t h32;     // [rows][Cr] fp32 residual stream
  size_t skip32;  // [rows][Cs] fp32 cumulative skip
  size_t hop;     // [rows][Cr] operand copy of the layer input (tc inference only; ff aliases h32)
  size_t gop;     // [rows][Cd] operand gate output (inference)
  // tc engine: the residual stream lives as a (hi, lo) pair of 16-bit slabs, h = hi + lo: hi is the
  // next GEMM's operand, and the pair is TMA-loaded into the residual GEMM's epilogue for the fp32 add
  size_t hi2[2];  // ping-pong [rows][Cr] (inference; training keeps hi per layer in `saved`)
  size_t lo2[2];  // ping-pong [rows][Cr]
  size_t gl[CMWG_MAX_DEPTH];  // per-layer gate outputs (inference; the skip GEMM reads all of them at the end)
  size_t flags;   // [depth][2][row tiles] dependency counters of the single-kernel forward (engine_mega.cuh)
  size_t ws_total;
  // saved (training): per layer
  size_t s_hin[CMWG_MAX_DEPTH], s_g[CMWG_MAX_DEPTH], s_a[CMWG_MAX_DEPTH], s_b[CMWG_MAX_DEPTH];
  size_t s_skip;  // final cumulative skip fp32 (for d end.weight)
  size_t saved_total;
};

// ---- backward workspace ------------------------------------------------------------------------
struct BwdLayout {
  size_t dskip_op;   // [rows][Cs] operand
  size_t dh32;       // [rows][Cr] fp32
  size_t dh_op;      // [rows][Cr] operand (tc only; ff aliases dh32)
  size_t dpre_op;    // [rows][2Cd] operand
  size_t dhi2[2], dlo2[2];          // tc: (hi, lo) residual-gradient pairs, ping-pong
  size_t dhi_l[CMWG_MAX_DEPTH];     // tc, single-kernel chain: per-layer hi halves (the weight-gradient GEMMs run afterwards)
  size_t flags;                     // [depth][2][row tiles] dependency counters of the single-kernel chain
  size_t dprel[CMWG_MAX_DEPTH];     // tc: per-layer dpre
  size_t partial;    // split-K partials / block partials
  size_t partial_bytes;
  size_t partial_start;  // block partials of the start conv backward (its own region: the gathered weight-gradient tiles in
                         // `partial` stay alive until the single weight-norm backward launch at the end)
  size_t dycl_lines; // [B][H][T][auxp] fp32 per-line conditioning gradient (2-D WN only; summed over lines afterwards)
  size_t dweff;      // fp32 effective-weight gradients of every conv (consumed by ONE weight-norm backward launch)
  size_t dweff_layer, dweff_start, dweff_end;  // in floats: per-layer stride, offsets of the start / end conv
  size_t gscale;     // 4 floats: gradient scale of the fp16-operand backward {max|dlst| bits, S, 1/S} (wn_kernels.cuh)
  size_t dl16;       // [rows][kb] operand: S * d(log_s, t) in the first 2 in_channels columns, zeros behind (folded `end` conv)
  size_t pred;       // [depth][Cd][2 in_channels] fp32: P_i = g_i^T d(log_s, t) (foldend_dw_kernel)
  size_t total;
};

inline void make_fwd_layout(const WnDims& d, int B, int T, FwdLayout* L) {
  size_t rows = (size_t)B * d.H * T;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  L->h32 = take(rows * d.Cr * 4);
  L->skip32 = take(rows * d.Cs * 4);
  L->hop = d.tc ? take(rows * d.Cr * d.opsize) : L->h32;
  L->gop = take(rows * d.Cd * d.opsize);
  if (d.tc) {
    for (int j = 0; j < 2; ++j) { L->hi2[j] = take(rows * d.Cr * 2); L->lo2[j] = take(rows * d.Cr * 2); }
    for (int i = 0; i < d.depth; ++i) L->gl[i] = take(rows * d.Cd * 2);
  }
  L->flags = take((size_t)d.depth * 2 * B * d.H * ceil_div(T, 256) * 4);
  L->ws_total = off;
  off = 0;
  for (int i = 0; i < d.depth; ++i) {
    L->s_hin[i] = take(rows * d.Cr * d.opsize);
    L->s_g[i] = take(rows * d.Cd * d.opsize);
    L->s_a[i] = d.tc ? L->s_g[i] : take(rows * d.Cd * d.opsize);   // tcgen05 engine: tanh is not saved (= g / sigmoid)
    L->s_b[i] = take(rows * d.Cd * d.opsize);
  }
  L->s_skip = take(rows * d.Cs * 4);
  L->saved_total = off;
}

// time-chunk length of one split of the weight-gradient GEMMs
inline int wgrad_chunk_len(int B, int T) {
  // aim for >= ~300 work items with a 4x2..4x6 tile grid per problem; keep chunks long to bound the
  // partial-sum traffic
  int target_splits = 24;
  int per_batch = ceil_div(target_splits, B);
  int lc = ceil_div(T, per_batch);
  lc = round_up(lc < 256 ? 256 : lc, 64);
  return lc;
}

inline void make_bwd_layout(const WnDims& d, int B, int T, BwdLayout* L) {
  size_t rows = (size_t)B * d.H * T;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  L->dskip_op = take(rows * d.Cs * d.opsize);
  L->dh32 = take(rows * d.Cr * 4);
  L->dh_op = d.tc ? take(rows * d.Cr * d.opsize) : L->dh32;
  L->dpre_op = take(rows * 2 * d.Cd * d.opsize);
  if (d.tc) {
    // residual-gradient stream as (hi, lo) 16-bit pairs, ping-pong; per-layer dpre for the deferred
    // conditioning-gradient GEMM (K-concatenated over layers)
    for (int j = 0; j < 2; ++j) { L->dhi2[j] = take(rows * d.Cr * 2); L->dlo2[j] = take(rows * d.Cr * 2); }
    L->dprel[0] = L->dpre_op;
    for (int i = 1; i < d.depth; ++i) L->dprel[i] = take(rows * 2 * d.Cd * 2);
    for (int i = 0; i < d.depth; ++i) L->dhi_l[i] = take(rows * d.Cr * 2);
    L->flags = take((size_t)d.depth * 2 * B * d.H * ceil_div(T, 256) * 4);
  }
  int lc = wgrad_chunk_len(B * d.H, T);
  size_t splits = (size_t)B * d.H * ceil_div(T, lc);
  // largest simultaneous partial set: all weight-gradient problems of one layer
  size_t per_layer = (size_t)2 * d.Cd * d.Cr * d.R + (size_t)(d.Cr + d.Cs) * d.Cd + (size_t)2 * d.Cd * d.auxp;
  // tc engine: at most TC_PLAN_PAIRS (74) splits per problem group, see tc_wgrad_plan
  if (d.tc) {
    size_t units = (size_t)B * d.H * ceil_div(T, 64);
    splits = units < 80 ? units : 80;
  }
  size_t p1 = splits * per_layer * 4;
  // start / end conv and bias-gradient block partials (32-row blocks)
  size_t blocks32 = (size_t)B * ceil_div(d.H * T, 32) + 1;
  size_t per_block = (size_t)d.Cr * d.cin + d.Cr;
  size_t pe = (size_t)2 * d.cin * d.Cs + 2 * d.cin;
  if (pe > per_block) per_block = pe;
  if ((size_t)2 * d.Cd > per_block) per_block = 2 * d.Cd;
  size_t p2 = (blocks32 + blocks32 / 64 + 2) * per_block * 4;  // + second-stage scratch
  L->partial_bytes = p1 > p2 ? p1 : p2;
  L->partial = take(L->partial_bytes);
  L->partial_start = take(p2);
  L->dycl_lines = d.H > 1 ? take(rows * d.auxp * 4) : 0;
  // effective-weight gradients: per layer dW_o, dW, dV_i side by side, then the start and end convs
  size_t per = (size_t)(d.Cr + d.Cs) * d.Cd + (size_t)2 * d.Cd * d.Cr * d.R + (size_t)2 * d.Cd * d.aux;
  per = align_up(per, 64);
  L->dweff_layer = per;
  L->dweff_start = per * d.depth;
  L->dweff_end = L->dweff_start + align_up((size_t)d.Cr * d.cin, 64);
  size_t total_f = L->dweff_end + align_up((size_t)2 * d.cin * d.Cs, 64);
  L->dweff = take(total_f * 4 + 4096);
  L->gscale = take(64);
  L->dl16 = d.tc ? take(rows * d.kb * 2) : 0;
  L->pred = take((size_t)d.depth * d.Cd * 2 * d.cin * 4);
  L->total = off;
}

}  // namespace cmwg
