// tcgen05 / TMEM / TMA GEMM engine for sm_100a (hand-written PTX; no CUTLASS).
//
// Persistent, warp-specialised CTAs (one per SM):
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor loads of the A (activation slab) and
//                      B (packed weight) tiles into a STAGES-deep shared-memory ring, 128B swizzle
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, K = 16 per
//                      instruction, fp32 accumulators in TMEM, double buffered (2 x BN columns)
//   warps 2..9         epilogue (two warps per TMEM lane quadrant, half of the columns each): tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue
//                      functor (gate / residual+skip / gate-backward / ...) -> global
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and TMEM full/empty mbarriers
// (MMA <-> epilogue) so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Two operand arrangements:
//   KMAJOR  (forward / dgrad GEMMs)  A = slab rows x channels (K = channels contiguous),
//           a dilated tap is a row shift of the TMA coordinate, out-of-range rows are zero filled by
//           the TMA unit (the conv's 'same' padding);  B = packed weights [N][K].
//   MNMAJOR (weight-gradient GEMMs)  A = slab[t][m], B = slab[t + shift][n], K = time: the very
//           same slabs read through MN-major UMMA descriptors, split-K over time chunks.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "engine_ff.cuh"  // GemmDesc / WgradProblem

namespace cmwg {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // 64 x 16-bit = 128 B = one swizzle row
constexpr int TC_A_BYTES = TC_BM * 128;   // 16 KB
constexpr int TC_EPI_WARPS = 8;            // 2 per TMEM lane quadrant, each owning half of the tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_MAX_WG = 8;              // weight-gradient problems per launch

struct alignas(64) TcGemmParams {
  CUtensorMap a_map[MAX_SEG];
  CUtensorMap b_map;
  int nseg;
  int seg_nkb[MAX_SEG];
  int seg_shift[MAX_SEG];
  int seg_koff[MAX_SEG];
  int B, T, tiles_per_batch, n_tiles, total_tiles, N;
  uint32_t idesc;
  uint32_t desc_lbo, desc_sbo;  // >>4 encoded; overridable by the self test
};

struct alignas(64) TcWgradParams {
  CUtensorMap a_map[TC_MAX_WG];
  CUtensorMap b_map[TC_MAX_WG];
  int nprob;
  int M[TC_MAX_WG], N[TC_MAX_WG], shift[TC_MAX_WG], a_c0[TC_MAX_WG], b_c0[TC_MAX_WG];
  int tile_begin[TC_MAX_WG + 1];  // prefix sum of (m_tiles * n_tiles) per problem
  int n_tiles_n[TC_MAX_WG];
  float* partial[TC_MAX_WG];
  int B, T, Lc, chunks_per_batch, splits, total_work;
  uint32_t idesc;
  uint32_t desc_lbo, desc_sbo;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor (sm_100 "version 1"), SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_enc, uint32_t sbo_enc) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_enc & 0x3FFF) << 16;
  d |= (uint64_t)(sbo_enc & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// instruction descriptor for kind::f16, fp32 accumulate
inline uint32_t make_idesc(int is_fp16, int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t fmt = is_fp16 ? 0u : 1u;  // 0 = F16, 1 = BF16
  uint32_t d = 0;
  d |= 1u << 4;                      // D format F32
  d |= fmt << 7;                     // A format
  d |= fmt << 10;                    // B format
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

struct TcSmem {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  uint32_t* tmem_ptr;
  float* stage;  // TC_EPI_WARPS x (32 rows x 32 fp32) transposition buffers of the epilogue warps
};

constexpr int TC_STAGE_FLOATS = 32 * 32;

// Epilogue transposition.  tcgen05.ld hands every thread ONE accumulator row (TMEM lane), so direct
// global accesses would touch 32 different 128-byte lines per warp instruction and saturate the L1
// tag stage.  Each warp therefore bounces its 32x32 fp32 chunk through shared memory (16-byte slots
// XOR-swizzled by row, conflict free both ways) and continues with the mapping
//   lane -> rows 4*i + lane/8 (i = 0..7), columns 4*(lane%8) .. +3
// in which 8 consecutive lanes cover one contiguous 128-byte row segment.
__device__ __forceinline__ void stage_write(float* sbuf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(sbuf + lane * 32 + ((j ^ (lane & 7)) << 2)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stage_read(const float* sbuf, int lane, int i, float (&o)[4]) {
  int r = 4 * i + (lane >> 3);
  float4 q = *reinterpret_cast<const float4*>(sbuf + r * 32 + (((lane & 7) ^ (r & 7)) << 2));
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}

template <int BN, int STAGES>
__device__ __forceinline__ TcSmem tc_carve(uint8_t* raw) {
  TcSmem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  constexpr int STAGE_BYTES = TC_A_BYTES + BN * 128;
  s.stages = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  s.full = bars;
  s.empty = bars + STAGES;
  s.tmem_full = bars + 2 * STAGES;
  s.tmem_empty = bars + 2 * STAGES + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  s.stage = reinterpret_cast<float*>(bars + 2 * STAGES + 6);  // 16-byte aligned: stages are 1 KB multiples
  return s;
}
template <int BN, int STAGES>
constexpr size_t tc_smem_bytes() {
  return (size_t)STAGES * (TC_A_BYTES + BN * 128) + (2 * STAGES + 6) * 8 + TC_EPI_WARPS * TC_STAGE_FLOATS * 4 + 1024;
}

template <int BN, int STAGES>
__device__ __forceinline__ void tc_setup(const TcSmem& s, int warp, int lane) {
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&s.tmem_full[i], 1); mbar_init(&s.tmem_empty[i], TC_EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(s.tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// ------------------------------------------------------------------------------------------------
// K-major GEMM with fused epilogue
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool PAIRED, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p, const Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem s = tc_carve<BN, STAGES>(smem_raw);
  constexpr int STAGE_BYTES = TC_A_BYTES + BN * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nseg; ++i) prefetch_tmap(&p.a_map[i]);
    prefetch_tmap(&p.b_map);
  }
  tc_setup<BN, STAGES>(s, warp, lane);
  const uint32_t tmem_base = *s.tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt = tile % p.n_tiles, rt = tile / p.n_tiles;
        int b = rt / p.tiles_per_batch, t0 = (rt % p.tiles_per_batch) * TC_BM, n0 = nt * BN;
        for (int sg = 0; sg < p.nseg; ++sg) {
          for (int kb = 0; kb < p.seg_nkb[sg]; ++kb) {
            mbar_wait(&s.empty[stage], phase ^ 1);
            uint8_t* sa = s.stages + stage * STAGE_BYTES;
            mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
            tma_load_3d(sa, &p.a_map[sg], &s.full[stage], kb * TC_BK, t0 + p.seg_shift[sg], b);
            tma_load_2d(sa + TC_A_BYTES, &p.b_map, &s.full[stage], p.seg_koff[sg] + kb * TC_BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int total_kb = 0;
      for (int sg = 0; sg < p.nseg; ++sg) total_kb += p.seg_nkb[sg];
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          uint32_t sa = smem_u32(s.stages + stage * STAGE_BYTES);
          uint64_t adesc = make_smem_desc(sa, p.desc_lbo, p.desc_sbo);
          uint64_t bdesc = make_smem_desc(sa + TC_A_BYTES, p.desc_lbo, p.desc_sbo);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // +32 bytes (>>4 = 2) per K=16 step inside the 128B swizzle row
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&s.empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&s.tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;          // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;  // which half of the tile's columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int nt = tile % p.n_tiles, rt = tile / p.n_tiles;
      int b = rt / p.tiles_per_batch, t0 = (rt % p.tiles_per_batch) * TC_BM, n0 = nt * BN;
      uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      float* sbuf = s.stage + (warp - 2) * TC_STAGE_FLOATS;
      const int tq = t0 + q * 32;                  // first row of this warp's quadrant
      const long long row0 = (long long)b * p.T + tq;
      const int cl = (lane & 7) * 4;                 // column offset inside a 32-wide chunk
      if constexpr (PAIRED) {
        mbar_wait(&s.tmem_full[acc], acc_phase);
        tc_fence_after();
        constexpr int G = BN / 2;
#pragma unroll 1
        for (int c = half * (G / 2); c < (half + 1) * (G / 2); c += 32) {
          float v[32], lo[8][4];
          tmem_ld32(taddr + c, v);
          stage_write(sbuf, lane, v);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) stage_read(sbuf, lane, i, lo[i]);
          __syncwarp();
          tmem_ld32(taddr + G + c, v);
          stage_write(sbuf, lane, v);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float hi[4];
            stage_read(sbuf, lane, i, hi);
            int r = 4 * i + (lane >> 3);
            if (tq + r < p.T) epi.template pair<4>(row0 + r, nt * G + c + cl, lo[i], hi);
          }
          __syncwarp();
        }
      } else {
        // The epilogue's global reads (residual / skip / saved gate values) are issued one chunk
        // AHEAD of the accumulator drain -- the first chunk even before the MMA of this tile has
        // finished -- so each lane keeps 8..16 independent 16-byte loads in flight.
        constexpr int NCH = (BN / 2) / 32;           // chunks per warp
        constexpr int AW = 4 * Epi::kAux;
        constexpr bool AHEAD = (Epi::kAux == 1);     // two aux sets fit the register budget
        float aux[AHEAD ? 2 : 1][8][AW];
        const int cbase = half * (BN / 2);
        auto issue = [&](int k, float (&dst)[8][AW]) {
          int col = n0 + cbase + 32 * k + cl;
          if (col < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              int r = 4 * i + (lane >> 3);
              if (tq + r < p.T) epi.template load<4>(row0 + r, col, dst[i]);
            }
          }
        };
        if constexpr (AHEAD) issue(0, aux[0]);
        mbar_wait(&s.tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          if constexpr (AHEAD) {
            if (k + 1 < NCH) issue(k + 1, aux[(k + 1) & 1]);
          } else {
            issue(k, aux[0]);
          }
          float v[32];
          tmem_ld32(taddr + cbase + 32 * k, v);
          stage_write(sbuf, lane, v);
          __syncwarp();
          float o[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) stage_read(sbuf, lane, i, o[i]);
          __syncwarp();
          int col = n0 + cbase + 32 * k + cl;
          if (col < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              int r = 4 * i + (lane >> 3);
              if (tq + r < p.T) epi.template apply<4>(row0 + r, col, o[i], aux[AHEAD ? (k & 1) : 0][i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// MN-major weight-gradient GEMM: D[m][n] = sum_t A[t][m] * B[t+shift][n], split over time chunks
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem s = tc_carve<BN, STAGES>(smem_raw);
  constexpr int STAGE_BYTES = TC_A_BYTES + BN * 128;
  constexpr int BOX_BYTES = 64 * 128;  // {64 channels, 64 time rows} x 16 bit
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nprob; ++i) { prefetch_tmap(&p.a_map[i]); prefetch_tmap(&p.b_map[i]); }
  }
  tc_setup<BN, STAGES>(s, warp, lane);
  const uint32_t tmem_base = *s.tmem_ptr;
  const int tiles_total = p.tile_begin[p.nprob];

  // work item -> (problem, m tile, n tile, split); tiles vary fastest so that one time chunk of the
  // operands is reused from L2 by neighbouring CTAs
  auto decode = [&](int w, int& pr, int& m0, int& n0, int& split) {
    split = w / tiles_total;
    int tl = w % tiles_total;
    pr = 0;
    while (tl >= p.tile_begin[pr + 1]) ++pr;
    tl -= p.tile_begin[pr];
    m0 = (tl / p.n_tiles_n[pr]) * TC_BM;
    n0 = (tl % p.n_tiles_n[pr]) * BN;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int pr, m0, n0, split;
        decode(w, pr, m0, n0, split);
        int b = split / p.chunks_per_batch, tc0 = (split % p.chunks_per_batch) * p.Lc;
        int nkb = (min(p.Lc, p.T - tc0) + TC_BK - 1) / TC_BK;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&s.empty[stage], phase ^ 1);
          uint8_t* sa = s.stages + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&s.full[stage], STAGE_BYTES);
          int t = tc0 + kb * TC_BK;
#pragma unroll
          for (int j = 0; j < TC_BM / 64; ++j)
            tma_load_3d(sa + j * BOX_BYTES, &p.a_map[pr], &s.full[stage], p.a_c0[pr] + m0 + j * 64, t, b);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_3d(sa + TC_A_BYTES + j * BOX_BYTES, &p.b_map[pr], &s.full[stage], p.b_c0[pr] + n0 + j * 64,
                        t + p.shift[pr], b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int pr, m0, n0, split;
        decode(w, pr, m0, n0, split);
        int tc0 = (split % p.chunks_per_batch) * p.Lc;
        int nkb = (min(p.Lc, p.T - tc0) + TC_BK - 1) / TC_BK;
        mbar_wait(&s.tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          uint32_t sa = smem_u32(s.stages + stage * STAGE_BYTES);
          uint64_t adesc = make_smem_desc(sa, p.desc_lbo, p.desc_sbo);
          uint64_t bdesc = make_smem_desc(sa + TC_A_BYTES, p.desc_lbo, p.desc_sbo);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // 16 time rows x 128 B = 2048 B (>>4 = 128) per K=16 step
            umma_f16(d_tmem, adesc + 128 * k, bdesc + 128 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&s.empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&s.tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      int pr, m0, n0, split;
      decode(w, pr, m0, n0, split);
      mbar_wait(&s.tmem_full[acc], acc_phase);
      tc_fence_after();
      const int mq = m0 + q * 32;
      const int M = p.M[pr], N = p.N[pr];
      float* out = p.partial[pr] + (long long)split * M * N;
      uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      float* sbuf = s.stage + (warp - 2) * TC_STAGE_FLOATS;
      const int cl = (lane & 7) * 4;
#pragma unroll 1
      for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
        float v[32];
        tmem_ld32(taddr + c, v);
        stage_write(sbuf, lane, v);
        __syncwarp();
        int n = n0 + c + cl;
        if (n < N) {  // N is a multiple of 4
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float o[4];
            stage_read(sbuf, lane, i, o);
            int m = mq + 4 * i + (lane >> 3);
            if (m < M) *reinterpret_cast<float4*>(out + (long long)m * N + n) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launches
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();

// 16-bit tensor map over a slab [B][T][C]: dims (C, T, B), box (box_c, box_t, 1), 128B swizzle, zero OOB fill
int make_slab_map(CUtensorMap* m, const void* ptr, int C, int T, int B, int box_c, int box_t, int is_fp16);
// 16-bit tensor map over a matrix [rows][ld]: dims (ld, rows), box (64, box_rows)
int make_matrix_map(CUtensorMap* m, const void* ptr, int ld, int rows, int box_rows, int is_fp16);

template <int BN, bool PAIRED, class Epi>
int tc_gemm_launch_bn(const GemmDesc& d, const Epi& epi, cudaStream_t st, int lbo_override = -1, int sbo_override = -1) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = d.nseg;
  for (int s = 0; s < d.nseg; ++s) {
    CMWG_REQUIRE(d.seg[s].K % TC_BK == 0, "tc_gemm: segment K=%d not a multiple of %d", d.seg[s].K, TC_BK);
    CMWG_PROPAGATE(make_slab_map(&p.a_map[s], d.seg[s].a, d.seg[s].lda, d.T, d.B, TC_BK, TC_BM, d.is_fp16));
    p.seg_nkb[s] = d.seg[s].K / TC_BK;
    p.seg_shift[s] = d.seg[s].shift;
    p.seg_koff[s] = d.seg[s].koff;
  }
  CMWG_PROPAGATE(make_matrix_map(&p.b_map, d.w, d.ldw, d.n_rows_w, BN, d.is_fp16));
  p.B = d.B; p.T = d.T; p.N = d.N;
  p.tiles_per_batch = ceil_div(d.T, TC_BM);
  p.n_tiles = ceil_div(d.N, BN);
  p.total_tiles = d.B * p.tiles_per_batch * p.n_tiles;
  p.idesc = make_idesc(d.is_fp16, TC_BM, BN, 0, 0);
  p.desc_lbo = lbo_override >= 0 ? (uint32_t)lbo_override : 1u;
  p.desc_sbo = sbo_override >= 0 ? (uint32_t)sbo_override : (1024u >> 4);
  if (p.total_tiles == 0) return CMWG_OK;
  auto kern = tc_gemm_kernel<BN, STAGES, PAIRED, Epi>;
  constexpr size_t smem = tc_smem_bytes<BN, STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    CMWG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = std::min(p.total_tiles, num_sms());
  ProfScope prof(st, d.tag);
  kern<<<grid, TC_THREADS, smem, st>>>(p, epi);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

template <bool PAIRED, class Epi>
int tc_gemm_launch(const GemmDesc& d, const Epi& epi, cudaStream_t st) {
  if (d.bn == 256) return tc_gemm_launch_bn<256, PAIRED, Epi>(d, epi, st);
  return tc_gemm_launch_bn<128, PAIRED, Epi>(d, epi, st);
}

int tc_wgrad_launch(const WgradProblem* probs, int nprob, int B, int T, int Lc, int is_fp16, cudaStream_t st,
                    int lbo_override = -1, int sbo_override = -1);

}  // namespace cmwg
