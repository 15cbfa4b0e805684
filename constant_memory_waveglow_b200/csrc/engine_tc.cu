// Host side of the tcgen05 engine: TMA descriptor construction (cached), pair launches, the
// weight-gradient launch and the self test.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "engine_tc.cuh"
#include "epilogues_tc.cuh"

namespace cmwg {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// ---- tensor-map cache: a training step re-encodes the same few hundred descriptors every step -----
struct MapKey {
  const void* ptr;
  int d0, d1, d2, d3, ld, b0, b1, flags;  // dims (d0 innermost), row pitch in elements, box, dtype/swizzle/rank
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && ld == o.ld && b0 == o.b0 &&
           b1 == o.b1 && flags == o.flags;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    auto mix = [&](int v) { h ^= (size_t)(unsigned)v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.d2); mix(k.d3); mix(k.ld); mix(k.b0); mix(k.b1); mix(k.flags);
    return h;
  }
};
static std::mutex g_map_mu;
static unsigned long long g_map_misses = 0, g_map_clears = 0;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

static bool dbg_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] == '1';
}

static int encode_cached(CUtensorMap* m, const MapKey& key, CUtensorMapDataType dt, int esize, int rank,
                         CUtensorMapSwizzle sw) {
  static const bool nocache = dbg_flag("CMWG_DEBUG_NOCACHE");
  if (!nocache) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *m = it->second;
      return CMWG_OK;
    }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CMWG_ERR_CUDA;
  ++g_map_misses;
  CMWG_REQUIRE((reinterpret_cast<uintptr_t>(key.ptr) & 15) == 0 && ((long long)key.ld * esize) % 16 == 0,
               "tensor map: pointer/row pitch not 16-byte aligned (pitch %d elements)", key.ld);
  cuuint64_t dims[4] = {(cuuint64_t)key.d0, (cuuint64_t)key.d1, (cuuint64_t)(rank >= 3 ? key.d2 : 1),
                        (cuuint64_t)(rank >= 4 ? key.d3 : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)key.ld * esize, (cuuint64_t)key.d1 * key.ld * esize,
                           (cuuint64_t)key.d2 * key.d1 * key.ld * esize};
  cuuint32_t box[4] = {(cuuint32_t)key.b0, (cuuint32_t)key.b1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dt, rank, const_cast<void*>(key.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(dims %d x %d x %d x %d, box %d x %d, flags %d) failed with CUresult %d", key.d0,
              key.d1, key.d2, key.d3, key.b0, key.b1, key.flags, (int)r);
    return CMWG_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  // a training step uses ~5000 descriptors; clearing costs one step of re-encoding (~60 ms), so the bound is generous
  // (2^18 entries ~ 50 MB of host memory) and only reached by workloads whose buffer addresses keep changing
  if (g_map_cache.size() > (1u << 18)) { g_map_cache.clear(); ++g_map_clears; }
  g_map_cache.emplace(key, *m);
  return CMWG_OK;
}

int get_slab_map(CUtensorMap* m, const void* ptr, int C, int ld, int T, int H, int B, int box_c, int box_t,
                 int is_fp16, int kind) {
  MapKey key{ptr, C, T, H, B, ld, box_c, box_t, (kind << 4) | (is_fp16 ? 1 : 0) | 2};
  if (kind == TC_MAP_CHUNK32)
    return encode_cached(m, key, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 4, CU_TENSOR_MAP_SWIZZLE_128B);
  return encode_cached(m, key, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 4,
                       kind == TC_MAP_CHUNK16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int get_matrix_map(CUtensorMap* m, const void* ptr, int ld, int rows, int box_rows, int is_fp16) {
  MapKey key{ptr, ld, rows, 1, 1, ld, TC_BK, box_rows, (is_fp16 ? 1 : 0)};
  return encode_cached(m, key, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 2,
                       CU_TENSOR_MAP_SWIZZLE_128B);
}

int tc_launch_pairs(const void* kern, size_t smem, int pairs, void** args, cudaStream_t st, int threads) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see common.cuh: both tc kernels pdl_wait()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CMWG_CHECK_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  static const bool dbg_sync = dbg_flag("CMWG_DEBUG_SYNC");
  if (dbg_sync) CMWG_CHECK_CUDA(cudaStreamSynchronize(st));
  return CMWG_OK;
}

// Task kernels (engine_mega.cuh) spin on counters that OTHER CTAs of the same grid advance: every CTA must be resident.
// On one stream that holds by construction (grid <= SM count / 2 pairs, one CTA per SM, and whatever else runs -- NCCL's
// kernels -- finishes on its own).  Two task kernels on DIFFERENT streams could each hold part of the machine and wait for
// the rest; CMWG_COOP=1 launches them cooperatively (the driver places the whole grid at once or not at all; no
// programmatic dependent launch then).  It is opt-in because Nsight Compute cannot replay cooperative cluster launches
// (LaunchFailed), and every measurement here goes through it.
int tc_launch_pairs_coresident(const void* kern, size_t smem, int pairs, void** args, cudaStream_t st, int threads) {
  const char* v = getenv("CMWG_COOP");
  if (!(v && v[0] == '1')) return tc_launch_pairs(kern, smem, pairs, args, st, threads);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  CMWG_CHECK_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  static const bool dbg_sync = dbg_flag("CMWG_DEBUG_SYNC");
  if (dbg_sync) CMWG_CHECK_CUDA(cudaStreamSynchronize(st));
  return CMWG_OK;
}

constexpr int TC_PLAN_PAIRS = 74;  // CTA pairs of a B200 (148 SMs); only load balance depends on it

static int wgrad_group_splits(const WgradProblem* probs, int nprob, int bn, int B, int T) {  // B = lines
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) tiles += ceil_div(probs[i].M, 2 * TC_BM) * ceil_div(probs[i].N, bn);
  const int total_units = B * ceil_div(T, TC_BK);
  if (tiles <= 0) return 1;
  // The launch takes ceil(tiles * s / pairs) rounds of ceil(units / s) k-blocks: 49 tiles on 74 pairs are ONE round of the
  // full K with s = 1 (a third of the pairs idle), two rounds of a third of K with s = 3.  Every split costs a partial tile
  // (written, then read by the reduce pass): charged as a few k-blocks so that equal round counts prefer fewer splits.
  const int smax = std::min(total_units, 80);
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= smax; ++s) {
    const long long rounds = ceil_div(tiles * s, TC_PLAN_PAIRS);
    const long long cost = rounds * ceil_div(total_units, s) + 8ll * s * rounds;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  // whole units per split; drop splits that would be empty
  const int ups = ceil_div(total_units, best);
  return ceil_div(total_units, ups);
}

void tc_wgrad_plan(const WgradProblem* probs, int nprob, int B, int H, int T, int force_splits, int* splits_out) {
  B *= (H > 0 ? H : 1);
  WgradProblem big[TC_MAX_WG], small[TC_MAX_WG];
  int nb = 0, ns = 0;
  for (int i = 0; i < nprob; ++i) {
    if (probs[i].N >= 256) big[nb++] = probs[i];
    else small[ns++] = probs[i];
  }
  const int total_units = B * ceil_div(T, TC_BK);
  int fs = force_splits > total_units ? total_units : force_splits;
  if (fs > 0) fs = ceil_div(total_units, ceil_div(total_units, fs));
  const int sb = fs > 0 ? fs : wgrad_group_splits(big, nb, 256, B, T);
  const int ss = fs > 0 ? fs : wgrad_group_splits(small, ns, 128, B, T);
  for (int i = 0; i < nprob; ++i) splits_out[i] = probs[i].N >= 256 ? sb : ss;
}

template <int BN>
static int tc_wgrad_launch_bn(const WgradProblem* probs, int nprob, int B, int H, int T, int splits, int is_fp16,
                              cudaStream_t st, int lbo_override, int sbo_override) {
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = nprob;
  p.B = B; p.T = T; p.H = H;
  p.units_per_batch = ceil_div(T, TC_BK);
  p.total_units = B * H * p.units_per_batch;
  p.splits = splits;
  p.units_per_split = ceil_div(p.total_units, splits);
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    const WgradProblem& q = probs[i];
    CMWG_REQUIRE(q.lda % 8 == 0 && q.ldb % 8 == 0 && q.a_c0 % 8 == 0 && q.b_c0 % 8 == 0 && q.N % 4 == 0,
                 "tc_wgrad: leading dimensions must be multiples of 8");
    CMWG_PROPAGATE(get_slab_map(&p.a_map[i], q.a, q.lda, q.lda, T, H, B, 64, TC_BK, is_fp16, TC_MAP_OPERAND));
    CMWG_PROPAGATE(get_slab_map(&p.b_map[i], q.b, q.ldb, q.ldb, T, q.bcast_h ? 1 : H, B, 64, TC_BK, is_fp16,
                                TC_MAP_OPERAND));
    CMWG_PROPAGATE(get_slab_map(&p.out_map[i], q.partial, q.N, q.N, q.M, 1, p.splits, 32, 32, 0, TC_MAP_CHUNK32));
    p.M[i] = q.M; p.N[i] = q.N; p.shift[i] = q.shift; p.a_c0[i] = q.a_c0; p.b_c0[i] = q.b_c0;
    p.shift_h[i] = q.shift_h; p.bcast[i] = q.bcast_h;
    p.n_tiles_n[i] = ceil_div(q.N, BN);
    p.tile_begin[i] = tiles;
    tiles += ceil_div(q.M, 2 * TC_BM) * p.n_tiles_n[i];
  }
  p.tile_begin[nprob] = tiles;
  p.total_work = tiles * p.splits;
  p.idesc = make_idesc(is_fp16, 2 * TC_BM, BN, 1, 1);
  p.desc_lbo = lbo_override >= 0 ? (uint32_t)lbo_override : (8192u >> 4);
  p.desc_sbo = sbo_override >= 0 ? (uint32_t)sbo_override : (1024u >> 4);
  if (p.total_work == 0) return CMWG_OK;
  auto kern = tc_wgrad_kernel<BN>;
  constexpr size_t smem = tc_smem_bytes<BN, WgradEpiShape>();
  static_assert(smem <= TC_SMEM_LIMIT, "shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    CMWG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int pairs = std::min(p.total_work, num_sms() / 2);
  ProfScope prof(st, CMWG_KCLASS_WGRAD);
  void* args[1] = {(void*)&p};
  CMWG_PROPAGATE(tc_launch_pairs((const void*)kern, smem, pairs, args, st));
  CMWG_COUNT_LAUNCH();
  return CMWG_OK;
}

int tc_wgrad_launch(const WgradProblem* probs, int nprob, int B, int H, int T, int is_fp16, cudaStream_t st,
                    int force_splits, int lbo_override, int sbo_override) {
  CMWG_REQUIRE(nprob >= 1 && nprob <= TC_MAX_WG, "tc_wgrad: %d problems (max %d)", nprob, TC_MAX_WG);
  if (H < 1) H = 1;
  int splits[TC_MAX_WG];
  tc_wgrad_plan(probs, nprob, B, H, T, force_splits, splits);
  // group by N tile width
  WgradProblem big[TC_MAX_WG], small[TC_MAX_WG];
  int nb = 0, ns = 0, sb = 1, ss = 1;
  for (int i = 0; i < nprob; ++i) {
    if (probs[i].N >= 256) { big[nb++] = probs[i]; sb = splits[i]; }
    else { small[ns++] = probs[i]; ss = splits[i]; }
  }
  if (nb) CMWG_PROPAGATE(tc_wgrad_launch_bn<256>(big, nb, B, H, T, sb, is_fp16, st, lbo_override, sbo_override));
  if (ns) CMWG_PROPAGATE(tc_wgrad_launch_bn<128>(small, ns, B, H, T, ss, is_fp16, st, lbo_override, sbo_override));
  return CMWG_OK;
}

}  // namespace cmwg

using namespace cmwg;

extern "C" int cmwg_wgrad_plan_splits(int tiles, int bn, int B, int T) {
  if (tiles < 1 || B < 1 || T < 1 || (bn != 128 && bn != 256)) return -1;
  WgradProblem pr[TC_MAX_WG];
  // `tiles` problems of one (256 x bn) tile each
  const int n = tiles < TC_MAX_WG ? tiles : TC_MAX_WG;
  for (int i = 0; i < n; ++i) { pr[i].M = 2 * TC_BM; pr[i].N = bn; }
  if (tiles > TC_MAX_WG) pr[0].M = 2 * TC_BM * (tiles - TC_MAX_WG + 1);   // the rest as more row tiles of the first problem
  return wgrad_group_splits(pr, n, bn, B, T);
}

extern "C" unsigned long long cmwg_debug_counter(int which) {
  return which == 0 ? cmwg::g_map_misses : cmwg::g_map_clears;
}

extern "C" int cmwg_selftest_tc_gemm(const void* a, const void* b, float* d, int M, int N, int K, int is_fp16,
                                     int variant, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // variant & 1      : 0 = K-major (A: M x K, B: N x K), 1 = MN-major (A: K x M, B: K x N)
  // variant & 2      : force BN = 128
  // variant & 4      : MN-major only: swap the LBO / SBO descriptor fields (diagnostic)
  if ((variant & 1) == 0) {
    CMWG_REQUIRE(K % 64 == 0, "selftest: K must be a multiple of 64");
    GemmDesc g;
    memset(&g, 0, sizeof(g));
    g.nseg = 1;
    g.seg[0].a = a; g.seg[0].lda = K; g.seg[0].K = K; g.seg[0].shift = 0; g.seg[0].koff = 0;
    g.w = b; g.ldw = K; g.N = N; g.n_rows_w = N; g.B = 1; g.T = M; g.is_fp16 = is_fp16;
    g.bn = ((variant & 2) || N < 256) ? 128 : 256;
    TcIo io;
    memset(&io, 0, sizeof(io));
    io.out[0] = TcStream{d, N, N, 1};
    StoreTcEpi epi{nullptr, nullptr};
    return tc_gemm_launch<StoreTcEpi>(g, io, epi, st);
  }
  WgradProblem pr;
  pr.a = a; pr.lda = M; pr.a_c0 = 0; pr.M = M;
  pr.b = b; pr.ldb = N; pr.b_c0 = 0; pr.N = N;
  pr.shift = 0; pr.shift_h = 0; pr.bcast_h = 0; pr.partial = d;
  if (variant & 4) return tc_wgrad_launch(&pr, 1, 1, 1, K, is_fp16, st, 1, 1024 >> 4, 8192 >> 4);
  return tc_wgrad_launch(&pr, 1, 1, 1, K, is_fp16, st, 1);
}
