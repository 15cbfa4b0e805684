// Host side of the tcgen05 engine: TMA descriptor construction, weight-gradient launch, self test.
#include "engine_tc.cuh"
#include "epilogues.cuh"

namespace cmwg {

PFN_encodeTiled get_encode_fn() {
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

int make_slab_map(CUtensorMap* m, const void* ptr, int C, int T, int B, int box_c, int box_t, int is_fp16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CMWG_ERR_CUDA;
  CMWG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (C * 2) % 16 == 0,
               "make_slab_map: pointer/row pitch not 16-byte aligned (C=%d)", C);
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_t, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(slab C=%d T=%d B=%d box=%dx%d) failed with CUresult %d", C, T, B, box_c, box_t,
              (int)r);
    return CMWG_ERR_CUDA;
  }
  return CMWG_OK;
}

int make_matrix_map(CUtensorMap* m, const void* ptr, int ld, int rows, int box_rows, int is_fp16) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return CMWG_ERR_CUDA;
  CMWG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0,
               "make_matrix_map: pointer/row pitch not 16-byte aligned (ld=%d)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(matrix ld=%d rows=%d box_rows=%d) failed with CUresult %d", ld, rows, box_rows,
              (int)r);
    return CMWG_ERR_CUDA;
  }
  return CMWG_OK;
}

template <int BN>
static int tc_wgrad_launch_bn(const WgradProblem* probs, int nprob, int B, int T, int Lc, int is_fp16, cudaStream_t st,
                              int lbo_override, int sbo_override) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = nprob;
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    const WgradProblem& q = probs[i];
    CMWG_REQUIRE(q.lda % 8 == 0 && q.ldb % 8 == 0 && q.a_c0 % 8 == 0 && q.b_c0 % 8 == 0,
                 "tc_wgrad: leading dimensions must be multiples of 8");
    CMWG_PROPAGATE(make_slab_map(&p.a_map[i], q.a, q.lda, T, B, 64, TC_BK, is_fp16));
    CMWG_PROPAGATE(make_slab_map(&p.b_map[i], q.b, q.ldb, T, B, 64, TC_BK, is_fp16));
    p.M[i] = q.M; p.N[i] = q.N; p.shift[i] = q.shift; p.a_c0[i] = q.a_c0; p.b_c0[i] = q.b_c0;
    p.partial[i] = q.partial;
    p.n_tiles_n[i] = ceil_div(q.N, BN);
    p.tile_begin[i] = tiles;
    tiles += ceil_div(q.M, TC_BM) * p.n_tiles_n[i];
  }
  p.tile_begin[nprob] = tiles;
  p.B = B; p.T = T; p.Lc = Lc;
  p.chunks_per_batch = ceil_div(T, Lc);
  p.splits = B * p.chunks_per_batch;
  p.total_work = tiles * p.splits;
  p.idesc = make_idesc(is_fp16, TC_BM, BN, 1, 1);
  p.desc_lbo = lbo_override >= 0 ? (uint32_t)lbo_override : (8192u >> 4);
  p.desc_sbo = sbo_override >= 0 ? (uint32_t)sbo_override : (1024u >> 4);
  if (p.total_work == 0) return CMWG_OK;
  auto kern = tc_wgrad_kernel<BN, STAGES>;
  constexpr size_t smem = tc_smem_bytes<BN, STAGES>();
  static bool attr_set = false;
  if (!attr_set) {
    CMWG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int grid = std::min(p.total_work, num_sms());
  ProfScope prof(st, CMWG_KCLASS_WGRAD);
  kern<<<grid, TC_THREADS, smem, st>>>(p);
  CMWG_COUNT_LAUNCH();
  CMWG_LAUNCH_CHECK();
  return CMWG_OK;
}

int tc_wgrad_launch(const WgradProblem* probs, int nprob, int B, int T, int Lc, int is_fp16, cudaStream_t st,
                    int lbo_override, int sbo_override) {
  CMWG_REQUIRE(nprob >= 1 && nprob <= TC_MAX_WG, "tc_wgrad: %d problems (max %d)", nprob, TC_MAX_WG);
  CMWG_REQUIRE(Lc % TC_BK == 0, "tc_wgrad: chunk length %d not a multiple of %d", Lc, TC_BK);
  // group by N tile width
  WgradProblem big[TC_MAX_WG], small[TC_MAX_WG];
  int nb = 0, ns = 0;
  for (int i = 0; i < nprob; ++i) {
    if (probs[i].N >= 256) big[nb++] = probs[i];
    else small[ns++] = probs[i];
  }
  if (nb) CMWG_PROPAGATE(tc_wgrad_launch_bn<256>(big, nb, B, T, Lc, is_fp16, st, lbo_override, sbo_override));
  if (ns) CMWG_PROPAGATE(tc_wgrad_launch_bn<128>(small, ns, B, T, Lc, is_fp16, st, lbo_override, sbo_override));
  return CMWG_OK;
}

}  // namespace cmwg

using namespace cmwg;

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
    StoreEpi epi{d, N, N};
    return tc_gemm_launch<false, StoreEpi>(g, epi, st);
  }
  WgradProblem pr;
  pr.a = a; pr.lda = M; pr.a_c0 = 0; pr.M = M;
  pr.b = b; pr.ldb = N; pr.b_c0 = 0; pr.N = N;
  pr.shift = 0; pr.partial = d;
  int Lc = round_up(K, 64);
  if (variant & 4) return tc_wgrad_launch(&pr, 1, 1, K, Lc, is_fp16, st, 1024 >> 4, 8192 >> 4);
  return tc_wgrad_launch(&pr, 1, 1, K, Lc, is_fp16, st);
}
