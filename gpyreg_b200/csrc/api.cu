// C ABI of gpyreg_b200 (see include/gpyreg_b200.h): host orchestration of the batched
// GP hot path on one B200.  One translation unit; build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include "../../include/gpyreg_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "chol.cuh"
#include "common.cuh"
#include "cov.cuh"
#include "gemm.cuh"
#include "predict.cuh"
#include "append.cuh"

using namespace gpb;

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
struct Bufs {              // device storage for `cap` batch slots
  // tensor maps of the four operand arrays [0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf] x [box of 132 / 68 rows]
  CUtensorMap tm[4][2];
  bool tm_ok = false;
  int cap = 0;
  bool has_w = false;
  bool cap_is_max = false;   // cap is limited by device memory, not by the request
  // W = L^-1 is its own allocation with its own capacity (slots 0 .. wcap-1): a gradient call of a few rows
  // after a large nlZ-only batch adds a small W next to the big arena instead of releasing and
  // re-partitioning it (cudaFree + cudaMalloc of ~100 GB was 0.6-1.0 s of a 10 s cfg3 fit)
  int wcap = 0;
  bool wcap_is_max = false;
  int Np = 0, Nt = 0, D = 0, P = 0, cov_n = 0;
  double *Abuf = nullptr, *Wbuf = nullptr, *Dbuf = nullptr, *DTbuf = nullptr;
  double *xs = nullptr, *resid = nullptr, *sn2v = nullptr, *bvec = nullptr, *zvec = nullptr,
         *alpha = nullptr, *logdet = nullptr, *mult = nullptr, *fmult = nullptr, *hyp = nullptr, *nlz = nullptr,
         *dnlz = nullptr, *gpart = nullptr;
  SlotP* sp = nullptr;
  int *fail = nullptr, *sel = nullptr, *sel2 = nullptr, *sel3 = nullptr, *sel4 = nullptr;
  double* prep_min = nullptr;        // prep_kernel scratch (per-CTA partial results)
  int *prep_nan = nullptr, *prep_ticket = nullptr;
  long long smat() const { return (long long)Np * Np; }
};

struct gpb_post;
struct gpb_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // look-ahead of the blocked Cholesky: the part of an outer trailing update the next outer
  // block does not touch runs on `aux` while the main stream factors that block
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int trtri = 2;             // env GPB_TRTRI: 2 recursive halving (default; as fast as the recurrence for large
                             // batches, 4-5x faster for one matrix), 0 column recurrence, 1 recursive up to trtri_max
  int trtri_max = 8;         // env GPB_TRTRI_MAX
  int lookahead = 2;         // env GPB_LOOKAHEAD: 0 off, 1 batches <= 8 only, 2 always (default: -10 % at B=9, -4 % at B=32,
                             // neutral at B=64)
  long long la_wide = 1500;  // env GPB_LA_WIDE: matrices x (remaining tile columns)^2 above which blocks stay wide
                             // (re-swept in round 2 with the faster chain: profiles/r02_small_batch_latency.txt)
  int la_chunk = 1 << 30;    // logical tiles per look-ahead launch (env GPB_LA_CHUNK; measured: uncut is best)
  int la_ob = 2;             // narrow outer block used with look-ahead (env GPB_LA_OB; 0 = 1 for <= 2 matrices, else 2)
  std::string err;
  Model md{};
  bool has_model = false, has_data = false;
  long long N = 0;
  int D = 0, Np = 0, Nt = 0;
  double *dX = nullptr, *dy = nullptr, *ds2 = nullptr;
  size_t ws_limit = 0;
  int gemm_bn = 64;          // 64: two CTAs per SM (default); 128: one (env GPB_GEMM_BN)
  int loader = 4;            // 0 cp.async (LDGSTS) | 1 TMA bulk copies + mbarriers | 2 bulk for launches that fill the
                             // GPU, cp.async for small ones | 3 tensor-map TMA | 4 (default) tensor-map TMA for launches
                             // that fill the GPU, cp.async for small ones (env GPB_LOADER=cpasync|tma|auto_bulk|tensor|auto)
  int outer_block = 4;       // tile columns per outer block of the two-level Cholesky (env GPB_OUTER_BLOCK)
  long long* diag_dbg = nullptr;   // env GPB_DIAG_DBG: phase clock stamps of the diagonal kernel
  Bufs ws;
  // Factor cache of the last single-chunk nlZ-only call: slot s still holds the Cholesky factor of
  // hyperparameter row s.  A later row that differs only in its MEAN hyperparameters re-uses it and
  // replays the O(N^2) forward solve (a slice sampler moves one coordinate at a time).
  struct {
    bool valid = false;
    int n = 0;
    std::vector<double> key;     // n x (cov_N + noise_N)
    std::vector<char> ok;
    long long hits = 0, misses = 0;
  } cache;
  bool cache_enabled = true;    // env GPB_NLZ_CACHE=0 disables
  bool grad_rows = true;        // env GPB_GRAD_ROWS=0: gradient by grad_kernel for every shape (A/B)
  int left_looking = 2;         // env GPB_LEFT: left-looking column updates inside wide outer blocks -- 0 never,
                                // 1 always, 2 (default) for batches of >= 24 matrices of >= 16 tile columns, with outer
                                // blocks of 8 tile columns (measured at N=5000: potrf of 64 matrices 89.2 -> 87.5 ms;
                                // slower below 16 matrices, where the chain diag -> panel -> next column decides, and
                                // for small matrices: N=1000, 32 matrices nlZ 1.10 -> 1.28 ms, N=1500 x 64 3.69 -> 3.83)
  bool outer_block_set = false; // GPB_OUTER_BLOCK given: use it for every batch size
  double timings[6] = {0, 0, 0, 0, 0, 0};
  long long launches = 0;
  cudaEvent_t ev[8] = {};
  // predict scratch
  double *pXs = nullptr, *pys = nullptr, *ps2s = nullptr, *pBt = nullptr, *pmu = nullptr,
         *pv = nullptr, *psamp = nullptr, *pout = nullptr;
  size_t pXs_n = 0, pBt_n = 0, ppart_n = 0, psamp_n = 0, pout_n = 0;
  // live posterior batches created from this context: they hold a pointer back to it, so
  // gpb_destroy releases the ones the caller has not freed (their handles die with the context)
  std::vector<gpb_post*> posts;
  const Bufs* cur = nullptr;     // buffers the tile GEMMs being enqueued work on (tensor maps for LOADER 2)
  bool quarter_tiles = true;     // env GPB_QUARTER=0: no 64x64 CTAs for small trailing updates
  bool pdl = true;               // env GPB_PDL=0: ordinary launches for the tile GEMMs and the diagonal kernel
};

struct gpb_post {
  gpb_ctx* ctx = nullptr;
  Bufs b;                   // b.cap == number of samples
  long long N = 0;
  Model md{};
  std::vector<SlotP> sp;    // host copy
  std::vector<int> status;  // 1 = Cholesky failed
  std::vector<double> hyp;  // host copy (B,P)
  bool w_ready = false;
  double* X = nullptr;      // device copy of the training inputs the factors belong to (N, D)
};

static std::string g_create_err;

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      char buf_[512];                                                                  \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
               cudaGetErrorString(e_));                                                \
      ctx->err = buf_;                                                                 \
      return GPB_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

// like CK, for functions that own temporaries released by a local `cleanup` lambda
#define CKC(call)                                   \
  do {                                              \
    cudaError_t e_ = (call);                        \
    if (e_ != cudaSuccess) {                        \
      ctx->err = cudaGetErrorString(e_);            \
      cleanup();                                    \
      return GPB_ECUDA;                             \
    }                                               \
  } while (0)

#define FAIL(code, msg) \
  do {                  \
    ctx->err = (msg);   \
    return (code);      \
  } while (0)

#define LAUNCHED(ctx) ((ctx)->launches++)

// NVTX range over the enclosing scope: labels the phases K1..K5 (and the jitter-retry loop) in
// nsys / ncu timelines
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

static inline int round_up(long long v, int m) { return (int)(((v + m - 1) / m) * m); }
static inline unsigned grid1d(long long n, int block = 256) {
  return (unsigned)std::min<long long>((n + block - 1) / block, 148LL * 16);
}

// CTA shapes of the tile GEMM (see gemm.cuh)
enum Shape { SHAPE_SPLIT_N = 0, SHAPE_SPLIT_M = 1, SHAPE_WIDE = 2 };

template <class Op, int BM_, int BN_>
static cudaError_t gemm_attr_shape() {
  cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<Op, BM_, BN_, 0>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)gemm_smem<BM_, BN_>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_nt_kernel<Op, BM_, BN_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)gemm_smem<BM_, BN_>());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_nt_kernel<Op, BM_, BN_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)gemm_smem<BM_, BN_>());
}
template <class Op>
static cudaError_t gemm_attr() {
  cudaError_t e = gemm_attr_shape<Op, 128, 128>();
  if (e != cudaSuccess) return e;
  return gemm_attr_shape<Op, 128, 64>();
}

static bool encode_map(CUtensorMap* m, double* base, long long rows, long long cols, int box_rows);

// Launch on ctx->stream; with ctx->pdl the kernel carries the programmatic-stream-serialization
// attribute (common.cuh: pdl_wait / pdl_launch), so its launch latency and prologue overlap the tail
// of its predecessor.  Only kernels that call pdl_wait() before their first dependent access may come
// through here.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(gpb_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// tensor maps of the operands of Op on the buffers the context is working on
template <class Op, int BM_, int BN_>
static bool tma_operands(const gpb_ctx* ctx, TmaOperands& o) {
  const Bufs* b = ctx->cur;
  if (Op::TMA_A < 0 || !b || !b->tm_ok) return false;
  if ((BM_ != BM && BM_ != BM / 2) || (BN_ != BN && BN_ != BN / 2)) return false;   // boxes exist for 128 and 64 rows
  double* bases[4] = {b->Abuf, b->Wbuf, b->Dbuf, b->DTbuf};
  const long long lds[4] = {b->Np, b->Np, T, T};
  constexpr int ia = (BM_ == BM) ? 0 : 1, ib = (BN_ == BN) ? 0 : 1;
  const int src[4] = {Op::TMA_A, Op::TMA_B, Op::TMA_A0 >= 0 ? Op::TMA_A0 : Op::TMA_A,
                      Op::TMA_B0 >= 0 ? Op::TMA_B0 : Op::TMA_B};
  for (int i = 0; i < 4; ++i)
    if (!bases[src[i]]) return false;
  o.a = b->tm[src[0]][ia];
  o.b = b->tm[src[1]][ib];
  o.a0 = b->tm[src[2]][ia];
  o.b0 = b->tm[src[3]][ib];
  for (int i = 0; i < 4; ++i) { o.base[i] = bases[src[i]]; o.ld[i] = lds[src[i]]; }
  o.b_alt = o.b; o.base[4] = o.base[1]; o.ld[4] = o.ld[1];
  return true;
}

// loader: 0 cp.async, 1 TMA bulk copies, 2 tensor-map TMA (falls back to 1 where the op has no maps)
template <class Op, int BM_, int BN_>
static void launch_shape(gpb_ctx* ctx, const Op& op, dim3 grid, int loader, const TmaOperands* custom = nullptr) {
  grid.x *= (BM / BM_) * (BN / BN_);
  TmaOperands tmo;
  if (loader == 2 && custom) tmo = *custom;
  else if (loader == 2 && !tma_operands<Op, BM_, BN_>(ctx, tmo)) loader = 1;
  if (loader == 2) launch_chain(ctx, gemm_nt_kernel<Op, BM_, BN_, 2>, grid, dim3(GEMM_THREADS), gemm_smem<BM_, BN_>(), op, tmo);
  else if (loader == 1) launch_chain(ctx, gemm_nt_kernel<Op, BM_, BN_, 1>, grid, dim3(GEMM_THREADS), gemm_smem<BM_, BN_>(), op, NoTma{});
  else launch_chain(ctx, gemm_nt_kernel<Op, BM_, BN_, 0>, grid, dim3(GEMM_THREADS), gemm_smem<BM_, BN_>(), op, NoTma{});
  LAUNCHED(ctx);
}

// ctx->loader: 0 cp.async | 1 bulk | 2 auto: bulk for launches that fill the GPU | 3 tensor maps | 4 auto with
// tensor maps (default).  TMA wins when the launch fills the machine (>= 2 CTAs per SM); a lone CTA on an SM
// hides latency better with per-thread cp.async (measured: B=1 triangular inverse)
static int pick_loader(const gpb_ctx* ctx, long long ctas) {
  switch (ctx->loader) {
    case 0: return 0;
    case 1: return 1;
    case 3: return 2;
    case 2: return ctas >= 2 * 148 ? 1 : 0;
    default: return ctas >= 2 * 148 ? 2 : 0;
  }
}

// grid.x counts logical 128x128 tiles; the split shapes launch two CTAs per tile.
template <class Op>
static void launch_gemm(gpb_ctx* ctx, const Op& op, dim3 grid, const TmaOperands* custom = nullptr) {
  if (Op::SLOT_MAJOR) grid = dim3(grid.y, grid.x);      // slot in x, tile in y (see gemm.cuh)
  const bool two = ctx->gemm_bn != 128;
  const long long ctas = (long long)grid.x * grid.y * grid.z * (two ? 2 : 1);
  const int loader = pick_loader(ctx, ctas);
  if constexpr (std::is_same<Op, OpSyrk>::value) {
    // a trailing update of a few tiles sits on the dependent chain of a small batch (diag -> panel ->
    // next column -> diag): quarter tiles put four CTAs on each, same arithmetic per element.  Only while
    // the quarters still fit one per SM: with two matrices the early steps are bound by the trailing
    // updates on the second stream, and more, smaller chain CTAs only queue behind them (measured)
    if (two && ctx->quarter_tiles && ctas * 2 <= 148) {
      launch_shape<Op, 64, 64>(ctx, op, grid, loader, custom);
      return;
    }
  }
  if (two) launch_shape<Op, 128, 64>(ctx, op, grid, loader, custom);
  else launch_shape<Op, 128, 128>(ctx, op, grid, loader, custom);
}

// predict / quad: A = the Ks scratch (Mcp rows x (samples*Np) columns), B = W or the explicit inverse
// of the posterior batch, chosen per sample
static bool pred_tma(const gpb_ctx* ctx, const Bufs& b, const double* Bt, int Mcp, long long cols, TmaOperands& o) {
  if (!b.tm_ok || !b.Wbuf) return false;
  const int ib = (ctx->gemm_bn != 128) ? 1 : 0;
  if (!encode_map(&o.a, const_cast<double*>(Bt), Mcp, cols, BM + 4)) return false;
  o.a0 = o.a;
  o.b = b.tm[1][ib];
  o.b0 = o.b;
  o.b_alt = b.tm[0][ib];
  o.base[0] = o.base[2] = Bt; o.ld[0] = o.ld[2] = Mcp;
  o.base[1] = o.base[3] = b.Wbuf; o.ld[1] = o.ld[3] = b.Np;
  o.base[4] = b.Abuf; o.ld[4] = b.Np;
  return true;
}

// the in-place potrf panel: row halves (64x128) keep it race-free with two CTAs per SM
static void launch_panel(gpb_ctx* ctx, const OpPanel& op, dim3 grid) {
  const bool two = ctx->gemm_bn != 128;
  const long long ctas = (long long)grid.x * grid.y * (two ? 2 : 1);
  const int loader = pick_loader(ctx, ctas);
  // a panel of a few tiles sits on the dependent chain of a small batch: row quarters put four CTAs
  // on each tile (the fused forward-substitution sum is shape-independent, see gemm.cuh)
  if (two && ctx->quarter_tiles && ctas * 2 <= 148) launch_shape<OpPanel, 32, 128>(ctx, op, grid, loader);
  else if (two) launch_shape<OpPanel, 64, 128>(ctx, op, grid, loader);
  else launch_shape<OpPanel, 128, 128>(ctx, op, grid, loader);
}

constexpr int COV_SMEM_MAX = (2 * MAXD * T + 2 * T + 8 * (MAXD + 2) + 8 * T) * 8;

template <int KIND>
static cudaError_t kind_attrs() {
  cudaError_t e;
  e = cudaFuncSetAttribute(build_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(ks_build_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  if (e) return e;
  e = cudaFuncSetAttribute(grad_kernel<KIND, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, COV_SMEM_MAX);
  return e;
}

static int init_attrs(gpb_ctx* ctx) {
  CK(gemm_attr<OpGeneric>());
  CK(gemm_attr<OpPlain>());
  CK(gemm_attr<OpPanel>());
  CK(gemm_attr<OpFwdZ>());
  CK((gemm_attr_shape<OpFwd, 64, 128>()));
  CK((gemm_attr_shape<OpFwdZ, 64, 128>()));
  CK((gemm_attr_shape<OpFwd, 128, 128>()));
  CK((gemm_attr_shape<OpPanel, 64, 128>()));
  CK((gemm_attr_shape<OpPanel, 32, 128>()));
  CK(gemm_attr<OpSyrk>());
  CK((gemm_attr_shape<OpSyrk, 64, 64>()));
  CK(gemm_attr<OpHpass>());
  CK(gemm_attr<OpWrec>());
  CK(gemm_attr<OpRecX>());
  CK(gemm_attr<OpRecW>());
  CK(gemm_attr<OpSyrk2>());
  CK(gemm_attr<OpPred>());
  CK(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
  CK(cudaFuncSetAttribute(quad_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (3 * MAXD * T + 10 * T) * 8));
  CK(kind_attrs<0>());
  CK(kind_attrs<1>());
  CK(kind_attrs<3>());
  CK(kind_attrs<5>());
  CK(kind_attrs<2>());
  return GPB_OK;
}

extern "C" int gpb_version(void) { return 100; }

extern "C" int gpb_create(int device, gpb_ctx** out) {
  if (!out) return GPB_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device available: ") + cudaGetErrorString(e);
    return GPB_ECUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device index out of range";
    return GPB_EINVAL;
  }
  gpb_ctx* ctx = new gpb_ctx();
  ctx->device = device;
  e = cudaSetDevice(device);
  int prio_lo = 0, prio_hi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  // the main stream carries the dependent chain of the factorisation: highest priority
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_hi);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice/StreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return GPB_ECUDA;
  }
  ctx->stream = ctx->own_stream;
  e = cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_lo);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    g_create_err = std::string("aux stream/events: ") + cudaGetErrorString(e);
    delete ctx;
    return GPB_ECUDA;
  }
  if (const char* pd = getenv("GPB_PDL")) ctx->pdl = atoi(pd) != 0;
  if (const char* qt = getenv("GPB_QUARTER")) ctx->quarter_tiles = atoi(qt) != 0;
  if (const char* la = getenv("GPB_LOOKAHEAD")) ctx->lookahead = atoi(la);
  if (const char* tr = getenv("GPB_TRTRI")) ctx->trtri = atoi(tr);
  if (const char* tm = getenv("GPB_TRTRI_MAX")) ctx->trtri_max = atoi(tm);
  if (const char* lc = getenv("GPB_LA_CHUNK")) ctx->la_chunk = std::max(1, atoi(lc));
  if (const char* lw = getenv("GPB_LA_WIDE")) ctx->la_wide = atoll(lw);
  if (const char* lo = getenv("GPB_LA_OB")) ctx->la_ob = std::max(0, atoi(lo));
  if (const char* bn = getenv("GPB_GEMM_BN")) ctx->gemm_bn = (atoi(bn) == 128) ? 128 : 64;
  if (const char* ld = getenv("GPB_LOADER"))
    ctx->loader = (strcmp(ld, "tma") == 0) ? 1 : (strcmp(ld, "cpasync") == 0) ? 0 : (strcmp(ld, "auto_bulk") == 0) ? 2
                : (strcmp(ld, "tensor") == 0) ? 3 : 4;
  if (const char* ob = getenv("GPB_OUTER_BLOCK")) { ctx->outer_block = std::max(1, atoi(ob)); ctx->outer_block_set = true; }
  if (const char* nc = getenv("GPB_NLZ_CACHE")) ctx->cache_enabled = atoi(nc) != 0;
  if (const char* gr = getenv("GPB_GRAD_ROWS")) ctx->grad_rows = atoi(gr) != 0;
  if (const char* ll = getenv("GPB_LEFT")) ctx->left_looking = atoi(ll);
  if (getenv("GPB_DIAG_DBG")) cudaMalloc(&ctx->diag_dbg, 64 * sizeof(long long));
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  int rc = init_attrs(ctx);
  if (rc != GPB_OK) {
    g_create_err = ctx->err;
    delete ctx;
    return rc;
  }
  *out = ctx;
  return GPB_OK;
}

static void free_bufs(Bufs& b);
static void release_post(gpb_post* post);

static void free_bufs(Bufs& b) {
  double** ptrs[] = {&b.Abuf, &b.Wbuf, &b.Dbuf, &b.DTbuf, &b.xs, &b.resid, &b.sn2v, &b.bvec, &b.zvec,
                     &b.alpha, &b.logdet, &b.mult, &b.fmult, &b.hyp, &b.nlz, &b.dnlz, &b.gpart};
  for (auto p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (b.sp) cudaFree(b.sp);
  b.sp = nullptr;
  if (b.prep_min) cudaFree(b.prep_min);
  b.prep_min = nullptr;
  int** ip[] = {&b.fail, &b.sel, &b.sel2, &b.sel3, &b.sel4, &b.prep_nan, &b.prep_ticket};
  for (auto p : ip) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  b.cap = 0;
  b.has_w = false;
  b.cap_is_max = false;
  b.wcap = 0;
  b.wcap_is_max = false;
}

extern "C" void gpb_destroy(gpb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  {
    std::vector<gpb_post*> live;
    live.swap(ctx->posts);
    for (gpb_post* p : live) release_post(p);
  }
  free_bufs(ctx->ws);
  double* p[] = {ctx->dX, ctx->dy, ctx->ds2, ctx->pXs, ctx->pys, ctx->ps2s, ctx->pBt, ctx->pmu,
                 ctx->pv, ctx->psamp, ctx->pout};
  for (auto q : p)
    if (q) cudaFree(q);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char* gpb_last_error(const gpb_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int gpb_set_stream(gpb_ctx* ctx, uint64_t s) {
  if (!ctx) return GPB_EINVAL;
  ctx->stream = s ? (cudaStream_t)(uintptr_t)s : ctx->own_stream;
  return GPB_OK;
}

extern "C" int gpb_set_workspace_limit(gpb_ctx* ctx, uint64_t bytes) {
  if (!ctx) return GPB_EINVAL;
  ctx->ws_limit = (size_t)bytes;
  return GPB_OK;
}

extern "C" int64_t gpb_launch_count(const gpb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gpb_cache_stats(const gpb_ctx* ctx, int64_t* hits, int64_t* misses) {
  if (!ctx || !hits || !misses) return GPB_EINVAL;
  *hits = ctx->cache.hits;
  *misses = ctx->cache.misses;
  return GPB_OK;
}

extern "C" int gpb_last_timings(const gpb_ctx* ctx, double out[6]) {
  if (!ctx || !out) return GPB_EINVAL;
  for (int i = 0; i < 6; ++i) out[i] = ctx->timings[i];
  return GPB_OK;
}

static int fill_model(Model& md, int cov_kind, int degree, int ard, int mean_kind, const int nz[3],
                      int D) {
  if (cov_kind < 0 || cov_kind > 2) return GPB_EINVAL;
  if (cov_kind == GPB_COV_MATERN && degree != 1 && degree != 3 && degree != 5) return GPB_EINVAL;
  if (mean_kind < 0 || mean_kind > 2) return GPB_EINVAL;
  if (nz[0] < 0 || nz[0] > 1 || nz[1] < 0 || nz[1] > 2 || nz[2] < 0 || nz[2] > 1) return GPB_EINVAL;
  md.cov_kind = cov_kind;
  md.degree = degree;
  md.ard = ard ? 1 : 0;
  md.mean_kind = mean_kind;
  md.nz0 = nz[0];
  md.nz1 = nz[1];
  md.nz2 = nz[2];
  md.D = D;
  md.cov_n = cov_count(cov_kind, md.ard, D);
  md.noise_n = noise_count(nz[0], nz[1], nz[2]);
  md.mean_n = mean_count(mean_kind, D);
  md.P = md.cov_n + md.noise_n + md.mean_n;
  return GPB_OK;
}

extern "C" int gpb_set_model(gpb_ctx* ctx, int cov_kind, int matern_degree, int ard, int mean_kind,
                             const int noise_flags[3]) {
  if (!ctx || !noise_flags) return GPB_EINVAL;
  Model md{};
  if (fill_model(md, cov_kind, matern_degree, ard, mean_kind, noise_flags, ctx->D) != GPB_OK)
    FAIL(GPB_EINVAL, "gpb_set_model: unsupported model descriptor");
  ctx->md = md;
  ctx->has_model = true;
  ctx->cache.valid = false;
  return GPB_OK;
}

extern "C" int gpb_set_data(gpb_ctx* ctx, const double* X, const double* y, const double* s2,
                            int64_t N, int D) {
  if (!ctx || !X || !y || N <= 0 || D <= 0) return GPB_EINVAL;
  if (D > MAXD) FAIL(GPB_EINVAL, "gpb_set_data: D > 64 is not supported by the fused kernels");
  if (N > 1000000) FAIL(GPB_EINVAL, "gpb_set_data: N too large");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->dX) cudaFree(ctx->dX);
  if (ctx->dy) cudaFree(ctx->dy);
  if (ctx->ds2) cudaFree(ctx->ds2);
  ctx->dX = ctx->dy = ctx->ds2 = nullptr;
  CK(cudaMalloc(&ctx->dX, sizeof(double) * N * D));
  CK(cudaMalloc(&ctx->dy, sizeof(double) * N));
  CK(cudaMemcpy(ctx->dX, X, sizeof(double) * N * D, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->dy, y, sizeof(double) * N, cudaMemcpyHostToDevice));
  if (s2) {
    CK(cudaMalloc(&ctx->ds2, sizeof(double) * N));
    CK(cudaMemcpy(ctx->ds2, s2, sizeof(double) * N, cudaMemcpyHostToDevice));
  }
  const int Np = round_up(N, T);
  if (Np != ctx->Np || D != ctx->D || N != ctx->N) free_bufs(ctx->ws);   // also re-zeroes the padding
  ctx->cache.valid = false;
  ctx->N = N;
  ctx->D = D;
  ctx->Np = Np;
  ctx->Nt = Np / T;
  ctx->has_data = true;
  if (ctx->has_model) {
    const int nz[3] = {ctx->md.nz0, ctx->md.nz1, ctx->md.nz2};
    fill_model(ctx->md, ctx->md.cov_kind, ctx->md.degree, ctx->md.ard, ctx->md.mean_kind, nz, D);
  }
  return GPB_OK;
}

// ---------------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------------
static size_t per_slot_bytes(int Np, int D, int P, int cov_n, bool with_w) {
  const size_t nt = Np / T;
  size_t b = (size_t)Np * Np * 8 * (with_w ? 2 : 1);
  b += 2 * nt * T * T * 8;                     // Dbuf, DTbuf
  b += (size_t)D * Np * 8 + 6 * (size_t)Np * 8;
  b += nt * 8 + (size_t)P * 16 + 64;
  b += nt * (nt + 1) / 2 * (size_t)cov_n * 8;  // gpart
  return b;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (TensorMapEncodeFn)p;
  }();
  return fn;
}

// 2-D map of a column-major array of `rows` x `cols` doubles; box = box_rows x BK
static bool encode_map(CUtensorMap* m, double* base, long long rows, long long cols, int box_rows) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc || !base) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  const cuuint64_t strides[1] = {(cuuint64_t)rows * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)BK};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void build_tensor_maps(Bufs& b) {
  b.tm_ok = false;
  double* bases[4] = {b.Abuf, b.Wbuf, b.Dbuf, b.DTbuf};
  const long long rows[4] = {b.Np, b.Np, T, T};
  const long long cols[4] = {(long long)b.cap * b.Np, (long long)b.wcap * b.Np, (long long)b.cap * b.Nt * T,
                             (long long)b.cap * b.Nt * T};
  bool ok = true;
  for (int i = 0; i < 4; ++i) {
    if (!bases[i]) { memset(b.tm[i], 0, sizeof b.tm[i]); continue; }      // no W: its ops are not launched
    ok = ok && encode_map(&b.tm[i][0], bases[i], rows[i], cols[i], BM + 4);
    ok = ok && encode_map(&b.tm[i][1], bases[i], rows[i], cols[i], BM / 2 + 4);
  }
  b.tm_ok = ok;
}

// (re)allocate W for `wcap` slots of an existing buffer set
static int alloc_w(gpb_ctx* ctx, Bufs& b, int wcap) {
  if (b.Wbuf) cudaFree(b.Wbuf);
  b.Wbuf = nullptr;
  b.wcap = 0;
  b.has_w = false;
  b.wcap_is_max = false;
  const size_t bytes = (size_t)b.Np * b.Np * 8 * (size_t)wcap;
  CK(cudaMalloc(&b.Wbuf, bytes));
  // padded rows/columns of W are never written by the (padding-skipping) kernels: zero once
  CK(cudaMemsetAsync(b.Wbuf, 0, bytes, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  b.wcap = wcap;
  b.has_w = true;
  build_tensor_maps(b);
  return GPB_OK;
}

static int alloc_bufs(gpb_ctx* ctx, Bufs& b, int cap, bool with_w, int Np, int D, const Model& md) {
  free_bufs(b);
  b.Np = Np;
  b.Nt = Np / T;
  b.D = D;
  b.P = md.P;
  b.cov_n = md.cov_n;
  const size_t smat = (size_t)Np * Np;
  const size_t nt = b.Nt;
  CK(cudaMalloc(&b.Abuf, smat * 8 * cap));
  CK(cudaMalloc(&b.Dbuf, nt * T * T * 8 * cap));
  CK(cudaMalloc(&b.DTbuf, nt * T * T * 8 * cap));
  // diag_kernel only stores the non-zero 32x32 blocks of D_k and D_k^T
  CK(cudaMemsetAsync(b.Dbuf, 0, nt * T * T * 8 * cap, ctx->stream));
  CK(cudaMemsetAsync(b.DTbuf, 0, nt * T * T * 8 * cap, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMalloc(&b.xs, (size_t)D * Np * 8 * cap));
  CK(cudaMalloc(&b.resid, (size_t)Np * 8 * cap));
  CK(cudaMalloc(&b.sn2v, (size_t)Np * 8 * cap));
  CK(cudaMalloc(&b.bvec, (size_t)Np * 8 * cap));
  CK(cudaMalloc(&b.zvec, (size_t)Np * 8 * cap));
  CK(cudaMalloc(&b.alpha, (size_t)Np * 8 * cap));
  CK(cudaMalloc(&b.logdet, nt * 8 * cap));
  CK(cudaMalloc(&b.mult, (size_t)8 * cap));
  CK(cudaMalloc(&b.fmult, (size_t)8 * cap));   // multiplier of the factor kept in each slot (cache)
  CK(cudaMalloc(&b.hyp, (size_t)std::max(md.P, 1) * 8 * cap));
  CK(cudaMalloc(&b.nlz, (size_t)8 * cap));
  CK(cudaMalloc(&b.dnlz, (size_t)std::max(md.P, 1) * 8 * cap));
  CK(cudaMalloc(&b.gpart, nt * (nt + 1) / 2 * (size_t)std::max(md.cov_n, 1) * 8 * cap));
  CK(cudaMalloc(&b.sp, sizeof(SlotP) * cap));
  CK(cudaMalloc(&b.fail, sizeof(int) * cap));
  CK(cudaMalloc(&b.sel, sizeof(int) * cap));
  CK(cudaMalloc(&b.sel2, sizeof(int) * cap));
  CK(cudaMalloc(&b.sel3, sizeof(int) * cap));
  CK(cudaMalloc(&b.sel4, sizeof(int) * cap));
  CK(cudaMalloc(&b.prep_min, sizeof(double) * PREP_MAX_CTAS * cap));
  CK(cudaMalloc(&b.prep_nan, sizeof(int) * PREP_MAX_CTAS * cap));
  CK(cudaMalloc(&b.prep_ticket, sizeof(int) * cap));
  CK(cudaMemsetAsync(b.prep_ticket, 0, sizeof(int) * cap, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  b.cap = cap;
  if (with_w) return alloc_w(ctx, b, cap);
  build_tensor_maps(b);
  return GPB_OK;
}

static int ensure_ws(gpb_ctx* ctx, long long B, bool with_w) {
  Bufs& w = ctx->ws;
  const Model& md = ctx->md;
  const size_t smat_b = (size_t)ctx->Np * ctx->Np * 8;
  auto realloc_all = [&](bool want_w) -> int {
    size_t freeb = 0, totalb = 0;
    CK(cudaMemGetInfo(&freeb, &totalb));
    size_t have = 0;
    if (w.cap) have = per_slot_bytes(w.Np, w.D, w.P, w.cov_n, false) * w.cap + (size_t)w.smat() * 8 * w.wcap;
    const size_t limit = ctx->ws_limit ? ctx->ws_limit : (size_t)((freeb + have) * 0.70);
    const size_t per = per_slot_bytes(ctx->Np, ctx->D, md.P, md.cov_n, want_w);
    const long long fit = (long long)(limit / per);
    if (fit < 1) FAIL(GPB_ENOMEM, "workspace for one matrix does not fit in device memory");
    const int want = (int)std::min<long long>(std::min<long long>(B, fit), 32768);
    ctx->cache.valid = false;
    int rc = alloc_bufs(ctx, w, want, want_w, ctx->Np, ctx->D, md);
    if (rc != GPB_OK) return rc;
    w.cap_is_max = (want == fit);
    w.wcap_is_max = want_w && w.cap_is_max;
    return GPB_OK;
  };
  const bool shape_ok = w.cap > 0 && w.Np == ctx->Np && w.D == ctx->D && w.P == md.P && w.cov_n == md.cov_n;
  // fast path (no cudaMemGetInfo): big enough, or already as big as memory allows.  A W left over from an
  // earlier gradient call stays where it is when an nlZ-only call grows the rest.
  if (!(shape_ok && (w.cap >= B || w.cap_is_max))) {
    int rc = realloc_all(with_w);
    if (rc != GPB_OK) return rc;
  }
  if (!with_w) return GPB_OK;
  const long long needw = std::min<long long>(B, w.cap);
  if (w.wcap >= needw || w.wcap_is_max) return GPB_OK;
  // grow W alone, next to the slots that are there
  size_t freeb = 0, totalb = 0;
  CK(cudaMemGetInfo(&freeb, &totalb));
  const size_t have_w = smat_b * (size_t)w.wcap;
  size_t avail;
  if (ctx->ws_limit) {
    const size_t rest = per_slot_bytes(w.Np, w.D, w.P, w.cov_n, false) * w.cap;
    avail = ctx->ws_limit > rest ? ctx->ws_limit - rest : 0;
  } else {
    avail = (size_t)((freeb + have_w) * 0.70);
  }
  const long long fitw = (long long)(avail / smat_b);
  if (fitw < 1) return realloc_all(true);        // no room beside the arena: re-partition it (A and W, same slots)
  const int wantw = (int)std::min<long long>(needw, fitw);
  int rc = alloc_w(ctx, w, wantw);
  if (rc != GPB_OK) return rc;
  w.wcap_is_max = (wantw == fitw) && wantw < needw;
  return GPB_OK;
}

// ---------------------------------------------------------------------------------
// pipeline pieces (all asynchronous on ctx->stream)
// ---------------------------------------------------------------------------------
static BatchBufs batch_bufs(const Bufs& b, const int* sel, long long N) {
  BatchBufs bb;
  bb.Abuf = b.Abuf;
  bb.Wbuf = b.Wbuf;
  bb.Dbuf = b.Dbuf;
  bb.DTbuf = b.DTbuf;
  bb.sel = sel;
  bb.smat = b.smat();
  bb.Np = b.Np;
  bb.Nt = b.Nt;
  bb.N = (int)N;
  return bb;
}

// CTAs per slot of prep_kernel: enough to spread one matrix's points, fewer when the batch already fills the GPU
static dim3 prep_grid(int nsel, int Np) {
  const int want = (Np + PREP_THREADS - 1) / PREP_THREADS;
  const int per = std::max(1, std::min(std::min(want, (int)PREP_MAX_CTAS), (2 * 148 + nsel - 1) / nsel));
  return dim3((unsigned)nsel, (unsigned)per);
}

template <int KIND>
static void launch_build(gpb_ctx* ctx, const BuildArgs& a, dim3 grid, size_t smem) {
  build_kernel<KIND><<<grid, 256, smem, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

static void run_prep_build(gpb_ctx* ctx, Bufs& b, const Model& md, long long N, const int* sel,
                           int nsel) {
  NvtxRange nv("gpb:K1 prep+build");
  PrepArgs pa;
  pa.md = md;
  pa.N = (int)N;
  pa.Np = b.Np;
  pa.X = ctx->dX;
  pa.y = ctx->dy;
  pa.s2 = ctx->ds2;
  pa.hyp = b.hyp;
  pa.sel = sel;
  pa.mult = b.mult;
  pa.xs = b.xs;
  pa.resid = b.resid;
  pa.sn2v = b.sn2v;
  pa.sp = b.sp;
  pa.part_min = b.prep_min;
  pa.part_nan = b.prep_nan;
  pa.ticket = b.prep_ticket;
  prep_kernel<<<prep_grid(nsel, b.Np), PREP_THREADS, 0, ctx->stream>>>(pa);
  LAUNCHED(ctx);

  BuildArgs ba;
  ba.D = md.D;
  ba.N = (int)N;
  ba.Np = b.Np;
  ba.Nt = b.Nt;
  ba.sel = sel;
  ba.xs = b.xs;
  ba.sn2v = b.sn2v;
  ba.sp = b.sp;
  ba.Abuf = b.Abuf;
  ba.smat = b.smat();
  ba.full = 0;
  dim3 grid((unsigned)(b.Nt * (b.Nt + 1) / 2), (unsigned)nsel);
  const size_t smem = (size_t)2 * md.D * T * 8;
  switch (kind_code(md.cov_kind, md.degree)) {
    case 0: launch_build<0>(ctx, ba, grid, smem); break;
    case 1: launch_build<1>(ctx, ba, grid, smem); break;
    case 3: launch_build<3>(ctx, ba, grid, smem); break;
    case 5: launch_build<5>(ctx, ba, grid, smem); break;
    default: launch_build<2>(ctx, ba, grid, smem); break;
  }
  // right-hand side of the forward solve: b = y - m (the listed slots only)
  copy_sel_kernel<<<dim3((unsigned)((b.Np + 255) / 256), (unsigned)nsel), 256, 0, ctx->stream>>>(
      b.bvec, b.resid, sel, b.Np);
  LAUNCHED(ctx);
}

// blocked right-looking Cholesky of every selected slot (lower, in place) with the forward
// substitution z = L^-1 (y - m) carried along.
static void run_potrf(gpb_ctx* ctx, Bufs& b, long long N, const int* sel, int nsel, bool write_w,
                      bool with_rhs = true) {
  NvtxRange nv("gpb:K2 potrf");
  ctx->cur = &b;
  const BatchBufs bb = batch_bufs(b, sel, N);
  const bool look0 = ctx->lookahead == 2 || (ctx->lookahead == 1 && nsel <= 8);
  const int Nt = b.Nt;                                 // outer block = OB tile columns
  // With look-ahead, the width of an outer block is chosen where the block starts: while the
  // trailing matrix is large the update is GEMM-bound and wide blocks (long K) pay; once it is
  // small, the dependent chain diag -> panel -> next column decides and narrow blocks shorten it
  // (measured, N=5000, one matrix: 5.0 -> 4.1 ms at width 1).  The choice never changes a bit of
  // the result: every element sees the same FP64 operations in the same order for any blocking.
  const int narrow = ctx->la_ob > 0 ? ctx->la_ob : (nsel <= 2 ? 1 : 2);
  const bool left_on = ctx->left_looking == 1 || (ctx->left_looking == 2 && nsel >= 24 && Nt >= 16);
  const int wide = (left_on && !ctx->outer_block_set) ? 8 : ctx->outer_block;
  auto width_at = [&](int k0) {
    if (!look0) return wide;
    const long long nrem = Nt - k0;
    return ((long long)nsel * nrem * nrem >= ctx->la_wide) ? std::max(narrow, wide) : narrow;
  };
  const bool look = look0;
  bool joined_pending = false;
  int ob0 = 0, obe = 0;                                // current outer block [ob0, obe)
  for (int k = 0; k < Nt; ++k) {
    if (k == obe) {
      ob0 = k;
      obe = std::min(k + width_at(k), Nt);
    }
    // left-looking inside a wide outer block: column k receives the block's earlier columns in ONE update
    // (K = (k - ob0)*128) right before it is factored, instead of one K = 128 update after each of them --
    // same operations in the same order for every element, but each C tile of the block is read and
    // written once per column instead of once per earlier column, and the K loop is longer
    const bool left = left_on && (obe - ob0) >= 3;
    if (left && k > ob0)
      launch_gemm(ctx, OpSyrk{bb, ob0, k - ob0, k, k + 1}, dim3((unsigned)(Nt - k), (unsigned)nsel));
    DiagArgs da;
    da.Abuf = b.Abuf;
    da.Wbuf = write_w ? b.Wbuf : nullptr;
    da.Dbuf = b.Dbuf;
    da.DTbuf = b.DTbuf;
    da.sel = sel;
    da.smat = b.smat();
    da.Np = b.Np;
    da.Nt = b.Nt;
    da.N = (int)N;
    da.k = k;
    da.bvec = with_rhs ? b.bvec : nullptr;
    da.zvec = with_rhs ? b.zvec : nullptr;
    da.logdet = b.logdet;
    da.fail = b.fail;
    da.dbg = (k == 0) ? ctx->diag_dbg : nullptr;
    launch_chain(ctx, diag_kernel, dim3((unsigned)nsel), dim3(256), DIAG_SMEM, da);
    LAUNCHED(ctx);
    const int n = Nt - k - 1;
    if (n <= 0) break;
    // panel  L_ik = A_ik D_k^T  with the forward-substitution update  b_i -= L_ik z_k  fused in
    launch_panel(ctx, OpPanel{bb, k, with_rhs ? b.zvec : nullptr, with_rhs ? b.bvec : nullptr},
                 dim3((unsigned)n, (unsigned)nsel));
    // two-level trailing update (see OpSyrk): inside the outer block only its own columns
    if (k + 1 < obe && !left) {
      const int cnt = OpSyrk::count(Nt, k + 1, obe);
      launch_gemm(ctx, OpSyrk{bb, k, 1, k + 1, obe}, dim3((unsigned)cnt, (unsigned)nsel));
    }
    if (k + 1 == obe && obe < Nt) {                    // outer block done: update the rest, long K
      const int kw = obe - ob0;
      const int obe2 = std::min(obe + width_at(obe), Nt);   // end of the NEXT outer block
      if (joined_pending) {                            // the previous remainder wrote these tiles
        cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
        joined_pending = false;
      }
      if (!look || obe2 >= Nt) {
        const int cnt = OpSyrk::count(Nt, obe, Nt);
        launch_gemm(ctx, OpSyrk{bb, ob0, kw, obe, Nt}, dim3((unsigned)cnt, (unsigned)nsel));
      } else {
        // look-ahead: the next outer block only needs its own tile columns [obe, obe2); the rest
        // of the update, columns [obe2, Nt), runs on the aux stream while the main stream
        // factors that block (one CTA per matrix in diag_kernel leaves the GPU idle otherwise)
        cudaEventRecord(ctx->ev_fork, ctx->stream);    // panels of [ob0, obe) are complete
        launch_gemm(ctx, OpSyrk{bb, ob0, kw, obe, obe2},
                    dim3((unsigned)OpSyrk::count(Nt, obe, obe2), (unsigned)nsel));
        cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0);
        cudaStream_t main_stream = ctx->stream;
        ctx->stream = ctx->aux;
        // (optionally in chunks; the low stream priority of aux already lets the main stream's
        // CTAs in first, and chunking only added launch tails when measured)
        const int total = OpSyrk::count(Nt, obe2, Nt);
        const int chunk = std::max(1, ctx->la_chunk / std::max(nsel, 1));
        for (int t0 = 0; t0 < total; t0 += chunk) {
          OpSyrk op{bb, ob0, kw, obe2, Nt};
          op.bx0 = t0;
          launch_gemm(ctx, op, dim3((unsigned)std::min(chunk, total - t0), (unsigned)nsel));
        }
        ctx->stream = main_stream;
        cudaEventRecord(ctx->ev_join, ctx->aux);
        joined_pending = true;
      }
    }
  }
  if (joined_pending) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
}

// alpha = L^-T z / sl
static void run_bwd(gpb_ctx* ctx, Bufs& b, const int* sel, int nsel) {
  NvtxRange nv("gpb:K2 backward solve");
  copy_sel_kernel<<<dim3((unsigned)((b.Np + 255) / 256), (unsigned)nsel), 256, 0, ctx->stream>>>(
      b.bvec, b.zvec, sel, b.Np);
  LAUNCHED(ctx);
  for (int i = b.Nt - 1; i >= 0; --i) {
    VecArgs va;
    va.Abuf = b.Abuf;
    va.DTbuf = b.DTbuf;
    va.sel = sel;
    va.smat = b.smat();
    va.Np = b.Np;
    va.Nt = b.Nt;
    va.k = i;
    va.bvec = b.bvec;
    va.zvec = b.zvec;
    va.alpha = b.alpha;
    va.sp = b.sp;
    bwd_step_kernel<<<dim3((unsigned)std::max(i, 1), (unsigned)nsel), 256, 0, ctx->stream>>>(va);
    LAUNCHED(ctx);
  }
}

// W = L^-1 (lower tiles of Wbuf; dual: also W^T in the upper tiles); then optionally
// Ainv = W^T W into the lower tiles of Abuf.
static void run_inverse(gpb_ctx* ctx, Bufs& b, long long N, const int* sel, int nsel, bool dual,
                        bool syrk2) {
  NvtxRange nv("gpb:K2 inverse (trtri + W^T W)");
  ctx->cur = &b;
  const BatchBufs bb = batch_bufs(b, sel, N);
  const int Nt = b.Nt;
  const bool recursive = ctx->trtri == 2 || (ctx->trtri == 1 && nsel <= ctx->trtri_max);
  if (Nt > 1 && recursive) {
    // recursive halving: few, wide launches (see OpRecX) -- the shape for a few matrices
    for (int s = 1; s < Nt; s *= 2) {
      const int P = (Nt - s + 2 * s - 1) / (2 * s);    // pairs whose second block is not empty
      const dim3 grid((unsigned)(s * s * P), (unsigned)nsel);
      launch_gemm(ctx, OpRecX{bb, s, P}, grid);
      launch_gemm(ctx, OpRecW{bb, s, P}, grid);
    }
  } else if (Nt > 1) {
    launch_gemm(ctx, OpHpass{bb}, dim3((unsigned)(Nt * (Nt - 1) / 2), (unsigned)nsel));
    for (int j = Nt - 2; j >= 0; --j)
      launch_gemm(ctx, OpWrec{bb, j, dual ? 1 : 0}, dim3((unsigned)(Nt - 1 - j), (unsigned)nsel));
  }
  if (syrk2) launch_gemm(ctx, OpSyrk2{bb}, dim3((unsigned)(Nt * (Nt + 1) / 2), (unsigned)nsel));
}

template <int KIND, int DP>
static void launch_grad_rows(gpb_ctx* ctx, const GradArgs& a, dim3 grid) {
  const size_t smem = ((size_t)T * (DP + 2) + 8 * (size_t)(DP + 2)) * 8;
  grad_rows_kernel<KIND, DP><<<grid, 256, smem, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

template <int KIND>
static void launch_grad(gpb_ctx* ctx, const GradArgs& a, dim3 grid, int ard, int D) {
  if (ard && D <= 16 && ctx->grad_rows) {          // row-per-thread form (cov.cuh)
    if (D <= 4) return launch_grad_rows<KIND, 4>(ctx, a, grid);
    if (D <= 6) return launch_grad_rows<KIND, 6>(ctx, a, grid);
    if (D <= 8) return launch_grad_rows<KIND, 8>(ctx, a, grid);
    if (D <= 10) return launch_grad_rows<KIND, 10>(ctx, a, grid);
    if (D <= 12) return launch_grad_rows<KIND, 12>(ctx, a, grid);
    return launch_grad_rows<KIND, 16>(ctx, a, grid);
  }
  const int dp = !ard ? 0 : (D <= 8 ? 8 : (D <= 12 ? 12 : (D <= 16 ? 16 : (D <= 32 ? 32 : 64))));
  const size_t smem = ((size_t)2 * D * T + 2 * T + 8 * (size_t)((dp ? dp : 1) + 2)) * 8;
  switch (dp) {
    case 0: grad_kernel<KIND, 0><<<grid, 256, smem, ctx->stream>>>(a); break;
    case 8: grad_kernel<KIND, 8><<<grid, 256, smem, ctx->stream>>>(a); break;
    case 12: grad_kernel<KIND, 12><<<grid, 256, smem, ctx->stream>>>(a); break;
    case 16: grad_kernel<KIND, 16><<<grid, 256, smem, ctx->stream>>>(a); break;
    case 32: grad_kernel<KIND, 32><<<grid, 256, smem, ctx->stream>>>(a); break;
    default: grad_kernel<KIND, 64><<<grid, 256, smem, ctx->stream>>>(a); break;
  }
  LAUNCHED(ctx);
}

static void run_grad(gpb_ctx* ctx, Bufs& b, const Model& md, long long N, const int* sel, int nsel) {
  NvtxRange nv("gpb:K3 gradient");
  GradArgs ga;
  ga.D = md.D;
  ga.N = (int)N;
  ga.Np = b.Np;
  ga.Nt = b.Nt;
  ga.cov_n = md.cov_n;
  ga.sel = sel;
  ga.xs = b.xs;
  ga.sp = b.sp;
  ga.Abuf = b.Abuf;
  ga.alpha = b.alpha;
  ga.gpart = b.gpart;
  ga.smat = b.smat();
  dim3 grid((unsigned)(b.Nt * (b.Nt + 1) / 2), (unsigned)nsel);
  switch (kind_code(md.cov_kind, md.degree)) {
    case 0: launch_grad<0>(ctx, ga, grid, md.ard, md.D); break;
    case 1: launch_grad<1>(ctx, ga, grid, md.ard, md.D); break;
    case 3: launch_grad<3>(ctx, ga, grid, md.ard, md.D); break;
    case 5: launch_grad<5>(ctx, ga, grid, md.ard, md.D); break;
    default: launch_grad<2>(ctx, ga, grid, md.ard, md.D); break;
  }
  GradFinalArgs fa;
  fa.md = md;
  fa.N = (int)N;
  fa.Np = b.Np;
  fa.Nt = b.Nt;
  fa.sel = sel;
  fa.X = ctx->dX;
  fa.y = ctx->dy;
  fa.s2 = ctx->ds2;
  fa.hyp = b.hyp;
  fa.sp = b.sp;
  fa.Abuf = b.Abuf;
  fa.alpha = b.alpha;
  fa.gpart = b.gpart;
  fa.dnlZ = b.dnlz;
  fa.smat = b.smat();
  grad_final_kernel<<<dim3((unsigned)nsel, (unsigned)(1 + md.noise_n + md.mean_n)), 256, 0, ctx->stream>>>(fa);
  LAUNCHED(ctx);
}

// Factor `n` slots (hyp already in b.hyp): build + potrf with the reference's x10 jitter
// retry per element (gaussian_process.py:2413-2421, :2430-2438).  On return b.sel holds the
// identity list, status_h[s] = 1 for slots that failed all 10 attempts.
static int factor_with_retry(gpb_ctx* ctx, Bufs& b, const Model& md, long long N,
                             const std::vector<int>& slots, bool write_w, std::vector<int>& status_h) {
  if (slots.empty()) return GPB_OK;
  const int maxslot = *std::max_element(slots.begin(), slots.end()) + 1;
  std::vector<int> failh(maxslot);
  CK(cudaMemcpyAsync(b.sel3, slots.data(), sizeof(int) * slots.size(), cudaMemcpyHostToDevice, ctx->stream));
  set_mult_kernel<<<(unsigned)((slots.size() + 255) / 256), 256, 0, ctx->stream>>>(b.mult, b.fail, b.sel3,
                                                                                 (int)slots.size(), 1.0);
  LAUNCHED(ctx);
  for (int s : slots) status_h[s] = 0;
  const int* sel = b.sel3;
  int nsel = (int)slots.size();
  std::vector<int> cur = slots;
  for (int attempt = 0; attempt < 10; ++attempt) {
    NvtxRange nv(attempt == 0 ? "gpb:factor" : "gpb:factor (jitter retry)");
    run_prep_build(ctx, b, md, N, sel, nsel);
    run_potrf(ctx, b, N, sel, nsel, write_w);
    CK(cudaGetLastError());                    // a refused launch must not pass for a result
    CK(cudaMemcpyAsync(failh.data(), b.fail, sizeof(int) * maxslot, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<int> next;
    for (int s : cur)
      if (failh[s]) next.push_back(s);
    if (next.empty()) break;
    if (attempt == 9) {
      for (int s : next) status_h[s] = 1;
      break;
    }
    // sn2_mult *= 10 for the failed slots, clear their flags, go again on the compacted list
    CK(cudaMemcpyAsync(b.sel2, next.data(), sizeof(int) * next.size(), cudaMemcpyHostToDevice,
                       ctx->stream));
    scale_mult_kernel<<<(unsigned)((next.size() + 255) / 256), 256, 0, ctx->stream>>>(
        b.mult, b.fail, b.sel2, (int)next.size());
    LAUNCHED(ctx);
    CK(cudaMemsetAsync(b.fail, 0, sizeof(int) * b.cap, ctx->stream));
    cur = next;
    sel = b.sel2;
    nsel = (int)next.size();
  }
  return GPB_OK;
}

// Replay only the forward solve z = L^-1 (y - m) on slots whose factor is cached (their
// covariance and noise hyperparameters are unchanged): O(N^2) instead of O(N^3).
static void run_solve_only(gpb_ctx* ctx, Bufs& b, const Model& md, long long N, const int* sel,
                           const int* fsel, int nsel) {
  NvtxRange nv("gpb:solve replay (cached factor)");
  // the factor's jitter multiplier belongs to the row as well
  copy_mult_kernel<<<(unsigned)((nsel + 255) / 256), 256, 0, ctx->stream>>>(b.mult, b.fmult, sel, fsel, nsel);
  LAUNCHED(ctx);
  PrepArgs pa;
  pa.md = md;
  pa.N = (int)N;
  pa.Np = b.Np;
  pa.X = ctx->dX;
  pa.y = ctx->dy;
  pa.s2 = ctx->ds2;
  pa.hyp = b.hyp;
  pa.sel = sel;
  pa.mult = b.mult;
  pa.xs = b.xs;
  pa.resid = b.resid;
  pa.sn2v = b.sn2v;
  pa.sp = b.sp;
  pa.part_min = b.prep_min;
  pa.part_nan = b.prep_nan;
  pa.ticket = b.prep_ticket;
  prep_kernel<<<prep_grid(nsel, b.Np), PREP_THREADS, 0, ctx->stream>>>(pa);
  LAUNCHED(ctx);
  copy_sel_kernel<<<dim3((unsigned)((b.Np + 255) / 256), (unsigned)nsel), 256, 0, ctx->stream>>>(
      b.bvec, b.resid, sel, b.Np);
  LAUNCHED(ctx);
  const BatchBufs bb = batch_bufs(b, sel, N);
  // block column k < Nt-1: one launch computes z_k = D_k b_k (in every CTA) and b_i -= L_ik z_k
  // with the CTA shape of the fused panel, so the reduction order (and every bit) is the same;
  // the last block only needs the solve
  for (int k = 0; k + 1 < b.Nt; ++k) {
    dim3 grid((unsigned)(b.Nt - k - 1), (unsigned)nsel);
    const OpFwdZ op{bb, k, b.zvec, b.bvec, fsel};
    if (ctx->gemm_bn != 128) launch_shape<OpFwdZ, 64, 128>(ctx, op, grid, 0);
    else launch_shape<OpFwdZ, 128, 128>(ctx, op, grid, 0);
  }
  DiagSolveArgs da;
  da.Dbuf = b.Dbuf;
  da.sel = sel;
  da.fsel = fsel;
  da.Np = b.Np;
  da.Nt = b.Nt;
  da.N = (int)N;
  da.k = b.Nt - 1;
  da.bvec = b.bvec;
  da.zvec = b.zvec;
  diag_solve_kernel<<<nsel, T, 0, ctx->stream>>>(da);
  LAUNCHED(ctx);
}

// ---------------------------------------------------------------------------------
// nlZ (+ gradient), batched
// ---------------------------------------------------------------------------------
static int nlz_batch_impl(gpb_ctx* ctx, const double* hyp, bool hyp_on_device, int64_t B,
                          int want_grad, double* nlZ, double* dnlZ, double* sn2_mult,
                          int32_t* status, bool out_on_device) {
  if (!ctx) return GPB_EINVAL;
  NvtxRange nv_call(want_grad ? "gpb_nlz_batch (nlZ + gradient)" : "gpb_nlz_batch (nlZ)");
  if (!ctx->has_model || !ctx->has_data) FAIL(GPB_ESTATE, "gpb_nlz_batch: set model and data first");
  if (!hyp || !nlZ || B <= 0) FAIL(GPB_EINVAL, "gpb_nlz_batch: bad arguments");
  if (want_grad && !dnlZ) FAIL(GPB_EINVAL, "gpb_nlz_batch: dnlZ is NULL");
  CK(cudaSetDevice(ctx->device));
  const Model md = ctx->md;
  const int P = md.P;
  int rc = ensure_ws(ctx, B, want_grad != 0);
  if (rc != GPB_OK) return rc;
  Bufs& b = ctx->ws;
  const cudaMemcpyKind kin = hyp_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  for (int i = 0; i < 6; ++i) ctx->timings[i] = 0.0;
  std::vector<int> status_h;
  std::vector<double> mult_h;

  const int chunk = want_grad ? std::min(b.cap, b.wcap) : b.cap;     // W may hold fewer slots than the rest
  for (int64_t row0 = 0; row0 < B; row0 += chunk) {
    const int n = (int)std::min<int64_t>(chunk, B - row0);
    if (P > 0) CK(cudaMemcpyAsync(b.hyp, hyp + row0 * P, sizeof(double) * n * P, kin, ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    {
      std::vector<int> ident(n);
      for (int i = 0; i < n; ++i) ident[i] = i;
      CK(cudaMemcpyAsync(b.sel, ident.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
      status_h.assign(n, 0);
      // which rows can re-use a factor still sitting in some workspace slot?  (Rows that differ
      // only in mean hyperparameters share one factor: all speculative proposals of a slice-
      // sampling move along a mean coordinate hit the factor of the current point.)
      const int kn = md.cov_n + md.noise_n;
      const bool cacheable = ctx->cache_enabled && !hyp_on_device && !want_grad && B <= b.cap;
      // An invalidated cache (new model/data/workspace, a gradient or multi-chunk call, a failed
      // call) must not leave entries behind: their slots were reallocated or overwritten, and the
      // re-keying below only covers the rows of THIS call.
      if (!ctx->cache.valid) {
        ctx->cache.n = 0;
        ctx->cache.key.clear();
        ctx->cache.ok.clear();
      }
      std::vector<int> miss, hit, hit_f;
      for (int sidx = 0; sidx < n; ++sidx) {
        int found = -1;
        if (cacheable && ctx->cache.valid) {
          const double* key = hyp + (row0 + sidx) * P;
          auto same = [&](int j) {
            return j < ctx->cache.n && ctx->cache.ok[j] &&
                   memcmp(&ctx->cache.key[(size_t)j * kn], key, sizeof(double) * kn) == 0;
          };
          if (same(sidx)) found = sidx;
          for (int j = 0; found < 0 && j < ctx->cache.n; ++j)
            if (same(j)) found = j;
        }
        if (found >= 0) { hit.push_back(sidx); hit_f.push_back(found); }
        else miss.push_back(sidx);
      }
      ctx->cache.hits += (long long)hit.size();
      ctx->cache.misses += (long long)miss.size();
      ctx->cache.valid = false;                  // stays invalid if anything below fails
      if (!hit.empty()) {                        // first: the misses below overwrite factor slots
        CK(cudaMemcpyAsync(b.sel4, hit.data(), sizeof(int) * hit.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(b.sel2, hit_f.data(), sizeof(int) * hit.size(), cudaMemcpyHostToDevice, ctx->stream));
        run_solve_only(ctx, b, md, ctx->N, b.sel4, b.sel2, (int)hit.size());
        NlzArgs nh;
        nh.sel = b.sel4;
        nh.fsel = b.sel2;
        nh.N = (int)ctx->N;
        nh.Np = b.Np;
        nh.Nt = b.Nt;
        nh.zvec = b.zvec;
        nh.logdet = b.logdet;
        nh.sp = b.sp;
        nh.nlz = b.nlz;
        nlz_kernel<<<(unsigned)hit.size(), 256, 0, ctx->stream>>>(nh);
        LAUNCHED(ctx);
        CK(cudaStreamSynchronize(ctx->stream));   // sel2 is reused by the retry loop below
      }
      rc = factor_with_retry(ctx, b, md, ctx->N, miss, want_grad != 0, status_h);
      if (rc != GPB_OK) return rc;
      if (cacheable) {
        ctx->cache.key.resize((size_t)std::max(n, ctx->cache.n) * kn);
        ctx->cache.ok.resize((size_t)std::max(n, ctx->cache.n), 0);
        for (int sidx : miss) {
          memcpy(&ctx->cache.key[(size_t)sidx * kn], hyp + (row0 + sidx) * P, sizeof(double) * kn);
          ctx->cache.ok[sidx] = status_h[sidx] == 0;
        }
        ctx->cache.n = std::max(n, ctx->cache.n);
        ctx->cache.valid = true;
      }
      // nlZ of the freshly factored rows; remember each new factor's jitter multiplier
      if (!miss.empty()) {
        CK(cudaMemcpyAsync(b.sel4, miss.data(), sizeof(int) * miss.size(), cudaMemcpyHostToDevice, ctx->stream));
        copy_mult_kernel<<<(unsigned)((miss.size() + 255) / 256), 256, 0, ctx->stream>>>(
            b.fmult, b.mult, b.sel4, b.sel4, (int)miss.size());
        LAUNCHED(ctx);
        NlzArgs nm;
        nm.sel = b.sel4;
        nm.fsel = nullptr;
        nm.N = (int)ctx->N;
        nm.Np = b.Np;
        nm.Nt = b.Nt;
        nm.zvec = b.zvec;
        nm.logdet = b.logdet;
        nm.sp = b.sp;
        nm.nlz = b.nlz;
        nlz_kernel<<<(unsigned)miss.size(), 256, 0, ctx->stream>>>(nm);
        LAUNCHED(ctx);
      }
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (want_grad) {
      CK(cudaEventRecord(ctx->ev[3], ctx->stream));
      run_inverse(ctx, b, ctx->N, b.sel, n, /*dual=*/true, /*syrk2=*/true);
      {                                        // alpha = W^T z / sl: W is there, no back substitution
        AlphaArgs aa;
        aa.Wbuf = b.Wbuf;
        aa.sel = b.sel;
        aa.smat = b.smat();
        aa.Np = b.Np;
        aa.N = (int)ctx->N;
        aa.zvec = b.zvec;
        aa.alpha = b.alpha;
        aa.sp = b.sp;
        alpha_gemv_kernel<<<dim3((unsigned)((b.Np + 7) / 8), (unsigned)n), 256, 0, ctx->stream>>>(aa);
        LAUNCHED(ctx);
      }
      CK(cudaEventRecord(ctx->ev[4], ctx->stream));
      run_grad(ctx, b, md, ctx->N, b.sel, n);
    } else {
      CK(cudaEventRecord(ctx->ev[3], ctx->stream));
      CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CK(cudaGetLastError());
    // results
    bool anyfail = false;
    for (int s = 0; s < n; ++s) anyfail |= status_h[s] != 0;
    if (anyfail) {
      // the reference raises for these rows; report NaN and status 1
      std::vector<double> nl(n);
      CK(cudaMemcpyAsync(nl.data(), b.nlz, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (int s = 0; s < n; ++s)
        if (status_h[s]) nl[s] = NAN;
      CK(cudaMemcpyAsync(b.nlz, nl.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMemcpyAsync(nlZ + row0, b.nlz, sizeof(double) * n, kout, ctx->stream));
    if (want_grad && P > 0)
      CK(cudaMemcpyAsync(dnlZ + row0 * P, b.dnlz, sizeof(double) * n * P, kout, ctx->stream));
    if (sn2_mult) CK(cudaMemcpyAsync(sn2_mult + row0, b.mult, sizeof(double) * n, kout, ctx->stream));
    if (status) {
      if (out_on_device) {
        CK(cudaMemcpyAsync(status + row0, status_h.data(), sizeof(int) * n, cudaMemcpyHostToDevice,
                           ctx->stream));
      } else {
        for (int s = 0; s < n; ++s) status[row0 + s] = status_h[s];
      }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    float ms;
    const int pairs[5][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {4, 5}};
    // {build+potrf (with retries), nlz, solve, inverse, gradient}
    for (int i = 0; i < 5; ++i) {
      CK(cudaEventElapsedTime(&ms, ctx->ev[pairs[i][0]], ctx->ev[pairs[i][1]]));
      ctx->timings[i] += ms;
    }
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]));
    ctx->timings[5] += ms;
  }
  return GPB_OK;
}

extern "C" int gpb_nlz_batch(gpb_ctx* ctx, const double* hyp, int64_t B, int want_grad, double* nlZ,
                             double* dnlZ, double* sn2_mult, int32_t* status) {
  return nlz_batch_impl(ctx, hyp, false, B, want_grad, nlZ, dnlZ, sn2_mult, status, false);
}

extern "C" int gpb_nlz_batch_dev(gpb_ctx* ctx, const double* d_hyp, int64_t B, int want_grad,
                                 double* d_nlZ, double* d_dnlZ, double* d_sn2_mult,
                                 int32_t* d_status) {
  return nlz_batch_impl(ctx, d_hyp, true, B, want_grad, d_nlZ, d_dnlZ, d_sn2_mult, d_status, true);
}

// ---------------------------------------------------------------------------------
// posteriors
// ---------------------------------------------------------------------------------
extern "C" int gpb_posterior_batch(gpb_ctx* ctx, const double* hyp, int64_t B, gpb_post** out) {
  if (!ctx || !out) return GPB_EINVAL;
  NvtxRange nv_call("gpb_posterior_batch");
  *out = nullptr;
  if (!ctx->has_model || !ctx->has_data) FAIL(GPB_ESTATE, "gpb_posterior_batch: set model and data first");
  if (!hyp || B <= 0 || B > 32768) FAIL(GPB_EINVAL, "gpb_posterior_batch: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const Model md = ctx->md;
  size_t freeb = 0, totalb = 0;
  CK(cudaMemGetInfo(&freeb, &totalb));
  const size_t need = per_slot_bytes(ctx->Np, ctx->D, md.P, md.cov_n, true) * (size_t)B;
  if (need > freeb * 0.95 && ctx->ws.cap > 0) {
    // the evaluation workspace is a cache: give it back before giving up
    CK(cudaStreamSynchronize(ctx->stream));
    free_bufs(ctx->ws);
    ctx->cache.valid = false;
    CK(cudaMemGetInfo(&freeb, &totalb));
  }
  if (need > freeb * 0.95)
    FAIL(GPB_ENOMEM, "gpb_posterior_batch: the posterior factors do not fit in device memory");
  gpb_post* post = new gpb_post();
  post->ctx = ctx;
  post->N = ctx->N;
  post->md = md;
  int rc = alloc_bufs(ctx, post->b, (int)B, true, ctx->Np, ctx->D, md);
  if (rc != GPB_OK) {
    free_bufs(post->b);
    delete post;
    return rc;
  }
  Bufs& b = post->b;
  post->hyp.assign(hyp, hyp + B * md.P);
  auto bail = [&](int code) {
    free_bufs(post->b);
    if (post->X) cudaFree(post->X);
    delete post;
    return code;
  };
  {
    cudaError_t e = cudaMalloc(&post->X, sizeof(double) * ctx->Np * ctx->D);   // room for appends
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(post->X, ctx->dX, sizeof(double) * ctx->N * ctx->D, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
  }
  if (md.P > 0) {
    cudaError_t e = cudaMemcpyAsync(b.hyp, hyp, sizeof(double) * B * md.P, cudaMemcpyHostToDevice,
                                    ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
  }
  {
    std::vector<int> ident((size_t)B);
    for (int i = 0; i < (int)B; ++i) ident[i] = i;
    cudaError_t e = cudaMemcpyAsync(b.sel, ident.data(), sizeof(int) * B, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
    post->status.assign((size_t)B, 0);
    rc = factor_with_retry(ctx, b, md, ctx->N, ident, true, post->status);
    if (rc != GPB_OK) return bail(rc);
  }
  run_bwd(ctx, b, b.sel, (int)B);
  post->sp.resize(B);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(post->sp.data(), b.sp, sizeof(SlotP) * B, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
  // low-noise samples store L = -(K + sn2_mult*diag(sn2))^-1 (gaussian_process.py:2440-2448):
  // build the full inverse now; L_chol samples get W = L^-1 lazily at the first predict.
  std::vector<int> low;
  for (int s = 0; s < (int)B; ++s)
    if (!post->sp[s].lchol && !post->status[s]) low.push_back(s);
  if (!low.empty()) {
    e = cudaMemcpyAsync(b.sel2, low.data(), sizeof(int) * low.size(), cudaMemcpyHostToDevice,
                        ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
    run_inverse(ctx, b, ctx->N, b.sel2, (int)low.size(), true, true);
    for (int s : low) {
      symmetrize_kernel<<<grid1d(b.smat()), 256, 0, ctx->stream>>>(b.Abuf + s * b.smat(), b.Np);
      LAUNCHED(ctx);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return bail(GPB_ECUDA); }
  }
  ctx->posts.push_back(post);
  *out = post;
  return GPB_OK;
}

extern "C" int64_t gpb_posterior_count(const gpb_post* post) { return post ? post->b.cap : 0; }

static void release_post(gpb_post* post) {
  free_bufs(post->b);
  if (post->X) cudaFree(post->X);
  delete post;
}

extern "C" void gpb_posterior_free(gpb_post* post) {
  if (!post) return;
  gpb_ctx* ctx = post->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  auto it = std::find(ctx->posts.begin(), ctx->posts.end(), post);
  if (it != ctx->posts.end()) ctx->posts.erase(it);
  release_post(post);
}

extern "C" int gpb_posterior_fetch(const gpb_post* post, int64_t s, int field, double* out) {
  if (!post || !out) return GPB_EINVAL;
  gpb_ctx* ctx = post->ctx;
  if (s < 0 || s >= post->b.cap) FAIL(GPB_EINVAL, "gpb_posterior_fetch: sample index out of range");
  CK(cudaSetDevice(ctx->device));
  const Bufs& b = post->b;
  const SlotP& p = post->sp[s];
  const long long N = post->N;
  switch (field) {
    case GPB_POST_ALPHA:
      CK(cudaMemcpyAsync(out, b.alpha + s * b.Np, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      return GPB_OK;
    case GPB_POST_SW:
      out[0] = 1.0 / sqrt(p.sn2_min * p.mult);     // gaussian_process.py:2517
      return GPB_OK;
    case GPB_POST_SN2MULT: out[0] = p.mult; return GPB_OK;
    case GPB_POST_LCHOL: out[0] = p.lchol ? 1.0 : 0.0; return GPB_OK;
    case GPB_POST_STATUS: out[0] = post->status[s]; return GPB_OK;
    case GPB_POST_L: {
      double* tmp = nullptr;
      CK(cudaMalloc(&tmp, sizeof(double) * N * N));
      fetch_L_kernel<<<grid1d(N * N), 256, 0, ctx->stream>>>(b.Abuf + s * b.smat(), b.Np, (int)N,
                                                             p.lchol, tmp);
      LAUNCHED(ctx);
      cudaError_t e = cudaMemcpyAsync(out, tmp, sizeof(double) * N * N, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      cudaFree(tmp);
      if (e != cudaSuccess) FAIL(GPB_ECUDA, cudaGetErrorString(e));
      return GPB_OK;
    }
    default: FAIL(GPB_EINVAL, "gpb_posterior_fetch: unknown field");
  }
}

// ---------------------------------------------------------------------------------
// predict
// ---------------------------------------------------------------------------------
static int grow(gpb_ctx* ctx, double** p, size_t* have, size_t need) {
  if (*have >= need) return GPB_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  CK(cudaMalloc(p, need * sizeof(double)));
  *have = need;
  return GPB_OK;
}

static int ensure_w(gpb_ctx* ctx, gpb_post* post) {
  if (post->w_ready) return GPB_OK;
  Bufs& b = post->b;
  std::vector<int> high;
  for (int s = 0; s < b.cap; ++s)
    if (post->sp[s].lchol && !post->status[s]) high.push_back(s);
  if (!high.empty()) {
    CK(cudaMemcpyAsync(b.sel2, high.data(), sizeof(int) * high.size(), cudaMemcpyHostToDevice, ctx->stream));
    run_inverse(ctx, b, post->N, b.sel2, (int)high.size(), false, false);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
  }
  post->w_ready = true;
  return GPB_OK;
}

template <int KIND>
static void launch_ks(gpb_ctx* ctx, const KsArgs& a, dim3 grid, size_t smem) {
  ks_build_kernel<KIND><<<grid, 256, smem, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

static int predict_impl(gpb_ctx* ctx, const gpb_post* cpost, const double* Xs, const double* ys,
                        const double* s2s, int64_t M, int add_noise, int separate, int want_lpd,
                        double* mu, double* s2, double* lpd, bool on_device) {
  if (!ctx || !cpost) return GPB_EINVAL;
  NvtxRange nv("gpb:K4 predict");
  gpb_post* post = const_cast<gpb_post*>(cpost);
  if (!Xs || !mu || !s2 || M <= 0) FAIL(GPB_EINVAL, "gpb_predict: bad arguments");
  if (want_lpd && (!ys || !lpd)) FAIL(GPB_EINVAL, "Cannot calculate log predictive density without y_star.");
  CK(cudaSetDevice(ctx->device));
  for (int st : post->status)
    if (st) FAIL(GPB_ESTATE, "gpb_predict: a posterior sample has no valid factorisation");
  int rc = ensure_w(ctx, post);
  if (rc != GPB_OK) return rc;
  const Bufs& b = post->b;
  const Model md = post->md;
  const int Ns = b.cap, D = md.D, Np = b.Np, Nt = b.Nt;
  const int need_ys2 = (add_noise || want_lpd) ? 1 : 0;
  const int Mc = (int)std::min<int64_t>(M, 8192);
  const int McpMax = round_up(Mc, T);
  const int ns = ctx->gemm_bn == 128 ? 1 : 2;
  // samples per launch: all of them when the per-sample scratch (Mcp x Np doubles) allows
  const size_t per_sample = (size_t)McpMax * Np;
  const int G = (int)std::max<size_t>(1, std::min<size_t>((size_t)Ns, ((size_t)2 << 30) / (per_sample * 8)));
  if ((rc = grow(ctx, &ctx->pXs, &ctx->pXs_n, (size_t)3 * McpMax * std::max(D, 1))) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pBt, &ctx->pBt_n, (size_t)G * per_sample)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pmu, &ctx->ppart_n, (size_t)G * 3 * Nt * McpMax)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->psamp, &ctx->psamp_n, (size_t)4 * Ns * McpMax)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pout, &ctx->pout_n, (size_t)3 * McpMax * (separate ? Ns : 1))) != GPB_OK) return rc;
  double* dXs = ctx->pXs;
  double* dys = ctx->pXs + (size_t)McpMax * D;
  double* ds2s = dys + McpMax;
  const int kc = kind_code(md.cov_kind, md.degree);
  const size_t ks_smem = ((size_t)2 * D * T + T + 8 * T) * 8;

  for (int64_t c0 = 0; c0 < M; c0 += Mc) {
    const int mc = (int)std::min<int64_t>(Mc, M - c0);
    const int Mcp = round_up(mc, T);
    const double *cXs, *cys = nullptr, *cs2s = nullptr;
    if (on_device) {
      cXs = Xs + c0 * D;
    } else {
      CK(cudaMemcpyAsync(dXs, Xs + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, ctx->stream));
      cXs = dXs;
      if (ys) {
        CK(cudaMemcpyAsync(dys, ys + c0, sizeof(double) * mc, cudaMemcpyHostToDevice, ctx->stream));
        cys = dys;
      }
      if (s2s) {
        CK(cudaMemcpyAsync(ds2s, s2s + c0, sizeof(double) * mc, cudaMemcpyHostToDevice, ctx->stream));
        cs2s = ds2s;
      }
    }
    double* mu_s = ctx->psamp;
    double* s2_s = mu_s + (size_t)Ns * Mcp;
    double* ys2_s = s2_s + (size_t)Ns * Mcp;
    double* lpd_s = ys2_s + (size_t)Ns * Mcp;
    const long long sBt = (long long)Mcp * Np, smu = (long long)Nt * Mcp, sv = (long long)Nt * ns * Mcp;
    double* mupart = ctx->pmu;
    double* vpart = ctx->pmu + (size_t)G * smu;
    // every kernel below covers a whole group of samples (grid z / y = sample): a predict call
    // is 3 launches + 1, not 3 per sample
    for (int s0 = 0; s0 < Ns; s0 += G) {
      const int g = std::min(G, Ns - s0);
      KsArgs ka;
      ka.md = md;
      ka.N = (int)post->N;
      ka.Np = Np;
      ka.Nt = Nt;
      ka.mc = mc;
      ka.Mcp = Mcp;
      ka.Xs = cXs;
      ka.hyp = b.hyp + (size_t)s0 * md.P;
      ka.xs = b.xs + (size_t)s0 * D * Np;
      ka.alpha = b.alpha + (size_t)s0 * Np;
      ka.sp = b.sp + s0;
      ka.Bt = ctx->pBt;
      ka.sBt = sBt;
      ka.mupart = mupart;
      ka.smu = smu;
      dim3 kgrid((unsigned)(Mcp / T), (unsigned)Nt, (unsigned)g);
      switch (kc) {
        case 0: launch_ks<0>(ctx, ka, kgrid, ks_smem); break;
        case 1: launch_ks<1>(ctx, ka, kgrid, ks_smem); break;
        case 3: launch_ks<3>(ctx, ka, kgrid, ks_smem); break;
        case 5: launch_ks<5>(ctx, ka, kgrid, ks_smem); break;
        default: launch_ks<2>(ctx, ka, kgrid, ks_smem); break;
      }
      OpPred op;
      op.Bt = ctx->pBt;
      op.ldbt = Mcp;
      op.sBt = sBt;
      op.Wbuf = b.Wbuf + (size_t)s0 * b.smat();
      op.Abuf = b.Abuf + (size_t)s0 * b.smat();
      op.smat = b.smat();
      op.sp = b.sp + s0;
      op.part = vpart;
      op.spart = sv;
      op.Mcp = Mcp;
      op.Np = Np;
      op.ns = ns;
      op.N = (int)post->N;
      op.mc = mc;
      {
        TmaOperands tmo;
        const bool have = pred_tma(ctx, b, ctx->pBt, Mcp, (long long)g * Np, tmo);
        launch_gemm(ctx, op, dim3((unsigned)(Mcp / T), (unsigned)Nt, (unsigned)g), have ? &tmo : nullptr);
      }
      FinishArgs fa;
      fa.md = md;
      fa.Nt = Nt;
      fa.nv = Nt * ns;
      fa.mc = mc;
      fa.Mcp = Mcp;
      fa.has_data = 1;
      fa.Xs = cXs;
      fa.ys = cys;
      fa.s2s = cs2s;
      fa.hyp = b.hyp + (size_t)s0 * md.P;
      fa.sp = b.sp + s0;
      fa.mupart = mupart;
      fa.smu = smu;
      fa.vpart = vpart;
      fa.sv = sv;
      fa.need_ys2 = need_ys2;
      fa.want_lpd = want_lpd && separate;
      fa.mu_s = mu_s + (size_t)s0 * Mcp;
      fa.s2_s = s2_s + (size_t)s0 * Mcp;
      fa.ys2_s = ys2_s + (size_t)s0 * Mcp;
      fa.lpd_s = lpd_s + (size_t)s0 * Mcp;
      pred_finish_kernel<<<dim3((unsigned)((mc + 255) / 256), (unsigned)g), 256, 0, ctx->stream>>>(fa);
      LAUNCHED(ctx);
    }
    const size_t ocols = separate ? Ns : 1;
    CombineArgs ca;
    ca.Ns = Ns;
    ca.mc = mc;
    ca.Mcp = Mcp;
    ca.add_noise = add_noise;
    ca.separate = separate;
    ca.want_lpd = want_lpd;
    ca.ys = cys;
    ca.mu_s = mu_s;
    ca.s2_s = s2_s;
    ca.ys2_s = ys2_s;
    ca.lpd_s = lpd_s;
    if (on_device) {
      ca.mu = mu + c0 * ocols;
      ca.s2 = s2 + c0 * ocols;
      ca.lpd = nullptr;
    } else {
      ca.mu = ctx->pout;
      ca.s2 = ctx->pout + (size_t)McpMax * ocols;
      ca.lpd = ctx->pout + (size_t)2 * McpMax * ocols;
    }
    pred_combine_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, ctx->stream>>>(ca);
    LAUNCHED(ctx);
    CK(cudaGetLastError());
    if (!on_device) {
      CK(cudaMemcpyAsync(mu + c0 * ocols, ca.mu, sizeof(double) * mc * ocols, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemcpyAsync(s2 + c0 * ocols, ca.s2, sizeof(double) * mc * ocols, cudaMemcpyDeviceToHost, ctx->stream));
      if (want_lpd)
        CK(cudaMemcpyAsync(lpd + c0 * ocols, ca.lpd, sizeof(double) * mc * ocols, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));   // scratch is reused by the next chunk
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return GPB_OK;
}

extern "C" int gpb_predict(gpb_ctx* ctx, const gpb_post* post, const double* Xs, const double* ys,
                           const double* s2s, int64_t M, int add_noise, int separate, int want_lpd,
                           double* mu, double* s2, double* lpd) {
  return predict_impl(ctx, post, Xs, ys, s2s, M, add_noise, separate, want_lpd, mu, s2, lpd, false);
}

extern "C" int gpb_predict_dev(gpb_ctx* ctx, const gpb_post* post, const double* d_Xs, int64_t M,
                               int add_noise, int separate, double* d_mu, double* d_s2) {
  return predict_impl(ctx, post, d_Xs, nullptr, nullptr, M, add_noise, separate, 0, d_mu, d_s2,
                      nullptr, true);
}

// ---------------------------------------------------------------------------------
// Bayesian quadrature (GP.quad)
// ---------------------------------------------------------------------------------
extern "C" int gpb_quad(gpb_ctx* ctx, const gpb_post* cpost, const double* mu, const double* sigma,
                        int64_t M, int compute_var, int separate, double* F, double* F_var) {
  if (!ctx || !cpost) return GPB_EINVAL;
  gpb_post* post = const_cast<gpb_post*>(cpost);
  if (!mu || !sigma || !F || M <= 0 || (compute_var && !F_var)) FAIL(GPB_EINVAL, "gpb_quad: bad arguments");
  const Model md = post->md;
  if (md.cov_kind != GPB_COV_SE)
    FAIL(GPB_EINVAL, "Bayesian quadrature only supports the squared exponential kernel.");
  if (md.nz0 != 1) FAIL(GPB_EINVAL, "gpb_quad: needs the constant noise term (hyp[cov_N] = log sigma)");
  CK(cudaSetDevice(ctx->device));
  for (int st : post->status)
    if (st) FAIL(GPB_ESTATE, "gpb_quad: a posterior sample has no valid factorisation");
  int rc = compute_var ? ensure_w(ctx, post) : GPB_OK;
  if (rc != GPB_OK) return rc;
  const Bufs& b = post->b;
  const int Ns = b.cap, D = md.D, Np = b.Np, Nt = b.Nt;
  const int Mc = (int)std::min<int64_t>(M, 8192);
  const int McpMax = round_up(Mc, T);
  const int ns = ctx->gemm_bn == 128 ? 1 : 2;
  const size_t per_sample = (size_t)McpMax * Np;
  const int G = (int)std::max<size_t>(1, std::min<size_t>((size_t)Ns, ((size_t)2 << 30) / (per_sample * 8)));
  if ((rc = grow(ctx, &ctx->pXs, &ctx->pXs_n, (size_t)3 * McpMax * std::max(D, 1))) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pBt, &ctx->pBt_n, (size_t)G * per_sample)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pmu, &ctx->ppart_n, (size_t)G * 3 * Nt * McpMax)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->psamp, &ctx->psamp_n, (size_t)4 * Ns * McpMax)) != GPB_OK) return rc;
  if ((rc = grow(ctx, &ctx->pout, &ctx->pout_n, (size_t)3 * McpMax * (separate ? Ns : 1))) != GPB_OK) return rc;
  double* dmu = ctx->pXs;
  double* dsg = ctx->pXs + (size_t)McpMax * D;
  const size_t smem = ((size_t)3 * D * T + 2 * T + 8 * T) * 8;
  for (int64_t c0 = 0; c0 < M; c0 += Mc) {
    const int mc = (int)std::min<int64_t>(Mc, M - c0);
    const int Mcp = round_up(mc, T);
    CK(cudaMemcpyAsync(dmu, mu + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dsg, sigma + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, ctx->stream));
    double* F_s = ctx->psamp;
    double* V_s = F_s + (size_t)Ns * Mcp;
    const long long sBt = (long long)Mcp * Np, smu = (long long)Nt * Mcp, sv = (long long)Nt * ns * Mcp;
    double* mupart = ctx->pmu;
    double* vpart = ctx->pmu + (size_t)G * smu;
    for (int s0 = 0; s0 < Ns; s0 += G) {               // a whole group of samples per launch
      const int g = std::min(G, Ns - s0);
      QuadArgs qa;
      qa.md = md;
      qa.N = (int)post->N;
      qa.Np = Np;
      qa.Nt = Nt;
      qa.mc = mc;
      qa.Mcp = Mcp;
      qa.mu = dmu;
      qa.sigma = dsg;
      qa.X = post->X;
      qa.hyp = b.hyp + (size_t)s0 * md.P;
      qa.alpha = b.alpha + (size_t)s0 * Np;
      qa.sp = b.sp + s0;
      qa.Bt = ctx->pBt;
      qa.sBt = sBt;
      qa.mupart = mupart;
      qa.smu = smu;
      quad_build_kernel<<<dim3((unsigned)(Mcp / T), (unsigned)Nt, (unsigned)g), 256, smem, ctx->stream>>>(qa);
      LAUNCHED(ctx);
      if (compute_var) {
        OpPred op;
        op.Bt = ctx->pBt;
        op.ldbt = Mcp;
        op.sBt = sBt;
        op.Wbuf = b.Wbuf + (size_t)s0 * b.smat();
        op.Abuf = b.Abuf + (size_t)s0 * b.smat();
        op.smat = b.smat();
        op.sp = b.sp + s0;
        op.part = vpart;
        op.spart = sv;
        op.Mcp = Mcp;
        op.Np = Np;
        op.ns = ns;
        op.N = (int)post->N;
        op.mc = mc;
        TmaOperands tmo;
        const bool have = pred_tma(ctx, b, ctx->pBt, Mcp, (long long)g * Np, tmo);
        launch_gemm(ctx, op, dim3((unsigned)(Mcp / T), (unsigned)Nt, (unsigned)g), have ? &tmo : nullptr);
      }
      QuadFinishArgs fa;
      fa.md = md;
      fa.Nt = Nt;
      fa.nv = Nt * ns;
      fa.mc = mc;
      fa.Mcp = Mcp;
      fa.compute_var = compute_var;
      fa.mu = dmu;
      fa.sigma = dsg;
      fa.hyp = b.hyp + (size_t)s0 * md.P;
      fa.mupart = mupart;
      fa.smu = smu;
      fa.vpart = vpart;
      fa.sv = sv;
      fa.F_s = F_s + (size_t)s0 * Mcp;
      fa.V_s = V_s + (size_t)s0 * Mcp;
      quad_finish_kernel<<<dim3((unsigned)((mc + 255) / 256), (unsigned)g), 256, 0, ctx->stream>>>(fa);
      LAUNCHED(ctx);
    }
    const size_t ocols = separate ? Ns : 1;
    if (!compute_var) CK(cudaMemsetAsync(V_s, 0, sizeof(double) * Ns * Mcp, ctx->stream));
    CombineArgs ca;
    ca.Ns = Ns;
    ca.mc = mc;
    ca.Mcp = Mcp;
    ca.add_noise = 0;
    ca.separate = separate;
    ca.want_lpd = 0;
    ca.ys = nullptr;
    ca.mu_s = F_s;
    ca.s2_s = V_s;
    ca.ys2_s = V_s;
    ca.lpd_s = nullptr;
    ca.mu = ctx->pout;
    ca.s2 = ctx->pout + (size_t)McpMax * ocols;
    ca.lpd = nullptr;
    pred_combine_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, ctx->stream>>>(ca);
    LAUNCHED(ctx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(F + c0 * ocols, ca.mu, sizeof(double) * mc * ocols, cudaMemcpyDeviceToHost, ctx->stream));
    if (compute_var)
      CK(cudaMemcpyAsync(F_var + c0 * ocols, ca.s2, sizeof(double) * mc * ocols, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaGetLastError());
  return GPB_OK;
}

// ---------------------------------------------------------------------------------
// plugin surface
// ---------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------
// full predictive covariance (GP.predict_full)
// ---------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------
// rank-one append (GP.update, gaussian_process.py:737-844)
// ---------------------------------------------------------------------------------
extern "C" int64_t gpb_posterior_size(const gpb_post* post) { return post ? post->N : 0; }

template <int KIND>
static void launch_append_ks(gpb_ctx* ctx, const AppendArgs& a, dim3 grid) {
  append_ks_kernel<KIND><<<grid, 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

static int predict_impl(gpb_ctx* ctx, const gpb_post* cpost, const double* Xs, const double* ys,
                        const double* s2s, int64_t M, int add_noise, int separate, int want_lpd,
                        double* mu, double* s2, double* lpd, bool on_device);

extern "C" int gpb_posterior_append(gpb_ctx* ctx, gpb_post* post, const double* x_new, double y_new,
                                    int32_t* status) {
  if (!ctx || !post || !x_new || !status) return GPB_EINVAL;
  NvtxRange nv("gpb:K5 rank-one append");
  if (post->ctx != ctx) FAIL(GPB_EINVAL, "gpb_posterior_append: posterior belongs to another context");
  const Model md = post->md;
  if (md.nz1 != 0 || md.nz2 != 0)
    FAIL(GPB_EAGAIN, "gpb_posterior_append: point-dependent noise, rebuild the batch");
  Bufs& b = post->b;
  const int N = (int)post->N, Np = b.Np, Ns = b.cap, D = md.D;
  if (N + 1 > Np) FAIL(GPB_EAGAIN, "gpb_posterior_append: no free row in the padded layout, rebuild the batch");
  for (int st : post->status)
    if (st) FAIL(GPB_ESTATE, "gpb_posterior_append: a posterior sample has no valid factorisation");
  CK(cudaSetDevice(ctx->device));
  // m*, v* of the new point under the current posterior, per sample (:756-758)
  double* dx = b.sn2v;                       // scratch of the posterior's own buffers
  double* mstar = b.nlz;
  double* vstar = b.logdet;                  // Nt >= 1 doubles per sample
  CK(cudaMemcpyAsync(dx, x_new, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
  int rc = predict_impl(ctx, post, dx, nullptr, nullptr, 1, /*add_noise=*/1, /*separate=*/1, 0, mstar,
                        vstar, nullptr, /*on_device=*/true);
  if (rc != GPB_OK) return rc;
  AppendArgs a;
  a.md = md;
  a.N = N;
  a.Np = Np;
  a.xnew = dx;
  a.ynew = y_new;
  a.hyp = b.hyp;
  a.sp = b.sp;
  a.xs = b.xs;
  a.Abuf = b.Abuf;
  a.Wbuf = b.Wbuf;
  a.smat = b.smat();
  a.alpha = b.alpha;
  a.kvec = b.bvec;
  a.cvec = b.zvec;
  a.avec = b.resid;
  a.mstar = mstar;
  a.vstar = vstar;
  a.status = b.fail;
  const dim3 kgrid((unsigned)((Np + 255) / 256), (unsigned)Ns);
  switch (kind_code(md.cov_kind, md.degree)) {
    case 0: launch_append_ks<0>(ctx, a, kgrid); break;
    case 1: launch_append_ks<1>(ctx, a, kgrid); break;
    case 3: launch_append_ks<3>(ctx, a, kgrid); break;
    case 5: launch_append_ks<5>(ctx, a, kgrid); break;
    default: launch_append_ks<2>(ctx, a, kgrid); break;
  }
  bool any_high = false, any_low = false;
  for (const SlotP& p : post->sp) (p.lchol ? any_high : any_low) = true;
  const int nt = (N + T - 1) / T;
  if (any_high) {
    append_gemv_n_kernel<<<dim3((unsigned)nt, (unsigned)Ns), 256, 0, ctx->stream>>>(a);
    LAUNCHED(ctx);
  }
  append_gemv_t_kernel<<<dim3((unsigned)((N + 7) / 8), (unsigned)Ns), 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
  if (any_low) {
    append_outer_kernel<<<dim3((unsigned)nt, (unsigned)nt, (unsigned)Ns), 256, 0, ctx->stream>>>(a);
    LAUNCHED(ctx);
  }
  append_finish_kernel<<<(unsigned)Ns, 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(post->X + (size_t)N * D, dx, sizeof(double) * D, cudaMemcpyDeviceToDevice, ctx->stream));
  std::vector<int> st((size_t)Ns);
  CK(cudaMemcpyAsync(st.data(), b.fail, sizeof(int) * Ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  bool ok = true;
  for (int s = 0; s < Ns; ++s) {
    status[s] = st[s];
    ok = ok && st[s] == 0;
  }
  // The stable samples now hold N+1 points.  An unstable one was left untouched: it is marked
  // unusable until the caller recomputes it on the extended data (gpb_posterior_rebuild), which is
  // what the reference does for exactly those samples (gaussian_process.py:864-868)
  (void)ok;
  post->N = N + 1;
  for (int s = 0; s < Ns; ++s)
    if (st[s]) post->status[s] = 1;
  return GPB_OK;
}

extern "C" int gpb_posterior_rebuild(gpb_ctx* ctx, gpb_post* post, const int32_t* slots, int64_t n) {
  if (!ctx || !post || !slots || n <= 0) return GPB_EINVAL;
  NvtxRange nv_call("gpb_posterior_rebuild");
  if (post->ctx != ctx) FAIL(GPB_EINVAL, "gpb_posterior_rebuild: posterior belongs to another context");
  Bufs& b = post->b;
  const Model md = post->md;
  if (!ctx->has_data || ctx->N != post->N || ctx->Np != b.Np || ctx->D != md.D)
    FAIL(GPB_ESTATE, "gpb_posterior_rebuild: the context must hold the data the posteriors cover (gpb_set_data first)");
  CK(cudaSetDevice(ctx->device));
  std::vector<int> list;
  for (int64_t i = 0; i < n; ++i) {
    if (slots[i] < 0 || slots[i] >= b.cap) FAIL(GPB_EINVAL, "gpb_posterior_rebuild: sample index out of range");
    list.push_back(slots[i]);
  }
  int rc = factor_with_retry(ctx, b, md, post->N, list, true, post->status);
  if (rc != GPB_OK) return rc;
  CK(cudaMemcpyAsync(b.sel3, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
  run_bwd(ctx, b, b.sel3, (int)list.size());
  std::vector<SlotP> all((size_t)b.cap);
  CK(cudaMemcpyAsync(all.data(), b.sp, sizeof(SlotP) * b.cap, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<int> low, high;
  for (int s : list) {
    post->sp[s] = all[s];
    if (post->status[s]) continue;
    (all[s].lchol ? high : low).push_back(s);
  }
  if (!low.empty()) {               // L = -(K + sn2_mult*diag(sn2))^-1, both triangles (gaussian_process.py:2440-2448)
    CK(cudaMemcpyAsync(b.sel2, low.data(), sizeof(int) * low.size(), cudaMemcpyHostToDevice, ctx->stream));
    run_inverse(ctx, b, post->N, b.sel2, (int)low.size(), true, true);
    for (int s : low) {
      symmetrize_kernel<<<grid1d(b.smat()), 256, 0, ctx->stream>>>(b.Abuf + s * b.smat(), b.Np);
      LAUNCHED(ctx);
    }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (!high.empty() && post->w_ready) {   // the other samples' W = L^-1 exists: keep the batch uniform
    CK(cudaMemcpyAsync(b.sel2, high.data(), sizeof(int) * high.size(), cudaMemcpyHostToDevice, ctx->stream));
    run_inverse(ctx, b, post->N, b.sel2, (int)high.size(), false, false);
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaGetLastError());
  return GPB_OK;
}

template <int KIND>
static void launch_full(gpb_ctx* ctx, const FullArgs& a) {
  full_finish_kernel<KIND><<<grid1d((long long)a.M * a.M), 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

extern "C" int gpb_predict_full(gpb_ctx* ctx, const gpb_post* cpost, const double* Xs, const double* ys,
                                const double* s2s, int64_t M, int add_noise, double* mu, double* cov) {
  if (!ctx || !cpost) return GPB_EINVAL;
  gpb_post* post = const_cast<gpb_post*>(cpost);
  if (!Xs || !mu || !cov || M <= 0) FAIL(GPB_EINVAL, "gpb_predict_full: bad arguments");
  if (M > 8192) FAIL(GPB_EINVAL, "gpb_predict_full: M > 8192 (the M x M covariance per sample is the limit)");
  CK(cudaSetDevice(ctx->device));
  for (int st : post->status)
    if (st) FAIL(GPB_ESTATE, "gpb_predict_full: a posterior sample has no valid factorisation");
  int rc = ensure_w(ctx, post);
  if (rc != GPB_OK) return rc;
  const Bufs& b = post->b;
  const Model md = post->md;
  const int Ns = b.cap, D = md.D, Np = b.Np, Nt = b.Nt;
  const int Mcp = round_up(M, T);
  double *dXs = nullptr, *dys = nullptr, *ds2s = nullptr, *Bt = nullptr, *Vt = nullptr, *C2 = nullptr,
         *mupart = nullptr, *dmu = nullptr, *dcov = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dXs, dys, ds2s, Bt, Vt, C2, mupart, dmu, dcov})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dXs, sizeof(double) * M * D));
  CKC(cudaMemcpyAsync(dXs, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, ctx->stream));
  if (ys) {
    CKC(cudaMalloc(&dys, sizeof(double) * M));
    CKC(cudaMemcpyAsync(dys, ys, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (s2s) {
    CKC(cudaMalloc(&ds2s, sizeof(double) * M));
    CKC(cudaMemcpyAsync(ds2s, s2s, sizeof(double) * M, cudaMemcpyHostToDevice, ctx->stream));
  }
  CKC(cudaMalloc(&Bt, sizeof(double) * Mcp * Np));
  CKC(cudaMalloc(&Vt, sizeof(double) * Mcp * Np));
  CKC(cudaMalloc(&C2, sizeof(double) * Mcp * Mcp));
  CKC(cudaMalloc(&mupart, sizeof(double) * Nt * Mcp));
  CKC(cudaMalloc(&dmu, sizeof(double) * M * Ns));
  CKC(cudaMalloc(&dcov, sizeof(double) * M * M));
  const int kc = kind_code(md.cov_kind, md.degree);
  const size_t ks_smem = ((size_t)2 * D * T + T + 8 * T) * 8;
  // the reference's noise is an (M,1) array as soon as a per-point term is active (noise_functions.py:258-278)
  const int per_point = ((md.nz1 > 0 && s2s) || (md.nz2 == 1 && ys)) ? 1 : 0;
  for (int s = 0; s < Ns; ++s) {
    const SlotP& p = post->sp[s];
    KsArgs ka;
    ka.md = md;
    ka.N = (int)post->N;
    ka.Np = Np;
    ka.Nt = Nt;
    ka.mc = (int)M;
    ka.Mcp = Mcp;
    ka.Xs = dXs;
    ka.hyp = b.hyp + (size_t)s * md.P;
    ka.xs = b.xs + (size_t)s * D * Np;
    ka.alpha = b.alpha + (size_t)s * Np;
    ka.sp = b.sp + s;
    ka.Bt = Bt;
    ka.sBt = 0;
    ka.mupart = mupart;
    ka.smu = 0;
    dim3 kgrid((unsigned)(Mcp / T), (unsigned)Nt, 1);
    switch (kc) {
      case 0: launch_ks<0>(ctx, ka, kgrid, ks_smem); break;
      case 1: launch_ks<1>(ctx, ka, kgrid, ks_smem); break;
      case 3: launch_ks<3>(ctx, ka, kgrid, ks_smem); break;
      case 5: launch_ks<5>(ctx, ka, kgrid, ks_smem); break;
      default: launch_ks<2>(ctx, ka, kgrid, ks_smem); break;
    }
    if (p.lchol) {
      // V^T = Bt W^T (Mcp x Np), then C2 = V^T V (Mcp x Mcp)
      // (W is lower triangular; the upper tiles of Wbuf hold W^T: restrict k per column tile)
      launch_gemm(ctx, OpPlain{Bt, b.Wbuf + (size_t)s * b.smat(), Vt, Mcp, Np, Mcp, Np, /*tri=*/1},
                  dim3((unsigned)(Mcp / T), (unsigned)Nt));
      launch_gemm(ctx, OpPlain{Vt, Vt, C2, Mcp, Mcp, Mcp, Np}, dim3((unsigned)(Mcp / T), (unsigned)(Mcp / T)));
    } else {
      // T1 = Ks^T Ainv (Mcp x Np), then C2 = T1 Ks (Mcp x Mcp)
      launch_gemm(ctx, OpPlain{Bt, b.Abuf + (size_t)s * b.smat(), Vt, Mcp, Np, Mcp, Np},
                  dim3((unsigned)(Mcp / T), (unsigned)Nt));
      launch_gemm(ctx, OpPlain{Vt, Bt, C2, Mcp, Mcp, Mcp, Np}, dim3((unsigned)(Mcp / T), (unsigned)(Mcp / T)));
    }
    FullArgs fa;
    fa.md = md;
    fa.Nt = Nt;
    fa.M = (int)M;
    fa.Mcp = Mcp;
    fa.Xs = dXs;
    fa.ys = dys;
    fa.s2s = ds2s;
    fa.hyp = b.hyp + (size_t)s * md.P;
    fa.sp = p;
    fa.mupart = mupart;
    fa.C2 = C2;
    fa.add_noise = add_noise;
    fa.per_point = per_point;
    fa.mu = dmu + s;
    fa.Ns = Ns;
    fa.cov = dcov;
    switch (kc) {
      case 0: launch_full<0>(ctx, fa); break;
      case 1: launch_full<1>(ctx, fa); break;
      case 3: launch_full<3>(ctx, fa); break;
      case 5: launch_full<5>(ctx, fa); break;
      default: launch_full<2>(ctx, fa); break;
    }
    CKC(cudaMemcpyAsync(cov + (size_t)s * M * M, dcov, sizeof(double) * M * M, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
  }
  CKC(cudaMemcpyAsync(mu, dmu, sizeof(double) * M * Ns, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  cleanup();
  return GPB_OK;
}

template <int KIND>
static void launch_cov(gpb_ctx* ctx, const CovArgs& a, long long total) {
  cov_plugin_kernel<KIND><<<grid1d(total), 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
}

extern "C" int gpb_cov(gpb_ctx* ctx, int cov_kind, int matern_degree, int ard, const double* hyp,
                       const double* X, int64_t N, int D, const double* Xs, int64_t M, int diag,
                       double* K, double* dK) {
  if (!ctx) return GPB_EINVAL;
  if (!hyp || !X || !K || N <= 0 || D <= 0) FAIL(GPB_EINVAL, "gpb_cov: bad arguments");
  if (dK && (Xs || diag)) FAIL(GPB_EINVAL, "X_star should be None when compute_grad is True.");
  Model md{};
  const int nz[3] = {0, 0, 0};
  if (fill_model(md, cov_kind, matern_degree, ard, 0, nz, D) != GPB_OK)
    FAIL(GPB_EINVAL, "gpb_cov: unsupported covariance descriptor");
  CK(cudaSetDevice(ctx->device));
  const long long cols = diag ? 1 : (Xs ? M : N);
  const long long total = N * cols;
  double *dh = nullptr, *dX = nullptr, *dXs = nullptr, *dKd = nullptr, *ddK = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dh, dX, dXs, dKd, ddK})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dh, sizeof(double) * md.cov_n));
  CKC(cudaMalloc(&dX, sizeof(double) * N * D));
  CKC(cudaMalloc(&dKd, sizeof(double) * total));
  CKC(cudaMemcpyAsync(dh, hyp, sizeof(double) * md.cov_n, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(dX, X, sizeof(double) * N * D, cudaMemcpyHostToDevice, ctx->stream));
  if (Xs && !diag) {
    CKC(cudaMalloc(&dXs, sizeof(double) * M * D));
    CKC(cudaMemcpyAsync(dXs, Xs, sizeof(double) * M * D, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (dK) CKC(cudaMalloc(&ddK, sizeof(double) * md.cov_n * N * N));
  CovArgs a;
  a.cov_kind = cov_kind;
  a.degree = matern_degree;
  a.ard = md.ard;
  a.D = D;
  a.N = N;
  a.M = M;
  a.hyp = dh;
  a.X = dX;
  a.Xs = diag ? nullptr : dXs;
  a.diag = diag;
  a.K = dKd;
  a.dK = ddK;
  switch (kind_code(cov_kind, matern_degree)) {
    case 0: launch_cov<0>(ctx, a, total); break;
    case 1: launch_cov<1>(ctx, a, total); break;
    case 3: launch_cov<3>(ctx, a, total); break;
    case 5: launch_cov<5>(ctx, a, total); break;
    default: launch_cov<2>(ctx, a, total); break;
  }
  CKC(cudaMemcpyAsync(K, dKd, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
  if (dK) CKC(cudaMemcpyAsync(dK, ddK, sizeof(double) * md.cov_n * N * N, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  cleanup();
  return GPB_OK;
}

extern "C" int gpb_mean(gpb_ctx* ctx, int mean_kind, const double* hyp, const double* X, int64_t N,
                        int D, double* m, double* dm) {
  if (!ctx) return GPB_EINVAL;
  if (!X || !m || N <= 0 || D <= 0 || mean_kind < 0 || mean_kind > 2) FAIL(GPB_EINVAL, "gpb_mean: bad arguments");
  const int mn = mean_count(mean_kind, D);
  if (mn > 0 && !hyp) FAIL(GPB_EINVAL, "gpb_mean: hyp is NULL");
  CK(cudaSetDevice(ctx->device));
  double *dh = nullptr, *dX = nullptr, *dm_ = nullptr, *ddm = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dh, dX, dm_, ddm})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dh, sizeof(double) * std::max(mn, 1)));
  CKC(cudaMalloc(&dX, sizeof(double) * N * D));
  CKC(cudaMalloc(&dm_, sizeof(double) * N));
  if (mn > 0) CKC(cudaMemcpyAsync(dh, hyp, sizeof(double) * mn, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(dX, X, sizeof(double) * N * D, cudaMemcpyHostToDevice, ctx->stream));
  if (dm && mn > 0) CKC(cudaMalloc(&ddm, sizeof(double) * N * mn));
  MeanArgs a;
  a.mean_kind = mean_kind;
  a.D = D;
  a.N = N;
  a.hyp = dh;
  a.X = dX;
  a.m = dm_;
  a.dm = ddm;
  mean_plugin_kernel<<<grid1d(N), 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
  CKC(cudaMemcpyAsync(m, dm_, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (ddm) CKC(cudaMemcpyAsync(dm, ddm, sizeof(double) * N * mn, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  cleanup();
  return GPB_OK;
}

extern "C" int gpb_noise(gpb_ctx* ctx, const int noise_flags[3], const double* hyp, const double* y,
                         const double* s2, int64_t N, double* sn2, double* dsn2) {
  if (!ctx) return GPB_EINVAL;
  if (!noise_flags || !sn2 || N <= 0) FAIL(GPB_EINVAL, "gpb_noise: bad arguments");
  const int nn = noise_count(noise_flags[0], noise_flags[1], noise_flags[2]);
  if (nn > 0 && !hyp) FAIL(GPB_EINVAL, "gpb_noise: hyp is NULL");
  CK(cudaSetDevice(ctx->device));
  double *dh = nullptr, *dy = nullptr, *ds2 = nullptr, *dsn = nullptr, *ddsn = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dh, dy, ds2, dsn, ddsn})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dh, sizeof(double) * std::max(nn, 1)));
  CKC(cudaMalloc(&dsn, sizeof(double) * N));
  if (nn > 0) CKC(cudaMemcpyAsync(dh, hyp, sizeof(double) * nn, cudaMemcpyHostToDevice, ctx->stream));
  if (y) {
    CKC(cudaMalloc(&dy, sizeof(double) * N));
    CKC(cudaMemcpyAsync(dy, y, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (s2) {
    CKC(cudaMalloc(&ds2, sizeof(double) * N));
    CKC(cudaMemcpyAsync(ds2, s2, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (dsn2 && nn > 0) CKC(cudaMalloc(&ddsn, sizeof(double) * N * nn));
  NoiseArgs a;
  a.nz0 = noise_flags[0];
  a.nz1 = noise_flags[1];
  a.nz2 = noise_flags[2];
  a.N = N;
  a.hyp = dh;
  a.y = dy;
  a.s2 = ds2;
  a.sn2 = dsn;
  a.dsn2 = ddsn;
  noise_plugin_kernel<<<grid1d(N), 256, 0, ctx->stream>>>(a);
  LAUNCHED(ctx);
  CKC(cudaMemcpyAsync(sn2, dsn, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
  if (ddsn) CKC(cudaMemcpyAsync(dsn2, ddsn, sizeof(double) * N * nn, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  cleanup();
  return GPB_OK;
}

// ---------------------------------------------------------------------------------
// test / measurement hooks
// ---------------------------------------------------------------------------------
extern "C" int gpb_debug_gemm_nt(gpb_ctx* ctx, const double* A, const double* B, double* C, int M,
                                 int N, int K, double alpha, double beta) {
  if (!ctx) return GPB_EINVAL;
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || M % BM || N % BN || K % BK)
    FAIL(GPB_EINVAL, "gpb_debug_gemm_nt: M,N must be multiples of 128 and K of 16");
  CK(cudaSetDevice(ctx->device));
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dA, dB, dC})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dA, sizeof(double) * M * K));
  CKC(cudaMalloc(&dB, sizeof(double) * N * K));
  CKC(cudaMalloc(&dC, sizeof(double) * M * N));
  CKC(cudaMemcpyAsync(dA, A, sizeof(double) * M * K, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(dB, B, sizeof(double) * N * K, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(dC, C, sizeof(double) * M * N, cudaMemcpyHostToDevice, ctx->stream));
  OpGeneric op{dA, dB, dC, M, N, M, K, alpha, alpha != 0.0 ? beta / alpha : 0.0};
  launch_gemm(ctx, op, dim3(M / BM, N / BN));
  CKC(cudaMemcpyAsync(C, dC, sizeof(double) * M * N, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  cleanup();
  return GPB_OK;
}

extern "C" int gpb_debug_gemm_bench(gpb_ctx* ctx, int M, int N, int K, int reps, double* ms) {
  if (!ctx) return GPB_EINVAL;
  if (!ms || M <= 0 || N <= 0 || K <= 0 || M % BM || N % BN || K % BK || reps <= 0)
    FAIL(GPB_EINVAL, "gpb_debug_gemm_bench: bad arguments");
  CK(cudaSetDevice(ctx->device));
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  auto cleanup = [&]() {
    for (double* p : {dA, dB, dC})
      if (p) cudaFree(p);
  };
  CKC(cudaMalloc(&dA, sizeof(double) * M * K));
  CKC(cudaMalloc(&dB, sizeof(double) * N * K));
  CKC(cudaMalloc(&dC, sizeof(double) * M * N));
  fill_kernel<<<grid1d((long long)M * K), 256, 0, ctx->stream>>>(dA, 0.5, (long long)M * K);
  fill_kernel<<<grid1d((long long)N * K), 256, 0, ctx->stream>>>(dB, 0.25, (long long)N * K);
  fill_kernel<<<grid1d((long long)M * N), 256, 0, ctx->stream>>>(dC, 0.0, (long long)M * N);
  OpGeneric op{dA, dB, dC, M, N, M, K, -1.0, -1.0};   // the trailing-update form C - A B^T
  for (int i = 0; i < 3; ++i) launch_gemm(ctx, op, dim3(M / BM, N / BN));
  CKC(cudaEventRecord(ctx->ev[6], ctx->stream));
  for (int i = 0; i < reps; ++i) launch_gemm(ctx, op, dim3(M / BM, N / BN));
  CKC(cudaEventRecord(ctx->ev[7], ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  CKC(cudaGetLastError());
  float t = 0;
  CKC(cudaEventElapsedTime(&t, ctx->ev[6], ctx->ev[7]));
  *ms = t / reps;
  cleanup();
  return GPB_OK;
}

extern "C" int gpb_debug_potrf(gpb_ctx* ctx, double* A, int n, int32_t* info) {
  if (!ctx) return GPB_EINVAL;
  if (!A || n <= 0 || !info) FAIL(GPB_EINVAL, "gpb_debug_potrf: bad arguments");
  CK(cudaSetDevice(ctx->device));
  Model md{};
  const int nz[3] = {1, 0, 0};
  fill_model(md, 0, 0, 1, 0, nz, 1);
  const int Np = round_up(n, T);
  Bufs b;
  int rc = alloc_bufs(ctx, b, 1, false, Np, 1, md);
  if (rc != GPB_OK) { free_bufs(b); return rc; }
  std::vector<double> pad((size_t)Np * Np, 0.0);
  for (int c = 0; c < Np; ++c)
    for (int r = 0; r < Np; ++r)
      pad[(size_t)c * Np + r] = (r < n && c < n) ? A[(size_t)c * n + r] : (r == c ? 1.0 : 0.0);
  int zero = 0;
  cudaError_t e = cudaMemcpy(b.Abuf, pad.data(), sizeof(double) * pad.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b.sel, &zero, sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(b.fail, 0, sizeof(int));
  if (e == cudaSuccess) {
    run_potrf(ctx, b, n, b.sel, 1, false, false);
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(pad.data(), b.Abuf, sizeof(double) * pad.size(), cudaMemcpyDeviceToHost);
  int failh = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&failh, b.fail, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && ctx->diag_dbg) {
    long long st[64];
    cudaMemcpy(st, ctx->diag_dbg, sizeof st, cudaMemcpyDeviceToHost);
    fprintf(stderr, "diag_kernel phase cycles (n=%d):", n);
    for (int i = 1; i < (int)st[0]; ++i) fprintf(stderr, " %lld", st[1 + i] - st[i]);
    fprintf(stderr, "  total %lld\n", st[st[0]] - st[1]);
    // the factor warp's own timeline: (start, end) of each 32x32 block factorisation, relative to the kernel's first stamp
    fprintf(stderr, "  factor warp (start..end of each block factorisation):");
    for (int i = 0; i + 1 < (int)st[40] && i + 1 < 16; i += 2)
      fprintf(stderr, " %lld..%lld (%lld = pivot loop %lld + 16x16 inverses %lld + off-diagonal block %lld)", st[41 + i] - st[1],
              st[42 + i] - st[1], st[42 + i] - st[41 + i], st[56 + i] - st[41 + i], st[57 + i] - st[56 + i],
              st[42 + i] - st[57 + i]);
    fprintf(stderr, "\n");
  }
  free_bufs(b);
  if (e != cudaSuccess) FAIL(GPB_ECUDA, cudaGetErrorString(e));
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) A[(size_t)c * n + r] = (r >= c) ? pad[(size_t)c * Np + r] : 0.0;
  *info = failh;
  return GPB_OK;
}

// Latency of the diagonal-tile kernel in a dependent chain: `reps` launches in stream order (no PDL), each
// on its own copy of one SPD 128x128 tile.  us = average microseconds per launch (CUDA events).
extern "C" int gpb_debug_diag_bench(gpb_ctx* ctx, int reps, int with_rhs, double* us) {
  if (!ctx || !us || reps <= 0 || reps > 4096) return GPB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  Model md{};
  const int nz[3] = {1, 0, 0};
  fill_model(md, 0, 0, 1, 0, nz, 1);
  Bufs b;
  int rc = alloc_bufs(ctx, b, reps, false, T, 1, md);
  if (rc != GPB_OK) { free_bufs(b); return rc; }
  std::vector<double> tile((size_t)T * T);
  for (int c = 0; c < T; ++c)
    for (int r = 0; r < T; ++r) tile[(size_t)c * T + r] = (r == c ? 2.0 : 0.0) + 0.5 / (1.0 + (r > c ? r - c : c - r));
  std::vector<int> ident((size_t)reps);
  for (int i = 0; i < reps; ++i) ident[i] = i;
  cudaError_t e = cudaMemcpy(b.sel, ident.data(), sizeof(int) * reps, cudaMemcpyHostToDevice);
  for (int i = 0; i < reps && e == cudaSuccess; ++i)
    e = cudaMemcpy(b.Abuf + (size_t)i * T * T, tile.data(), sizeof(double) * T * T, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(b.fail, 0, sizeof(int) * reps);
  if (e == cudaSuccess) e = cudaMemset(b.bvec, 0, sizeof(double) * T * reps);
  const bool pdl0 = ctx->pdl;
  ctx->pdl = false;
  auto launch = [&](int s) {
    DiagArgs da{};
    da.Abuf = b.Abuf; da.Wbuf = nullptr; da.Dbuf = b.Dbuf; da.DTbuf = b.DTbuf;
    da.sel = b.sel + s; da.smat = b.smat(); da.Np = T; da.Nt = 1; da.N = T; da.k = 0;
    da.bvec = with_rhs ? b.bvec : nullptr; da.zvec = with_rhs ? b.zvec : nullptr;
    da.logdet = b.logdet; da.fail = b.fail; da.dbg = ctx->diag_dbg;
    launch_chain(ctx, diag_kernel, dim3(1), dim3(256), DIAG_SMEM, da);
  };
  if (e == cudaSuccess) {
    for (int s = 0; s < std::min(reps, 4); ++s) launch(s);
    cudaEventRecord(ctx->ev[6], ctx->stream);
    for (int s = 0; s < reps; ++s) launch(s);
    cudaEventRecord(ctx->ev[7], ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
  }
  ctx->pdl = pdl0;
  float ms = 0;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
  int failh = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&failh, b.fail, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && ctx->diag_dbg) {
    long long st[33];
    cudaMemcpy(st, ctx->diag_dbg, sizeof st, cudaMemcpyDeviceToHost);
    fprintf(stderr, "diag_kernel phase cycles:");
    for (int i = 1; i < (int)st[0]; ++i) fprintf(stderr, " %lld", st[1 + i] - st[i]);
    fprintf(stderr, "  total %lld\n", st[st[0]] - st[1]);
    long long s0[24];
    cudaMemcpy(s0, ctx->diag_dbg + 40, sizeof s0, cudaMemcpyDeviceToHost);
    fprintf(stderr, "factor warp (cycles since kernel start; begin/end of each block factorisation):");
    for (int i = 0; i < (int)s0[0] && i < 16; ++i) fprintf(stderr, " %lld", s0[1 + i] - st[1]);
    fprintf(stderr, "\n");
  }
  free_bufs(b);
  if (e != cudaSuccess) FAIL(GPB_ECUDA, cudaGetErrorString(e));
  if (failh) FAIL(GPB_ESTATE, "gpb_debug_diag_bench: the test tile did not factor");
  *us = 1e3 * ms / reps;
  return GPB_OK;
}
