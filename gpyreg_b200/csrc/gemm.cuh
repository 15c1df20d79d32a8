// FP64 tensor-core tile GEMM for sm_100a:  C(m,n) = alpha * sum_k A(m,k) * B(n,k) + beta * C(m,n)
//
// Blackwell's tcgen05/UMMA path has no f64 kind; the FP64 tensor instruction on
// sm_100a is the warp-level DMMA (PTX mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4).
// Every dense contraction of the GP hot path (Cholesky panel + trailing update,
// triangular inverse, K^-1 = W^T W, predictive-variance solve) is phrased as this ONE
// "NT" product on column-major operands, so both operand tiles are contiguous along
// M / N and stream into shared memory with 16-byte async copies.
//
// CTA tile 128x128, K step 16, 8 warps (4 along M x 2 along N, 32x64 per warp),
// 4-stage cp.async ring (132 KB of the 227 KB shared memory), one CTA per SM.
// Shared tiles are stored [k][m] with a pitch of 132 doubles so the DMMA fragment
// loads (lane -> (m = lane/4, k = lane%4)) hit 16 distinct 8-byte banks per half warp.
#pragma once
#include <cuda.h>      // CUtensorMap (the encode function is fetched at run time: no libcuda link)

#include "common.cuh"

namespace gpb {

constexpr int BM = 128, BN = 128, BK = 16, NSTAGE = 4;

// Tensor maps of the operand sources of one launch (LOADER = 2).  Every operand of the factorisation
// lives in one of four per-batch buffers, each of which is ONE 2-D column-major array: Abuf / Wbuf
// are Np rows x (slots*Np) columns, Dbuf / DTbuf are T rows x (slots*Nt*T) columns.  A stage of an
// operand is then a single box of (tile rows + 4) x BK elements: the 4 extra rows are the padding
// of the shared-memory pitch (132 / 68 doubles, conflict-free DMMA fragment loads), fetched as
// real neighbouring rows or zero-filled beyond the array -- nothing reads them.
struct alignas(64) TmaOperands {
  CUtensorMap a, b, a0, b0;          // A, B and their alternate sources for k in [0, T)
  CUtensorMap b_alt;                 // second B array, chosen per tile (predict: W or the explicit inverse)
  const double* base[5];             // base address of the array behind each map
  long long ld[5];                   // its leading dimension (elements)
};
constexpr int PITCH = BM + 4;
constexpr int GEMM_THREADS = 256;

struct GemmTile {
  const double* A;  long long lda;    // A(m,k) at A[m + k*lda]
  const double* B;  long long ldb;    // B(n,k) at B[n + k*ldb]
  const double* A0; long long lda0;   // optional alternate source for k in [0,T)
  const double* B0; long long ldb0;
  double* C;  long long ldc;          // C(m,n) at C[m + n*ldc]       (may be null)
  double* Ct; long long ldct;         // also/only store C(m,n) at Ct[n + m*ldct]
  const double* E; long long lde;     // reduce mode: rowsum[m] = sum_n C(m,n)*E(m,n) (E null: C^2)
  double* rowsum; long long rs_half;  // rs_half: offset between the two column halves
  const double* vdot; double* vdst;   // GM_ROWDOT: vdst[m] -= sum_n C(m,n) * vdot[n]
  const double* zD; const double* zb; double* zout; int znact;   // GM_ZSOLVE: D_k, b_k, where z_k goes (or null)
  int K;
  int b_alt;                          // LOADER 2: B lives in the launch's alternate B array (TmaOperands::b_alt)
  int mvalid, nvalid;                 // rows / columns of the 128x128 tile that hold data (rest is padding)
  // Triangular operand tiles (the diagonal tiles D_k, D_k^T, W_ii): the K steps in which a warp's rows or
  // columns only meet the stored zeros of the triangle are skipped (adding 0 * x changes no bit).
  //   tri_lo bit 0: first tile (k < T), operand zero for k < m;  bit 1: zero for k < n
  //   hi_m_off / hi_n_off >= 0: in the tile that starts at this k offset the operand is zero for
  //   k - off > m  /  k - off > n
  int tri_lo, hi_m_off, hi_n_off;
  // diagonal output tile of a symmetric result (potrf trailing update, W^T W): only its lower triangle is ever
  // read, so the warps whose 32 x WN block lies strictly above the diagonal sit the product out
  int lower_only;
  double alpha, cscale;               // result = alpha * (cscale * C + A B^T); cscale = beta / alpha
  bool valid;
};

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* smem, const double* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- TMA bulk-copy + mbarrier helpers (cp.async.bulk -> SASS UBLKCP) ---------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// one contiguous run of `bytes` (multiple of 16) global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// kernel parameter carrying the maps: only the LOADER = 2 instantiation pays for the 700 bytes
struct NoTma {};
template <int LOADER> struct TmaParam { typedef NoTma type; };
template <> struct TmaParam<2> { typedef TmaOperands type; };

// one box of a 2-D tensor map -> shared memory (cp.async.bulk.tensor -> SASS UTMALDG), completion
// counted on `bar`; c0 = coordinate along the contiguous dimension (row), c1 = column
__device__ __forceinline__ void tma_box_2d(double* dst, const CUtensorMap* map, int c0, int c1,
                                           unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// MODE bits
constexpr int GM_BETA = 1;     // C enters the product: accumulators start at (beta/alpha) * C
constexpr int GM_STORE = 2;    // normal store
constexpr int GM_STORET = 4;   // transposed store
constexpr int GM_REDUCE = 8;   // row-sum epilogue (no store)

constexpr int GM_ZSOLVE = 32;  // before the row-dot: every CTA first computes vdot = z_k = D_k b_k itself (solve-only replay)
constexpr int GM_ROWDOT = 16;  // after the store: vdst[m] -= sum_n C(m,n) * vdot[n]   (fused forward solve)

template <int BM_, int BN_>
constexpr size_t gemm_smem() {
  return (size_t)NSTAGE * BK * (BM_ + 4 + BN_ + 4) * sizeof(double) + 2 * NSTAGE * sizeof(unsigned long long);
}

// CTA shapes (rows x columns of the logical 128x128 tile handled by one CTA):
//   128x128  one CTA per SM, 32x64 warp tiles (only where nothing else fits);
//   128x64   the tile is split into two column halves, 32x32 warp tiles, <= 128 registers and
//            100 KB of shared memory, so TWO CTAs share an SM and one CTA's prologue/epilogue
//            (C tile read/write, pipeline fill) overlaps the other's DMMA main loop;
//    64x128  split into two row halves instead: each CTA reads and writes only its own rows,
//            which keeps an IN-PLACE product (potrf panel: C aliases A) race-free.
// blockIdx.x = tile * parts + part.
// LOADER = 0: every thread streams its share of the operand tiles with 16-byte cp.async
// (LDGSTS).  LOADER = 1: warp 0 issues TMA bulk copies (cp.async.bulk, one lane per tile
// column: 32 copies per stage) and the ring is synchronised with full/empty mbarriers, so the
// other 7 warps issue no load instructions at all.
// LOADER = 2: like 1, but a stage is TWO tensor-map TMA copies (one box per operand) issued by one
// elected lane, instead of 32 bulk copies issued by 32 lanes.
template <class Op, int BM_, int BN_, int LOADER>
__global__ void __launch_bounds__(GEMM_THREADS, ((BM_ * BN_ < BM * BN) ? 2 : 1))
gemm_nt_kernel(const __grid_constant__ Op op, const __grid_constant__ typename TmaParam<LOADER>::type tm) {
  extern __shared__ __align__(128) double gsm[];
  constexpr int NSN = BN / BN_, NSM = BM / BM_;       // column / row parts per logical tile
  constexpr int PA = BM_ + 4, PB = BN_ + 4;           // pitches: fragment loads hit 16 distinct banks
  constexpr int MW = BM_ / 32, NW = 8 / MW;           // warp grid
  constexpr int WN = BN_ / NW, NI = WN / 8;           // warp tile 32 x WN
  constexpr int MODE = Op::MODE;
  // default order: x = (tile, part), y = batch slot.  SLOT_MAJOR ops (tiles of unequal K) put
  // the slot in x and the tile in y, so the block scheduler hands out the longest tiles of
  // ALL matrices first (longest-processing-time order) instead of matrix after matrix.
  const int part = (int)(blockIdx.x % (NSN * NSM));
  const int hn = part % NSN, hm = part / NSN;
  const int bxq = (int)(blockIdx.x / (NSN * NSM));
  GemmTile t = Op::SLOT_MAJOR ? op.resolve((int)blockIdx.y, bxq) : op.resolve(bxq, (int)blockIdx.y);
  if (!t.valid) return;
  if (NSN > 1) {
    const long long off = (long long)hn * BN_;
    t.B += off;
    if (t.B0) t.B0 += off;
    if (t.C) t.C += off * t.ldc;
    if (t.Ct) t.Ct += off;
    if (t.E) t.E += off * t.lde;
    if (t.rowsum) t.rowsum += (long long)hn * t.rs_half;
  }
  if (NSM > 1) {
    const long long off = (long long)hm * BM_;
    t.A += off;
    if (t.A0) t.A0 += off;
    if (t.C) t.C += off;
    if (t.Ct) t.Ct += off * t.ldct;
    if (t.E) t.E += off;
    if (t.rowsum) t.rowsum += off;
    if (t.vdst) t.vdst += off;
  }
  // Padding: a CTA whose rows or columns are all padding has nothing to do; inside a CTA the
  // warps whose 32 x WN sub-tile is all padding skip their loads, DMMAs and stores (the padded
  // part of every operand is zero / identity, so the skipped results would be unchanged).
  const int nv = t.nvalid - hn * BN_;
  const int mv = t.mvalid - hm * BM_;
  if (!(MODE & GM_REDUCE) && (mv <= 0 || nv <= 0)) return;
  if (t.lower_only && (hm + 1) * BM_ <= hn * BN_) return;      // this part of a diagonal tile is all above the diagonal

  double* As = gsm;
  double* Bs = gsm + NSTAGE * BK * PA;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  // warp -> sub-tile: warps w and w+4 share a scheduler; the second four are permuted so that a scheduler's
  // two warps sit at opposite ends of the tile -- with a triangular operand one of them skips many K steps,
  // the other few, and every scheduler ends up with the same amount of DMMA work
  const int lw = (warp < 4) ? warp : ((MW == 2) ? (warp ^ 2) : 11 - warp);
  const int wm = (lw % MW) * 32, wn = (lw / MW) * WN;
  const int KT = t.K / BK;
  const bool wact = (wm < mv) && (wn < nv) && !(t.lower_only && hm * BM_ + wm + 32 <= hn * BN_ + wn);
  // K steps this warp needs [klo, khi) and this CTA loads [cklo, ckhi)
  int klo = 0, khi = KT, cklo = 0, ckhi = KT;
  if (t.tri_lo & 1) { klo = (hm * BM_ + wm) / BK; cklo = hm * BM_ / BK; }
  if (t.tri_lo & 2) { klo = max(klo, (hn * BN_ + wn) / BK); cklo = max(cklo, hn * BN_ / BK); }
  if (t.hi_m_off >= 0) {
    khi = min(khi, (t.hi_m_off + hm * BM_ + wm + 32) / BK);
    ckhi = min(ckhi, (t.hi_m_off + (hm + 1) * BM_) / BK);
  }
  if (t.hi_n_off >= 0) {
    khi = min(khi, (t.hi_n_off + hn * BN_ + wn + WN) / BK);
    ckhi = min(ckhi, (t.hi_n_off + (hn + 1) * BN_) / BK);
  }
  const int KT2 = max(ckhi - cklo, 0);                // K steps of this CTA; step `it` is k tile cklo + it

  auto load_stage = [&](int kt, int stage) {
    const int k0 = kt * BK;
    const bool alt = (k0 < T);
    const double* Ap = (alt && t.A0) ? t.A0 : t.A;
    const long long la = (alt && t.A0) ? t.lda0 : t.lda;
    const double* Bp = (alt && t.B0) ? t.B0 : t.B;
    const long long lb = (alt && t.B0) ? t.ldb0 : t.ldb;
#pragma unroll
    for (int i = 0; i < BM_ / 32; ++i) {
      const int c = tid + GEMM_THREADS * i;
      const int kk = c / (BM_ / 2), mc = (c % (BM_ / 2)) * 2;
      cp_async16(As + (stage * BK + kk) * PA + mc, Ap + (long long)(k0 + kk) * la + mc);
    }
#pragma unroll
    for (int i = 0; i < BN_ / 32; ++i) {
      const int c = tid + GEMM_THREADS * i;
      const int kk = c / (BN_ / 2), nc = (c % (BN_ / 2)) * 2;
      cp_async16(Bs + (stage * BK + kk) * PB + nc, Bp + (long long)(k0 + kk) * lb + nc);
    }
  };

  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(gsm + NSTAGE * BK * (PA + PB));
  unsigned long long* empty_bar = full_bar + NSTAGE;
  constexpr unsigned STAGE_BYTES = BK * (BM_ + BN_) * sizeof(double);
  // called by all 32 lanes of warp 0: lane l copies tile column (l & 15) of A (l < 16) or B
  auto bulk_stage = [&](int kt, int stage) {
    const int k0 = kt * BK;
    const bool alt = (k0 < T);
    const double* Ap = (alt && t.A0) ? t.A0 : t.A;
    const long long la = (alt && t.A0) ? t.lda0 : t.lda;
    const double* Bp = (alt && t.B0) ? t.B0 : t.B;
    const long long lb = (alt && t.B0) ? t.ldb0 : t.ldb;
    if (lane == 0) mbar_expect_tx(full_bar + stage, STAGE_BYTES);
    __syncwarp();
    const int kk = lane & (BK - 1);
    if (lane < BK)
      bulk_g2s(As + (stage * BK + kk) * PA, Ap + (long long)(k0 + kk) * la, BM_ * 8, full_bar + stage);
    else
      bulk_g2s(Bs + (stage * BK + kk) * PB, Bp + (long long)(k0 + kk) * lb, BN_ * 8, full_bar + stage);
  };
  // everything above is index arithmetic on launch parameters and host-written slot lists: from here
  // on the kernel reads what its predecessors in the stream wrote
  pdl_wait();
  pdl_launch();
  // LOADER 2: box coordinates of the operands in their arrays (once per CTA)
  int rA = 0, cA = 0, rB = 0, cB = 0, rA0 = 0, cA0 = 0, rB0 = 0, cB0 = 0;
  if constexpr (LOADER == 2) {
    if (warp == 0 && lane == 0) {
      long long o = t.A - tm.base[0];
      rA = (int)(o % tm.ld[0]); cA = (int)(o / tm.ld[0]);
      const int ib = t.b_alt ? 4 : 1;
      o = t.B - tm.base[ib];
      rB = (int)(o % tm.ld[ib]); cB = (int)(o / tm.ld[ib]);
      if (t.A0) { o = t.A0 - tm.base[2]; rA0 = (int)(o % tm.ld[2]); cA0 = (int)(o / tm.ld[2]); }
      if (t.B0) { o = t.B0 - tm.base[3]; rB0 = (int)(o % tm.ld[3]); cB0 = (int)(o / tm.ld[3]); }
    }
  }
  constexpr unsigned STAGE_BYTES_BOX = BK * (PA + PB) * sizeof(double);
  // called by warp 0; only lane 0 acts
  auto tma_stage = [&](int kt, int stage) {
    if constexpr (LOADER == 2) {
      if (lane != 0) return;
      const int k0 = kt * BK;
      const bool alt = (k0 < T);
      mbar_expect_tx(full_bar + stage, STAGE_BYTES_BOX);
      if (alt && t.A0) tma_box_2d(As + stage * BK * PA, &tm.a0, rA0, cA0 + k0, full_bar + stage);
      else tma_box_2d(As + stage * BK * PA, &tm.a, rA, cA + k0, full_bar + stage);
      if (alt && t.B0) tma_box_2d(Bs + stage * BK * PB, &tm.b0, rB0, cB0 + k0, full_bar + stage);
      else tma_box_2d(Bs + stage * BK * PB, t.b_alt ? &tm.b_alt : &tm.b, rB, cB + k0, full_bar + stage);
    }
  };
  if (LOADER >= 1) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(full_bar + s, 1);
        mbar_init(empty_bar + s, GEMM_THREADS / 32);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
      for (int s = 0; s < NSTAGE - 1; ++s)
        if (s < KT2) { if (LOADER == 2) tma_stage(cklo + s, s); else bulk_stage(cklo + s, s); }
    }
  } else {
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
      if (s < KT2) load_stage(cklo + s, s);
      cp_async_commit();
    }
  }

  // thread owns C(m = wm + mi*8 + g, n = wn + ni*8 + 2*tq + {0,1})
  double acc[4][NI][2];
  if ((MODE & GM_BETA) && wact) {
    // start from cscale * C: the loads go straight into the accumulator registers and overlap
    // the pipeline fill, instead of a latency-bound read-modify-write epilogue
    const double f = t.cscale;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int m = wm + mi * 8 + g, n = wn + ni * 8 + 2 * tq;
        acc[mi][ni][0] = f * t.C[m + (long long)n * t.ldc];
        acc[mi][ni][1] = f * t.C[m + (long long)(n + 1) * t.ldc];
      }
  } else {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  }

  for (int it = 0; it < KT2; ++it) {
    const int kt = cklo + it;
    if (LOADER >= 1) {
      const int nk = it + NSTAGE - 1;
      if (warp == 0 && nk < KT2) {
        const int ns = nk % NSTAGE;
        // the slot was last read in iteration it-1: wait until all 8 warps have released it
        if (nk >= NSTAGE) mbar_wait(empty_bar + ns, ((nk / NSTAGE) - 1) & 1);
        if (LOADER == 2) tma_stage(cklo + nk, ns); else bulk_stage(cklo + nk, ns);
      }
      mbar_wait(full_bar + (it % NSTAGE), (it / NSTAGE) & 1);
    } else {
      cp_async_wait<NSTAGE - 2>();
      __syncthreads();
      const int nk = it + NSTAGE - 1;
      if (nk < KT2) load_stage(cklo + nk, nk % NSTAGE);
      cp_async_commit();
    }
    const double* as = As + (it % NSTAGE) * BK * PA;
    const double* bs = Bs + (it % NSTAGE) * BK * PB;
    if (wact && kt >= klo && kt < khi) {
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        double a[4], b[NI];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = as[(k4 * 4 + tq) * PA + wm + mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) b[ni] = bs[(k4 * 4 + tq) * PB + wn + ni * 8 + g];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
    if (LOADER >= 1) {                       // this warp is done with the slot
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + (it % NSTAGE));
    }
  }
  if (LOADER == 0) cp_async_wait<0>();

  // ---- epilogue
  if (MODE & GM_REDUCE) {
    __syncthreads();            // all warps done with the ring; reuse it for the N-direction reduce
    double* red = gsm;          // [NW][BM_]
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int m = wm + mi * 8 + g;
      double s = 0.0;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = wn + ni * 8 + 2 * tq + j;
          const double c = t.alpha * acc[mi][ni][j];
          const double e = t.E ? t.E[m + (long long)n * t.lde] : c;
          s += c * e;
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (tq == 0) red[(lw / MW) * BM_ + m] = s;
    }
    __syncthreads();
    if (tid < BM_) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += red[w * BM_ + tid];
      t.rowsum[tid] = s;
    }
    return;
  }
  if (wact) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int m = wm + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const int n = wn + ni * 8 + 2 * tq;
        const double c0 = t.alpha * acc[mi][ni][0];
        const double c1 = t.alpha * acc[mi][ni][1];
        if (MODE & GM_STORE) {
          t.C[m + (long long)n * t.ldc] = c0;
          t.C[m + (long long)(n + 1) * t.ldc] = c1;
        }
        if (MODE & GM_STORET) {
          if (t.Ct) {
            double2 v = make_double2(c0, c1);
            *reinterpret_cast<double2*>(t.Ct + n + (long long)m * t.ldct) = v;
          }
        }
      }
    }
  }
  if (MODE & GM_ZSOLVE) {
    // z_k = D_k b_k, recomputed by every CTA of the step (D_k is L2-resident after the first):
    // one launch per block column instead of a solve launch and an update launch.  The lower
    // triangle of D_k is staged in the (unused: K = 0) operand ring with all loads in flight at
    // once -- columns 0..63 with all rows, columns 64..127 with rows 64..127 -- and summed in the
    // order of zsolve_row (element c into accumulator c mod 4, increasing c): bit-identical.
    double* zs = gsm + 256;                    // clear of the reduction scratch gsm[0..256)
    double* zbs = gsm + 384;
    double* D1 = gsm + 512;                    // [64][128]
    double* D2 = D1 + 64 * T;                  // [64][64]
    for (int e = tid; e < 64 * 64; e += GEMM_THREADS) {
      const int col = e >> 6, r2 = (e & 63) * 2;
      if (r2 + 1 >= col) cp_async16(D1 + col * T + r2, t.zD + col * T + r2);
    }
    for (int e = tid; e < 64 * 32; e += GEMM_THREADS) {
      const int col = e >> 5, r2 = (e & 31) * 2;
      if (r2 + 1 >= col) cp_async16(D2 + col * 64 + r2, t.zD + (64 + col) * T + 64 + r2);
    }
    cp_async_commit();
    if (tid < T) zbs[tid] = t.zb[tid];
    cp_async_wait<0>();
    __syncthreads();
    if (tid < T) {
      const int rr = tid;
      double s4[4] = {0.0, 0.0, 0.0, 0.0};
      if (rr < t.znact) {
        const int c1 = min(rr, 63);
        for (int c0 = 0; c0 <= c1; c0 += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c0 + u <= c1) s4[u] = __fma_rn(D1[(c0 + u) * T + rr], zbs[c0 + u], s4[u]);
        }
        for (int c0 = 64; c0 <= rr; c0 += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c0 + u <= rr) s4[u] = __fma_rn(D2[(c0 + u - 64) * 64 + rr - 64], zbs[c0 + u], s4[u]);
        }
      }
      const double z = __dadd_rn(__dadd_rn(s4[0], s4[1]), __dadd_rn(s4[2], s4[3]));
      zs[tid] = z;
      if (t.zout) t.zout[tid] = z;
    }
    __syncthreads();
    t.vdot = zs;
  }
  if (MODE & GM_ROWDOT) {
    // fused forward substitution: vdst[m] -= sum_n C(m,n) * vdot[n]  (needs all columns of the
    // tile in this CTA: BN_ == BN).  The order of the sum does not depend on the CTA shape: the 128
    // columns are taken in eight groups of 16 (two 8-column DMMA blocks: FMA chain over the thread's
    // four values, then the xor tree over the four threads of a row), and the groups are added in
    // increasing order -- so a 32x128, 64x128 or 128x128 CTA gives the same bits.
    if (t.vdot == nullptr) return;
    __syncthreads();
    double* red = (MODE & GM_ZSOLVE) ? gsm + 512 : gsm;    // [8][BM_]; after the z solve D_k's staging area is free
    static_assert(NI % 2 == 0, "row-dot epilogue works on groups of two 8-column blocks");
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int m = wm + mi * 8 + g;
#pragma unroll
      for (int gq = 0; gq < NI / 2; ++gq) {
        double s = 0.0;
        if (wact) {
#pragma unroll
          for (int ni = 2 * gq; ni < 2 * gq + 2; ++ni) {
            const int n = wn + ni * 8 + 2 * tq;
            // explicit FMAs: every instantiation (fused panel, solve-only replay) rounds identically
            s = __fma_rn(t.alpha * acc[mi][ni][0], t.vdot[n], s);
            s = __fma_rn(t.alpha * acc[mi][ni][1], t.vdot[n + 1], s);
          }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tq == 0) red[((wn >> 4) + gq) * BM_ + m] = s;
      }
    }
    __syncthreads();
    if (tid < BM_ && tid < mv) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < BN / 16; ++w) s += red[w * BM_ + tid];
      t.vdst[tid] -= s;
    }
  }
}

// ------------------------------------------------------------------------------------
// Tile -> problem maps.  All matrices are column-major, padded to Np = Nt*T with an
// identity block, one Np x Np buffer per batch slot (stride Np*Np); `sel` lists the
// slots this launch works on (blockIdx.y).
// ------------------------------------------------------------------------------------
struct BatchBufs {
  double* Abuf;     // A -> L (lower) ; later H^T (upper) ; later K^-1 (lower)
  double* Wbuf;     // W = L^-1 (lower), W^T (upper, gradient path)
  double* Dbuf;     // [slot][Nt][T*T]  D_k = inv(L_kk), column-major, zeros above the diagonal
  double* DTbuf;    // [slot][Nt][T*T]  D_k^T
  const int* sel;
  long long smat;   // Np*Np
  int Np, Nt;
  int N;            // rows that hold data; N16 = N rounded up to the K step
  __device__ int n16() const { return (N + BK - 1) / BK * BK; }
};

__device__ __forceinline__ GemmTile empty_tile() {
  GemmTile t;
  t.A = t.B = t.A0 = t.B0 = nullptr;
  t.lda = t.ldb = t.lda0 = t.ldb0 = 0;
  t.C = t.Ct = nullptr;
  t.ldc = t.ldct = 0;
  t.E = nullptr;
  t.lde = 0;
  t.rowsum = nullptr;
  t.rs_half = 0;
  t.vdot = nullptr;
  t.vdst = nullptr;
  t.zD = t.zb = nullptr;
  t.zout = nullptr;
  t.znact = 0;
  t.K = 0;
  t.b_alt = 0;
  t.mvalid = BM;
  t.nvalid = BN;
  t.tri_lo = 0;
  t.hi_m_off = t.hi_n_off = -1;
  t.lower_only = 0;
  t.alpha = 1.0;
  t.cscale = 0.0;
  t.valid = true;
  return t;
}

// debug / benchmark: plain C = alpha A B^T + beta C
struct OpGeneric {
  static constexpr int TMA_A = -1, TMA_B = -1, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_BETA | GM_STORE;
  const double* A; const double* B; double* C;
  long long lda, ldb, ldc;
  int K;
  double alpha, cscale;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    t.A = A + (long long)bx * BM; t.lda = lda;
    t.B = B + (long long)by * BN; t.ldb = ldb;
    t.C = C + (long long)bx * BM + (long long)by * BN * ldc; t.ldc = ldc;
    t.K = K; t.alpha = alpha; t.cscale = cscale;
    return t;
  }
};

// plain product C = A B^T (C is not read).  tri = 1: B is LOWER TRIANGULAR in tile storage whose upper
// tiles hold something else (Wbuf keeps W^T there): column tile `by` of the product only runs over
// k < (by+1)*T, the tiles on and below B's diagonal.
struct OpPlain {
  static constexpr int TMA_A = -1, TMA_B = -1, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_STORE;
  const double* A; const double* B; double* C;
  long long lda, ldb, ldc;
  int K;
  int tri = 0;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    t.A = A + (long long)bx * BM; t.lda = lda;
    t.B = B + (long long)by * BN; t.ldb = ldb;
    t.C = C + (long long)bx * BM + (long long)by * BN * ldc; t.ldc = ldc;
    t.K = tri ? min((by + 1) * BN, K) : K;
    return t;
  }
};

// potrf panel, step k:  L_ik = A_ik * D_k^T  (i > k), in place; with a right-hand side the
// forward-substitution update  b_i -= L_ik z_k  rides in the epilogue
struct OpPanel {
  static constexpr int TMA_A = 0, TMA_B = 2, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_STORE | GM_ROWDOT;
  BatchBufs b; int k;
  const double* zvec; double* bvec;      // [nslots][Np] or null
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int i = k + 1 + bx;
    double* tile = b.Abuf + slot * b.smat + (long long)i * T + (long long)k * T * b.Np;
    t.A = tile; t.lda = b.Np;
    t.B = b.Dbuf + ((long long)slot * b.Nt + k) * T * T; t.ldb = T;
    t.C = tile; t.ldc = b.Np;
    t.K = T;
    t.hi_n_off = 0;                        // D_k(n, q) = 0 for q > n
    t.mvalid = b.N - i * T;
    if (zvec) {
      t.vdot = zvec + (long long)slot * b.Np + (long long)k * T;
      t.vdst = bvec + (long long)slot * b.Np + (long long)i * T;
    }
    return t;
  }
};

// forward substitution only, on an existing factor:  b_i -= L_ik z_k  through the SAME epilogue
// as the fused panel (K = 0: the accumulators are just the stored L_ik), so replaying the solve
// for a new right-hand side is bit-identical to a full evaluation
struct OpFwd {
  static constexpr int TMA_A = -1, TMA_B = -1, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_BETA | GM_ROWDOT;
  BatchBufs b; int k;
  const double* zvec; double* bvec;
  const int* fsel;                       // slot whose factor entry `by` uses (several rows may share one)
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int fslot = fsel ? fsel[by] : slot;
    const int i = k + 1 + bx;
    double* tile = b.Abuf + fslot * b.smat + (long long)i * T + (long long)k * T * b.Np;
    t.A = tile; t.lda = b.Np;
    t.B = tile; t.ldb = b.Np;
    t.C = tile; t.ldc = b.Np;
    t.K = 0; t.cscale = 1.0;
    t.mvalid = b.N - i * T;
    t.vdot = zvec + (long long)slot * b.Np + (long long)k * T;
    t.vdst = bvec + (long long)slot * b.Np + (long long)i * T;
    return t;
  }
};

// solve-only replay, one launch per block column: like OpFwd, but every CTA first computes
// z_k = D_k b_k itself (the CTAs of the first tile row also store it)
struct OpFwdZ {
  static constexpr int TMA_A = -1, TMA_B = -1, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_BETA | GM_ROWDOT | GM_ZSOLVE;
  BatchBufs b; int k;
  double* zvec; double* bvec;
  const int* fsel;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int fslot = fsel ? fsel[by] : slot;
    const int i = k + 1 + bx;
    double* tile = b.Abuf + fslot * b.smat + (long long)i * T + (long long)k * T * b.Np;
    t.A = tile; t.lda = b.Np;
    t.B = tile; t.ldb = b.Np;
    t.C = tile; t.ldc = b.Np;
    t.K = 0; t.cscale = 1.0;
    t.mvalid = b.N - i * T;
    t.zD = b.Dbuf + ((long long)fslot * b.Nt + k) * T * T;
    t.zb = bvec + (long long)slot * b.Np + (long long)k * T;
    t.zout = (bx == 0) ? zvec + (long long)slot * b.Np + (long long)k * T : nullptr;
    t.znact = min(T, b.N - k * T);
    t.vdst = bvec + (long long)slot * b.Np + (long long)i * T;
    return t;
  }
};

// potrf trailing update:  A_ij -= sum_{k in [k0, k0+kw)} L_ik L_jk^T  for tile columns
// j in [jlo, jhi), rows i >= j.  Two-level blocking: inside an outer block of `kw` tile columns
// the update is restricted to that block's columns (K = 128 per step); once the outer block is
// factored, ONE update with K = kw*128 brings the rest of the matrix up to date, so most of the
// flops run with a long K loop and each C tile is read and written once per outer block.
struct OpSyrk {
  static constexpr int TMA_A = 0, TMA_B = 0, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_BETA | GM_STORE;
  BatchBufs b; int k0, kw, jlo, jhi;
  int bx0 = 0;     // first tile of this launch (a long update can be issued in chunks)
  // tile enumeration: column-major over the trapezoid {(i,j): jlo <= j < jhi, j <= i < Nt}
  __host__ __device__ static int count(int Nt, int jlo, int jhi) {
    const int nc = jhi - jlo;
    return nc * (Nt - jlo) - nc * (nc - 1) / 2;
  }
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    int j = jlo, rem = bx + bx0;
    while (rem >= b.Nt - j) { rem -= b.Nt - j; ++j; }      // at most jhi-jlo steps
    const int i = j + rem;
    double* base = b.Abuf + slot * b.smat;
    t.A = base + (long long)i * T + (long long)k0 * T * b.Np; t.lda = b.Np;
    t.B = base + (long long)j * T + (long long)k0 * T * b.Np; t.ldb = b.Np;
    t.C = base + (long long)i * T + (long long)j * T * b.Np; t.ldc = b.Np;
    t.K = kw * T; t.alpha = -1.0; t.cscale = -1.0;
    t.mvalid = b.N - i * T;
    t.nvalid = b.N - j * T;
    t.lower_only = (i == j);
    return t;
  }
};

// H pass:  upper tile (j,i) <- (L_ij * D_j)^T  for every i > j
struct OpHpass {
  static constexpr int TMA_A = 0, TMA_B = 3, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_STORET;
  BatchBufs b;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    int a, c;
    tri_decode(bx, a, c);
    const int i = a + 1, j = c;
    double* base = b.Abuf + slot * b.smat;
    t.A = base + (long long)i * T + (long long)j * T * b.Np; t.lda = b.Np;
    t.B = b.DTbuf + ((long long)slot * b.Nt + j) * T * T; t.ldb = T;
    t.Ct = base + (long long)j * T + (long long)i * T * b.Np; t.ldct = b.Np;
    t.K = T;
    return t;
  }
};

// triangular inverse, block column j (descending):  W_ij = - sum_{k=j+1..i} W_ik H_kj
struct OpWrec {
  static constexpr int TMA_A = 1, TMA_B = 0, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = true;
  static constexpr int MODE = GM_STORE | GM_STORET;
  BatchBufs b; int j; int dual;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int i = b.Nt - 1 - bx;            // bx = 0 is the longest product (K = (i-j)*T)
    double* W = b.Wbuf + slot * b.smat;
    const double* H = b.Abuf + slot * b.smat;
    t.A = W + (long long)i * T + (long long)(j + 1) * T * b.Np; t.lda = b.Np;
    t.B = H + (long long)j * T + (long long)(j + 1) * T * b.Np; t.ldb = b.Np;
    t.C = W + (long long)i * T + (long long)j * T * b.Np; t.ldc = b.Np;
    if (dual) { t.Ct = W + (long long)j * T + (long long)i * T * b.Np; t.ldct = b.Np; }
    t.K = min((i - j) * T, b.n16() - (j + 1) * T); t.alpha = -1.0;
    t.mvalid = b.N - i * T;
    return t;
  }
};

// Triangular inverse by recursive halving, bottom-up (block size s = 1, 2, 4, ... tiles).  A pair
// of adjacent diagonal blocks [a, mid) and [mid, end) whose inverses are known is merged:
//     W_21 = -W_22 (L_21 W_11)
// in two passes, every pair of a level in the same launch -- 2*ceil(log2 Nt) launches in all,
// with Nt^2/4 tiles in the top ones, where the column recurrence (OpWrec) needs Nt-1 launches whose
// longest tile has K = Nt*T: the better shape for a few matrices.
//   pass 1 (OpRecX): X = L_21 W_11, stored transposed in the (free) upper triangle of Abuf
//   pass 2 (OpRecW): W_21 = -W_22 X into the lower tiles of Wbuf and transposed into the upper
// Tile order: longest K first; bx = (kidx * P + pair) * s + other.
struct OpRecX {
  static constexpr int TMA_A = 0, TMA_B = 1, TMA_A0 = -1, TMA_B0 = 3;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = true;
  static constexpr int MODE = GM_STORET;
  BatchBufs b; int s, P;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int il = bx % s, pr = (bx / s) % P, jl = bx / (s * P);
    const int a = 2 * s * pr, mid = a + s;
    const int i = mid + il, j = a + jl;                  // X(i, j) = sum_{k=j}^{mid-1} L(i,k) W(k,j)
    if (i >= b.Nt || i * T >= b.N) { t.valid = false; return t; }
    double* A = b.Abuf + slot * b.smat;
    const double* W = b.Wbuf + slot * b.smat;
    t.A = A + (long long)i * T + (long long)j * T * b.Np; t.lda = b.Np;
    t.B = W + (long long)j * T + (long long)j * T * b.Np; t.ldb = b.Np;     // W^T: upper tiles
    t.B0 = b.DTbuf + ((long long)slot * b.Nt + j) * T * T; t.ldb0 = T;      // W(j,j)^T = D_j^T
    t.Ct = A + (long long)j * T + (long long)i * T * b.Np; t.ldct = b.Np;
    t.K = min((mid - j) * T, b.n16() - j * T);
    t.tri_lo = 2;                          // first K tile: D_j^T(n, q) = 0 for q < n
    return t;
  }
};

struct OpRecW {
  static constexpr int TMA_A = 1, TMA_B = 0, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = true;
  static constexpr int MODE = GM_STORE | GM_STORET;
  BatchBufs b; int s, P;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    const int jl = bx % s, pr = (bx / s) % P, il = s - 1 - bx / (s * P);
    const int a = 2 * s * pr, mid = a + s;
    const int i = mid + il, j = a + jl;                  // W(i, j) = -sum_{k=mid}^{i} W(i,k) X(k,j)
    if (i >= b.Nt || i * T >= b.N) { t.valid = false; return t; }
    double* W = b.Wbuf + slot * b.smat;
    const double* A = b.Abuf + slot * b.smat;
    t.A = W + (long long)i * T + (long long)mid * T * b.Np; t.lda = b.Np;   // last K tile: W(i,i) = D_i
    t.B = A + (long long)j * T + (long long)mid * T * b.Np; t.ldb = b.Np;   // X^T
    t.C = W + (long long)i * T + (long long)j * T * b.Np; t.ldc = b.Np;
    t.Ct = W + (long long)j * T + (long long)i * T * b.Np; t.ldct = b.Np;
    t.K = min((i - mid + 1) * T, b.n16() - mid * T); t.alpha = -1.0;
    t.hi_m_off = (i - mid) * T;            // last K tile: W(i,i)(m, q) = D_i(m, q) = 0 for q > m
    return t;
  }
};

// K^-1 = W^T W (lower tiles a >= c) written over the lower triangle of Abuf
struct OpSyrk2 {
  static constexpr int TMA_A = 1, TMA_B = 1, TMA_A0 = 3, TMA_B0 = 3;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = true;
  static constexpr int MODE = GM_STORE;
  BatchBufs b;
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int slot = b.sel[by];
    int a, c;
    tri_decode(bx, a, c);
    const double* W = b.Wbuf + slot * b.smat;
    const double* DTa = b.DTbuf + ((long long)slot * b.Nt + a) * T * T;
    // k runs over rows q = a*T .. Np-1 of W; the first T of them are the diagonal tile
    t.A = W + (long long)a * T + (long long)a * T * b.Np; t.lda = b.Np;
    t.B = W + (long long)c * T + (long long)a * T * b.Np; t.ldb = b.Np;
    t.A0 = DTa; t.lda0 = T;
    if (a == c) { t.B0 = DTa; t.ldb0 = T; }
    t.lower_only = (a == c);
    t.tri_lo = (a == c) ? 3 : 1;           // first K tile: D_a^T(m, q) = 0 for q < m (and for q < n on the diagonal)
    t.C = b.Abuf + slot * b.smat + (long long)a * T + (long long)c * T * b.Np; t.ldc = b.Np;
    t.K = min((b.Nt - a) * T, b.n16() - a * T);
    t.mvalid = b.N - a * T;
    t.nvalid = b.N - c * T;
    return t;
  }
};

// predictive variance:  part[nt][j] = sum_{m in tile nt} ( sum_k Bt(j,k) Wm(m,k) )^2   (L_chol)
//                    or sum_{m in tile nt} ( sum_k Bt(j,k) X(m,k) ) * Bt(j,m)            (low noise)
struct OpPred {
  static constexpr int TMA_A = -1, TMA_B = -1, TMA_A0 = -1, TMA_B0 = -1;   // operand arrays: 0 Abuf, 1 Wbuf, 2 Dbuf, 3 DTbuf
  static constexpr bool SLOT_MAJOR = false;
  static constexpr int MODE = GM_REDUCE;
  // blockIdx.z = posterior sample within the group handled by this launch
  const double* Bt; long long ldbt, sBt;   // per sample (Mcp x Np) column-major, stride sBt
  const double* Wbuf; const double* Abuf;  // W = L^-1 (L_chol samples) / Ainv (low noise), stride smat
  long long smat;
  const SlotP* sp;                         // per-sample scalars: sp[z].lchol picks the operand
  double* part; long long spart;           // per sample [Nt*ns][Mcp]
  int Mcp, Np, ns;                         // ns = column halves per tile (BN / BN_)
  int N, mc;                               // training rows / test points that hold data
  __device__ GemmTile resolve(int bx, int by) const {
    GemmTile t = empty_tile();
    const int z = blockIdx.z;
    const int tri = sp[z].lchol;             // W lower triangular -> K = (nt+1)*T
    const double* bt = Bt + z * sBt;
    const int jt = bx, nt = (int)gridDim.y - 1 - by;   // longest K first
    t.A = bt + (long long)jt * BM; t.lda = ldbt;
    t.B = (tri ? Wbuf : Abuf) + z * smat + (long long)nt * BN; t.ldb = Np;
    t.b_alt = tri ? 0 : 1;
    const int n16 = (N + BK - 1) / BK * BK;
    t.K = tri ? min((nt + 1) * T, n16) : n16;
    if (tri) t.hi_n_off = nt * T;            // last K tile: the diagonal tile of W, zero for k > m
    t.mvalid = mc - jt * BM;
    t.nvalid = N - nt * BN;
    if (!tri) { t.E = bt + (long long)jt * BM + (long long)nt * BN * ldbt; t.lde = ldbt; }
    t.rowsum = part + z * spart + (long long)nt * ns * Mcp + (long long)jt * BM;
    t.rs_half = Mcp;
    return t;
  }
};

}  // namespace gpb
