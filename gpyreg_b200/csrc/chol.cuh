// Panel-level kernels of the batched blocked Cholesky / triangular solves.
//   diag_kernel       factor one 128x128 diagonal tile in shared memory, invert it, write
//                     D_k, D_k^T, log-det partial, and the forward-solve block z_k = D_k b_k
//                     (factor_block32: the 32x32 diagonal blocks in the registers of one warp)
//   diag_solve_kernel z_k = D_k b_k on an existing factor (solve-only replay of the last block;
//                     the other blocks are fused into the tile GEMM launch, OpFwdZ)
//   bwd_step_kernel   w_i = D_i^T b_i ; b_j -= L_ij^T w_i (backward substitution, j < i; posterior path)
//   alpha_gemv_kernel alpha = W^T z / sl once W = L^-1 exists (gradient path)
//   nlz_kernel        nlZ = z^T z/(2 sl) + sum log L_ii + N log(2 pi sl)/2
// The dense O(N^3) work between these is the tile GEMM in gemm.cuh.
#pragma once
#include <utility>

#include "common.cuh"

namespace gpb {

constexpr int DP_PITCH = T + 4;                                   // 132: 16-byte rows, conflict-free DMMA fragments
constexpr int SB = 32;                                            // sub-block of the tile factorisation
constexpr int IVP = 36;                                           // pitch of the 32x32 scratch blocks
constexpr size_t DIAG_SMEM =
    ((size_t)T * DP_PITCH + 8 * SB * IVP + 2 * T) * sizeof(double);

struct DiagArgs {
  double* Abuf; double* Wbuf;      // Wbuf may be null (nlZ-only path)
  double* Dbuf; double* DTbuf;
  const int* sel;
  long long smat;
  int Np, Nt, N, k;
  double* bvec;                    // [nslots][Np] running right-hand side (forward solve)
  double* zvec;                    // [nslots][Np] z = L^-1 r
  double* logdet;                  // [nslots][Nt] partial sums of log L_ii
  int* fail;                       // [nslots]  set to 1 on a pivot <= 0 or NaN
  long long* dbg;                  // optional: clock64() stamps of the phases (thread 0 of CTA 0)
};

__device__ __forceinline__ void dmma_t(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}


// --- 32x32 diagonal block in registers, one row per lane; every register index is a
// compile-time constant (the steps are instantiated through a fold expression).
//  * The lane's own diagonal entry lives in a separate register `dg`: its update
//    dg -= L(lane,J)^2 needs no other lane, so the pivot -> rsqrt -> scale -> next pivot chain is one
//    shuffle, one MUFU and five dependent FP64 operations per step.
//  * rsqrt is written out (MUFU seed, relative error 2^-22.4, then one third-order step -- the
//    arithmetic of the library's main path) so that the scaling of the column is folded into its
//    last FMA: L(lane,J) = a*y0 + (a*y0*e)*p.
//  * The finished column is broadcast through shared memory (one 8-byte store per lane, 16-byte
//    loads): a 64-bit warp shuffle per element costs two slots of the slow shuffle pipe, which
//    was what bounded this loop.
//  * A pivot <= 0 or NaN only raises the flag (LAPACK dpotrf: info > 0): the NaNs it produces
//    flow through harmlessly and the host retries that matrix with more jitter.
template <int J>
__device__ __forceinline__ void chol32_step(double (&arow)[SB], double& dg, double& rdiag, double& piv,
                                            int lane, int& failed, double* colbuf) {
  if (!(piv > 0.0)) failed = 1;
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(piv));
  const double a = (lane == J) ? dg : arow[J];        // the diagonal lane scales the pivot itself
  const double ay0 = a * y0;
  const double e = fma(-(y0 * y0), piv, 1.0);
  const double p = fma(e, 0.375, 0.5);
  const double lj = fma(p, ay0 * e, ay0);             // a / sqrt(piv)
  // selects, no branch: a divergent `if (lane == J)` region costs a reconvergence per step
  const double yfull = fma(p, y0 * e, y0);            // 1 / sqrt(piv)
  rdiag = (lane == J) ? yfull : rdiag;                // 1 / L(J, J) on the diagonal lane
  const double dnew = fma(-lj, lj, dg);
  dg = (lane == J) ? lj : ((lane > J) ? dnew : dg);   // L(J, J) | updated diagonal | final
  arow[J] = lj;                                       // lanes <= J: unused garbage
  if (J + 1 < SB) {
    piv = __shfl_sync(0xffffffffu, dnew, J + 1);      // next pivot first (lane J+1 > J holds dnew): it heads the chain
    double* cb = colbuf + (J & 1) * SB;               // double-buffered: no WAR hazard across steps
    cb[lane] = lj;                                    // L(lane, J), valid for lanes > J
    __syncwarp();
    constexpr int CS = (J + 1) & ~1;
#pragma unroll
    // unpredicated: entries on and above the diagonal of the register block (lane <= c) collect
    // garbage that nothing reads -- the diagonal lives in dg, the write-back and the inverse use the
    // strictly lower part only.  The loop is issue-bound: one FMA per column instead of
    // compare + FMA + two selects
    for (int c = CS; c < SB; c += 2) {
      const double2 v = *reinterpret_cast<const double2*>(cb + c);
      if (c > J) arow[c] = fma(-lj, v.x, arow[c]);
      arow[c + 1] = fma(-lj, v.y, arow[c + 1]);
    }
  }
}
template <int... Js>
__device__ __forceinline__ void chol32_all(double (&arow)[SB], double& dg, double& rdiag, double& piv,
                                           int lane, int& failed, double* colbuf,
                                           std::integer_sequence<int, Js...>) {
  (chol32_step<Js>(arow, dg, rdiag, piv, lane, failed, colbuf), ...);
}

// Inverse of the 32x32 factor, both 16x16 diagonal halves at once in one warp: lane (h, c), h = lane & 16,
// c = lane & 15, owns column c of the inverse of the half-block L_hh: inv[q] = Inv_hh(q, c).
// Right-looking: once x_Q = inv[Q] is final it is pushed into the partial sums of the rows below,
// then row Q+1 is finished -- the chain from x_Q to x_{Q+1} is one FMA and one multiply.  The factor
// is read back from shared memory (it was just stored there), two rows per 16-byte load.  The
// off-diagonal block Inv_21 = -Inv_22 L_21 Inv_11 is two 16x16x16 products on the tensor pipe.
constexpr int HB = SB / 2;
template <int Q>
__device__ __forceinline__ void inv16_step(const double* Lcol, const double* rd, double (&inv)[HB],
                                           int lane) {
  if (Q >= HB - 1) return;
  // Lcol -> L(h, h) of this lane's half; column h+Q of the factor, rows h+Q+1 .. h+15
  const double* col = Lcol + Q * DP_PITCH;
  const int cc = lane & (HB - 1);
  constexpr int RS = (Q + 1) & ~1;
#pragma unroll
  for (int r = RS; r < HB; r += 2) {
    const double2 v = *reinterpret_cast<const double2*>(col + r);     // L(h+r, h+Q), L(h+r+1, h+Q)
    // unpredicated: for rows r <= cc the term is L(r,Q) * x_Q with Q < cc, and x_Q = 0 there
    if (r > Q) inv[r] = fma(v.x, inv[Q], inv[r]);
    inv[r + 1] = fma(v.y, inv[Q], inv[r + 1]);
  }
  const double dn = rd[Q + 1];                                        // 1 / L(h+Q+1, h+Q+1)
  if (cc < Q + 1) inv[Q + 1] = -inv[Q + 1] * dn;
}
template <int... Qs>
__device__ __forceinline__ void inv16_all(const double* Lcol, const double* rd, double (&inv)[HB],
                                          int lane, std::integer_sequence<int, Qs...>) {
  (inv16_step<Qs>(Lcol, rd, inv, lane), ...);
}

// Factor the 32x32 diagonal block at (c0, c0) of the tile S (one warp): Cholesky in registers,
// factor back to S, inverse of the factor to Ivp.  Kept out of line so that its register
// allocation and instruction schedule do not depend on the rest of diag_kernel.
__device__ __noinline__ int factor_block32(double* S, double* Ivp, double* Tm, int c0, int lane, long long* tdbg) {
  // tdbg (GPB_DIAG_DBG only): clock stamps after the pivot loop and after the 16x16 inverses
#define FSTAMP(i) do { if (tdbg) { if (lane == 0) tdbg[i] = clock64(); __syncwarp(); } } while (0)
  const int g = lane >> 2, tq = lane & 3;
  double* colbuf = Tm;               // [2][32] column broadcast buffers
  double* rd = Tm + 2 * SB;          // [32] 1 / L_rr
  int failed = 0;
  double rdiag = 0.0;                // 1 / L_rr of the lane's own row
  double arow[SB];                   // strictly lower part of the lane's row; the diagonal is dg
#pragma unroll
  for (int c = 0; c < SB; ++c) arow[c] = (c < lane) ? S[(c0 + c) * DP_PITCH + c0 + lane] : 0.0;
  double dg = S[(c0 + lane) * DP_PITCH + c0 + lane];
  double piv = __shfl_sync(0xffffffffu, dg, 0);
  chol32_all(arow, dg, rdiag, piv, lane, failed, colbuf, std::make_integer_sequence<int, SB>{});
#pragma unroll
  for (int c = 0; c < SB; ++c)
    if (c < lane) S[(c0 + c) * DP_PITCH + c0 + lane] = arow[c];
  S[(c0 + lane) * DP_PITCH + c0 + lane] = dg;
  rd[lane] = rdiag;
  __syncwarp();
  FSTAMP(0);
  // inverses of the two 16x16 diagonal halves
  const int h = lane & HB;
  double inv[HB];
#pragma unroll
  for (int q = 0; q < HB; ++q) inv[q] = (q == (lane & (HB - 1))) ? rdiag : 0.0;
  inv16_all(S + (c0 + h) * DP_PITCH + c0 + h, rd + h, inv, lane, std::make_integer_sequence<int, HB>{});
#pragma unroll
  for (int q = 0; q < HB; ++q) Ivp[lane * IVP + h + q] = inv[q];     // column `lane`, rows h..h+15
  __syncwarp();
  FSTAMP(1);
#undef FSTAMP
  // Inv_21 = -Inv_22 (L_21 Inv_11): four 8x8 output blocks, K = 16, on DMMA.  k is the OUTER loop: the DMMAs are
  // volatile asm statements and keep their program order, so with the output block outermost the four dependent
  // chains of a stage ran one after the other (16 x the DMMA latency: 1.5 k cycles for the two stages)
  double t[2][2][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) t[mi][ni][0] = t[mi][ni][1] = 0.0;
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    double af[2], bf[2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) af[mi] = S[(c0 + k4 * 4 + tq) * DP_PITCH + c0 + HB + mi * 8 + g];   // L_21(m, k)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) bf[ni] = Ivp[(ni * 8 + g) * IVP + k4 * 4 + tq];                     // Inv_11(k, n)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma_t(t[mi][ni][0], t[mi][ni][1], af[mi], bf[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      Tm[(ni * 8 + 2 * tq) * IVP + mi * 8 + g] = t[mi][ni][0];
      Tm[(ni * 8 + 2 * tq + 1) * IVP + mi * 8 + g] = t[mi][ni][1];
    }
  __syncwarp();
  double d[2][2][2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) d[mi][ni][0] = d[mi][ni][1] = 0.0;
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    double af[2], bf[2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) af[mi] = -Ivp[(HB + k4 * 4 + tq) * IVP + HB + mi * 8 + g];          // -Inv_22(m, k)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) bf[ni] = Tm[(ni * 8 + g) * IVP + k4 * 4 + tq];                      // T(k, n)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma_t(d[mi][ni][0], d[mi][ni][1], af[mi], bf[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      Ivp[(ni * 8 + 2 * tq) * IVP + HB + mi * 8 + g] = d[mi][ni][0];
      Ivp[(ni * 8 + 2 * tq + 1) * IVP + HB + mi * 8 + g] = d[mi][ni][1];
    }
  return failed;
}

// Out of line (one copy each): the kernel's code must stay inside the SM's instruction cache -- with
// these bodies inlined at every call site the block factorisation's unrolled code was evicted between
// launches and its first call ran five times slower.
// rows m0..m0+7 below block c0: L21 = P * Inv^T   (C(m,n) = sum_k P(m,k) Inv(n,k))
__device__ __noinline__ void diag_panel_rows(double* S, const double* Ivp, int c0, int mb, int g, int tq) {
  const int m0 = c0 + SB + mb * 8;
  double af[8];
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) af[k4] = S[(c0 + k4 * 4 + tq) * DP_PITCH + m0 + g];
  double acc[4][2];
#pragma unroll
  for (int n8 = 0; n8 < 4; ++n8) acc[n8][0] = acc[n8][1] = 0.0;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    double bf[4];
#pragma unroll
    for (int n8 = 0; n8 < 4; ++n8) bf[n8] = Ivp[(k4 * 4 + tq) * IVP + n8 * 8 + g];
#pragma unroll
    for (int n8 = 0; n8 < 4; ++n8) dmma_t(acc[n8][0], acc[n8][1], af[k4], bf[n8]);   // 4 chains
  }
#pragma unroll
  for (int n8 = 0; n8 < 4; ++n8) {
    S[(c0 + n8 * 8 + 2 * tq) * DP_PITCH + m0 + g] = acc[n8][0];
    S[(c0 + n8 * 8 + 2 * tq + 1) * DP_PITCH + m0 + g] = acc[n8][1];
  }
}

// S -= L21 L21^T on one 16x16 super-block (2x2 DMMA blocks) of the trailing lower triangle: two A and
// two B fragments per k-step feed four DMMAs
__device__ __noinline__ void diag_update_super(double* S, int c0, int mblocks, int idx, int g, int tq) {
  int si, sj;
  tri_decode(idx, si, sj);
  const int r0 = c0 + SB + si * 16, q0 = c0 + SB + sj * 16;
  const bool okr = 2 * si + 1 < mblocks, okq = 2 * sj + 1 < mblocks;   // second half inside the tile
  const bool diag = si == sj;                                          // block (0,1) is above the diagonal
  double x00[2], x01[2], x10[2], x11[2];
  x00[0] = S[(q0 + 2 * tq) * DP_PITCH + r0 + g];
  x00[1] = S[(q0 + 2 * tq + 1) * DP_PITCH + r0 + g];
  x10[0] = okr ? S[(q0 + 2 * tq) * DP_PITCH + r0 + 8 + g] : 0.0;
  x10[1] = okr ? S[(q0 + 2 * tq + 1) * DP_PITCH + r0 + 8 + g] : 0.0;
  x01[0] = (okq && !diag) ? S[(q0 + 8 + 2 * tq) * DP_PITCH + r0 + g] : 0.0;
  x01[1] = (okq && !diag) ? S[(q0 + 8 + 2 * tq + 1) * DP_PITCH + r0 + g] : 0.0;
  x11[0] = (okr && okq) ? S[(q0 + 8 + 2 * tq) * DP_PITCH + r0 + 8 + g] : 0.0;
  x11[1] = (okr && okq) ? S[(q0 + 8 + 2 * tq + 1) * DP_PITCH + r0 + 8 + g] : 0.0;
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const int col = (c0 + k4 * 4 + tq) * DP_PITCH;
    const double a0 = -S[col + r0 + g];
    const double a1 = okr ? -S[col + r0 + 8 + g] : 0.0;
    const double b0 = S[col + q0 + g];
    const double b1 = okq ? S[col + q0 + 8 + g] : 0.0;
    dmma_t(x00[0], x00[1], a0, b0);
    dmma_t(x10[0], x10[1], a1, b0);
    dmma_t(x11[0], x11[1], a1, b1);
    if (!diag) dmma_t(x01[0], x01[1], a0, b1);
  }
  S[(q0 + 2 * tq) * DP_PITCH + r0 + g] = x00[0];
  S[(q0 + 2 * tq + 1) * DP_PITCH + r0 + g] = x00[1];
  if (okr) {
    S[(q0 + 2 * tq) * DP_PITCH + r0 + 8 + g] = x10[0];
    S[(q0 + 2 * tq + 1) * DP_PITCH + r0 + 8 + g] = x10[1];
  }
  if (okq && !diag) {
    S[(q0 + 8 + 2 * tq) * DP_PITCH + r0 + g] = x01[0];
    S[(q0 + 8 + 2 * tq + 1) * DP_PITCH + r0 + g] = x01[1];
  }
  if (okr && okq) {
    S[(q0 + 8 + 2 * tq) * DP_PITCH + r0 + 8 + g] = x11[0];
    S[(q0 + 8 + 2 * tq + 1) * DP_PITCH + r0 + 8 + g] = x11[1];
  }
}

// One CTA factors a 128x128 diagonal tile and inverts the factor.
//   S(r,c) lives at S[c*132 + r] (column-major, conflict-free along r, 16-byte aligned columns).
//   The tile is processed in four 32-wide block columns.  Per block column:
//     1. warp 0 holds the 32x32 diagonal block one row per lane IN REGISTERS, runs the
//        right-looking Cholesky with warp shuffles (pivot check like LAPACK dpotrf: a pivot
//        <= 0 or NaN marks the matrix as failed) and inverts the 32x32 factor the same way;
//     2. all 8 warps: panel below  L21 = P * Inv^T  on the FP64 tensor pipe (DMMA);
//     3. all 8 warps: trailing update  S -= L21 L21^T  on DMMA.
//   Then D = L^-1 is assembled block by block (D_ij = -Inv_i * sum_k L_ik D_kj, DMMA); the
//   off-diagonal blocks of D are kept transposed in the unused upper triangle of S.
__global__ void __launch_bounds__(256) diag_kernel(DiagArgs a) {
  extern __shared__ __align__(16) double dsm[];
  double* S = dsm;                              // [T][129]
  double* Iv = S + T * DP_PITCH;                // [4][32][36]  Inv_p(r,c) at Iv[p][c*36 + r]
  double* Tm = Iv + 4 * SB * IVP;               // [4][32][36] scratch blocks, T(m,n) at Tm[n*36 + m]
  double* bsh = Tm + 4 * SB * IVP;              // b_k
  double* lg = bsh + T;                         // log L_jj
  __shared__ int s_failed;
  __shared__ long long stamps[32];
  int nst = 0;
  // stamped by warp 7 (not by the factor warp: a divergent branch in front of factor_block32 would send it
  // down its slow not-converged path and distort what is being measured)
#define STAMP() do { if (a.dbg && threadIdx.x == 224 && nst < 32) stamps[nst++] = clock64(); } while (0)
  // second timeline, kept by the factor warp itself (explicitly re-converged before it goes on)
  __shared__ long long stamps0[16];
  __shared__ long long fstamps[8];       // inside factor_block32: end of the pivot loop, end of the 16x16 inverses
  int nst0 = 0;
#define STAMP0() do { if (a.dbg) { if (threadIdx.x == 0 && nst0 < 16) stamps0[nst0++] = clock64(); __syncwarp(); } } while (0)
  STAMP();
  const int slot = a.sel[blockIdx.x];
  const int k = a.k, Np = a.Np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  double* A = a.Abuf + slot * a.smat + (long long)k * T + (long long)k * T * Np;
  const int nact = min(T, a.N - k * T);         // rows/cols holding data (the rest is identity)
  const int nact8 = (nact + 7) & ~7;

  for (int e = tid; e < 4 * SB * IVP; e += 256) {     // diagonal-block inverses start as identity
    const int within = e % (SB * IVP);
    Iv[e] = (within / IVP == within % IVP) ? 1.0 : 0.0;
  }
  if (tid == 0) s_failed = 0;
  // the tile (and the running right-hand side) are written by the preceding trailing update
  pdl_wait();
  pdl_launch();
  // lower triangle of the tile with 16-byte async copies (all in flight at once); nothing reads
  // the strict upper triangle of S before it is overwritten
  for (int e = tid; e < T * T / 2; e += 256) {
    const int r2 = (e & 63) * 2, c = e >> 6;
    if (r2 + 1 < c) continue;                         // chunk entirely above the diagonal
    const unsigned dst = (unsigned)__cvta_generic_to_shared(S + c * DP_PITCH + r2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(A + (long long)c * Np + r2));
  }
  asm volatile("cp.async.commit_group;\n" ::);
  if (tid < T) {
    bsh[tid] = a.bvec ? a.bvec[(long long)slot * Np + k * T + tid] : 0.0;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  STAMP();

  // Warp-specialised block loop.  The serial part is the 32x32 block factorisation in warp 0
  // (factor_block32, ~7.4 k cycles); everything else of a step is DMMA work that only the NEXT
  // step's panel needs.  So per step p:
  //   X  warps 0-3: panel rows of block p+1 (one 8-row block each), then the update of the diagonal
  //      block (p+1, p+1) -- all the next factorisation needs;
  //   Y  warps 4-7: the panel rows below block p+1;
  //   warp 0 goes straight on to factor block p+1 while warps 1-7 apply the rest of the trailing
  //   update (every 16x16 super-block except the three of block (p+1, p+1)).
  // Named barriers: 1, 2 = warps 0-3 (128 threads); 3 = all panel rows written (warp 0 only
  // arrives); 4 = end of step.  Every element sees the same operations in the same order as in a
  // plain sequential sweep.
  auto bar_sync = [](int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); };
  auto bar_arrive = [](int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); };
  for (int p = 0; p < T / SB; ++p) {
    const int c0 = p * SB;                               // block factored in this round
    if (c0 >= nact) break;
    const int cp = c0 - SB;                              // block whose panel / trailing update is due (p > 0)
    const int mblocks = (p > 0) ? (nact8 - c0) / 8 : 0;  // 8-row blocks below block p-1 (>= 1: block p exists)
    const int ms = (mblocks + 1) / 2;
    const int nsup = ms * (ms + 1) / 2;                  // 16x16 super-blocks of the trailing lower triangle
    if (p > 0) {
      const double* Ivp = Iv + (p - 1) * SB * IVP;
      // ---- X / Y: the panel of block p-1
      if (warp < 4) {
        if (warp < mblocks) diag_panel_rows(S, Ivp, cp, warp, g, tq);
        bar_sync(1, 128);                                // rows of block p complete
        if (warp < 3 && warp < nsup) diag_update_super(S, cp, mblocks, warp, g, tq);      // block (p, p)
        bar_sync(2, 128);                                // block (p, p) up to date
      } else {
        for (int mb = warp; mb < mblocks; mb += 4) diag_panel_rows(S, Ivp, cp, mb, g, tq);
      }
    }
    if (warp == 0) {
      if (p > 0) bar_arrive(3, 256);
      STAMP0();
      // ---- the serial part: diagonal block p in registers, lane r owns row r (single call site: the
      // function is 40 KB of unrolled code)
      if (factor_block32(S, Iv + p * SB * IVP, Tm, c0, lane, a.dbg ? fstamps + 2 * p : nullptr)) s_failed = 1;
      STAMP0();
    } else if (p > 0) {
      bar_sync(3, 256);                                  // every panel row of block p-1 is written
      // warp 4 shares its scheduler with warp 0: it sits the overlap out so that the (issue-bound)
      // factorisation keeps its issue slots; the other six warps carry the trailing update
      if (warp != 4) {
        const int w = (warp < 4) ? warp - 1 : warp - 2;  // 0..5
        for (int idx = 3 + w; idx < nsup; idx += 6) diag_update_super(S, cp, mblocks, idx, g, tq);
      }
    }
    bar_sync(4, 256);
    STAMP();
  }
  __syncthreads();
  STAMP();

  if (tid < T) lg[tid] = (tid < nact) ? log(S[tid * DP_PITCH + tid]) : 0.0;
  __syncthreads();
  // write L_kk back (lower triangle), two rows per 16-byte store
  for (int e = tid; e < T * T / 2; e += 256) {
    const int r2 = (e & 63) * 2, c = e >> 6;
    if (r2 + 1 < c) continue;
    const double2 v = *reinterpret_cast<const double2*>(S + c * DP_PITCH + r2);
    if (r2 >= c) *reinterpret_cast<double2*>(A + (long long)c * Np + r2) = v;
    else A[(long long)c * Np + r2 + 1] = v.y;
  }
  if (warp == 0) {                                     // sum of log L_jj, fixed order
    double sl = (lg[lane] + lg[lane + 32]) + (lg[lane + 64] + lg[lane + 96]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
    if (lane == 0) {
      a.logdet[(long long)slot * a.Nt + k] = sl;
      if (s_failed) a.fail[slot] = 1;
    }
  }

  __syncthreads();
  STAMP();
  // ---- D = L^-1 from the 32x32 blocks, recursively: D_jj = Inv_j; then the two 64x64 diagonal
  // halves get their off-diagonal block (round 1: blocks (1,0) and (3,2)), then the 64x64 block below
  // (round 2: blocks (2,0) (2,1) (3,0) (3,1)):  D_21 = -D_22 (L_21 D_11).  Four stages instead of the
  // twelve of a block-by-block sweep, with all eight warps busy in each.
  // D(R,C) for R in a later block than C is stored transposed at S[R*132 + C] (operand layout of
  // the next stage) AND in natural layout over L(R,C), which nothing reads any more at that point
  // (L is already written back): the output pass then reads D without bank conflicts.
  {
    auto put_D = [&](int R, int C, double v0, double v1) {     // D(R, C) and D(R, C+1)
      S[R * DP_PITCH + C] = v0;
      S[R * DP_PITCH + C + 1] = v1;
      S[C * DP_PITCH + R] = v0;
      S[(C + 1) * DP_PITCH + R] = v1;
    };
    // round 1, stage A: T_p = L(i,j) Inv_j for the pairs p = 0: (1,0), p = 1: (3,2)
    {
      const int pr = warp >> 2, ni = warp & 3, i = 2 * pr + 1, j = 2 * pr;
      if (i * SB < nact) {
        double acc[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const int q = k4 * 4 + tq;
          const double bf = Iv[j * SB * IVP + (ni * 8 + g) * IVP + q];                    // Inv_j(q, n)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
            dmma_t(acc[mi][0], acc[mi][1], S[(j * SB + q) * DP_PITCH + i * SB + mi * 8 + g], bf);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          Tm[pr * SB * IVP + (ni * 8 + 2 * tq) * IVP + mi * 8 + g] = acc[mi][0];
          Tm[pr * SB * IVP + (ni * 8 + 2 * tq + 1) * IVP + mi * 8 + g] = acc[mi][1];
        }
      }
    }
    __syncthreads();
    // round 1, stage B: D(i,j) = -Inv_i T_p
    {
      const int pr = warp >> 2, ni = warp & 3, i = 2 * pr + 1, j = 2 * pr;
      if (i * SB < nact) {
        double acc[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const int q = k4 * 4 + tq;
          const double bf = Tm[pr * SB * IVP + (ni * 8 + g) * IVP + q];                   // T_p(q, n)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
            dmma_t(acc[mi][0], acc[mi][1], -Iv[i * SB * IVP + q * IVP + mi * 8 + g], bf);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
          put_D(i * SB + mi * 8 + g, j * SB + ni * 8 + 2 * tq, acc[mi][0], acc[mi][1]);
      }
    }
    __syncthreads();
    // round 2, stage A: X(i,j) = sum_{kb=j..1} L(i,kb) D11(kb,j) for i in {2,3}, j in {0,1}
    {
      const int o = warp >> 1, nh = warp & 1, i = 2 + (o >> 1), j = o & 1;
      if (i * SB < nact) {
        double acc[2][4][2];
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) acc[n2][mi][0] = acc[n2][mi][1] = 0.0;
        for (int kb = j; kb < 2; ++kb) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int q = k4 * 4 + tq;
            double bf[2];
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2) {
              const int n = (2 * nh + n2) * 8 + g;
              bf[n2] = (kb == j) ? Iv[j * SB * IVP + n * IVP + q]                         // Inv_j(q, n)
                                 : S[(SB + q) * DP_PITCH + n];                            // D(1,0)(q, n)
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
              const double af = S[(kb * SB + q) * DP_PITCH + i * SB + mi * 8 + g];       // L(i,kb)(m, q)
              dmma_t(acc[0][mi][0], acc[0][mi][1], af, bf[0]);
              dmma_t(acc[1][mi][0], acc[1][mi][1], af, bf[1]);
            }
          }
        }
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            const int n = (2 * nh + n2) * 8 + 2 * tq;
            Tm[o * SB * IVP + n * IVP + mi * 8 + g] = acc[n2][mi][0];
            Tm[o * SB * IVP + (n + 1) * IVP + mi * 8 + g] = acc[n2][mi][1];
          }
      }
    }
    __syncthreads();
    // round 2, stage B: D(i,j) = -sum_{k=2..i} D22(i,k) X(k,j)
    {
      const int o = warp >> 1, nh = warp & 1, i = 2 + (o >> 1), j = o & 1;
      if (i * SB < nact) {
        double acc[2][4][2];
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) acc[n2][mi][0] = acc[n2][mi][1] = 0.0;
        for (int k = 2; k <= i; ++k) {
          const double* X = Tm + ((k - 2) * 2 + j) * SB * IVP;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int q = k4 * 4 + tq;
            double bf[2];
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2) bf[n2] = X[((2 * nh + n2) * 8 + g) * IVP + q];   // X(k,j)(q, n)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
              const int m = mi * 8 + g;
              const double af = (k == i) ? -Iv[i * SB * IVP + q * IVP + m]                // -Inv_i(m, q)
                                         : -S[(2 * SB + q) * DP_PITCH + 3 * SB + m];     // -D(3,2)(m, q), natural copy
              dmma_t(acc[0][mi][0], acc[0][mi][1], af, bf[0]);
              dmma_t(acc[1][mi][0], acc[1][mi][1], af, bf[1]);
            }
          }
        }
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
            put_D(i * SB + mi * 8 + g, j * SB + (2 * nh + n2) * 8 + 2 * tq, acc[n2][mi][0], acc[n2][mi][1]);
      }
    }
  }
  __syncthreads();
  STAMP();
  // element (rr, c) of D, rr >= c
  auto Dval = [&](int rr, int c) -> double {
    if ((rr >> 5) == (c >> 5)) return Iv[(rr >> 5) * SB * IVP + (c & 31) * IVP + (rr & 31)];
    return (rr < nact) ? S[c * DP_PITCH + rr] : 0.0;      // natural copy written by pass B
  };
  // ---- outputs: D_k (column-major, zeros above the diagonal), D_k^T, W diag tile, z_k
  double* Dk = a.Dbuf + ((long long)slot * a.Nt + k) * T * T;
  double* DTk = a.DTbuf + ((long long)slot * a.Nt + k) * T * T;
  double* Wd = a.Wbuf ? a.Wbuf + slot * a.smat + (long long)k * T + (long long)k * T * Np : nullptr;
  {
    // thread = (row rr, column parity): per 32-column block the source is warp-uniform, so the
    // shared-memory loads of a block are issued together ahead of the global stores
    const int rr = tid & (T - 1), cpar = tid >> 7, rb = rr >> 5;
#pragma unroll 1
    for (int cb = 0; cb < T / SB; ++cb) {
      double v[16], w[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int c = cb * SB + 2 * u + cpar;
        // D(rr, c), rr >= c
        v[u] = (cb == rb) ? ((rr >= c) ? Iv[cb * SB * IVP + (c & 31) * IVP + (rr & 31)] : 0.0)
                          : ((cb < rb && rr < nact) ? S[c * DP_PITCH + rr] : 0.0);
        // D^T(rr, c) = D(c, rr), c >= rr -- only the gradient / posterior paths (those that also
        // keep W) read D_k^T: the nlZ-only path skips it
        w[u] = !Wd ? 0.0
             : (cb == rb) ? ((c >= rr) ? Iv[cb * SB * IVP + (rr & 31) * IVP + (c & 31)] : 0.0)
                          : ((cb > rb && c < nact) ? S[c * DP_PITCH + rr] : 0.0);
      }
      // blocks that are identically zero (D above, D^T below the block diagonal) are not stored:
      // the buffers are zeroed when they are allocated and nothing else writes there
      if (cb <= rb) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = cb * SB + 2 * u + cpar;
          Dk[c * T + rr] = v[u];
          if (Wd) Wd[(long long)c * Np + rr] = v[u];
        }
      }
      if (cb >= rb && Wd) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = cb * SB + 2 * u + cpar;
          DTk[c * T + rr] = w[u];
        }
      }
    }
  }
  if (a.zvec && tid < T) {
    const int rr = tid;
    // four interleaved partial sums (c mod 4), same order as diag_solve_kernel
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    if (rr < nact) {
      int c = 0;
      for (; c + 3 <= rr; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s4[u] = __fma_rn(Dval(rr, c + u), bsh[c + u], s4[u]);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (c + u <= rr) s4[u] = __fma_rn(Dval(rr, c + u), bsh[c + u], s4[u]);
    }
    a.zvec[(long long)slot * Np + k * T + rr] = __dadd_rn(__dadd_rn(s4[0], s4[1]), __dadd_rn(s4[2], s4[3]));
  }
  __syncthreads();
  STAMP();
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    a.dbg[40] = nst0;
    for (int i = 0; i < nst0; ++i) a.dbg[41 + i] = stamps0[i];
    for (int i = 0; i < 8; ++i) a.dbg[56 + i] = fstamps[i];
  }
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 224) {
    a.dbg[0] = nst;
    for (int i = 0; i < nst; ++i) a.dbg[1 + i] = stamps[i];
  }
#undef STAMP
#undef STAMP0
}

// z_k = D_k b_k on an existing factor (solve-only replay; same arithmetic as diag_kernel's tail)
struct DiagSolveArgs {
  const double* Dbuf; const int* sel;
  const int* fsel;                 // slot holding the factor used by entry t (null = its own)
  int Np, Nt, N, k;
  const double* bvec; double* zvec;
};

__global__ void __launch_bounds__(T) diag_solve_kernel(DiagSolveArgs a) {
  __shared__ double bsh[T];
  const int slot = a.sel[blockIdx.x];
  const int rr = threadIdx.x;
  const int nact = min(T, a.N - a.k * T);
  bsh[rr] = a.bvec[(long long)slot * a.Np + a.k * T + rr];
  __syncthreads();
  const int fslot = a.fsel ? a.fsel[blockIdx.x] : slot;
  const double* Dk = a.Dbuf + ((long long)fslot * a.Nt + a.k) * T * T;
  a.zvec[(long long)slot * a.Np + a.k * T + rr] = zsolve_row(Dk, bsh, rr, nact);
}

// dst[slot][:] = src[slot][:] for the listed slots only
__global__ void copy_sel_kernel(double* dst, const double* src, const int* sel, int Np) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  const long long o = (long long)sel[blockIdx.y] * Np + i;
  dst[o] = src[o];
}

struct VecArgs {
  const double* Abuf; const double* DTbuf;
  const int* sel;
  long long smat;
  int Np, Nt, k;
  double* bvec; const double* zvec;
  double* alpha;             // [nslots][Np]
  const SlotP* sp;
};

// backward substitution, block row i = a.k (descending): every CTA recomputes
// w_i = D_i^T b_i; CTA j < i applies b_j -= L_ij^T w_i; CTA 0 stores alpha_i = w_i / sl.
// 256 threads; loads are issued in batches so many are in flight (the kernel is pure latency).
__global__ void __launch_bounds__(256) bwd_step_kernel(VecArgs a) {
  __shared__ double w[T];
  __shared__ double bi[T];
  __shared__ double part[2][T];
  const int slot = a.sel[blockIdx.y];
  const int i = a.k, j = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < T) bi[tid] = a.bvec[(long long)slot * a.Np + i * T + tid];
  __syncthreads();
  {
    // (D^T b)(n) = sum_m DT(n,m) b(m): thread (n, h) sums the m of parity h
    const double* DT = a.DTbuf + ((long long)slot * a.Nt + i) * T * T;
    const int n = tid & (T - 1), h = tid >> 7;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (int m = h; m < T; m += 8) {
      s0 += DT[m * T + n] * bi[m];
      s1 += DT[(m + 2) * T + n] * bi[m + 2];
      s2 += DT[(m + 4) * T + n] * bi[m + 4];
      s3 += DT[(m + 6) * T + n] * bi[m + 6];
    }
    part[h][n] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (tid < T) {
    const double s = part[0][tid] + part[1][tid];
    w[tid] = s;
    if (j == 0) a.alpha[(long long)slot * a.Np + i * T + tid] = s / a.sp[slot].sl;   // :2455-2465
  }
  __syncthreads();
  if (j >= i) return;
  // (L_ij^T w)(n) = sum_m L(iT+m, jT+n) w(m): 8 warps x 16 columns, lanes over m, 4 columns per pass
  const double* L = a.Abuf + slot * a.smat + (long long)i * T + (long long)j * T * a.Np;
  const int lane = tid & 31, warp = tid >> 5;
  const double w0 = w[lane], w1 = w[lane + 32], w2 = w[lane + 64], w3 = w[lane + 96];
  for (int n0 = warp * 16; n0 < warp * 16 + 16; n0 += 4) {
    double v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double* col = L + (long long)(n0 + c) * a.Np;
      v[c] = col[lane] * w0 + col[lane + 32] * w1 + col[lane + 64] * w2 + col[lane + 96] * w3;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    }
    if (lane < 4) {
      const double r = (lane == 0) ? v[0] : (lane == 1 ? v[1] : (lane == 2 ? v[2] : v[3]));
      a.bvec[(long long)slot * a.Np + j * T + n0 + lane] -= r;
    }
  }
}

// alpha = L^-T z / sl as W^T z once W = L^-1 exists (gradient path): one launch of column dot
// products (one warp per column, coalesced along the column) instead of the Nt dependent launches
// of the backward substitution.
struct AlphaArgs {
  const double* Wbuf; const int* sel;
  long long smat;
  int Np, N;
  const double* zvec; double* alpha;
  const SlotP* sp;
};

__global__ void __launch_bounds__(256) alpha_gemv_kernel(AlphaArgs a) {
  const int slot = a.sel[blockIdx.y];
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= a.Np) return;
  double* alpha = a.alpha + (long long)slot * a.Np;
  if (j >= a.N) {
    if (lane == 0) alpha[j] = 0.0;
    return;
  }
  const double* col = a.Wbuf + slot * a.smat + (long long)j * a.Np;
  const double* z = a.zvec + (long long)slot * a.Np;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = (j & ~31) + lane;                    // aligned start: the loads coalesce
  if (i < j) i += 32;
  for (; i + 96 < a.N; i += 128) {
    s0 = fma(col[i], z[i], s0);
    s1 = fma(col[i + 32], z[i + 32], s1);
    s2 = fma(col[i + 64], z[i + 64], s2);
    s3 = fma(col[i + 96], z[i + 96], s3);
  }
  for (; i < a.N; i += 32) s0 = fma(col[i], z[i], s0);
  double t = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) alpha[j] = t / a.sp[slot].sl;                       // :2455-2465
}

struct NlzArgs {
  const int* sel;
  const int* fsel;   // optional: slot whose factor (log-det partials) entry t uses; null = its own
  int N, Np, Nt;
  const double* zvec; const double* logdet;
  const SlotP* sp;
  double* nlz;       // [nslots]
};

__global__ void __launch_bounds__(256) nlz_kernel(NlzArgs a) {
  __shared__ double sh[256];
  const int slot = a.sel[blockIdx.x];
  const double* z = a.zvec + (long long)slot * a.Np;
  double s = 0.0;
  for (int i = threadIdx.x; i < a.N; i += 256) s += z[i] * z[i];
  s = block_sum<256>(s, sh);
  if (threadIdx.x == 0) {
    const SlotP p = a.sp[slot];
    double ld = 0.0;
    const int fslot = a.fsel ? a.fsel[blockIdx.x] : slot;
    for (int k = 0; k < a.Nt; ++k) ld += a.logdet[(long long)fslot * a.Nt + k];
    // gaussian_process.py:2469-2473 with (y-m)^T alpha = z^T z / sl
    a.nlz[slot] = s / p.sl / 2 + ld + a.N * log(2 * M_PI * p.sl) / 2;
  }
}

// copy helpers -------------------------------------------------------------------
// Posterior.L as the reference stores it: row-major (N,N) upper factor U = L^T
// (zeros below the diagonal), or -Ainv (symmetric) for the low-noise branch.
__global__ void fetch_L_kernel(const double* Abuf, int Np, int N, int lchol, double* out) {
  const long long total = (long long)N * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);    // out[i][j]
    double v;
    if (lchol) v = (j >= i) ? Abuf[(long long)i * Np + j] : 0.0;          // U(i,j) = L(j,i)
    else v = -((j >= i) ? Abuf[(long long)i * Np + j] : Abuf[(long long)j * Np + i]);
    out[e] = v;
  }
}

// mirror the lower triangle into the upper one (tile pairs), for the low-noise predict GEMM
__global__ void symmetrize_kernel(double* Abuf, int Np) {
  const long long total = (long long)Np * Np;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e % Np), c = (int)(e / Np);
    if (r > c) Abuf[(long long)r * Np + c] = Abuf[e];
  }
}

__global__ void fill_kernel(double* p, double v, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) p[e] = v;
}
__global__ void copy_kernel(double* dst, const double* src, long long n) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) dst[e] = src[e];
}
// out[dst[t]] = in[src[t]]  (jitter multipliers: per-row value <-> value of the factor in a slot)
__global__ void copy_mult_kernel(double* out, const double* in, const int* dst, const int* src, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[dst[t]] = in[src[t]];
}
__global__ void set_mult_kernel(double* mult, int* fail, const int* sel, int nsel, double v) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nsel) { mult[sel[t]] = v; fail[sel[t]] = 0; }
}
__global__ void scale_mult_kernel(double* mult, const int* fail, const int* sel, int nsel) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nsel) { const int s = sel[t]; if (fail[s]) mult[s] *= 10.0; }
}

}  // namespace gpb
