// Covariance / mean / noise kernels:
//   prep_kernel   per-slot hyperparameter decoding, pre-scaled inputs, residual y-m, noise vector
//   build_kernel  K1: fused covariance builder  A = K/sl + diag(sn2/sn2_div)  (or K + mult*diag(sn2))
//   grad_kernel   K3: fused gradient reduction  sum_ij Q_ij dK_ij/dtheta  without storing dK
//   grad_final    noise / mean gradient + deterministic tile reduction
//   plugin kernels behind covariance/mean/noise .compute()
#pragma once
#include "common.cuh"

namespace gpb {

// scaled coordinate exactly as each reference kernel scales its inputs
//   SE:        X / ell                      covariance_functions.py:165 ; isotropic:203
//   Matern:    X * (sqrt(nu)/ell)  (ARD)    covariance_functions.py:252
//              (X * sqrt(nu)) / ell (iso)   isotropic_covariance_functions.py:134
//   RQ:        X * (1/ell)                  covariance_functions.py:332
__device__ __forceinline__ double scale_coord(int cov_kind, int ard, int degree, double x,
                                              double ell) {
  if (cov_kind == 0) return x / ell;
  if (cov_kind == 2) return x * (1.0 / ell);
  const double sq = sqrt((double)degree);
  if (ard) return x * (sq / ell);
  return (x * sq) / ell;
}

// mean m(x) for one point (mean_functions.py:126, :254-255, :384-388)
__device__ __forceinline__ double mean_value(int mean_kind, int D, const double* hm,
                                             const double* x) {
  if (mean_kind == 0) return 0.0;
  if (mean_kind == 1) return hm[0];
  double s = 0.0;
  for (int k = 0; k < D; ++k) {
    const double z = (x[k] - hm[1 + k]) / exp(hm[1 + D + k]);
    s += z * z;
  }
  return hm[0] - 0.5 * s;
}

// noise variance sn2_i for one point (noise_functions.py:248-278); y/s2 may be absent
__device__ __forceinline__ double noise_value(int nz0, int nz1, int nz2, const double* hn,
                                              bool has_y, double y, bool has_s2, double s2) {
  int i = 0;
  double sn2;
  if (nz0 == 0) {
    sn2 = 2.220446049250313e-16;   // np.spacing(1.0)
  } else {
    sn2 = exp(2 * hn[i]);
    ++i;
  }
  const double s2v = has_s2 ? s2 : 0.0;
  if (nz1 == 1) {
    sn2 += s2v;
  } else if (nz1 == 2) {
    sn2 += exp(hn[i]) * s2v;
    ++i;
  }
  if (nz2 == 1) {
    if (has_y) {
      const double zz = fmax(0.0, hn[i] - y);
      sn2 += exp(2 * hn[i + 1]) * (zz * zz);
    }
  }
  return sn2;
}

struct PrepArgs {
  Model md;
  int N, Np;
  const double* X;      // (N,D) row-major
  const double* y;      // (N)
  const double* s2;     // (N) or null
  const double* hyp;    // [nslots][P]
  const int* sel;
  const double* mult;   // [nslots]
  double* xs;           // [nslots][D][Np]  pre-scaled inputs, dimension-major
  double* resid;        // [nslots][Np]     y - m
  double* sn2v;         // [nslots][Np]
  SlotP* sp;            // [nslots]
  double* part_min;     // [nslots][PREP_MAX_CTAS] scratch: per-CTA min of sn2
  int* part_nan;        // [nslots][PREP_MAX_CTAS]
  int* ticket;          // [nslots] zero between launches
};

// Several CTAs per slot (a lone CTA over N*D divisions was 80 us of a 3.6 ms one-matrix nlZ): each
// handles a strided share of the points and leaves its partial min / NaN flag in scratch; the CTA
// that finishes last (per-slot ticket) combines them -- min is order-independent -- and writes SlotP.
constexpr int PREP_THREADS = 256;
constexpr int PREP_MAX_CTAS = 32;
__global__ void __launch_bounds__(PREP_THREADS) prep_kernel(PrepArgs a) {
  __shared__ double sh[PREP_THREADS];
  __shared__ int shnan[PREP_THREADS];
  __shared__ int s_last;
  const int slot = a.sel[blockIdx.x];
  const Model& md = a.md;
  const int D = md.D, N = a.N, Np = a.Np;
  const double* h = a.hyp + (long long)slot * md.P;
  const double* hn = h + md.cov_n;
  const double* hm = h + md.cov_n + md.noise_n;
  const int nl = md.ard ? D : 1;
  double* xs = a.xs + (long long)slot * D * Np;
  double* resid = a.resid + (long long)slot * Np;
  double* sn2v = a.sn2v + (long long)slot * Np;

  // exp() of the length scales and of the mean's scales once per CTA, not once per point
  __shared__ double ells[MAXD], oms[MAXD];
  if (threadIdx.x < D) {
    ells[threadIdx.x] = exp(h[md.ard ? threadIdx.x : 0]);
    oms[threadIdx.x] = (md.mean_kind == 2) ? exp(hm[1 + D + threadIdx.x]) : 1.0;
  }
  __syncthreads();
  double vmin = INFINITY;
  int anynan = 0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < Np; i += gridDim.y * blockDim.x) {
    if (i < N) {
      const double* x = a.X + (long long)i * D;
      for (int k = 0; k < D; ++k)
        xs[(long long)k * Np + i] = scale_coord(md.cov_kind, md.ard, md.degree, x[k], ells[k]);
      double m = 0.0;                                  // mean_value() with the hoisted exp
      if (md.mean_kind == 1) m = hm[0];
      else if (md.mean_kind == 2) {
        double sq = 0.0;
        for (int k = 0; k < D; ++k) {
          const double z = (x[k] - hm[1 + k]) / oms[k];
          sq += z * z;
        }
        m = hm[0] - 0.5 * sq;
      }
      resid[i] = a.y[i] - m;
      const double v = noise_value(md.nz0, md.nz1, md.nz2, hn, true, a.y[i], a.s2 != nullptr,
                                   a.s2 ? a.s2[i] : 0.0);
      sn2v[i] = v;
      if (v != v) anynan = 1;
      vmin = fmin(vmin, v);
    } else {
      for (int k = 0; k < D; ++k) xs[(long long)k * Np + i] = 0.0;
      resid[i] = 0.0;
      sn2v[i] = 1.0;
    }
  }
  sh[threadIdx.x] = vmin;
  shnan[threadIdx.x] = anynan;
  __syncthreads();
  for (int s = PREP_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
      shnan[threadIdx.x] |= shnan[threadIdx.x + s];
    }
    __syncthreads();
  }
  double* pmin = a.part_min + (long long)slot * PREP_MAX_CTAS;
  int* pnan = a.part_nan + (long long)slot * PREP_MAX_CTAS;
  if (threadIdx.x == 0) {
    pmin[blockIdx.y] = sh[0];
    pnan[blockIdx.y] = shnan[0];
    __threadfence();
    s_last = (atomicAdd(a.ticket + slot, 1) == (int)gridDim.y - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  double mn = INFINITY;
  int nan = 0;
  for (int c = 0; c < (int)gridDim.y; ++c) {
    mn = fmin(mn, ((volatile double*)pmin)[c]);
    nan |= ((volatile int*)pnan)[c];
  }
  a.ticket[slot] = 0;                            // ready for the next launch
  SlotP p;
  p.sf2 = exp(2 * h[nl]);
  p.rq_a = (md.cov_kind == 2) ? exp(h[nl + 1]) : 1.0;
  p.sn2_min = nan ? NAN : mn;                    // np.min propagates NaN
  p.mult = a.mult[slot];
  p.lchol = (p.sn2_min >= 1e-6) ? 1 : 0;         // gaussian_process.py:2404
  p.sl = p.lchol ? p.sn2_min * p.mult : 1.0;     // :2422 / :2439
  p.pad = 0;
  a.sp[slot] = p;
}

// ---------------------------------------------------------------------------------
// K1: build the (scaled) covariance matrix, lower tiles only.
// thread (tx, ty): rows tx + 32a (a < 4), columns ty*16 + b (b < 16) of a 128x128 tile.
// ---------------------------------------------------------------------------------
struct BuildArgs {
  int D, N, Np, Nt;
  const int* sel;
  const double* xs;
  const double* sn2v;
  const SlotP* sp;
  double* Abuf;
  long long smat;
  int full;          // 1: every tile (i,j) (grid.x = Nt*Nt); 0: lower tiles only
};

template <int KIND>
__global__ void __launch_bounds__(256) build_kernel(BuildArgs a) {
  extern __shared__ double bsm[];
  const int slot = a.sel[blockIdx.y];
  int ti, tj;
  if (a.full) { ti = blockIdx.x % a.Nt; tj = blockIdx.x / a.Nt; }
  else tri_decode(blockIdx.x, ti, tj);
  const int D = a.D, Np = a.Np, N = a.N;
  double* xr = bsm;              // [D][128]
  double* xc = bsm + D * T;      // [D][128]
  const double* xs = a.xs + (long long)slot * D * Np;
  for (int e = threadIdx.x; e < D * T; e += blockDim.x) {
    const int k = e / T, i = e % T;
    xr[e] = xs[(long long)k * Np + ti * T + i];
    xc[e] = xs[(long long)k * Np + tj * T + i];
  }
  __syncthreads();
  const SlotP p = a.sp[slot];
  const double inv_sl = 1.0 / p.sl;      // one reciprocal per CTA instead of a division per element
  const double* sn2v = a.sn2v + (long long)slot * Np;
  double* A = a.Abuf + slot * a.smat;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

  for (int b0 = 0; b0 < 16; b0 += 4) {
    double r2[4][4];
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) r2[aa][bb] = 0.0;
    for (int k = 0; k < D; ++k) {
      double xa[4], xb[4];
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) xa[aa] = xr[k * T + tx + 32 * aa];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) xb[bb] = xc[k * T + ty * 16 + b0 + bb];
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const double d = xa[aa] - xb[bb];
          // sequential, unfused sum like scipy's sqeuclidean kernel
          r2[aa][bb] = __dadd_rn(r2[aa][bb], __dmul_rn(d, d));
        }
    }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int gj = tj * T + ty * 16 + b0 + bb;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        const int gi = ti * T + tx + 32 * aa;
        double v;
        if (gi >= N || gj >= N) {
          v = (gi == gj) ? 1.0 : 0.0;
        } else {
          const double K = kern_value<KIND>(r2[aa][bb], p.sf2, p.rq_a);
          if (p.lchol) {
            v = K * inv_sl;                                          // gaussian_process.py:2416
            if (gi == gj) v += sn2v[gi] / p.sn2_min;                 // :2409-2412
          } else {
            v = K;                                                   // :2433
            if (gi == gj) v += p.mult * sn2v[gi];
          }
        }
        A[(long long)gj * Np + gi] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// K3: gradient reduction over the lower tiles of Q = Ainv/sl - alpha alpha^T.
//   gpart[slot][tile][p] = sum over the tile of  w_ij * Q_ij * dK_ij/dtheta_p,
//   w = 1 below the diagonal, 1/2 on it  (so the full-matrix  1/2 sum_ij  is reproduced,
//   gaussian_process.py:2487-2488).  dK is regenerated from the pre-scaled inputs.
// DP = compile-time bound on the number of ARD length scales (0 -> isotropic).
// ---------------------------------------------------------------------------------
struct GradArgs {
  int D, N, Np, Nt, cov_n;
  const int* sel;
  const double* xs;
  const SlotP* sp;
  const double* Abuf;     // lower tiles hold Ainv
  const double* alpha;    // [nslots][Np]
  double* gpart;          // [nslots][ntiles][cov_n]
  long long smat;
};

// four CTAs per SM up to 12 length scales (64 registers; measured against three: see DESIGN.md)
template <int KIND, int DP>
__global__ void __launch_bounds__(256, (DP <= 12 ? 4 : (DP <= 16 ? 3 : (DP <= 32 ? 2 : 1)))) grad_kernel(GradArgs a) {
  extern __shared__ double bsm[];
  constexpr int NACC = (DP > 0 ? DP : 1) + 2;      // length scales | sf | rq shape
  const int slot = a.sel[blockIdx.y];
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  const int D = a.D, Np = a.Np, N = a.N;
  double* xr = bsm;
  double* xc = bsm + D * T;
  double* al_r = bsm + 2 * D * T;       // [128]
  double* al_c = al_r + T;              // [128]
  double* red = al_c + T;               // [8][NACC]
  const double* xs = a.xs + (long long)slot * D * Np;
  for (int e = threadIdx.x; e < D * T; e += blockDim.x) {
    const int k = e / T, i = e % T;
    xr[e] = xs[(long long)k * Np + ti * T + i];
    xc[e] = xs[(long long)k * Np + tj * T + i];
  }
  const double* alpha = a.alpha + (long long)slot * Np;
  if (threadIdx.x < T) {
    al_r[threadIdx.x] = alpha[ti * T + threadIdx.x];
    al_c[threadIdx.x] = alpha[tj * T + threadIdx.x];
  }
  __syncthreads();
  const SlotP p = a.sp[slot];
  const double inv_sl = 1.0 / p.sl;      // one reciprocal per CTA instead of a division per element
  const double* Ainv = a.Abuf + slot * a.smat;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

  double acc[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q] = 0.0;

  for (int b = 0; b < 16; ++b) {
    const int jl = ty * 16 + b;
    const int gj = tj * T + jl;
    if (gj >= N) continue;
    double r2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < D; ++k) {
      const double xb = xc[k * T + jl];
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        const double d = xr[k * T + tx + 32 * aa] - xb;
        // fused here (two FP64 instructions per dimension instead of three): the gradient tolerates
        // the one-ulp difference from the matrix builder's unfused, scipy-ordered sum
        r2[aa] = fma(d, d, r2[aa]);
      }
    }
    double cw[4];
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
      const int il = tx + 32 * aa;
      const int gi = ti * T + il;
      cw[aa] = 0.0;
      if (gi >= N || gi < gj) continue;
      const double w = (gi == gj) ? 0.5 : 1.0;
      const double Q = w * (Ainv[(long long)gj * Np + gi] * inv_sl - al_r[il] * al_c[jl]);   // :2477-2484
      double K, c, dshape;
      kern_value_grad<KIND>(r2[aa], p.sf2, p.rq_a, K, c, dshape);
      acc[NACC - 2] += Q * (2 * K);                  // dK/dlog(sf) = 2K
      if (KIND == 2) acc[NACC - 1] += Q * dshape;
      if (DP == 0) acc[0] += (Q * c) * r2[aa];       // isotropic: dK/dlog(ell) = c * r^2
      cw[aa] = Q * c;
    }
    if (DP > 0) {
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        if (k < D) {
          const double xb = xc[k * T + jl];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            const int gi = ti * T + tx + 32 * aa;
            if (gi < N && gi >= gj) {
              const double d = xr[k * T + tx + 32 * aa] - xb;
              acc[k] += cw[aa] * (d * d);
            }
          }
        }
      }
    }
  }
  // deterministic reduction: warp xor-tree, then warps in order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NACC + q] = v;
  }
  __syncthreads();
  if (threadIdx.x < a.cov_n) {
    // map hyperparameter index -> accumulator
    const int pidx = threadIdx.x;
    const int nl = (DP > 0) ? D : 1;
    int q;
    if (pidx < nl) q = pidx;
    else if (pidx == nl) q = NACC - 2;
    else q = NACC - 1;
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w * NACC + q];
    const long long ntiles = (long long)a.Nt * (a.Nt + 1) / 2;
    a.gpart[((long long)slot * ntiles + blockIdx.x) * a.cov_n + pidx] = s;
  }
}

// ---------------------------------------------------------------------------------
// K3, row-per-thread form (ARD with D <= 16: the shapes a fit spends its time in).
// grad_kernel above re-reads its 4 rows' coordinates from shared memory for every column and
// evaluates the radial function inside divergent regions (the library sqrt/exp carry slow-path
// branches, so the four pairs of a round never overlap): 25 shared loads and ~95 FP64 instructions
// per pair, FP64 pipe 48-52 % busy.  Here a thread owns ONE row of the tile: its D coordinates live
// in registers, a column's coordinates (and alpha_j) are one contiguous, warp-uniform run in shared
// memory (broadcast 128-bit loads: D/2+1 per pair), the squared differences are kept from the
// distance sum for the length-scale accumulators (4 FP64 instructions per dimension instead of 5),
// and the radial functions are branch-free (below), so the whole pair is one basic block.
// Masked pairs (upper triangle of a diagonal tile, padding) run with Q = 0.  sf^2 and the factor 2
// of dK/dlog(sf) are applied once per tile after the reduction.
// ---------------------------------------------------------------------------------

// exp(-x) for x >= 0 without a slow path: Cody-Waite reduction by ln2 (hi/lo), degree-12 Taylor
// polynomial on |r| <= ln2/2 (truncation 1.7e-16), 2^n added into the exponent field.  Arguments above
// 708 are evaluated at 708 (3e-308 instead of a denormal or 0: nothing downstream can tell at the
// gradient tolerance); NaN stays NaN.
__constant__ double EXP_C[16] = {
    1.4426950408889634, 6755399441055744.0 /* 1.5 * 2^52 */, -6.93147180559945286e-01, -2.31904681384629956e-17,
    2.08767569878680990e-09 /* 1/12! */, 2.50521083854417188e-08, 2.75573192239858907e-07,
    2.75573192239858907e-06, 2.48015873015873016e-05, 1.98412698412698413e-04, 1.38888888888888889e-03,
    8.33333333333333333e-03, 4.16666666666666667e-02, 1.66666666666666667e-01, 0.5, 1.0};
// (the constants sit in constant memory so that each is an operand of its FMA: as immediates the
// compiler rebuilt them with two uniform moves per use, 22 extra issue slots per pair)
// exp(-x) for x >= 0
__device__ __forceinline__ double exp_neg(double x) {
  x = (x > 708.0) ? 708.0 : x;
  const double t = fma(x, -EXP_C[0], EXP_C[1]);
  const int n = __double2loint(t);
  const double nf = t - EXP_C[1];                 // -round(x / ln 2)
  double r = fma(nf, EXP_C[2], -x);
  r = fma(nf, EXP_C[3], r);
  double p = EXP_C[4];
#pragma unroll
  for (int q = 5; q < 16; ++q) p = fma(p, r, EXP_C[q]);
  p = fma(p, r, EXP_C[15]);
  return __hiloint2double(__double2hiint(p) + (int)((unsigned)n << 20), __double2loint(p));
}

// s = sqrt(x), rinv = 1/sqrt(x) for x >= 0 without a slow path: hardware seed (2^-22), two coupled
// Goldschmidt steps.  x = 0 gives s = 0, rinv = +inf.
__device__ __forceinline__ void sqrt_rsqrt_nonneg(double x, double& s, double& rinv) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const bool zero = (__double2hiint(x) | __double2loint(x)) == 0;
  s = zero ? 0.0 : g;
  rinv = zero ? __longlong_as_double(0x7ff0000000000000LL) : 2.0 * h;
}

// radial factors WITHOUT sf^2:  K = sf2*kf,  dK/dlog(ell_k) = sf2*cf*Delta_k^2,  dK/dlog(shape) = sf2*sh
template <int KIND>
__device__ __forceinline__ void radial_factors(double r2, double a_rq, double half_over_a,
                                               double& kf, double& cf, double& sh) {
  sh = 0.0;
  if (KIND == 0) {
    kf = exp_neg(0.5 * r2);
    cf = kf;
  } else if (KIND == 2) {
    const double Mq = fma(r2, half_over_a, 1.0);
    const double lg = log(Mq);
    kf = exp(-a_rq * lg);
    cf = kf / Mq;
    sh = kf * (0.5 * r2 / Mq - a_rq * lg);
  } else {
    double r, rinv;
    sqrt_rsqrt_nonneg(r2, r, rinv);
    const double e = exp_neg(r);
    if (KIND == 1) { kf = e; cf = rinv * e; }                       // inf at r = 0, as the reference
    else if (KIND == 3) { kf = fma(r, e, e); cf = e; }
    else {
      const double third = 1.0 / 3;
      kf = fma(r, fma(r, third, 1.0), 1.0) * e;
      cf = fma(r, third, third) * e;
    }
  }
}

template <int KIND, int DP>
__global__ void __launch_bounds__(256, 2) grad_rows_kernel(GradArgs a) {
  extern __shared__ double bsm[];
  constexpr int NACC = DP + 2;                     // length scales | sf | rq shape
  constexpr int S = DP + 2;                        // shared row: DP coordinates, alpha_j, pad (16-byte rows)
  constexpr int PF = 6;                            // Ainv loads in flight per warp
  const int slot = a.sel[blockIdx.y];
  int ti, tj;
  tri_decode(blockIdx.x, ti, tj);
  const int D = a.D, Np = a.Np, N = a.N;
  double* xcs = bsm;                               // [128][S]
  double* red = bsm + T * S;                       // [8][NACC]
  const double* xs = a.xs + (long long)slot * D * Np;
  const double* alpha = a.alpha + (long long)slot * Np;
  for (int e = threadIdx.x; e < DP * T; e += blockDim.x) {
    const int k = e / T, j = e % T;
    xcs[j * S + k] = (k < D) ? xs[(long long)k * Np + tj * T + j] : 0.0;
  }
  if (threadIdx.x < T) {
    xcs[threadIdx.x * S + DP] = alpha[tj * T + threadIdx.x];
    xcs[threadIdx.x * S + DP + 1] = 0.0;
  }
  const int il = threadIdx.x & (T - 1), half = threadIdx.x >> 7;
  const int gi = ti * T + il;
  double xr[DP];
#pragma unroll
  for (int k = 0; k < DP; ++k) xr[k] = (k < D) ? xs[(long long)k * Np + gi] : 0.0;
  const double ar = alpha[gi];
  const SlotP p = a.sp[slot];
  const double inv_sl = 1.0 / p.sl;
  const double half_over_a = 0.5 / p.rq_a;
  __syncthreads();

  // columns this warp visits: its half of the tile, cut at N and (diagonal tile) at its last row
  const bool diag = (ti == tj);
  const int j0 = half * 64;
  int jend = min(64, N - tj * T - j0);
  if (diag) jend = min(jend, (il | 31) + 1 - j0);
  const bool rowok = gi < N;
  const double* Acol = a.Abuf + slot * a.smat + (long long)(tj * T + j0) * Np + gi;

  double acc[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q] = 0.0;

  auto pair = [&](const int jl, const double av) {
    const double2* xc2 = reinterpret_cast<const double2*>(xcs + jl * S);
    double d2[DP];
#pragma unroll
    for (int k = 0; k < DP; k += 2) {
      const double2 c = xc2[k / 2];
      const double da = xr[k] - c.x, db = xr[k + 1] - c.y;
      d2[k] = da * da;
      d2[k + 1] = db * db;
    }
    const double ac = xcs[jl * S + DP];
    double r2 = d2[0];
#pragma unroll
    for (int k = 1; k < DP; ++k) r2 += d2[k];
    double kf, cf, sh;
    radial_factors<KIND>(r2, p.rq_a, half_over_a, kf, cf, sh);
    double Q = fma(av, inv_sl, -(ar * ac));                         // gaussian_process.py:2477-2484
    if (diag && il == jl) Q *= 0.5;
    const bool valid = rowok && (!diag || il >= jl);
    Q = valid ? Q : 0.0;
    acc[NACC - 2] = fma(Q, kf, acc[NACC - 2]);
    if (KIND == 2) acc[NACC - 1] = fma(Q, sh, acc[NACC - 1]);
    double cw = Q * cf;
    if (KIND == 1) cw = valid ? cw : 0.0;                           // cf = inf on masked r = 0 pairs
#pragma unroll
    for (int k = 0; k < DP; ++k) acc[k] = fma(cw, d2[k], acc[k]);
  };
  // The elements of Ainv arrive one 256-byte row segment per warp and column, PF loads in flight per warp.
  // The column loop is unrolled PF times so that every slot of the ring is a fixed register that is
  // re-loaded right after it has been consumed: a rotating ring (or a single look-ahead element) is
  // implemented with register moves FROM the load's destination, which wait for the load -- ncu showed
  // 32 % of all stall samples on that move and the same run time with or without the look-ahead.
  double ring[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) ring[u] = (jend > 0) ? Acol[(long long)u * Np] : 0.0;
  for (int jj = 0; jj < jend; jj += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const double av = ring[u];
      ring[u] = Acol[(long long)min(jj + u + PF, 63) * Np];          // unconditional: the slot IS the destination
      if (jj + u < jend) pair(j0 + jj + u, av);                      // warp-uniform
    }
  }
  // deterministic reduction: warp xor-tree, then warps in order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NACC + q] = v;
  }
  __syncthreads();
  if (threadIdx.x < a.cov_n) {
    const int pidx = threadIdx.x;
    const int q = (pidx < D) ? pidx : (pidx == D ? NACC - 2 : NACC - 1);
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w * NACC + q];
    s *= (pidx == D) ? 2 * p.sf2 : p.sf2;                           // dK/dlog(sf) = 2K
    const long long ntiles = (long long)a.Nt * (a.Nt + 1) / 2;
    a.gpart[((long long)slot * ntiles + blockIdx.x) * a.cov_n + pidx] = s;
  }
}

struct GradFinalArgs {
  Model md;
  int N, Np, Nt;
  const int* sel;
  const double* X; const double* y; const double* s2;
  const double* hyp;       // [nslots][P]
  const SlotP* sp;
  const double* Abuf;      // diag of Ainv
  const double* alpha;
  const double* gpart;
  double* dnlZ;            // [nslots][P]
  long long smat;
};

__global__ void __launch_bounds__(256) grad_final_kernel(GradFinalArgs a) {
  __shared__ double sh[256];
  const int slot = a.sel[blockIdx.x];
  const Model& md = a.md;
  const int D = md.D, N = a.N, Np = a.Np;
  const double* h = a.hyp + (long long)slot * md.P;
  const double* hn = h + md.cov_n;
  const double* hm = h + md.cov_n + md.noise_n;
  const SlotP p = a.sp[slot];
  const double* Ainv = a.Abuf + slot * a.smat;
  const double* alpha = a.alpha + (long long)slot * Np;
  double* out = a.dnlZ + (long long)slot * md.P;
  const long long ntiles = (long long)a.Nt * (a.Nt + 1) / 2;

  // blockIdx.y = task: 0 -> covariance hyperparameters, then one CTA per noise / mean
  // hyperparameter (they are independent sums over the N points)
  const int task = blockIdx.y;
  // covariance hyperparameters: ordered sum of the tile partials
  if (task == 0) {
    if ((int)threadIdx.x < md.cov_n) {
      const double* gp = a.gpart + (long long)slot * ntiles * md.cov_n + threadIdx.x;
      double s = 0.0;
      for (long long t = 0; t < ntiles; ++t) s += gp[t * md.cov_n];
      out[threadIdx.x] = s;
    }
    return;
  }
  // noise hyperparameters: 0.5 * sn2_mult * sum_i dsn2_iq * Q_ii   (:2491-2504)
  if (task <= md.noise_n) {
    const int q = task - 1;
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double Qii = Ainv[(long long)i * Np + i] / p.sl - alpha[i] * alpha[i];
      // derivative of sn2_i w.r.t. noise hyperparameter q (noise_functions.py:253-277)
      int idx = 0;
      double d = 0.0;
      if (md.nz0 == 1) { if (q == idx) d = 2 * exp(2 * hn[idx]); ++idx; }
      if (md.nz1 == 2) { if (q == idx) d = exp(hn[idx]) * (a.s2 ? a.s2[i] : 0.0); ++idx; }
      if (md.nz2 == 1) {
        const double thr = hn[idx], w2 = exp(2 * hn[idx + 1]);
        const double zz = fmax(0.0, thr - a.y[i]);
        if (q == idx) d = 2 * w2 * (thr - a.y[i]) * (zz > 0 ? 1.0 : 0.0);
        if (q == idx + 1) d = 2 * w2 * (zz * zz);
      }
      s += d * Qii;
    }
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[md.cov_n + q] = 0.5 * p.mult * s;
    return;
  }
  // mean hyperparameters: -dm^T alpha   (:2507-2508; mean_functions.py:258, :390-395)
  {
    const int q = task - 1 - md.noise_n;
    if (q >= md.mean_n) return;
    const int k = (q == 0) ? 0 : (q - 1) % D;
    const double om = (q == 0) ? 1.0 : exp(hm[1 + D + k]);
    const double xm = (q == 0) ? 0.0 : hm[1 + k];
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      double d;
      if (q == 0) d = 1.0;
      else {
        const double diff = a.X[(long long)i * D + k] - xm;
        if (q <= D) d = diff / (om * om);
        else { const double z = diff / om; d = z * z; }
      }
      s += d * alpha[i];
    }
    s = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[md.cov_n + md.noise_n + q] = -s;
  }
}

// ---------------------------------------------------------------------------------
// Plugin surface kernels (one hyperparameter vector, un-padded, row-major outputs)
// ---------------------------------------------------------------------------------
struct CovArgs {
  int cov_kind, degree, ard, D;
  long long N, M;
  const double* hyp;     // device, cov_n
  const double* X;       // (N,D)
  const double* Xs;      // (M,D) or null -> pairwise on X
  int diag;
  double* K;             // (N,M) / (N,N) / (N)
  double* dK;            // (cov_n,N,N) or null
};

template <int KIND>
__global__ void __launch_bounds__(256) cov_plugin_kernel(CovArgs a) {
  const int D = a.D;
  const int nl = a.ard ? D : 1;
  const double sf2 = exp(2 * a.hyp[nl]);
  const double a_rq = (a.cov_kind == 2) ? exp(a.hyp[nl + 1]) : 1.0;
  const long long cols = a.diag ? 1 : (a.Xs ? a.M : a.N);
  const long long total = a.N * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / cols, j = e % cols;
    if (a.diag) {
      a.K[i] = kern_value<KIND>(0.0, sf2, a_rq);     // tmp = zeros((N,1)), covariance_functions.py:163
      continue;
    }
    const double* xi = a.X + i * D;
    const double* xj = (a.Xs ? a.Xs : a.X) + j * D;
    double r2 = 0.0;
    for (int k = 0; k < D; ++k) {
      const double ell = exp(a.hyp[a.ard ? k : 0]);
      const double d = scale_coord(a.cov_kind, a.ard, a.degree, xi[k], ell) -
                       scale_coord(a.cov_kind, a.ard, a.degree, xj[k], ell);
      r2 = __dadd_rn(r2, __dmul_rn(d, d));
    }
    if (!a.dK) {
      a.K[e] = kern_value<KIND>(r2, sf2, a_rq);
      continue;
    }
    double K, c, dshape;
    kern_value_grad<KIND>(r2, sf2, a_rq, K, c, dshape);
    a.K[e] = K;
    const long long nn = a.N * a.N;
    if (a.ard) {
      for (int k = 0; k < D; ++k) {
        const double ell = exp(a.hyp[k]);
        const double d = scale_coord(a.cov_kind, 1, a.degree, xi[k], ell) -
                         scale_coord(a.cov_kind, 1, a.degree, xj[k], ell);
        a.dK[k * nn + e] = c * (d * d);
      }
    } else {
      a.dK[e] = c * r2;
    }
    a.dK[nl * nn + e] = 2 * K;
    if (KIND == 2) a.dK[(nl + 1) * nn + e] = dshape;
  }
}

struct MeanArgs {
  int mean_kind, D;
  long long N;
  const double* hyp; const double* X;
  double* m; double* dm;     // dm (N,mean_n) row-major or null
};

__global__ void __launch_bounds__(256) mean_plugin_kernel(MeanArgs a) {
  const int D = a.D;
  const int mean_n = mean_count(a.mean_kind, D);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N;
       i += (long long)gridDim.x * blockDim.x) {
    const double* x = a.X + i * D;
    a.m[i] = mean_value(a.mean_kind, D, a.hyp, x);
    if (a.dm && mean_n > 0) {
      double* d = a.dm + i * mean_n;
      d[0] = 1.0;
      if (a.mean_kind == 2) {
        for (int k = 0; k < D; ++k) {
          const double om = exp(a.hyp[1 + D + k]);
          const double diff = x[k] - a.hyp[1 + k];
          d[1 + k] = diff / (om * om);
          const double z = diff / om;
          d[1 + D + k] = z * z;
        }
      }
    }
  }
}

struct NoiseArgs {
  int nz0, nz1, nz2;
  long long N;
  const double* hyp; const double* y; const double* s2;
  double* sn2; double* dsn2;   // dsn2 (N,noise_n) row-major or null
};

__global__ void __launch_bounds__(256) noise_plugin_kernel(NoiseArgs a) {
  const int nn = noise_count(a.nz0, a.nz1, a.nz2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.N;
       i += (long long)gridDim.x * blockDim.x) {
    const double yi = a.y ? a.y[i] : 0.0;
    const double s2i = a.s2 ? a.s2[i] : 0.0;
    a.sn2[i] = noise_value(a.nz0, a.nz1, a.nz2, a.hyp, a.y != nullptr, yi, a.s2 != nullptr, s2i);
    if (a.dsn2 && nn > 0) {
      double* d = a.dsn2 + i * nn;
      for (int q = 0; q < nn; ++q) d[q] = 0.0;
      int idx = 0;
      if (a.nz0 == 1) { d[idx] = 2 * exp(2 * a.hyp[idx]); ++idx; }
      if (a.nz1 == 2) { d[idx] = exp(a.hyp[idx]) * s2i; ++idx; }
      if (a.nz2 == 1 && a.y) {
        const double thr = a.hyp[idx], w2 = exp(2 * a.hyp[idx + 1]);
        const double zz = fmax(0.0, thr - yi);
        d[idx] = 2 * w2 * (thr - yi) * (zz > 0 ? 1.0 : 0.0);
        d[idx + 1] = 2 * w2 * (zz * zz);
      }
    }
  }
}

}  // namespace gpb
