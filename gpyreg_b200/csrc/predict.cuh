// K4: batched prediction (reference: GP.predict, gaussian_process.py:1663-1816).
// Per hyperparameter sample and chunk of test points:
//   ks_build_kernel   Bt(j,k) = sW * k(x_k, x*_j)  (test-point-contiguous) and the partial
//                     sums of  Ks^T alpha  per training tile
//   tile GEMM OpPred  partial  sum_m (W Bt^T)(m,j)^2            (gemm.cuh)
//   pred_finish       mu, s2 = kss - sum, clamp, noise, lpd for this sample
//   pred_combine      separate-samples layout or the across-sample average (:1793-1811)
#pragma once
#include "common.cuh"
#include "cov.cuh"

namespace gpb {

struct KsArgs {
  Model md;
  int N, Np, Nt;
  int mc, Mcp;                 // valid test points in this chunk, padded to 128
  const double* Xs;            // (mc, D) row-major, this chunk
  const double* hyp;           // this sample's hyperparameter row
  const double* xs;            // this sample's pre-scaled training inputs [D][Np]
  const double* alpha;         // [Np]
  SlotP sp;
  double scale;                // sW (L_chol) or 1
  double* Bt;                  // (Mcp, Np) column-major
  double* mupart;              // [Nt][Mcp]
};

// tile: rows = test points j (tx + 32a), columns = training points k (ty*16 + b)
template <int KIND>
__global__ void __launch_bounds__(256) ks_build_kernel(KsArgs a) {
  extern __shared__ double bsm[];
  const Model& md = a.md;
  const int D = md.D, Np = a.Np;
  const int jt = blockIdx.x, kt = blockIdx.y;
  double* xr = bsm;               // [D][128] scaled test points
  double* xc = bsm + D * T;       // [D][128] scaled training points
  double* al = bsm + 2 * D * T;   // [128]
  double* red = al + T;           // [8][128]
  for (int e = threadIdx.x; e < D * T; e += blockDim.x) {
    const int k = e / T, i = e % T;
    const int j = jt * T + i;
    double v = 0.0;
    if (j < a.mc) {
      const double ell = exp(a.hyp[md.ard ? k : 0]);
      v = scale_coord(md.cov_kind, md.ard, md.degree, a.Xs[(long long)j * D + k], ell);
    }
    xr[e] = v;
    xc[e] = a.xs[(long long)k * Np + kt * T + i];
  }
  if (threadIdx.x < T) al[threadIdx.x] = a.alpha[kt * T + threadIdx.x];
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double musum[4] = {0.0, 0.0, 0.0, 0.0};
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
          const double d = xb[bb] - xa[aa];       // cdist(X, X_star): train minus test
          r2[aa][bb] = __dadd_rn(r2[aa][bb], __dmul_rn(d, d));
        }
    }
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int kl = ty * 16 + b0 + bb;
      const int gk = kt * T + kl;
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        const int jl = tx + 32 * aa;
        const int gj = jt * T + jl;
        double K = 0.0;
        if (gk < a.N && gj < a.mc) K = kern_value<KIND>(r2[aa][bb], a.sp.sf2, a.sp.rq_a);
        a.Bt[(long long)gk * a.Mcp + gj] = a.scale * K;
        musum[aa] += K * al[kl];
      }
    }
  }
#pragma unroll
  for (int aa = 0; aa < 4; ++aa) red[ty * T + tx + 32 * aa] = musum[aa];
  __syncthreads();
  if (threadIdx.x < T) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w * T + threadIdx.x];
    a.mupart[(long long)kt * a.Mcp + jt * T + threadIdx.x] = s;
  }
}

struct FinishArgs {
  Model md;
  int Nt, nv, mc, Mcp;         // nv = number of variance partials (Nt * column halves)
  int has_data;                // 0: GP without training data -> prior mean / variance
  int lchol;
  const double* Xs; const double* ys; const double* s2s;    // chunk pointers (ys/s2s may be null)
  const double* hyp;
  SlotP sp;
  const double* mupart; const double* vpart;
  int need_ys2, want_lpd;
  double* mu_s; double* s2_s; double* ys2_s; double* lpd_s;   // this sample's rows, [Mcp]
};

__global__ void __launch_bounds__(256) pred_finish_kernel(FinishArgs a) {
  const Model& md = a.md;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.mc) return;
  const double* hn = a.hyp + md.cov_n;
  const double* hm = a.hyp + md.cov_n + md.noise_n;
  double mu = mean_value(md.mean_kind, md.D, hm, a.Xs + (long long)j * md.D);   // :1734-1739
  // kss = covariance.compute(hyp, x_star, compute_diag=True) = sf2 * f(0) * exp(0)   (:1741)
  const int kc = kind_code(md.cov_kind, md.degree);
  double kss;
  if (kc == 2) kss = kern_value<2>(0.0, a.sp.sf2, a.sp.rq_a);
  else if (kc == 0) kss = kern_value<0>(0.0, a.sp.sf2, a.sp.rq_a);
  else kss = kern_value<3>(0.0, a.sp.sf2, a.sp.rq_a);
  double s2 = kss;
  if (a.has_data) {
    double m = 0.0, v = 0.0;
    for (int t = 0; t < a.Nt; ++t) m += a.mupart[(long long)t * a.Mcp + j];
    for (int t = 0; t < a.nv; ++t) v += a.vpart[(long long)t * a.Mcp + j];
    mu += m;                                        // :1747
    s2 = kss - v;                                   // :1758 / :1762 (L = -Ainv)
  }
  s2 = fmax(s2, 0.0);                               // :1770
  a.mu_s[j] = mu;
  a.s2_s[j] = s2;
  if (a.need_ys2) {
    const double sn2 = noise_value(md.nz0, md.nz1, md.nz2, hn, a.ys != nullptr,
                                   a.ys ? a.ys[j] : 0.0, a.s2s != nullptr,
                                   a.s2s ? a.s2s[j] : 0.0);
    const double ys2 = s2 + sn2 * a.sp.mult;        // :1779
    a.ys2_s[j] = ys2;
    if (a.want_lpd) {
      const double dlt = a.ys[j] - mu;
      a.lpd_s[j] = -0.5 * (dlt * dlt) / ys2 - 0.5 * log(2 * M_PI * ys2);   // :1783-1787
    }
  }
}

struct CombineArgs {
  int Ns, mc, Mcp;
  int add_noise, separate, want_lpd;
  const double* ys;
  const double* mu_s; const double* s2_s; const double* ys2_s; const double* lpd_s;   // [Ns][Mcp]
  double* mu; double* s2; double* lpd;        // chunk outputs: (mc) or (mc, Ns) row-major
};

__global__ void __launch_bounds__(256) pred_combine_kernel(CombineArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.mc) return;
  const int Ns = a.Ns;
  const double* var_s = a.add_noise ? a.ys2_s : a.s2_s;      // :1789-1790
  if (a.separate) {
    for (int s = 0; s < Ns; ++s) {
      a.mu[(long long)j * Ns + s] = a.mu_s[(long long)s * a.Mcp + j];
      a.s2[(long long)j * Ns + s] = var_s[(long long)s * a.Mcp + j];
      if (a.want_lpd) a.lpd[(long long)j * Ns + s] = a.lpd_s[(long long)s * a.Mcp + j];
    }
    return;
  }
  double mu, s2, v = 0.0;
  if (Ns > 1) {                                               // :1794-1798
    double sm = 0.0, sv = 0.0;
    for (int s = 0; s < Ns; ++s) {
      sm += a.mu_s[(long long)s * a.Mcp + j];
      sv += var_s[(long long)s * a.Mcp + j];
    }
    const double mbar = sm / Ns;
    for (int s = 0; s < Ns; ++s) {
      const double d = a.mu_s[(long long)s * a.Mcp + j] - mbar;
      v += d * d;
    }
    v /= (Ns - 1);
    mu = mbar;
    s2 = sv / Ns + v;
  } else {
    mu = a.mu_s[j];
    s2 = var_s[j];
  }
  a.mu[j] = mu;
  a.s2[j] = s2;
  if (a.want_lpd) {
    double pv = s2;                                           // :1803-1806
    if (!a.add_noise) {                                       // :1807-1811
      double sy = 0.0;
      for (int s = 0; s < Ns; ++s) sy += a.ys2_s[(long long)s * a.Mcp + j];
      pv = sy / Ns + v;
    }
    const double dlt = a.ys[j] - mu;
    a.lpd[j] = -0.5 * (dlt * dlt) / pv - 0.5 * log(2 * M_PI * pv);
  }
}

}  // namespace gpb
