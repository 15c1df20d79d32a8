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
  // blockIdx.z = posterior sample within the group; all per-sample arrays are strided
  const double* hyp;           // hyperparameter rows, stride md.P
  const double* xs;            // pre-scaled training inputs [D][Np] per sample, stride D*Np
  const double* alpha;         // [Np] per sample, stride Np
  const SlotP* sp;             // per-sample scalars
  double* Bt; long long sBt;   // (Mcp, Np) column-major per sample:  sW * Ks  (L_chol) or Ks
  double* mupart; long long smu;   // [Nt][Mcp] per sample
};

// tile: rows = test points j (tx + 32a), columns = training points k (ty*16 + b)
template <int KIND>
__global__ void __launch_bounds__(256) ks_build_kernel(KsArgs a) {
  extern __shared__ double bsm[];
  const Model& md = a.md;
  const int D = md.D, Np = a.Np;
  const int jt = blockIdx.x, kt = blockIdx.y, z = blockIdx.z;
  const double* hyp = a.hyp + (long long)z * md.P;
  const double* xs = a.xs + (long long)z * D * Np;
  const double* alpha = a.alpha + (long long)z * Np;
  const SlotP sp = a.sp[z];
  const double scale = sp.lchol ? 1.0 / sqrt(sp.sn2_min * sp.mult) : 1.0;   // sW, gaussian_process.py:2517
  double* Bt = a.Bt + z * a.sBt;
  double* mupart = a.mupart + z * a.smu;
  double* xr = bsm;               // [D][128] scaled test points
  double* xc = bsm + D * T;       // [D][128] scaled training points
  double* al = bsm + 2 * D * T;   // [128]
  double* red = al + T;           // [8][128]
  for (int e = threadIdx.x; e < D * T; e += blockDim.x) {
    const int k = e / T, i = e % T;
    const int j = jt * T + i;
    double v = 0.0;
    if (j < a.mc) {
      const double ell = exp(hyp[md.ard ? k : 0]);
      v = scale_coord(md.cov_kind, md.ard, md.degree, a.Xs[(long long)j * D + k], ell);
    }
    xr[e] = v;
    xc[e] = xs[(long long)k * Np + kt * T + i];
  }
  if (threadIdx.x < T) al[threadIdx.x] = alpha[kt * T + threadIdx.x];
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
        if (gk < a.N && gj < a.mc) K = kern_value<KIND>(r2[aa][bb], sp.sf2, sp.rq_a);
        Bt[(long long)gk * a.Mcp + gj] = scale * K;
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
    mupart[(long long)kt * a.Mcp + jt * T + threadIdx.x] = s;
  }
}

struct FinishArgs {
  Model md;
  int Nt, nv, mc, Mcp;         // nv = number of variance partials (Nt * column halves)
  int has_data;                // 0: GP without training data -> prior mean / variance
  const double* Xs; const double* ys; const double* s2s;    // chunk pointers (ys/s2s may be null)
  // blockIdx.y = posterior sample within the group
  const double* hyp;           // stride md.P
  const SlotP* sp;
  const double* mupart; long long smu;
  const double* vpart; long long sv;
  int need_ys2, want_lpd;
  double* mu_s; double* s2_s; double* ys2_s; double* lpd_s;   // [samples][Mcp], first sample of the group
};

__global__ void __launch_bounds__(256) pred_finish_kernel(FinishArgs a) {
  const Model& md = a.md;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.mc) return;
  const int z = blockIdx.y;
  const double* hyp = a.hyp + (long long)z * md.P;
  const SlotP sp = a.sp[z];
  const double* mupart = a.mupart + z * a.smu;
  const double* vpart = a.vpart + z * a.sv;
  const double* hn = hyp + md.cov_n;
  const double* hm = hyp + md.cov_n + md.noise_n;
  double mu = mean_value(md.mean_kind, md.D, hm, a.Xs + (long long)j * md.D);   // :1734-1739
  // kss = covariance.compute(hyp, x_star, compute_diag=True) = sf2 * f(0) * exp(0)   (:1741)
  const int kc = kind_code(md.cov_kind, md.degree);
  double kss;
  if (kc == 2) kss = kern_value<2>(0.0, sp.sf2, sp.rq_a);
  else if (kc == 0) kss = kern_value<0>(0.0, sp.sf2, sp.rq_a);
  else kss = kern_value<3>(0.0, sp.sf2, sp.rq_a);
  double s2 = kss;
  if (a.has_data) {
    double m = 0.0, v = 0.0;
    for (int t = 0; t < a.Nt; ++t) m += mupart[(long long)t * a.Mcp + j];
    for (int t = 0; t < a.nv; ++t) v += vpart[(long long)t * a.Mcp + j];
    mu += m;                                        // :1747
    s2 = kss - v;                                   // :1758 / :1762 (L = -Ainv)
  }
  s2 = fmax(s2, 0.0);                               // :1770
  const long long o = (long long)z * a.Mcp + j;
  a.mu_s[o] = mu;
  a.s2_s[o] = s2;
  if (a.need_ys2) {
    const double sn2 = noise_value(md.nz0, md.nz1, md.nz2, hn, a.ys != nullptr,
                                   a.ys ? a.ys[j] : 0.0, a.s2s != nullptr,
                                   a.s2s ? a.s2s[j] : 0.0);
    const double ys2 = s2 + sn2 * sp.mult;          // :1779
    a.ys2_s[o] = ys2;
    if (a.want_lpd) {
      const double dlt = a.ys[j] - mu;
      a.lpd_s[o] = -0.5 * (dlt * dlt) / ys2 - 0.5 * log(2 * M_PI * ys2);   // :1783-1787
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

// ---------------------------------------------------------------------------------
// Bayesian quadrature against Gaussian measures N(mu_j, diag(sigma_j^2)), SE-ARD kernel only
// (reference: GP.quad, gaussian_process.py:1818-1981).  Same structure as predict with the
// cross-covariance replaced by  z(j,i) = exp(lnnf_j - 1/2 sum_k ((mu_jk - x_ik)/tau_jk)^2),
// tau_jk = sqrt(sigma_jk^2 + ell_k^2)  (:1925-1937).
// ---------------------------------------------------------------------------------
struct QuadArgs {
  Model md;
  int N, Np, Nt, mc, Mcp;
  const double* mu;            // (mc, D) chunk
  const double* sigma;         // (mc, D) chunk
  const double* X;             // (N, D) training inputs, row-major
  // blockIdx.z = posterior sample within the group; per-sample arrays are strided
  const double* hyp;           // stride md.P
  const double* alpha;         // [Np] per sample
  const SlotP* sp;
  double* Bt; long long sBt;   // (Mcp, Np) column-major per sample
  double* mupart; long long smu;   // [Nt][Mcp] per sample
};

__global__ void __launch_bounds__(256) quad_build_kernel(QuadArgs a) {
  extern __shared__ double bsm[];
  const int D = a.md.D;
  const int jt = blockIdx.x, kt = blockIdx.y, z = blockIdx.z;
  const double* hyp = a.hyp + (long long)z * a.md.P;
  const double* alpha = a.alpha + (long long)z * a.Np;
  const SlotP sp = a.sp[z];
  // 1/sqrt(sn2_eff), sn2_eff = exp(2 hyp[cov_N]) * sn2_mult  (gaussian_process.py:1919-1920)
  const double scale = sp.lchol ? 1.0 / sqrt(exp(2 * hyp[a.md.cov_n]) * sp.mult) : 1.0;
  double* Bt = a.Bt + z * a.sBt;
  double* mupart = a.mupart + z * a.smu;
  double* mr = bsm;                 // [D][128] measure means
  double* it = bsm + D * T;         // [D][128] 1 / tau
  double* xc = bsm + 2 * D * T;     // [D][128] training inputs
  double* lnnf = bsm + 3 * D * T;   // [128]
  double* al = lnnf + T;            // [128]
  double* red = al + T;             // [8][128]
  for (int e = threadIdx.x; e < D * T; e += blockDim.x) {
    const int k = e / T, i = e % T;
    const int j = jt * T + i, gi = kt * T + i;
    double m = 0.0, itau = 0.0;
    if (j < a.mc) {
      const double sg = a.sigma[(long long)j * D + k], ell = exp(hyp[a.md.ard ? k : 0]);
      m = a.mu[(long long)j * D + k];
      itau = 1.0 / sqrt(sg * sg + ell * ell);
    }
    mr[e] = m;
    it[e] = itau;
    xc[e] = (gi < a.N) ? a.X[(long long)gi * D + k] : 0.0;
  }
  if (threadIdx.x < T) {
    const int j = jt * T + threadIdx.x;
    double v = 0.0;
    if (j < a.mc) {
      double sl = 0.0, st = 0.0;            // sum log ell, sum log tau
      for (int k = 0; k < D; ++k) {
        // isotropic kernel: the one length scale serves every dimension (the reference indexes
        // hyp[0:D], which is the same thing for D = 1, the only isotropic case its tests cover)
        const double lh = hyp[a.md.ard ? k : 0];
        const double sg = a.sigma[(long long)j * D + k], ell = exp(lh);
        sl += lh;
        st += log(sqrt(sg * sg + ell * ell));
      }
      v = 2 * hyp[a.md.ard ? D : 1] + sl - st;            // :1925-1927
    }
    lnnf[threadIdx.x] = v;
    al[threadIdx.x] = alpha[kt * T + threadIdx.x];
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double fsum[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = 0; b < 16; ++b) {
    const int il = ty * 16 + b, gi = kt * T + il;
    double d2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < D; ++k) {
      const double x = xc[k * T + il];
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        const double q = (mr[k * T + tx + 32 * aa] - x) * it[k * T + tx + 32 * aa];
        d2[aa] += q * q;
      }
    }
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
      const int jl = tx + 32 * aa, gj = jt * T + jl;
      double zz = 0.0;
      if (gi < a.N && gj < a.mc) zz = exp(lnnf[jl] - 0.5 * d2[aa]);    // :1937
      Bt[(long long)gi * a.Mcp + gj] = scale * zz;
      fsum[aa] += zz * al[il];
    }
  }
#pragma unroll
  for (int aa = 0; aa < 4; ++aa) red[ty * T + tx + 32 * aa] = fsum[aa];
  __syncthreads();
  if (threadIdx.x < T) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w * T + threadIdx.x];
    mupart[(long long)kt * a.Mcp + jt * T + threadIdx.x] = s;
  }
}

struct QuadFinishArgs {
  Model md;
  int Nt, nv, mc, Mcp, compute_var;
  const double* mu; const double* sigma;
  const double* hyp;             // stride md.P, blockIdx.y = sample
  const double* mupart; long long smu;
  const double* vpart; long long sv;
  double* F_s; double* V_s;      // [samples][Mcp], first sample of the group
};

__global__ void __launch_bounds__(256) quad_finish_kernel(QuadFinishArgs a) {
  const Model& md = a.md;
  const int D = md.D;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.mc) return;
  const int z = blockIdx.y;
  const double* hyp = a.hyp + (long long)z * md.P;
  const double* mupart = a.mupart + z * a.smu;
  const double* vpart = a.vpart + z * a.sv;
  const double* hm = hyp + md.cov_n + md.noise_n;
  const double* mu = a.mu + (long long)j * D;
  const double* sg = a.sigma + (long long)j * D;
  double F = (md.mean_kind == 0) ? 0.0 : hm[0];                      // :1906-1909
  for (int t = 0; t < a.Nt; ++t) F += mupart[(long long)t * a.Mcp + j];
  if (md.mean_kind == 2) {                                           // :1939-1946
    double nu = 0.0;
    for (int k = 0; k < D; ++k) {
      const double xm = hm[1 + k], om = exp(hm[1 + D + k]);
      nu += 1.0 / (om * om) * (mu[k] * mu[k] + sg[k] * sg[k] - 2 * mu[k] * xm + xm * xm);
    }
    F += -0.5 * nu;
  }
  const long long o = (long long)z * a.Mcp + j;
  a.F_s[o] = F;
  if (a.compute_var) {
    double sl = 0.0, st = 0.0;
    for (int k = 0; k < D; ++k) {
      const double lh = hyp[md.ard ? k : 0];
      const double ell = exp(lh);
      sl += lh;
      st += log(sqrt(2 * sg[k] * sg[k] + ell * ell));
    }
    const double nf_kk = exp(2 * hyp[md.ard ? D : 1] + sl - st);     // :1949-1950
    double v = 0.0;
    for (int t = 0; t < a.nv; ++t) v += vpart[(long long)t * a.Mcp + j];
    a.V_s[o] = fmax(2.220446049250313e-16, nf_kk - v);               // :1966-1969
  }
}

// ---------------------------------------------------------------------------------
// Full predictive covariance (reference: GP.predict_full, gaussian_process.py:1561-1661):
//   cov_s = K(x*,x*) - C2, C2 = V^T V (L_chol) or Ks^T Ainv Ks (low noise), symmetrised,
//   plus the noise term of :1655-1659.
// ---------------------------------------------------------------------------------
struct FullArgs {
  Model md;
  int Nt, M, Mcp;
  const double* Xs; const double* ys; const double* s2s;
  const double* hyp;
  SlotP sp;
  const double* mupart;        // [Nt][Mcp]
  const double* C2;            // (Mcp, Mcp) column-major
  int add_noise, per_point;    // per_point: the reference's noise is an (M,1) array
  double* mu;                  // this sample's column of (M, Ns): mu[j*Ns]
  int Ns;
  double* cov;                 // this sample's (M, M) row-major block
};

template <int KIND>
__global__ void __launch_bounds__(256) full_finish_kernel(FullArgs a) {
  const Model& md = a.md;
  const int D = md.D;
  const long long total = (long long)a.M * a.M;
  const double* hn = a.hyp + md.cov_n;
  const double* hm = a.hyp + md.cov_n + md.noise_n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / a.M), j = (int)(e % a.M);
    const double* xi = a.Xs + (long long)i * D;
    const double* xj = a.Xs + (long long)j * D;
    double r2 = 0.0;
    for (int k = 0; k < D; ++k) {
      const double ell = exp(a.hyp[md.ard ? k : 0]);
      const double d = scale_coord(md.cov_kind, md.ard, md.degree, xi[k], ell) -
                       scale_coord(md.cov_kind, md.ard, md.degree, xj[k], ell);
      r2 = __dadd_rn(r2, __dmul_rn(d, d));
    }
    const double K = kern_value<KIND>(r2, a.sp.sf2, a.sp.rq_a);
    const double cij = K - a.C2[(long long)j * a.Mcp + i];
    const double cji = K - a.C2[(long long)i * a.Mcp + j];
    double v = (cij + cji) / 2;                                     // :1645
    if (a.add_noise) {
      const double sn2 = noise_value(md.nz0, md.nz1, md.nz2, hn, a.ys != nullptr,
                                     a.ys ? a.ys[i] : 0.0, a.s2s != nullptr, a.s2s ? a.s2s[i] : 0.0);
      // np.dot(np.eye(M), sn2_star) * sn2_mult: a scalar sn2 lands on the diagonal, an (M,1)
      // sn2 broadcasts over the whole row (:1655-1659)
      if (a.per_point || i == j) v += sn2 * a.sp.mult;
    }
    a.cov[e] = v;
    if (j == 0) {
      double m = mean_value(md.mean_kind, D, hm, xi);
      for (int t = 0; t < a.Nt; ++t) m += a.mupart[(long long)t * a.Mcp + i];
      a.mu[(long long)i * a.Ns] = m;
    }
  }
}

}  // namespace gpb
