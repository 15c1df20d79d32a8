// K5: rank-one posterior append (reference: GP.update rank-1 branch, gaussian_process.py:737-844).
// One new training point is appended to EVERY posterior sample of a device-resident batch:
//   append_ks_kernel     k_i = k(x_i, x_new)                       (:770-771)
//   append_gemv_n_kernel c = L^-1 k = W k          (L_chol)        (:779-781; lower L = upper^T)
//   append_gemv_t_kernel a = W^T c  (L_chol) or u = Ainv k (low noise)   (:800-807, :820)
//   append_finish_kernel sqrt_arg test, new row of L and W, alpha update  (:784-843)
//   append_outer_kernel  Ainv += u u^T / v*        (low noise)     (:821-827)
// All memory-bound: 2 passes over the lower triangle of W (or one over Ainv plus the update).
// Only models whose noise variance does not depend on the point (constant term only) take this
// path -- there sn2_eff of the new point equals the scaling sl of the stored factor and the
// reference's formulas are exact; the host layer rebuilds the batch otherwise.
#pragma once
#include "common.cuh"
#include "cov.cuh"

namespace gpb {

struct AppendArgs {
  Model md;
  int N, Np;                 // training points BEFORE the append; padded pitch
  const double* xnew;        // [D] device
  double ynew;
  const double* hyp;         // [Ns][P]
  const SlotP* sp;           // [Ns]
  double* xs;                // [Ns][D][Np] pre-scaled inputs (entry N is written)
  double* Abuf; double* Wbuf; long long smat;
  double* alpha;             // [Ns][Np]
  double* kvec; double* cvec; double* avec;   // [Ns][Np] scratch
  const double* mstar; const double* vstar;   // [Ns] predict(x_new, add_noise=True), :756-758
  int* status;               // [Ns] 1 = sqrt_arg <= 0 (:790-798), nothing written for that sample
};

template <int KIND>
__global__ void __launch_bounds__(256) append_ks_kernel(AppendArgs a) {
  __shared__ double xn[MAXD];
  const Model& md = a.md;
  const int D = md.D, Np = a.Np, s = blockIdx.y;
  const double* hyp = a.hyp + (long long)s * md.P;
  double* xs = a.xs + (long long)s * D * Np;
  if (threadIdx.x < D) {
    const double ell = exp(hyp[md.ard ? threadIdx.x : 0]);
    xn[threadIdx.x] = scale_coord(md.cov_kind, md.ard, md.degree, a.xnew[threadIdx.x], ell);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  const SlotP sp = a.sp[s];
  double v = 0.0;
  if (i < a.N) {
    double r2 = 0.0;
    for (int k = 0; k < D; ++k) {
      const double d = xs[(long long)k * Np + i] - xn[k];      // cdist(self.X, X_new)
      r2 = __dadd_rn(r2, __dmul_rn(d, d));
    }
    v = kern_value<KIND>(r2, sp.sf2, sp.rq_a);
  } else if (i == a.N) {
    for (int k = 0; k < D; ++k) xs[(long long)k * Np + i] = xn[k];
  }
  a.kvec[(long long)s * Np + i] = v;
}

// c_i = sum_{j<=i} W(i,j) k_j for the L_chol samples; rows of one 128-tile per CTA, two column
// interleaves reduced in a fixed order.
__global__ void __launch_bounds__(256) append_gemv_n_kernel(AppendArgs a) {
  __shared__ double red[2][T];
  const int s = blockIdx.y, it = blockIdx.x, Np = a.Np;
  if (!a.sp[s].lchol) return;
  const double* W = a.Wbuf + s * a.smat;
  const double* k = a.kvec + (long long)s * Np;
  const int r = threadIdx.x & (T - 1), h = threadIdx.x >> 7;
  const int i = it * T + r;
  const int jmax = min(i, a.N - 1);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < a.N) {
    int j = h;
    for (; j + 6 <= jmax; j += 8) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(W[i + (long long)(j + 2 * u) * Np], k[j + 2 * u], acc[u]);
    }
    for (; j <= jmax; j += 2) acc[0] = fma(W[i + (long long)j * Np], k[j], acc[0]);
  }
  red[h][r] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (h == 0) a.cvec[(long long)s * Np + i] = (i < a.N) ? red[0][r] + red[1][r] : 0.0;
}

// column dot products: out_j = sum_{i=lo..N-1} M(i,j) v_i ; one warp per column.
//   L_chol: M = W (lower), v = c, lo = j  -> a = W^T c
//   low noise: M = Ainv (symmetric, both triangles stored), v = k, lo = 0 -> u = Ainv k
__global__ void __launch_bounds__(256) append_gemv_t_kernel(AppendArgs a) {
  const int s = blockIdx.y, Np = a.Np;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= a.N) return;
  const int tri = a.sp[s].lchol;
  const double* M = (tri ? a.Wbuf : a.Abuf) + s * a.smat + (long long)j * Np;
  const double* v = (tri ? a.cvec : a.kvec) + (long long)s * Np;
  const int lo = tri ? j : 0;
  double acc0 = 0.0, acc1 = 0.0;
  int i = (lo & ~31) + lane;                 // aligned start so that the loads coalesce
  if (i < lo) i += 32;
  for (; i + 32 < a.N; i += 64) {
    acc0 = fma(M[i], v[i], acc0);
    acc1 = fma(M[i + 32], v[i + 32], acc1);
  }
  if (i < a.N) acc0 = fma(M[i], v[i], acc0);
  double t = acc0 + acc1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) a.avec[(long long)s * Np + j] = t;
}

__global__ void __launch_bounds__(256) append_finish_kernel(AppendArgs a) {
  __shared__ double sh[256];
  const int s = blockIdx.x, Np = a.Np, N = a.N;
  const SlotP sp = a.sp[s];
  double* A = a.Abuf + s * a.smat;
  double* alpha = a.alpha + (long long)s * Np;
  const double* av = a.avec + (long long)s * Np;
  const double m = a.mstar[s], v = a.vstar[s];
  const double coef = (m - a.ynew) / v;                                   // :838-843
  if (sp.lchol) {
    const double* c = a.cvec + (long long)s * Np;
    double* W = a.Wbuf + s * a.smat;
    double part = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) part = fma(c[i], c[i], part);
    const double cc = block_sum<256>(part, sh);
    const double sn2_eff = sp.sn2_min * sp.mult;                          // :766-767
    const double sqrt_arg = sn2_eff * sn2_eff + sp.sf2 * sn2_eff - cc;    // :786-790, K(x,x) = sf2
    if (sqrt_arg <= 0.0) {
      if (threadIdx.x == 0) a.status[s] = 1;
      return;
    }
    const double d = sqrt(sqrt_arg) / sn2_eff;                            // :812-815
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const double aj = av[j] / sn2_eff;                                  // alpha_update, :800-807
      A[N + (long long)j * Np] = c[j] / sn2_eff;                          // new row of lower L
      W[N + (long long)j * Np] = -aj / d;                                 // new row of W = L^-1
      alpha[j] = alpha[j] + coef * aj;
    }
    if (threadIdx.x == 0) {
      A[N + (long long)N * Np] = d;
      W[N + (long long)N * Np] = 1.0 / d;
      alpha[N] = 0.0 + coef * -1.0;
      a.status[s] = 0;
    }
  } else {
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const double w = -av[j] / v;                                        // -v of :821
      A[N + (long long)j * Np] = w;
      A[j + (long long)N * Np] = w;
      alpha[j] = alpha[j] + coef * av[j];
    }
    if (threadIdx.x == 0) {
      A[N + (long long)N * Np] = 1.0 / v;                                 // -(-1/v*), :825
      alpha[N] = 0.0 + coef * -1.0;
      a.status[s] = 0;
    }
  }
}

// low-noise samples: Ainv(i,j) += u_i u_j / v*   for i, j < N   (:822-823 with L = -Ainv)
__global__ void __launch_bounds__(256) append_outer_kernel(AppendArgs a) {
  const int s = blockIdx.z, Np = a.Np, N = a.N;
  if (a.sp[s].lchol) return;
  double* A = a.Abuf + s * a.smat;
  const double* u = a.avec + (long long)s * Np;
  const double v = a.vstar[s];
  const int i0 = blockIdx.x * T, j0 = blockIdx.y * T;
  const int r = threadIdx.x & (T - 1), h = threadIdx.x >> 7;
  const int i = i0 + r;
  if (i >= N) return;
  const double ui = -(u[i] / v);           // v_i = -alpha_update_i / v*, then L += v alpha_update^T
  for (int jj = h; jj < T; jj += 2) {
    const int j = j0 + jj;
    if (j < N) A[i + (long long)j * Np] -= ui * u[j];
  }
}

}  // namespace gpb
