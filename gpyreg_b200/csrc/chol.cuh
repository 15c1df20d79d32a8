// Panel-level kernels of the batched blocked Cholesky / triangular solves.
//   diag_kernel       factor one 128x128 diagonal tile in shared memory, invert it, write
//                     D_k, D_k^T, log-det partial, and the forward-solve block z_k = D_k b_k
//   fwd_update_kernel b_i -= L_ik z_k                     (forward substitution, i > k)
//   bwd_step_kernel   w_i = D_i^T b_i ; b_j -= L_ij^T w_i (backward substitution, j < i)
//   nlz_kernel        nlZ = z^T z/(2 sl) + sum log L_ii + N log(2 pi sl)/2
// The dense O(N^3) work between these is the tile GEMM in gemm.cuh.
#pragma once
#include "common.cuh"

namespace gpb {

constexpr int DP_PITCH = T + 1;                                   // 129
constexpr size_t DIAG_SMEM = (size_t)T * DP_PITCH * sizeof(double) + 5 * T * sizeof(double);

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
};

// S(r,c) lives at S[c*129 + r]: column-major, conflict-free along r.
__global__ void __launch_bounds__(256) diag_kernel(DiagArgs a) {
  extern __shared__ double dsm[];
  double* S = dsm;                       // [T][129]
  double* dinv = dsm + T * DP_PITCH;     // 1 / L_jj
  double* bsh = dinv + T;                // b_k
  double* lg = bsh + T;                  // log L_jj
  double* cj = lg + T;                   // scaled pivot column
  double* dd = cj + T;                   // diagonal of L
  const int slot = a.sel[blockIdx.x];
  const int k = a.k, Np = a.Np;
  const int tid = threadIdx.x;
  double* A = a.Abuf + slot * a.smat + (long long)k * T + (long long)k * T * Np;
  const int nact = min(T, a.N - k * T);  // rows/cols of this tile that hold data (rest: identity)

  // load the lower triangle of the tile
  for (int e = tid; e < T * T; e += 256) {
    const int r = e & (T - 1), c = e >> 7;
    S[c * DP_PITCH + r] = (r >= c) ? A[(long long)c * Np + r] : 0.0;
  }
  if (tid < T) bsh[tid] = a.bvec ? a.bvec[(long long)slot * Np + k * T + tid] : 0.0;
  __syncthreads();

  // ---- right-looking Cholesky of the active block, all 256 threads on the rank-1 update.
  // The scaled pivot column lives in its own array (cj) so the update loop's loads do not
  // alias its stores and can be software-pipelined; the diagonal goes to dd[].
  int failed = 0;
  const int r = tid & (T - 1), half = tid >> 7;
  for (int j = 0; j < nact; ++j) {
    double piv = S[j * DP_PITCH + j];
    if (!(piv > 0.0)) { failed = 1; piv = 1.0; }     // LAPACK dpotrf: ajj <= 0 or NaN -> info > 0
    const double d = sqrt(piv);
    if (half == 0) {
      if (r == j) dd[j] = d;
      else if (r > j && r < nact) {
        const double l = S[j * DP_PITCH + r] / d;
        S[j * DP_PITCH + r] = l;
        cj[r] = l;
      }
    }
    __syncthreads();
    if (r > j && r < nact) {
      const double lrj = cj[r];
      double* Sr = S + r;
      int c = j + 1 + half;
      for (; c + 6 <= r; c += 8) {
        const double u0 = cj[c], u1 = cj[c + 2], u2 = cj[c + 4], u3 = cj[c + 6];
        double s0 = Sr[c * DP_PITCH], s1 = Sr[(c + 2) * DP_PITCH], s2 = Sr[(c + 4) * DP_PITCH],
               s3 = Sr[(c + 6) * DP_PITCH];
        s0 -= lrj * u0; s1 -= lrj * u1; s2 -= lrj * u2; s3 -= lrj * u3;
        Sr[c * DP_PITCH] = s0; Sr[(c + 2) * DP_PITCH] = s1; Sr[(c + 4) * DP_PITCH] = s2;
        Sr[(c + 6) * DP_PITCH] = s3;
      }
      for (; c <= r; c += 2) Sr[c * DP_PITCH] -= lrj * cj[c];
    }
    __syncthreads();
  }
  if (tid < nact) S[tid * DP_PITCH + tid] = dd[tid];
  __syncthreads();
  if (tid < T) {
    const double ljj = S[tid * DP_PITCH + tid];      // 1.0 in the padded part
    dinv[tid] = 1.0 / ljj;
    lg[tid] = (tid < nact) ? log(ljj) : 0.0;
  }
  __syncthreads();
  // write L_kk back (lower triangle)
  for (int e = tid; e < T * T; e += 256) {
    const int rr = e & (T - 1), c = e >> 7;
    if (rr >= c) A[(long long)c * Np + rr] = S[c * DP_PITCH + rr];
  }
  if (tid == 0) {
    double s = 0.0;
    for (int j = 0; j < nact; ++j) s += lg[j];       // fixed order
    a.logdet[(long long)slot * a.Nt + k] = s;
    if (failed) a.fail[slot] = 1;
  }

  // ---- D = L^-1 by forward substitution, one thread per column c; D(q,c), q > c, is kept
  // transposed in the (free) strict upper triangle: S(c,q).  No barrier needed: a thread
  // only reads L (final) and its own column.
  // The loops run over (rr, q) uniformly so a warp reads L(rr,q) as a broadcast and
  // D(q,c) from consecutive banks.
  if (tid < T) {
    const int c = tid;
    for (int rr = 1; rr < nact; ++rr) {
      if (c < rr) {
        double acc0 = S[c * DP_PITCH + rr] * dinv[c];  // L(rr,c) * D(c,c)
        double acc1 = 0.0;
        int q = rr - 1;
        for (; q - 1 > c; q -= 2) {
          acc0 += S[q * DP_PITCH + rr] * S[q * DP_PITCH + c];
          acc1 += S[(q - 1) * DP_PITCH + rr] * S[(q - 1) * DP_PITCH + c];
        }
        if (q > c) acc0 += S[q * DP_PITCH + rr] * S[q * DP_PITCH + c];
        S[rr * DP_PITCH + c] = -(acc0 + acc1) * dinv[rr];
      }
    }
  }
  __syncthreads();
  // ---- outputs: D_k (column-major, zeros above the diagonal), D_k^T, W diag tile, z_k
  double* Dk = a.Dbuf + ((long long)slot * a.Nt + k) * T * T;
  double* DTk = a.DTbuf + ((long long)slot * a.Nt + k) * T * T;
  double* Wd = a.Wbuf ? a.Wbuf + slot * a.smat + (long long)k * T + (long long)k * T * Np : nullptr;
  for (int e = tid; e < T * T; e += 256) {
    const int rr = e & (T - 1), c = e >> 7;          // element (rr, c) of D
    double v;
    if (rr == c) v = dinv[c];
    else if (rr > c) v = (rr < nact) ? S[rr * DP_PITCH + c] : 0.0;
    else v = 0.0;
    Dk[c * T + rr] = v;
    if (Wd) Wd[(long long)c * Np + rr] = v;
  }
  for (int e = tid; e < T * T; e += 256) {
    const int rr = e & (T - 1), c = e >> 7;          // element (rr, c) of D^T = D(c, rr)
    double v;
    if (rr == c) v = dinv[c];
    else if (c > rr) v = (c < nact) ? S[c * DP_PITCH + rr] : 0.0;
    else v = 0.0;
    DTk[c * T + rr] = v;
  }
  if (a.zvec && tid < T) {
    const int rr = tid;
    double s = 0.0;
    if (rr < nact) {
      for (int c = 0; c < rr; ++c) s += S[rr * DP_PITCH + c] * bsh[c];
      s += dinv[rr] * bsh[rr];
    }
    a.zvec[(long long)slot * Np + k * T + rr] = s;
  }
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

// forward substitution update after step k: b_i -= L_ik z_k, i = k+1+blockIdx.x
__global__ void __launch_bounds__(T) fwd_update_kernel(VecArgs a) {
  __shared__ double z[T];
  const int slot = a.sel[blockIdx.y];
  const int i = a.k + 1 + blockIdx.x;
  z[threadIdx.x] = a.zvec[(long long)slot * a.Np + a.k * T + threadIdx.x];
  __syncthreads();
  const double* L = a.Abuf + slot * a.smat + (long long)i * T + (long long)a.k * T * a.Np;
  double s = 0.0;
#pragma unroll 8
  for (int q = 0; q < T; ++q) s += L[(long long)q * a.Np + threadIdx.x] * z[q];
  a.bvec[(long long)slot * a.Np + i * T + threadIdx.x] -= s;
}

// backward substitution, block row i = a.k (descending): every CTA recomputes
// w_i = D_i^T b_i; CTA j < i applies b_j -= L_ij^T w_i; CTA 0 stores alpha_i = w_i / sl.
__global__ void __launch_bounds__(T) bwd_step_kernel(VecArgs a) {
  __shared__ double w[T];
  __shared__ double bi[T];
  const int slot = a.sel[blockIdx.y];
  const int i = a.k, j = blockIdx.x;
  const int tid = threadIdx.x;
  bi[tid] = a.bvec[(long long)slot * a.Np + i * T + tid];
  __syncthreads();
  const double* DT = a.DTbuf + ((long long)slot * a.Nt + i) * T * T;
  double s = 0.0;
#pragma unroll 8
  for (int m = 0; m < T; ++m) s += DT[m * T + tid] * bi[m];       // (D^T b)(n) = sum_m DT(n,m) b(m)
  w[tid] = s;
  if (j == 0) a.alpha[(long long)slot * a.Np + i * T + tid] = s / a.sp[slot].sl;   // :2455-2465
  __syncthreads();
  if (j >= i) return;
  // (L_ij^T w)(n) = sum_m L(iT+m, jT+n) w(m): one warp per group of columns, lanes over m
  const double* L = a.Abuf + slot * a.smat + (long long)i * T + (long long)j * T * a.Np;
  const int lane = tid & 31, warp = tid >> 5;
  for (int n = warp; n < T; n += 4) {
    const double* col = L + (long long)n * a.Np;
    double v = col[lane] * w[lane] + col[lane + 32] * w[lane + 32] + col[lane + 64] * w[lane + 64] +
               col[lane + 96] * w[lane + 96];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.bvec[(long long)slot * a.Np + j * T + n] -= v;
  }
}

struct NlzArgs {
  const int* sel;
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
    for (int k = 0; k < a.Nt; ++k) ld += a.logdet[(long long)slot * a.Nt + k];
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
__global__ void scale_mult_kernel(double* mult, const int* fail, const int* sel, int nsel) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nsel) { const int s = sel[t]; if (fail[s]) mult[s] *= 10.0; }
}

}  // namespace gpb
