// Shared definitions for the gpyreg_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace gpb {

constexpr int T = 128;          // tile edge of the blocked factorisation
constexpr int MAXD = 64;        // max input dimension supported by the fused kernels (staging tiles of 2*D*128
                                // doubles in shared memory, one gradient accumulator per ARD length scale in registers)

// Model descriptor: which plugin objects the GP was built from
// (reference: gaussian_process.py:43-62).
struct Model {
  int cov_kind;      // 0 SE, 1 Matern, 2 RQ
  int degree;        // Matern 1/3/5
  int ard;           // 1 ARD, 0 isotropic
  int mean_kind;     // 0 zero, 1 const, 2 negative quadratic
  int nz0, nz1, nz2; // GaussianNoise.parameters (noise_functions.py:33-41)
  int D;
  int cov_n, noise_n, mean_n, P;
};

// Per-slot scalars derived from one hyperparameter row.
struct SlotP {
  double sf2;        // exp(2*h[nl])
  double rq_a;       // RQ shape alpha = exp(h[D+1])
  double sn2_min;    // min_i sn2_i  (NaN if any NaN)
  double sl;         // sn2_div*sn2_mult (L_chol) or 1
  double mult;       // sn2_mult
  int lchol;         // np.min(sn2) >= 1e-6  (gaussian_process.py:2404)
  int pad;
};

__host__ __device__ inline int cov_count(int kind, int ard, int D) {
  if (!ard) return kind == 2 ? 3 : 2;     // isotropic RQ: log ell, log sf, log shape
  return D + (kind == 2 ? 2 : 1);
}
__host__ __device__ inline int noise_count(int nz0, int nz1, int nz2) {
  return (nz0 == 1) + (nz1 == 2) + 2 * (nz2 == 1);
}
__host__ __device__ inline int mean_count(int kind, int D) {
  return kind == 0 ? 0 : (kind == 1 ? 1 : 1 + 2 * D);
}

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; it must call pdl_wait()
// before it touches anything the predecessor writes (the wait returns when the predecessor grid has
// completed and its writes are visible).  pdl_launch() lets the NEXT kernel in the stream begin its
// launch; every kernel here calls it only after its own pdl_wait(), so whatever a successor does
// before ITS wait runs with everything up to the predecessor's predecessor complete.  Both are
// no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// lower-triangle tile enumeration: idx -> (i, j), i >= j, idx = i(i+1)/2 + j
__device__ inline void tri_decode(int idx, int& i, int& j) {
  int r = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
  while ((long long)(r + 1) * (r + 2) / 2 <= idx) ++r;
  while ((long long)r * (r + 1) / 2 > idx) --r;
  i = r;
  j = idx - r * (r + 1) / 2;
}

// ---- covariance radial functions -----------------------------------------------
// Inputs are the squared distance r2 accumulated on pre-scaled coordinates
// (covariance_functions.py:165, :251-257, :332).  KIND: 0 SE, 1/3/5 Matern, 2 RQ.
template <int KIND>
__device__ __forceinline__ double kern_value(double r2, double sf2, double a_rq) {
  if (KIND == 0) {
    return sf2 * exp(-r2 / 2);                       // covariance_functions.py:169
  } else if (KIND == 2) {
    double Mq = 1 + 0.5 * r2 / a_rq;                 // :338
    return sf2 * pow(Mq, -a_rq);                     // :339
  } else {
    double r = sqrt(r2);
    double f = (KIND == 1) ? 1.0 : (KIND == 3 ? 1 + r : 1 + r * (1 + r * (1.0 / 3)));   // :210-218
    return sf2 * f * exp(-r);                        // :259
  }
}

// radial factor c such that dK/dlog(ell_k) = c * Delta_k^2  (and K itself)
template <int KIND>
__device__ __forceinline__ void kern_value_grad(double r2, double sf2, double a_rq,
                                                double& K, double& c, double& dshape) {
  dshape = 0.0;
  if (KIND == 0) {
    K = sf2 * exp(-r2 / 2);
    c = K;                                           // :177-181  dK_k = K * Delta_k^2
  } else if (KIND == 2) {
    double Mq = 1 + 0.5 * r2 / a_rq;
    K = sf2 * pow(Mq, -a_rq);
    c = sf2 * pow(Mq, -a_rq - 1);                    // :357
    dshape = K * (0.5 * r2 / Mq - a_rq * log(Mq));   // :363
  } else {
    double r = sqrt(r2);
    double e = exp(-r);
    double f = (KIND == 1) ? 1.0 : (KIND == 3 ? 1 + r : 1 + r * (1 + r * (1.0 / 3)));
    double df = (KIND == 1) ? 1.0 / r : (KIND == 3 ? 1.0 : (1 + r) * (1.0 / 3));   // :210-218
    K = sf2 * f * e;
    c = sf2 * (df * e);                              // :280 (inf*0 -> NaN for Matern-1 at r=0)
  }
}

// z_rr = sum_{c <= rr} D(rr, c) b_c for one row of a column-major 128x128 D_k: element c goes to
// accumulator c mod 4 in increasing c, combined as (s0 + s1) + (s2 + s3) -- the order of
// diag_kernel's tail, so a replayed solve is bit-identical.  Loads are issued 16 at a time.
__device__ __forceinline__ double zsolve_row(const double* Dk, const double* bsh, int rr, int nact) {
  double s4[4] = {0.0, 0.0, 0.0, 0.0};
  if (rr < nact) {
    for (int c0 = 0; c0 <= rr; c0 += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (c0 + u <= rr) ? Dk[(c0 + u) * T + rr] : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (c0 + u <= rr) s4[u & 3] = __fma_rn(v[u], bsh[c0 + u], s4[u & 3]);
    }
  }
  return __dadd_rn(__dadd_rn(s4[0], s4[1]), __dadd_rn(s4[2], s4[3]));
}

__host__ __device__ inline int kind_code(int cov_kind, int degree) {
  return cov_kind == 0 ? 0 : (cov_kind == 2 ? 2 : degree);   // 0,1,3,5,2
}

// deterministic block reduction (fixed tree), result valid in thread 0
template <int NT>
__device__ inline double block_sum(double v, double* sh) {
  int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
#pragma unroll
  for (int s = NT / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  double r = sh[0];
  __syncthreads();
  return r;
}

}  // namespace gpb
