/*
 * gpyreg_b200 -- C ABI of the B200-native GP hot path.
 *
 * Drop-in boundary for the path BASELINE.json's north_star names:
 * covariance assembly -> Cholesky posterior -> nlZ + gradient -> prediction,
 * batched over hyperparameter vectors.  Every entry point takes plain pointers
 * and sizes; all host buffers are caller-owned IEEE float64, C-contiguous, and
 * are never retained after the call returns.  Return value: 0 = ok, otherwise a
 * GPB_E* code; gpb_last_error() gives the text.  Per-element numerical failure
 * (Cholesky still failing after the reference's 10 jitter retries) is reported
 * in status[], never by aborting the batch.
 *
 * The reference (acerbilab/gpyreg) is pure Python; "file:line" below names the
 * reference code each entry point replaces (paths under gpyreg/).
 */
#ifndef GPYREG_B200_H
#define GPYREG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* covariance_functions.py:131 / :189 / :288 (+ isotropic_covariance_functions.py:86,:164 with ard=0) */
enum { GPB_COV_SE = 0, GPB_COV_MATERN = 1, GPB_COV_RQ = 2 };
/* mean_functions.py:6 / :134 / :263 */
enum { GPB_MEAN_ZERO = 0, GPB_MEAN_CONST = 1, GPB_MEAN_NEGQUAD = 2 };
/* error codes */
enum {
  GPB_OK = 0,
  GPB_EINVAL = 1,    /* bad argument / unsupported shape */
  GPB_ECUDA = 2,     /* CUDA runtime error (text in gpb_last_error) */
  GPB_ENOMEM = 3,    /* workspace does not fit on the device */
  GPB_ESTATE = 4,    /* call order: model/data not set */
  GPB_EAGAIN = 5     /* gpb_posterior_append: not applicable here, do the full rebuild */
};
/* gpb_posterior_fetch fields -- the members of Posterior, gaussian_process.py:2568-2586 */
enum {
  GPB_POST_ALPHA = 0,   /* (N)      alpha                                        */
  GPB_POST_L = 1,       /* (N,N)    row-major: upper Cholesky factor U (U^T U = A) if L_chol,
                                    else -(K + sn2_mult*diag(sn2))^-1            */
  GPB_POST_SW = 2,      /* (1)      1/sqrt(min(sn2)*sn2_mult); the reference tiles it to (N,1) */
  GPB_POST_SN2MULT = 3, /* (1)                                                   */
  GPB_POST_LCHOL = 4,   /* (1)      1.0 / 0.0                                    */
  GPB_POST_STATUS = 5   /* (1)      0 ok, 1 = "Singular matrix for L Cholesky decomposition" */
};

typedef struct gpb_ctx gpb_ctx;
typedef struct gpb_post gpb_post;

/* One context per process and GPU (one host thread per context).  gpb_destroy also releases
 * every posterior batch created from the context that is still alive: their handles must
 * not be used (or freed) afterwards. */
int gpb_create(int device, gpb_ctx** out);
void gpb_destroy(gpb_ctx* ctx);
const char* gpb_last_error(const gpb_ctx* ctx);   /* ctx may be NULL: last create error */
int gpb_version(void);
/* Launch all work of this context on an existing CUDA stream (a cudaStream_t
 * passed as an integer, e.g. torch.cuda.current_stream().cuda_stream). 0 = the
 * context's own stream. */
int gpb_set_stream(gpb_ctx* ctx, uint64_t cuda_stream);
/* Cap (bytes) on the batched-matrix workspace; 0 = 70% of free device memory. */
int gpb_set_workspace_limit(gpb_ctx* ctx, uint64_t bytes);

/* Which plugin objects the GP was built with: GP(D, covariance, mean, noise),
 * gaussian_process.py:43-62.  noise_flags = GaussianNoise.parameters,
 * noise_functions.py:33-41: {constant_add, user_provided (0/1/2=scaled), rectified}. */
int gpb_set_model(gpb_ctx* ctx, int cov_kind, int matern_degree, int ard,
                  int mean_kind, const int noise_flags[3]);
/* Training data (gp.X, gp.y, gp.s2; gaussian_process.py:1008-1017, :846-862).
 * X is (N,D) row-major; y is (N); s2 is (N) or NULL.  Uploaded once. */
int gpb_set_data(gpb_ctx* ctx, const double* X, const double* y, const double* s2,
                 int64_t N, int D);

/* Batched GP.__compute_nlZ(hyp, compute_grad, compute_prior=False)
 * = GP.__core_computation(hyp, 1, want_grad), gaussian_process.py:1520-1538, :2357-2512,
 * for B hyperparameter rows at once (the reference loops: f_min_fill.py:174-176).
 *   hyp (B,P) row-major, P = cov_N + noise_N + mean_N; nlZ (B); dnlZ (B,P) or NULL;
 *   sn2_mult (B) or NULL; status (B) or NULL: 0 ok, 1 = Cholesky failed 10 times
 *   (the reference raises LinAlgError, :2450-2453). */
int gpb_nlz_batch(gpb_ctx* ctx, const double* hyp, int64_t B, int want_grad,
                  double* nlZ, double* dnlZ, double* sn2_mult, int32_t* status);
/* Same, with DEVICE pointers (inputs already resident in HBM); results stay on
 * the device.  Used for kernel-only timing and by multi-GPU callers. */
int gpb_nlz_batch_dev(gpb_ctx* ctx, const double* d_hyp, int64_t B, int want_grad,
                      double* d_nlZ, double* d_dnlZ, double* d_sn2_mult,
                      int32_t* d_status);

/* Batched GP.update full-recompute loop: posteriors[i] = __core_computation(hyp[i],0,0),
 * gaussian_process.py:870-879.  The factors stay on the device. */
int gpb_posterior_batch(gpb_ctx* ctx, const double* hyp, int64_t B, gpb_post** out);
int64_t gpb_posterior_count(const gpb_post* post);
int gpb_posterior_fetch(const gpb_post* post, int64_t b, int field, double* out);
void gpb_posterior_free(gpb_post* post);
/* Number of training points the factors of `post` currently cover. */
int64_t gpb_posterior_size(const gpb_post* post);
/* Rank-one update of GP.update, gaussian_process.py:737-844: append ONE training point
 * (x_new (D), y_new, no s2) to every sample of `post`, in place on the device.
 *   status[s] = 1  <=>  the reference's stability test fails for sample s
 *   (sqrt_arg <= 0, :784-798, "Rank-one update of Cholesky factor unstable"): that sample
 *   is left untouched and unusable until the caller recomputes it with gpb_posterior_rebuild.
 * Returns GPB_EAGAIN (nothing changed) when the in-place update does not apply: the noise
 * variance depends on the point (user-provided or output-dependent terms), or the padded
 * device layout has no free row left (every 128 points) -- rebuild with gpb_posterior_batch. */
int gpb_posterior_append(gpb_ctx* ctx, gpb_post* post, const double* x_new, double y_new,
                         int32_t* status);
/* "Compute full update where rank-1 failed", gaussian_process.py:864-868: recompute the listed
 * samples of `post` from scratch (sn2_mult restarts at 1) on the data the context holds, which
 * must be the data the other samples cover (call gpb_set_data with the extended X, y first).
 * The other samples keep their rank-one factors. */
int gpb_posterior_rebuild(gpb_ctx* ctx, gpb_post* post, const int32_t* slots, int64_t n);

/* GP.predict, gaussian_process.py:1663-1816, over all samples of `post`.
 *   Xs (M,D); ys (M) or NULL; s2s (M) or NULL; outputs mu, s2 [, lpd]:
 *   (M) when separate == 0 (averaged over samples, :1793-1811), else (M,Ns) row-major.
 *   lpd may be NULL when want_lpd == 0. */
int gpb_predict(gpb_ctx* ctx, const gpb_post* post, const double* Xs, const double* ys,
                const double* s2s, int64_t M, int add_noise, int separate, int want_lpd,
                double* mu, double* s2, double* lpd);
/* Device-pointer variant (Xs, mu, s2 on the device; no ys/s2s/lpd). */
int gpb_predict_dev(gpb_ctx* ctx, const gpb_post* post, const double* d_Xs, int64_t M,
                    int add_noise, int separate, double* d_mu, double* d_s2);

/* GP.predict_full, gaussian_process.py:1561-1661: posterior mean and FULL covariance at M test
 * points (M <= 8192), per hyperparameter sample.  mu (M,Ns) row-major; cov (Ns,M,M) C-order
 * (the Python layer returns cov.transpose(1,2,0) like the reference, :1661). */
int gpb_predict_full(gpb_ctx* ctx, const gpb_post* post, const double* Xs, const double* ys,
                     const double* s2s, int64_t M, int add_noise, double* mu, double* cov);

/* GP.quad, gaussian_process.py:1818-1981: Bayesian quadrature of the GP against M Gaussian
 * measures N(mu_j, diag(sigma_j^2)) (squared-exponential ARD kernel only).
 *   mu, sigma (M,D); F [, F_var]: (M) when separate == 0, else (M,Ns) row-major;
 *   F_var may be NULL when compute_var == 0. */
int gpb_quad(gpb_ctx* ctx, const gpb_post* post, const double* mu, const double* sigma, int64_t M,
             int compute_var, int separate, double* F, double* F_var);

/* Plugin surface -------------------------------------------------------------------
 * covariance.compute(hyp, X, X_star, compute_diag, compute_grad),
 * covariance_functions.py:135-186, :221-285, :301-367; isotropic_...py:104-161, :173-221.
 *   K:  (N,N) | (N,M) if Xs != NULL | (N) if diag.   dK: (cov_N,N,N) C-order or NULL
 *   (the Python layer returns dK.transpose(1,2,0) like the reference, :184). */
int gpb_cov(gpb_ctx* ctx, int cov_kind, int matern_degree, int ard, const double* hyp,
            const double* X, int64_t N, int D, const double* Xs, int64_t M, int diag,
            double* K, double* dK);
/* mean.compute(hyp, X, compute_grad), mean_functions.py:82-131, :210-260, :340-397.
 *   m (N); dm (N,mean_N) row-major or NULL. */
int gpb_mean(gpb_ctx* ctx, int mean_kind, const double* hyp, const double* X, int64_t N,
             int D, double* m, double* dm);
/* noise.compute(hyp, X, y, s2, compute_grad), noise_functions.py:179-283.
 *   y, s2 may be NULL.  sn2 (N) always per point (the Python layer collapses it to a
 *   scalar when the reference would); dsn2 (N,noise_N) row-major or NULL. */
int gpb_noise(gpb_ctx* ctx, const int noise_flags[3], const double* hyp, const double* y,
              const double* s2, int64_t N, double* sn2, double* dsn2);

/* Test / measurement hooks (not part of the reference surface) ----------------------
 * C = alpha * A * B^T + beta * C on the FP64 tensor-core tile kernel; A (M,K), B (N,K),
 * C (M,N) column-major host buffers, M,N multiples of 128, K a multiple of 16. */
int gpb_debug_gemm_nt(gpb_ctx* ctx, const double* A, const double* B, double* C,
                      int M, int N, int K, double alpha, double beta);
/* In-place lower Cholesky of a column-major (n,n) host matrix through the blocked
 * batched path (padding handled inside); info = 1 if a pivot was <= 0 or NaN. */
int gpb_debug_potrf(gpb_ctx* ctx, double* A, int n, int32_t* info);
/* Latency of the diagonal-tile kernel in a dependent chain of `reps` launches (microseconds per launch). */
int gpb_debug_diag_bench(gpb_ctx* ctx, int reps, int with_rhs, double* us);
/* Time `reps` launches of the tile GEMM on device-resident random data and return
 * the average ms per launch (CUDA events). */
int gpb_debug_gemm_bench(gpb_ctx* ctx, int M, int N, int K, int reps, double* ms);
/* Per-phase device timings (ms) of the last gpb_nlz_batch[_dev] call:
 * {prep+build, potrf, solve, inverse, gradient, total}. */
int gpb_last_timings(const gpb_ctx* ctx, double out[6]);
/* Number of kernels launched by this context so far. */
int64_t gpb_launch_count(const gpb_ctx* ctx);
/* Factor cache of gpb_nlz_batch (nlZ-only calls): a row whose covariance and noise
 * hyperparameters equal those of the row last evaluated in the same batch position re-uses that
 * Cholesky factor and only replays the O(N^2) forward solve (bit-identical result).  Counts rows. */
int gpb_cache_stats(const gpb_ctx* ctx, int64_t* hits, int64_t* misses);

#ifdef __cplusplus
}
#endif
#endif /* GPYREG_B200_H */
