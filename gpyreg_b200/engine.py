"""Thin object wrapper over the C ABI: one Engine = one gpb_ctx = one GPU.

NumPy float64 in, NumPy float64 out.  The Engine mirrors what the reference's
numerical core needs from ``GP`` (gaussian_process.py:2357-2521): a model
descriptor, the training data, and batched evaluation over hyperparameter rows.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from ._lib import GpbError, f64, ptr

FIELDS = {"alpha": 0, "L": 1, "sW": 2, "sn2_mult": 3, "L_chol": 4, "status": 5}


def default_device():
    """LOCAL_RANK under torchrun, else 0 (one process per GPU)."""
    return int(os.environ.get("GPYREG_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


class PosteriorBatch:
    """Device-resident posteriors of a batch of hyperparameter samples."""

    def __init__(self, engine, handle, count, N):
        self.engine, self._h, self.count, self.N = engine, handle, count, N
        engine._batches.add(self)

    def fetch(self, s, field):
        n = {"alpha": self.N, "L": self.N * self.N}.get(field, 1)
        out = np.empty(n)
        self.engine._check(self.engine.lib.gpb_posterior_fetch(self._h, s, FIELDS[field], ptr(out)))
        if field == "L":
            return out.reshape(self.N, self.N)
        return out if field == "alpha" else out[0]

    def free(self):
        # a closed engine has already released its batches (gpb_destroy): never touch the handle then
        if self._h is not None and self.engine.lib is not None and self.engine._h is not None:
            self.engine.lib.gpb_posterior_free(self._h)
        self._h = None
        self.engine._batches.discard(self)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, device=None):
        self.lib = _lib.load()
        self.device = default_device() if device is None else device
        h = C.c_void_p()
        rc = self.lib.gpb_create(self.device, C.byref(h))
        if rc != 0:
            raise GpbError(rc, self.lib.gpb_last_error(None).decode())
        self._h = h
        self.model = None
        self.N = self.D = 0
        self._batches = weakref.WeakSet()      # live PosteriorBatch objects of this context
        self.data_owner = None                 # token of whoever uploaded the current data (GP objects share engines)

    def close(self):
        if getattr(self, "_h", None) is not None:
            for b in list(self._batches):      # their device buffers die with the context
                b.free()
            self.lib.gpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GpbError(rc, self.lib.gpb_last_error(self._h).decode())

    # -- configuration ---------------------------------------------------------
    def set_model(self, cov_kind, degree, ard, mean_kind, noise_params):
        nz = (C.c_int * 3)(*[int(v) for v in noise_params])
        self._check(self.lib.gpb_set_model(self._h, cov_kind, degree, int(ard), mean_kind, nz))
        self.model = (cov_kind, degree, int(ard), mean_kind, tuple(int(v) for v in noise_params))

    def set_data(self, X, y, s2=None):
        X = f64(X)
        N, D = X.shape
        y = f64(y, (N,))
        s2 = None if s2 is None else f64(s2, (N,))
        self._check(self.lib.gpb_set_data(self._h, ptr(X), ptr(y), ptr(s2), N, D))
        self.N, self.D = N, D
        if self.model is not None:          # parameter counts depend on D
            self.set_model(*self.model)

    def set_stream(self, cuda_stream):
        self._check(self.lib.gpb_set_stream(self._h, int(cuda_stream)))

    def set_workspace_limit(self, nbytes):
        self._check(self.lib.gpb_set_workspace_limit(self._h, int(nbytes)))

    # -- hot path ---------------------------------------------------------------
    def hyp_n(self):
        """Length of a hyperparameter row for the current model and data: cov | noise | mean."""
        from .spec import ModelSpec
        ck, deg, ard, mk, nz = self.model
        return ModelSpec(D=self.D, cov_kind=ck, degree=deg, ard=bool(ard), mean_kind=mk, noise_params=nz).hyp_n

    def _check_hyp(self, hyp):
        hyp = f64(hyp)
        if hyp.ndim == 1:
            hyp = hyp[None, :]
        if self.model is None or self.N == 0:
            raise GpbError(4, "set_model and set_data first")
        if hyp.ndim != 2 or hyp.shape[1] != self.hyp_n():
            raise ValueError(f"hyperparameter rows must have {self.hyp_n()} entries, got shape {hyp.shape}")
        return hyp

    def nlz_batch(self, hyp, want_grad=False):
        """-> nlZ (B,), dnlZ (B,P) or None, sn2_mult (B,), status (B,) int32."""
        hyp = self._check_hyp(hyp)
        B, P = hyp.shape
        nlz = np.empty(B)
        dnlz = np.empty((B, P)) if want_grad else None
        mult = np.empty(B)
        status = np.zeros(B, dtype=np.int32)
        self._check(self.lib.gpb_nlz_batch(self._h, ptr(hyp), B, int(want_grad), ptr(nlz), ptr(dnlz),
                                           ptr(mult), ptr(status)))
        return nlz, dnlz, mult, status

    def nlz_batch_dev(self, d_hyp, B, want_grad, d_nlz, d_dnlz, d_mult=0, d_status=0):
        """Device-pointer variant (ints from tensor.data_ptr())."""
        self._check(self.lib.gpb_nlz_batch_dev(self._h, d_hyp, B, int(want_grad), d_nlz, d_dnlz or None,
                                               d_mult or None, d_status or None))

    def posterior_batch(self, hyp):
        hyp = self._check_hyp(hyp)
        h = C.c_void_p()
        self._check(self.lib.gpb_posterior_batch(self._h, ptr(hyp), hyp.shape[0], C.byref(h)))
        return PosteriorBatch(self, h, hyp.shape[0], self.N)

    def posterior_append(self, post, x_new, y_new):
        """Rank-one append of one training point to every sample of ``post``, in place on the
        device (gaussian_process.py:737-844).  Returns the per-sample status (1 = the reference's
        stability test failed; the batch must then be rebuilt), or None when the in-place update
        does not apply (GPB_EAGAIN: point-dependent noise / no free row in the padded layout)."""
        x_new = f64(x_new).reshape(-1)
        status = np.zeros((post.count,), dtype=np.int32)
        rc = self.lib.gpb_posterior_append(self._h, post._h, ptr(x_new), float(y_new), ptr(status))
        if rc == 5:
            return None
        self._check(rc)
        post.N = int(self.lib.gpb_posterior_size(post._h))
        return status

    def posterior_rebuild(self, post, slots):
        """Recompute the listed samples of ``post`` from scratch on the data this engine holds (the
        reference's "full update where rank-1 failed", gaussian_process.py:864-868)."""
        slots = np.ascontiguousarray(slots, dtype=np.int32).reshape(-1)
        self._check(self.lib.gpb_posterior_rebuild(self._h, post._h, ptr(slots), slots.size))

    def predict(self, post, Xs, ys=None, s2s=None, add_noise=False, separate=False, want_lpd=False):
        Xs = f64(Xs)
        M = Xs.shape[0]
        ys = None if ys is None else f64(ys, (M,))
        s2s = None if s2s is None else f64(s2s, (M,))
        cols = post.count if separate else 1
        mu, s2 = np.empty((M, cols)), np.empty((M, cols))
        lpd = np.empty((M, cols)) if want_lpd else None
        self._check(self.lib.gpb_predict(self._h, post._h, ptr(Xs), ptr(ys), ptr(s2s), M, int(add_noise),
                                         int(separate), int(want_lpd), ptr(mu), ptr(s2), ptr(lpd)))
        return (mu, s2, lpd) if want_lpd else (mu, s2)

    def predict_full(self, post, Xs, ys=None, s2s=None, add_noise=False):
        """-> mu (M, Ns), cov (M, M, Ns) like GP.predict_full."""
        Xs = f64(Xs)
        M = Xs.shape[0]
        ys = None if ys is None else f64(ys, (M,))
        s2s = None if s2s is None else f64(s2s, (M,))
        mu = np.empty((M, post.count))
        cov = np.empty((post.count, M, M))
        self._check(self.lib.gpb_predict_full(self._h, post._h, ptr(Xs), ptr(ys), ptr(s2s), M,
                                              int(add_noise), ptr(mu), ptr(cov)))
        return mu, cov.transpose(1, 2, 0)

    def quad(self, post, mu, sigma, compute_var=False, separate=False):
        mu, sigma = f64(mu), f64(sigma)
        M = mu.shape[0]
        cols = post.count if separate else 1
        F = np.empty((M, cols))
        Fv = np.empty((M, cols)) if compute_var else None
        self._check(self.lib.gpb_quad(self._h, post._h, ptr(mu), ptr(sigma), M, int(compute_var),
                                      int(separate), ptr(F), ptr(Fv)))
        return (F, Fv) if compute_var else F

    def predict_dev(self, post, d_Xs, M, add_noise, separate, d_mu, d_s2):
        self._check(self.lib.gpb_predict_dev(self._h, post._h, d_Xs, M, int(add_noise), int(separate),
                                             d_mu, d_s2))

    # -- plugin surface ------------------------------------------------------------
    def cov(self, cov_kind, degree, ard, hyp, X, Xs=None, diag=False, grad=False):
        X = f64(X)
        N, D = X.shape
        hyp = f64(hyp)
        cov_n = hyp.size
        M = 0
        if Xs is not None:
            Xs = f64(Xs)
            M = Xs.shape[0]
        K = np.empty((N, 1)) if diag else np.empty((N, M if Xs is not None else N))
        dK = np.empty((cov_n, N, N)) if grad else None
        self._check(self.lib.gpb_cov(self._h, cov_kind, degree, int(ard), ptr(hyp), ptr(X), N, D,
                                     ptr(Xs), M, int(diag), ptr(K), ptr(dK)))
        return (K, dK) if grad else K

    def mean(self, mean_kind, hyp, X, grad=False):
        X = f64(X)
        N, D = X.shape
        hyp = f64(hyp)
        m = np.empty(N)
        dm = np.empty((N, hyp.size)) if (grad and hyp.size) else None
        self._check(self.lib.gpb_mean(self._h, mean_kind, ptr(hyp) if hyp.size else None, ptr(X), N, D,
                                      ptr(m), ptr(dm)))
        return (m, dm) if grad else m

    def noise(self, noise_params, hyp, N, y=None, s2=None, grad=False):
        hyp = f64(hyp)
        nz = (C.c_int * 3)(*[int(v) for v in noise_params])
        y = None if y is None else f64(y, (N,))
        s2 = None if s2 is None else f64(s2, (N,))
        sn2 = np.empty(N)
        dsn2 = np.empty((N, hyp.size)) if (grad and hyp.size) else None
        self._check(self.lib.gpb_noise(self._h, nz, ptr(hyp) if hyp.size else None, ptr(y), ptr(s2), N,
                                       ptr(sn2), ptr(dsn2)))
        return (sn2, dsn2) if grad else sn2

    # -- hooks ---------------------------------------------------------------------
    def debug_gemm_nt(self, A, B, Cm, alpha=1.0, beta=0.0):
        """A (M,K), B (N,K), C (M,N) as NumPy arrays; returns alpha*A@B.T + beta*C."""
        A, B, Cm = f64(A), f64(B), f64(Cm)
        M, K = A.shape
        N = B.shape[0]
        Af, Bf, Cf = np.asfortranarray(A), np.asfortranarray(B), np.asfortranarray(Cm).copy(order="F")
        self._check(self.lib.gpb_debug_gemm_nt(self._h, Af.ctypes.data, Bf.ctypes.data, Cf.ctypes.data,
                                               M, N, K, alpha, beta))
        return np.ascontiguousarray(Cf)

    def debug_potrf(self, A):
        """Lower Cholesky of a symmetric (n,n) matrix through the batched blocked path."""
        Af = np.asfortranarray(f64(A)).copy(order="F")
        info = np.zeros(1, dtype=np.int32)
        self._check(self.lib.gpb_debug_potrf(self._h, Af.ctypes.data, Af.shape[0], ptr(info)))
        return np.ascontiguousarray(Af), int(info[0])

    def debug_diag_bench(self, reps=200, with_rhs=True):
        """Microseconds per launch of the diagonal-tile kernel in a dependent chain."""
        us = np.zeros(1)
        self._check(self.lib.gpb_debug_diag_bench(self._h, int(reps), int(with_rhs), ptr(us)))
        return float(us[0])

    def debug_gemm_bench(self, M, N, K, reps=10):
        ms = np.zeros(1)
        self._check(self.lib.gpb_debug_gemm_bench(self._h, M, N, K, reps, ptr(ms)))
        return float(ms[0])

    def last_timings(self):
        out = np.zeros(6)
        self._check(self.lib.gpb_last_timings(self._h, ptr(out)))
        return dict(zip(("factor", "nlz", "solve", "inverse", "gradient", "total"), out))

    def launch_count(self):
        return int(self.lib.gpb_launch_count(self._h))

    def cache_stats(self):
        """(hits, misses) of the nlZ factor cache, in hyperparameter rows."""
        out = np.zeros(2, dtype=np.int64)
        self._check(self.lib.gpb_cache_stats(self._h, out.ctypes.data, out.ctypes.data + 8))
        return int(out[0]), int(out[1])


_default = {}


def get_engine(device=None):
    """The process-wide engine of a device (created on first use; raises without a GPU).  GP objects
    and the plugin ``compute`` methods share it: a context sizes its evaluation workspace to the
    free device memory, so one context per GP would make several live GPs starve each other."""
    dev = default_device() if device is None else device
    eng = _default.get(dev)
    if eng is None or eng._h is None:
        eng = _default[dev] = Engine(dev)
    return eng
