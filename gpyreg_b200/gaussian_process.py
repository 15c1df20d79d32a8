"""Drop-in ``GP`` for the path BASELINE.json's north_star names, behind the reference's
API (gpyreg/gaussian_process.py): ``update``, ``fit``, ``predict``, the private nlZ /
posterior entry points, hyperparameter / bounds / prior plumbing.

Every numerical evaluation goes to the B200 through :class:`gpyreg_b200.Engine`
(C ABI, include/gpyreg_b200.h); there is no NumPy fallback for the hot path.  The three
places where the reference loops over hyperparameter vectors one at a time are batched:
the ``f_min_fill`` design (f_min_fill.py:174-176), the posterior rebuild in ``update``
(gaussian_process.py:870-879) and the per-sample loop of ``predict`` (:1727).

A one-point ``update`` takes the reference's rank-one path (:737-844) in place on the device
(``gpb_posterior_append``) when the noise variance does not depend on the point; otherwise, and
every 128th point (padded layout full), the batch is rebuilt -- same posterior up to rounding.

Not built: ``plot`` (matplotlib).
"""
import math
import warnings

import numpy as np
import scipy as sp
import scipy.linalg
import scipy.optimize
import scipy.special
import scipy.stats

from .batched_drivers import MultiChainSliceSampler, minimize_lockstep
from .engine import Engine
from .f_min_fill import (f_min_fill, smoothbox_cdf, smoothbox_student_t_cdf)
from .sharding import sharded_nlz_device, sharded_predict_device, sharded_rows
from .slice_sample import SliceSampler
from .spec import ModelSpec


def _world_size():
    """Number of ranks when the process runs under torch.distributed (one process per GPU)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size()
    except ImportError:
        pass
    return 1


def _scipy_at_least(major, minor):
    try:
        parts = sp.__version__.split(".")
        return (int(parts[0]), int(parts[1])) >= (major, minor)
    except (ValueError, IndexError):
        return False


# SciPy >= 1.15 ships L-BFGS-B as re-entrant C (the older Fortran kept state in SAVE variables)
_SCIPY_LBFGSB_REENTRANT = _scipy_at_least(1, 15)


def _robust_cholesky(sigma):
    """Upper factor T with T^T T = sigma; falls back to an eigen-decomposition when sigma is
    only positive semi-definite (gaussian_process.py:2331-2355)."""
    try:
        return sp.linalg.cholesky(sigma, check_finite=False)
    except sp.linalg.LinAlgError:
        w, U = sp.linalg.eigh((sigma + sigma.T) / 2)
        keep = np.abs(w) > np.abs(np.spacing(np.max(w))) * w.shape[0]
        w, U = w[keep], U[:, keep]
        if np.any(w < 0):
            return np.zeros(sigma.shape)
        U = U * np.where(U[np.argmax(np.abs(U), axis=0), np.arange(U.shape[1])] < 0, -1.0, 1.0)
        return np.dot(np.diag(np.sqrt(w)), U.T)


class Posterior:
    """The record the reference keeps per hyperparameter sample (gaussian_process.py:2568-2586):
    ``hyp, alpha, sW, L, sn2_mult, L_chol``.  Here the factors live on the GPU; ``alpha``,
    ``sW`` and ``L`` are fetched into NumPy arrays the first time they are read."""

    _FIELDS = ("alpha", "sW", "L", "sn2_mult", "L_chol")

    def __init__(self, hyp, alpha, sW, L, sn2_mult, Lchol, _batch=None, _index=0):
        self.hyp = hyp
        self._batch, self._index = _batch, _index
        self._val = {"alpha": alpha, "sW": sW, "L": L, "sn2_mult": sn2_mult, "L_chol": Lchol}
        self._have = {k: (_batch is None) for k in self._FIELDS}

    def _get(self, name):
        if not self._have[name]:
            b, s = self._batch, self._index
            if b is None:
                # a record of a copied / unpickled GP whose factor was not carried along: the owning
                # GP re-creates the device factors of all its samples and this field is read from them
                gp = self._owner() if getattr(self, "_owner", None) is not None else None
                b = gp._device_batch() if gp is not None else None
                if b is None:
                    raise RuntimeError("Posterior: the factor of this record is gone (its GP was deleted)")
            if name == "alpha":
                v = b.fetch(s, "alpha").reshape(-1, 1)
            elif name == "sW":
                v = np.ones((b.N, 1)) * b.fetch(s, "sW")      # constant vector, :2517
            elif name == "L":
                v = b.fetch(s, "L")
            elif name == "sn2_mult":
                m = b.fetch(s, "sn2_mult")
                v = int(m) if float(m).is_integer() else float(m)
            else:
                v = bool(b.fetch(s, "L_chol"))
            self._val[name], self._have[name] = v, True
        return self._val[name]

    def _set(self, name, v):
        self._val[name], self._have[name] = v, True

    def _detached(self, with_factor=True):
        """A plain host-side copy of the record with no tie to the device: what ``copy.deepcopy`` /
        ``pickle`` of a reference Posterior would hold.  ``with_factor=False`` (copies of a whole GP)
        leaves the (N, N) factor behind unless it has already been fetched: the copied GP rebuilds its
        device factors on first use and the field is read from them then."""
        skip = () if with_factor or self._have["L"] else ("L",)
        vals = {k: (None if k in skip else self._get(k)) for k in self._FIELDS}
        vals = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in vals.items()}
        new = Posterior(np.array(self.hyp, copy=True), vals["alpha"], vals["sW"], vals["L"], vals["sn2_mult"],
                        vals["L_chol"], _index=self._index)
        for k in skip:
            new._have[k] = False
        return new

    def __deepcopy__(self, memo):
        if self._batch is None and getattr(self, "_owner", None) is None:
            # already a plain host record (possibly one of GP.__getstate__'s, with its factor left
            # behind): copy it as it is
            new = _posterior_from_state(np.array(self.hyp, copy=True),
                                        {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self._val.items()},
                                        self._have, self._index)
        else:
            new = self._detached()
        memo[id(self)] = new
        return new

    def __reduce__(self):
        d = self if (self._batch is None and getattr(self, "_owner", None) is None) else self._detached()
        return (_posterior_from_state, (d.hyp, dict(d._val), dict(d._have), d._index))

    alpha = property(lambda s: s._get("alpha"), lambda s, v: s._set("alpha", v))
    sW = property(lambda s: s._get("sW"), lambda s, v: s._set("sW", v))
    L = property(lambda s: s._get("L"), lambda s, v: s._set("L", v))
    sn2_mult = property(lambda s: s._get("sn2_mult"), lambda s, v: s._set("sn2_mult", v))
    L_chol = property(lambda s: s._get("L_chol"), lambda s, v: s._set("L_chol", v))


def _posterior_from_state(hyp, val, have, index):
    p = Posterior(hyp, val["alpha"], val["sW"], val["L"], val["sn2_mult"], val["L_chol"], _index=index)
    p._have = dict(have)
    return p


def _spec_of(D, covariance, mean, noise):
    """Read the model descriptor off the plugin objects."""
    try:
        return ModelSpec(D=D, cov_kind=covariance._cov_kind, degree=int(getattr(covariance, "degree", 0)),
                         ard=bool(covariance._ard), mean_kind=mean._mean_kind,
                         noise_params=tuple(int(v) for v in noise.parameters))
    except AttributeError as e:
        raise TypeError("gpyreg_b200.GP needs covariance/mean/noise objects from gpyreg_b200 "
                        "(the fused CUDA kernels implement exactly those families)") from e


def gp_from_spec(spec):
    """A GP built from the plugin classes a ModelSpec names (benchmarks, tools)."""
    from . import covariance_functions as cf
    from . import isotropic_covariance_functions as icf
    from . import mean_functions as mf
    from .noise_functions import GaussianNoise
    if spec.cov_kind == 0:
        cov = cf.SquaredExponential() if spec.ard else icf.SquaredExponentialIsotropic()
    elif spec.cov_kind == 1:
        cov = cf.Matern(spec.degree) if spec.ard else icf.MaternIsotropic(spec.degree)
    else:
        cov = cf.RationalQuadraticARD() if spec.ard else icf.RationalQuadraticIsotropic()
    mean = (mf.ZeroMean, mf.ConstantMean, mf.NegativeQuadratic)[spec.mean_kind]()
    p = spec.noise_params
    return GP(spec.D, cov, mean, GaussianNoise(p[0] == 1, p[1] >= 1, p[1] == 2, p[2] == 1))


class GP:
    """A single Gaussian-process model (gaussian_process.py:15-62)."""

    SHARD_MIN_N = 512      # below this a factorisation is cheaper than the all-gather's latency

    MAX_D = 64             # the fused kernels stage 2*D*128 doubles in shared memory and keep one gradient accumulator
                           # per ARD length scale in registers (csrc/common.cuh MAXD)

    def __init__(self, D, covariance, mean, noise):
        if D > self.MAX_D:
            # fail at construction, not at the first upload deep inside fit()/update()
            raise ValueError(f"gpyreg_b200.GP supports input dimension D <= {self.MAX_D} (got D = {D}): the fused "
                             "covariance / gradient kernels hold one accumulator per dimension in registers. "
                             "The plugin compute() methods have no such limit.")
        self.D = D
        self.covariance, self.mean, self.noise = covariance, mean, noise
        self._spec = _spec_of(D, covariance, mean, noise)
        self.s2 = self.X = self.y = None
        self.posteriors = None
        self.no_prior = None
        self.normalization_constants = None
        self._engine = None
        self._data_key = None
        self._token = object()               # identifies this GP's upload on the shared engine
        self._post_batch = None
        self.set_bounds()
        self.set_priors()
        self.temporary_data = {}

    # ------------------------------------------------------------------ copying
    _DEVICE_STATE = ("_engine", "_post_batch", "_token", "_data_key")

    def __getstate__(self):
        """Everything but the device handles (PyVBMC deep-copies and pickles its GPs): the copy's
        posteriors are plain host records, and its factors are rebuilt on the GPU at the first
        predict / quad / rank-one update."""
        d = {k: v for k, v in self.__dict__.items() if k not in self._DEVICE_STATE}
        if self.posteriors is not None:
            posts = np.empty(self.posteriors.shape, dtype=object)
            for i, p in enumerate(self.posteriors):
                posts[i] = p._detached(with_factor=False)     # 8 N^2 bytes per sample stay on the device
            d["posteriors"] = posts
        return d

    def __setstate__(self, d):
        import weakref
        self.__dict__.update(d)
        self._engine = self._post_batch = self._data_key = None
        self._token = object()
        if self.posteriors is not None:
            for p in self.posteriors:
                p._owner = weakref.ref(self)

    def __deepcopy__(self, memo):
        import copy
        new = type(self).__new__(type(self))
        memo[id(self)] = new
        new.__setstate__(copy.deepcopy(self.__getstate__(), memo))
        return new

    def _device_batch(self):
        """The device-resident posterior batch behind ``self.posteriors``; rebuilt from the stored
        hyperparameter samples when this GP holds host records only (a copied / unpickled GP)."""
        batch = self._post_batch
        if batch is not None and batch._h is not None and self.posteriors is not None and \
                all(p._batch is batch for p in self.posteriors):
            return batch
        if self.posteriors is None or self.X is None or self.y is None or \
                any(p._have["alpha"] and p._val["alpha"] is None for p in self.posteriors):
            return None                      # cleaned, or never computed: the caller must update()
        hyp = np.stack([np.asarray(p.hyp, dtype=float) for p in self.posteriors])
        self.posteriors, self._post_batch = self._posteriors_for(hyp)
        return self._post_batch

    # ------------------------------------------------------------------ layout helpers
    def _hyper_info(self):
        """[(name, count)] in the order cov | noise | mean (gaussian_process.py:174)."""
        return (self.covariance.hyperparameter_info(self.D) + self.noise.hyperparameter_info()
                + self.mean.hyperparameter_info(self.D))

    def _counts(self):
        return (self.covariance.hyperparameter_count(self.D), self.noise.hyperparameter_count(),
                self.mean.hyperparameter_count(self.D))

    def _hyp_n(self):
        return sum(self._counts())

    def _slices(self):
        lo = 0
        for name, n in self._hyper_info():
            yield name, slice(lo, lo + n)
            lo += n

    def __str__(self):
        cov_n, noise_n, mean_n = self._counts()

        def plural(n):
            return f", {n} parameter\n" if n == 1 else f", {n} parameters\n"

        cov = "Covariance function: " + type(self.covariance).__name__
        if type(self.covariance).__name__ == "Matern":
            cov += "(degree=" + str(self.covariance.degree) + ")\n"
        cov += plural(cov_n)
        mean = "Mean function: " + type(self.mean).__name__ + plural(mean_n)
        noise = "Noise function: " + type(self.noise).__name__
        p = self.noise.parameters
        if np.any(p):
            flags = []
            if p[0] == 1:
                flags.append("constant_add=True")
            if p[1] == 1:
                flags.append("user_provided_add=True")
            if p[1] == 2:
                flags.append("scale_user_provided=True")
            if p[2] == 1:
                flags.append("rectified_linear_output_dependent_add=True")
            noise += "(" + ", ".join(flags) + ")"
        noise += plural(noise_n)
        priors = "Hyperparameter priors: " + ("none\n" if self.no_prior else "present\n")
        samples = "Hyperparameter samples: " + str(0 if self.posteriors is None else np.size(self.posteriors))
        body = "Dimension: " + str(self.D) + "\n" + cov + mean + noise + priors + samples
        return "GP:\n" + "".join("    " + ln for ln in body.splitlines(True))

    def __repr__(self):
        """``self.<attribute> = <summary>`` for the attributes the reference lists first
        (gaussian_process.py:64-80), then every other public attribute in sorted order; arrays with
        fewer than 10 elements are printed in full, larger ones by shape, dictionaries by identity
        (the layout of the reference's formatting.full_repr)."""
        def brief(v):
            if isinstance(v, np.ndarray):
                if np.prod(v.shape) < 10:
                    text = np.array2string(v, precision=4, suppress_small=True, separator=", ")
                    if "\n" in text:
                        text = "\n" + "\n".join("    " + ln for ln in text.splitlines())
                    return f"{text} : {type(v).__name__}"
                return f"{v.shape} {type(v).__name__}"
            if type(v) is dict:
                return object.__repr__(v)
            return repr(v)
        first = ["D", "covariance", "mean", "noise", "X", "y", "s2", "lower_bounds", "upper_bounds",
                 "posteriors"]
        rest = sorted(k for k in self.__dict__ if k not in first and not k.startswith("_"))
        lines = [f"self.{n} = {brief(getattr(self, n, None))}" for n in first + rest]
        return "GP:\n" + "\n".join("    " + ln for ln in ",\n".join(lines).splitlines())

    # ------------------------------------------------------------------ bounds
    def set_bounds(self, bounds=None):
        """gaussian_process.py:147-205"""
        n = self._hyp_n()
        lb, ub = np.full((n,), np.nan), np.full((n,), np.nan)
        for name, sl in self._slices():
            if bounds is None:
                continue
            if name not in bounds:
                raise ValueError("Missing hyperparameter " + name)
            if bounds[name] is not None:
                lb[sl], ub[sl] = bounds[name]
        self.lower_bounds, self.upper_bounds = lb, ub
        if self.no_prior is not None:
            self.__recompute_normalization_constants()

    def get_bounds(self):
        return self.bounds_to_dict(self.lower_bounds, self.upper_bounds)

    def bounds_to_dict(self, lower_bounds, upper_bounds):
        return {name: (lower_bounds[sl], upper_bounds[sl]) for name, sl in self._slices()}

    def get_recommended_bounds(self, lower_bounds=None, upper_bounds=None):
        """gaussian_process.py:245-359: NaN entries are replaced by the plugins' bounds."""
        if self.X is None or self.y is None:
            raise ValueError("GP does not have X or y set!")

        def resolve(v, current):
            if isinstance(v, (list, tuple, np.ndarray)):
                return np.array(v, dtype=float)
            if isinstance(v, str) and v == "current":
                return current.copy()
            if v is None or (isinstance(v, str) and v == "recommended"):
                return np.full_like(current, np.nan)
            raise ValueError("`lower_bounds` should be 'recommended'/`None`, 'current', or an array.")

        lb = resolve(lower_bounds, self.lower_bounds)
        ub = resolve(upper_bounds, self.upper_bounds)
        info = [self.covariance.get_bounds_info(self.X, self.y),
                self.noise.get_bounds_info(self.X, self.y),
                self.mean.get_bounds_info(self.X, self.y)]
        rec_lb = np.concatenate([i["LB"] for i in info])
        rec_ub = np.concatenate([i["UB"] for i in info])
        lb = np.where(np.isnan(lb), rec_lb, lb)
        ub = np.where(np.isnan(ub), rec_ub, ub)
        return self.bounds_to_dict(lb, np.maximum(lb, ub))

    # ------------------------------------------------------------------ priors
    def set_priors(self, priors=None):
        """gaussian_process.py:421-514"""
        n = self._hyp_n()
        hp = {k: np.full((n,), np.nan) for k in ("mu", "sigma", "df", "a", "b")}
        any_prior = False
        for name, sl in self._slices():
            if priors is None:
                continue
            if name not in priors:
                raise ValueError("Missing hyperparameter " + name)
            if priors[name] is None:
                continue
            any_prior = True
            kind, par = priors[name]
            if kind == "gaussian":
                hp["mu"][sl], hp["sigma"][sl] = par
                hp["df"][sl] = 0
            elif kind == "student_t":
                hp["mu"][sl], hp["sigma"][sl], hp["df"][sl] = par
            elif kind == "smoothbox":
                hp["a"][sl], hp["b"][sl], hp["sigma"][sl] = par
                hp["df"][sl] = 0
            elif kind == "smoothbox_student_t":
                hp["a"][sl], hp["b"][sl], hp["sigma"][sl], hp["df"][sl] = par
            else:
                raise ValueError("Unknown hyperprior type " + kind)
        self.hyper_priors = hp
        self.no_prior = not any_prior
        self.__recompute_normalization_constants()

    def get_priors(self):
        """gaussian_process.py:361-419"""
        hp = self.hyper_priors
        out = {}
        for name, sl in self._slices():
            mu, sigma, df, a, b = (hp[k][sl].copy() for k in ("mu", "sigma", "df", "a", "b"))
            val = None
            if np.all(np.isfinite(a)) and np.all(np.isfinite(b)) and np.all(np.isfinite(sigma)):
                if np.all(df == 0) or np.all(df == np.inf):
                    val = ("smoothbox", (a, b, sigma))
                elif np.all(df > 0):
                    val = ("smoothbox_student_t", (a, b, sigma, df))
            elif np.all(np.isfinite(mu)) and np.all(np.isfinite(sigma)):
                if np.all(df == 0) or np.all(df == np.inf):
                    val = ("gaussian", (mu, sigma))
                elif np.all(df > 0):
                    val = ("student_t", (mu, sigma, df))
            out[name] = val
        return out

    def __recompute_normalization_constants(self):
        """Mass of each prior inside [LB, UB] (gaussian_process.py:1234-1273)."""
        hp = self.hyper_priors
        nc = np.full(self.lower_bounds.shape, 1.0)
        for i in range(nc.size):
            mu, sigma, df = hp["mu"][i], np.abs(hp["sigma"])[i], hp["df"][i]
            a, b, lb, ub = hp["a"][i], hp["b"][i], self.lower_bounds[i], self.upper_bounds[i]
            if lb == ub or (not np.isfinite(lb) and not np.isfinite(ub)):
                continue
            if not np.isfinite(mu) and not np.isfinite(sigma):
                continue
            gaussian_tails = df == 0 or not np.isfinite(df)
            if np.isfinite(a) and np.isfinite(b):
                if gaussian_tails:
                    lo, hi = smoothbox_cdf(lb, sigma, a, b), smoothbox_cdf(ub, sigma, a, b)
                else:
                    lo = smoothbox_student_t_cdf(lb, df, sigma, a, b)
                    hi = smoothbox_student_t_cdf(ub, df, sigma, a, b)
            elif gaussian_tails:
                lo, hi = sp.stats.norm.cdf([lb, ub], loc=mu, scale=sigma)
            else:
                lo, hi = sp.stats.t.cdf([lb, ub], df, loc=mu, scale=sigma)
            nc[i] = hi - lo
        self.normalization_constants = nc

    def _log_priors_batch(self, hyp, compute_grad):
        """log prior (and gradient) for a (B, P) array of hyperparameter rows: the
        reference's per-vector formulas (gaussian_process.py:1275-1466) vectorised over B."""
        hyp = np.atleast_2d(np.asarray(hyp, dtype=float))
        B, P = hyp.shape
        hp = self.hyper_priors
        mu, sigma, df, a, b = hp["mu"], np.abs(hp["sigma"]), hp["df"], hp["a"], hp["b"]
        lb, ub = self.lower_bounds, self.upper_bounds
        fin = np.isfinite
        # NB the reference writes `df == 0 | ~np.isfinite(df)`, which Python parses as
        # `df == (0 | ~isfinite(df))`: true only where df == 0 (:1293, :1309).  Kept.
        df0 = df == 0
        fixed = lb == ub
        box = fin(a) & fin(b) & df0 & ~fin(mu) & fin(sigma)
        box_t = fin(a) & fin(b) & (df > 0) & ~fin(mu) & fin(sigma) & fin(df)
        flat = ~fin(mu) & ~fin(sigma)
        gauss = ~flat & ~box & df0 & fin(sigma)
        stud = ~flat & ~box_t & (df > 0) & fin(df)
        lp = np.zeros(B)
        dlp = np.zeros((B, P)) if compute_grad else None
        with np.errstate(all="ignore"):
            if np.any(fixed):
                lp[np.any(hyp[:, fixed] != lb[fixed], axis=1)] = -np.inf
                if compute_grad:
                    dlp[:, fixed] = np.nan
            for idx, heavy in ((box, False), (box_t, True)):
                if not np.any(idx):
                    continue
                h, s, aa, bb = hyp[:, idx], sigma[idx], a[idx], b[idx]
                below, above = h < aa, h > bb
                z2 = np.where(below, ((h - aa) / s) ** 2, 0.0) + np.where(above, ((h - bb) / s) ** 2, 0.0)
                out = below | above
                if heavy:
                    nu = df[idx]
                    C = 1.0 + (bb - aa) * sp.special.gamma(0.5 * (nu + 1)) / (
                        sp.special.gamma(0.5 * nu) * s * np.sqrt(nu * np.pi))
                    base = (sp.special.gammaln(0.5 * (nu + 1)) - sp.special.gammaln(0.5 * nu)
                            - 0.5 * np.log(np.pi * nu) - np.log(C * s))
                    lp += np.sum(base + np.where(out, -0.5 * (nu + 1) * np.log1p(z2 / nu), 0.0), axis=1)
                    if compute_grad:
                        edge = np.where(below, aa, bb)
                        g = -(nu + 1) / nu / (1 + z2 / nu) * (h - edge) / s ** 2
                        dlp[:, idx] = np.where(out, g, 0.0)
                else:
                    C = 1.0 + (bb - aa) / (s * np.sqrt(2 * np.pi))
                    t_out = -0.5 * (np.log(C ** 2 * 2 * np.pi * s ** 2) + z2)
                    t_in = -(np.log(C * s) + np.log(np.sqrt(2 * np.pi)))
                    lp += np.sum(np.where(out, t_out, t_in), axis=1)
                    if compute_grad:
                        edge = np.where(below, aa, bb)
                        dlp[:, idx] = np.where(out, -(h - edge) / s ** 2, 0.0)
            if np.any(gauss):
                h, s, m = hyp[:, gauss], sigma[gauss], mu[gauss]
                lp -= 0.5 * np.sum(np.log(2 * np.pi * s ** 2) + ((h - m) / s) ** 2, axis=1)
                if compute_grad:
                    dlp[:, gauss] = -(h - m) / s ** 2
            if np.any(stud):
                h, s, m, nu = hyp[:, stud], sigma[stud], mu[stud], df[stud]
                z2 = ((h - m) / s) ** 2
                lp += np.sum(sp.special.gammaln(0.5 * (nu + 1)) - sp.special.gammaln(0.5 * nu)
                             - 0.5 * np.log(np.pi * nu) - np.log(s)
                             - 0.5 * (nu + 1) * np.log1p(z2 / nu), axis=1)
                if compute_grad:
                    dlp[:, stud] = -(nu + 1) / nu / (1 + z2 / nu) * (h - m) / s ** 2
            lp -= np.sum(np.log(self.normalization_constants))
        return (lp, dlp) if compute_grad else lp

    def __compute_log_priors(self, hyp, compute_grad):
        out = self._log_priors_batch(np.asarray(hyp, dtype=float)[None, :], compute_grad)
        if compute_grad:
            return out[0][0], out[1][0]
        return out[0]

    # ------------------------------------------------------------------ hyperparameters
    def get_hyperparameters(self, as_array=False):
        if self.posteriors is None:
            hyp = np.full((1, self._hyp_n()), np.nan)
        else:
            hyp = np.stack([np.array(p.hyp, dtype=float) for p in self.posteriors])
        return hyp if as_array else self.hyperparameters_to_dict(hyp)

    def set_hyperparameters(self, hyp_new, compute_posterior=True):
        if isinstance(hyp_new, np.ndarray):
            if hyp_new.ndim == 1:
                hyp_new = np.reshape(hyp_new, (1, -1))
            if hyp_new.shape[1] != self._hyp_n():
                raise ValueError("Input hyperparameter array is the wrong shape!")
        else:
            hyp_new = self.hyperparameters_from_dict(hyp_new)
        self.update(hyp=hyp_new, compute_posterior=compute_posterior)

    def hyperparameters_to_dict(self, hyp_arr):
        hyp_arr = np.asarray(hyp_arr)
        if hyp_arr.ndim == 1:
            hyp_arr = np.reshape(hyp_arr, (1, -1))
        if hyp_arr.shape[1] != self._hyp_n():
            raise ValueError("Input hyperparameter array is the wrong shape!")
        return [{name: row[sl].copy() for name, sl in self._slices()} for row in hyp_arr]

    def hyperparameters_from_dict(self, hyp_dict_list):
        if isinstance(hyp_dict_list, dict):
            hyp_dict_list = [hyp_dict_list]
        out = np.zeros((len(hyp_dict_list), self._hyp_n()))
        for i, d in enumerate(hyp_dict_list):
            for name, sl in self._slices():
                out[i, sl] = d[name]
        return out

    # ------------------------------------------------------------------ device plumbing
    @property
    def engine(self):
        """The device's shared engine (one context per GPU, not per GP: a context sizes its workspace
        to the free device memory, so per-GP contexts would starve each other)."""
        if self._engine is None or self._engine._h is None:
            from .engine import get_engine
            self._engine = get_engine()
        return self._engine

    @staticmethod
    def _fingerprint(a):
        """Exact fingerprint of an array's bits: sum_i w_i * bits_i mod 2^64 with fixed odd
        weights.  Every element enters through a bijection of Z/2^64, so ANY in-place edit of a
        single element changes it, and an edit of several elements escapes only by a 2^-64
        coincidence -- at the cost of one pass over the data (~20 us for 5000 x 10)."""
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        n = a.size
        w = GP._fp_weights
        if w.size < n:
            rng = np.random.Generator(np.random.PCG64(0x9E3779B97F4A7C15))
            w = GP._fp_weights = rng.integers(0, 2 ** 63, size=max(n, 2 * w.size), dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        return int(np.dot(a.view(np.uint64), w[:n]))

    _fp_weights = np.zeros(0, dtype=np.uint64)

    def _sync_engine(self):
        """Upload (X, y, s2) when they changed since the last call, or when another GP used the
        shared engine in between.  The check reads every element (the reference reads self.X,
        self.y, self.s2 afresh on every evaluation, so in-place edits must be seen)."""
        eng = self.engine
        key = tuple((a.shape, self._fingerprint(a)) if isinstance(a, np.ndarray) else None
                    for a in (self.X, self.y, self.s2))
        if key != self._data_key or eng.data_owner is not self._token:
            X = np.ascontiguousarray(self.X, dtype=float)
            y = np.ascontiguousarray(self.y, dtype=float)
            s2 = None if self.s2 is None else np.ascontiguousarray(self.s2, dtype=float)
            sp_ = self._spec
            eng.set_model(sp_.cov_kind, sp_.degree, sp_.ard, sp_.mean_kind, sp_.noise_params)
            eng.set_data(X, y.reshape(-1), None if s2 is None else s2.reshape(-1))
            eng.data_owner = self._token
            self._data_key = key
        return eng

    def _convert_shapes(self, X, y, s2):
        """gaussian_process.py:2523-2565"""
        if X is None and y is None and s2 is None:
            return X, y, s2
        if X is not None:
            if X.ndim == 1:
                X = X[None, :]
            if X.ndim != 2:
                raise AssertionError("X need to be an array of shape (N, D)")
            N, D = X.shape
            if D != self.D:
                raise AssertionError(f"The dimension of input data {D}"
                                     f"doesn't match GP's input dimension {self.D}.")
        else:
            try:
                N, D = self.X.shape
            except AttributeError:
                raise AttributeError(f"self.X is not a numpy array, self.X = {self.X}")
        if y is not None:
            y = y.reshape(N, 1)
        if isinstance(s2, (float, int)):
            s2 = s2 * np.ones((N, 1))
        elif isinstance(s2, np.ndarray):
            s2 = s2.reshape(N, 1)
        elif s2 is not None:
            raise TypeError("s2 type need to be Union[np.ndarray, float, int, None].")
        return X, y, s2

    # ------------------------------------------------------------------ core numerics
    def _nlz_batch(self, hyp, compute_grad=False, compute_prior=False):
        """Batched ``__compute_nlZ``: (B, P) rows -> nlZ (B,) [, dnlZ (B, P)].  Raises the
        reference's LinAlgError if any row's Cholesky fails all 10 jitter retries."""
        hyp = np.atleast_2d(np.asarray(hyp, dtype=float))
        eng = self._sync_engine()
        world = _world_size()
        if world > 1 and hyp.shape[0] >= world and self.X.shape[0] >= self.SHARD_MIN_N:
            # one process per GPU, same batch on every rank (same seeds): each rank evaluates its
            # block of rows on the device and one all-gather returns all of them -- results are
            # bitwise those of a single-GPU run, because a row's value does not depend on the batch
            # it is in.  Design batches, lock-step L-BFGS starts and slice-sampling chains all come
            # through here, so they spread over the GPUs as soon as there is a row per rank.
            nlz, dnlz, _, status = sharded_nlz_device(eng, hyp, compute_grad)
        else:
            nlz, dnlz, _, status = eng.nlz_batch(hyp, want_grad=compute_grad)
        if status.any():
            raise sp.linalg.LinAlgError("Singular matrix for L Cholesky decomposition")
        if compute_prior:
            if compute_grad:
                lp, dlp = self._log_priors_batch(hyp, True)
                nlz, dnlz = nlz - lp, dnlz - dlp
            else:
                nlz = nlz - self._log_priors_batch(hyp, False)
        return (nlz, dnlz) if compute_grad else nlz

    def __compute_nlZ(self, hyp, compute_grad, compute_prior):
        """gaussian_process.py:1520-1538"""
        out = self._nlz_batch(np.asarray(hyp, dtype=float).reshape(1, -1), bool(compute_grad),
                              bool(compute_prior))
        if compute_grad:
            return out[0][0], out[1][0]
        return out[0]

    _compute_nlZ = __compute_nlZ          # alias named by BASELINE.json

    def __core_computation(self, hyp, compute_nlZ, compute_nlZ_grad):
        """gaussian_process.py:2357-2521: nlZ, (nlZ, dnlZ) or a Posterior."""
        if compute_nlZ:
            return self.__compute_nlZ(hyp, bool(compute_nlZ_grad), False)
        return self._posteriors_for(np.asarray(hyp, dtype=float).reshape(1, -1))[0][0]

    def _compute_posterior(self, hyp):
        """Alias named by BASELINE.json: the posterior record(s) for hyp (1-D or (B, P))."""
        hyp = np.asarray(hyp, dtype=float)
        posts, _ = self._posteriors_for(np.atleast_2d(hyp))
        return posts[0] if hyp.ndim == 1 else posts

    def _posteriors_for(self, hyp):
        eng = self._sync_engine()
        batch = eng.posterior_batch(hyp)
        for s in range(batch.count):
            if batch.fetch(s, "status") != 0:
                raise sp.linalg.LinAlgError("Singular matrix for L Cholesky decomposition")
        posts = np.empty((batch.count,), dtype=object)
        for s in range(batch.count):
            posts[s] = Posterior(hyp[s].copy(), None, None, None, None, None, _batch=batch, _index=s)
        return posts, batch

    def __gp_obj_fun(self, hyp, compute_grad, swap_sign):
        """gaussian_process.py:1540-1559"""
        out = self.__compute_nlZ(hyp, compute_grad, self.no_prior is not True)
        sign = -1 if swap_sign else 1
        if compute_grad:
            return sign * out[0], sign * out[1]
        return sign * out

    def log_likelihood(self, hyp, compute_grad=False):
        if isinstance(hyp, dict):
            hyp = self.hyperparameters_from_dict(hyp)
        out = self.__compute_nlZ(np.asarray(hyp, dtype=float).reshape(-1), compute_grad, False)
        return (-out[0], -out[1]) if compute_grad else -out

    def log_posterior(self, hyp, compute_grad=False):
        if isinstance(hyp, dict):
            hyp = self.hyperparameters_from_dict(hyp)
        out = self.__compute_nlZ(np.asarray(hyp, dtype=float).reshape(-1), compute_grad, True)
        return (-out[0], -out[1]) if compute_grad else -out

    # ------------------------------------------------------------------ update / clean
    def update(self, X_new=None, y_new=None, s2_new=None, hyp=None, compute_posterior=True):
        """Add data and/or replace the hyperparameter samples (gaussian_process.py:691-884).
        One new point without ``s2`` on a GP that already holds posteriors takes the reference's
        rank-one path (:737-844), done in place on the device for all samples at once; everything
        else rebuilds the posteriors of all samples in ONE batched GPU call."""
        X_new, y_new, s2_new = self._convert_shapes(X_new, y_new, s2_new)
        rank_one = (X_new is not None and y_new is not None and compute_posterior
                    and self.X is not None and self.y is not None
                    and X_new.shape[0] == 1 and y_new.shape[0] == 1 and s2_new is None)   # :738-748
        appended, unstable = False, ()
        if rank_one:
            appended, unstable = self._rank_one_append(X_new, y_new)
        if X_new is not None:
            self.X = X_new.copy() if self.X is None else np.concatenate((self.X, X_new))
        if y_new is not None:
            self.y = y_new.copy() if self.y is None else np.concatenate((self.y, y_new))
        if s2_new is not None:
            self.s2 = s2_new.copy() if self.s2 is None else np.concatenate((self.s2, s2_new))
        if appended:
            if len(unstable):
                # "Compute full update where rank-1 failed" (:864-868): only those samples, on the extended
                # data; the others keep their rank-one factors and jitter multipliers
                self._sync_engine().posterior_rebuild(self._post_batch, unstable)
                for s in unstable:
                    p = self.posteriors[s]
                    for k in Posterior._FIELDS:
                        p._have[k] = False
                        p._val[k] = None
            return
        if rank_one and self.posteriors is not None:
            # the rank-one branch keeps the current samples and ignores ``hyp`` (:864-868)
            hyp = np.array([p.hyp for p in self.posteriors], dtype=float)
        else:
            hyp = self.get_hyperparameters(as_array=True) if hyp is None else np.array(hyp, dtype=float)
        if hyp.ndim == 1:
            hyp = hyp.reshape(1, -1)
        if compute_posterior and self.X is not None and self.y is not None:
            self.posteriors, self._post_batch = self._posteriors_for(hyp)
        else:
            self._post_batch = None
            self.posteriors = np.empty((hyp.shape[0],), dtype=object)
            for i in range(hyp.shape[0]):
                self.posteriors[i] = Posterior(hyp[i, :], None, None, None, None, None)

    def _rank_one_append(self, X_new, y_new):
        """The device rank-one update -> (applied, unstable samples).  ``applied`` is False when the
        in-place update does not apply (the caller then rebuilds all samples in one batched call);
        samples whose update failed the reference's stability test (:784-798) were left untouched and
        are recomputed by the caller once the data is extended (:864-868)."""
        batch = self._device_batch()
        if batch is None:
            return False, ()
        status = batch.engine.posterior_append(batch, X_new[0], float(y_new[0, 0]))
        if status is None:
            return False, ()
        unstable = np.flatnonzero(status)
        for s in unstable:
            warnings.warn("Rank-one update of Cholesky factor unstable "
                          + f"for posterior {s}. Reverting to full update.", stacklevel=3)
        for p in self.posteriors:                 # alpha, sW, L changed on the device
            for k in ("alpha", "sW", "L"):
                p._have[k] = False
                p._val[k] = None
        return True, unstable

    def clean(self):
        """Drop the factors (gaussian_process.py:886-905); ``update()`` rebuilds them."""
        self.temporary_data = {}
        if self.posteriors is not None:
            for p in self.posteriors:
                p._batch = None
                for k in Posterior._FIELDS:
                    p._set(k, None)
        if self._post_batch is not None:
            self._post_batch.free()
            self._post_batch = None

    # ------------------------------------------------------------------ predict
    def predict(self, x_star, y_star=None, s2_star=None, add_noise=False, separate_samples=False,
                return_lpd=False):
        """Posterior mean and variance at x_star over all hyperparameter samples
        (gaussian_process.py:1663-1816); one GPU call for all samples and points."""
        x_star, y_star, s2_star = self._convert_shapes(x_star, y_star, s2_star)
        if return_lpd and y_star is None:
            raise ValueError("Cannot calculate log predictive density without y_star.")
        if self.y is None:
            return self._predict_prior(x_star, y_star, s2_star, add_noise, separate_samples, return_lpd)
        batch = self._device_batch()
        if batch is None:
            raise RuntimeError("GP.predict: the posteriors hold no factors; call "
                               "update(compute_posterior=True) first")
        ys = None if y_star is None else y_star.reshape(-1)
        s2s = None if s2_star is None else s2_star.reshape(-1)

        def run(lo, hi):
            return self.engine.predict(batch, x_star[lo:hi], None if ys is None else ys[lo:hi],
                                       None if s2s is None else s2s[lo:hi], add_noise=add_noise,
                                       separate=separate_samples, want_lpd=return_lpd)
        world = _world_size()
        if world > 1 and x_star.shape[0] >= 4096 * world:
            # test points sharded over the GPUs (every rank holds all posterior samples)
            if ys is None and s2s is None and not return_lpd:
                return sharded_predict_device(self.engine, batch, x_star, add_noise, separate_samples)
            return sharded_rows(run, x_star.shape[0])
        return run(0, x_star.shape[0])

    def _predict_prior(self, x_star, y_star, s2_star, add_noise, separate, return_lpd):
        """GP without training data: prior mean and variance through the plugin kernels
        (gaussian_process.py:1765-1767)."""
        cov_n, noise_n, mean_n = self._counts()
        s_N = self.posteriors.size
        M = x_star.shape[0]
        mu, s2, ys2 = np.zeros((M, s_N)), np.zeros((M, s_N)), np.zeros((M, s_N))
        for s, post in enumerate(self.posteriors):
            h = np.asarray(post.hyp, dtype=float)
            mu[:, s] = np.reshape(self.mean.compute(h[cov_n + noise_n:cov_n + noise_n + mean_n], x_star), -1)
            s2[:, s] = np.maximum(self.covariance.compute(h[:cov_n], x_star, compute_diag=True)[:, 0], 0)
            if return_lpd or add_noise:
                sn2 = self.noise.compute(h[cov_n:cov_n + noise_n], x_star, y_star, s2_star)
                mult = post.sn2_mult if post.sn2_mult is not None else 1
                ys2[:, s] = s2[:, s] + np.reshape(sn2 * mult, -1)
        lpd = None
        if return_lpd and separate:
            lpd = -0.5 * (y_star - mu) ** 2 / ys2 - 0.5 * np.log(2 * np.pi * ys2)
        if add_noise:
            s2 = ys2
        if not separate:
            v = 0
            if s_N > 1:
                mbar = mu.sum(1, keepdims=True) / s_N
                v = np.sum((mu - mbar) ** 2, 1) / (s_N - 1)
                s2 = np.reshape(s2.sum(1) / s_N + v, (-1, 1))
                mu = mbar
            if return_lpd:
                pv = s2 if add_noise else np.reshape(ys2.sum(1) / s_N + v, (-1, 1))
                lpd = -0.5 * (y_star - mu) ** 2 / pv - 0.5 * np.log(2 * np.pi * pv)
        return (mu, s2, lpd) if return_lpd else (mu, s2)

    # ------------------------------------------------------------------ fit
    def fit(self, X=None, y=None, s2=None, hyp0=None, options=None):
        """Train the hyperparameters (gaussian_process.py:910-1232): space-filling design
        (evaluated as ONE batch on the GPU), L-BFGS-B from the best ``opts_N`` starts, then
        slice sampling; finally the posteriors of all kept samples are built in one batch."""
        options = options or {}
        opts_N = options.get("opts_N", 3)
        init_N = options.get("init_N", 2 ** 10)
        init_method = options.get("init_method", "sobol")
        thin = options.get("thin", 5)
        df_base = options.get("df_base", 7)
        widths = options.get("widths", None)
        tol_opt = options.get("tol_opt", 1e-5)
        tol_opt_mcmc = options.get("tol_opt_mcmc", 1e-3)
        sampler_name = options.get("sampler", "slicesample")
        s_N = options.get("n_samples", 10)
        burn_in = options.get("burn", thin * s_N)
        lower_bounds = options.get("lower_bounds", "current")
        upper_bounds = options.get("upper_bounds", "current")
        # extra, optional keys (defaults reproduce the reference's sequential drivers):
        n_chains = int(options.get("n_chains", 1))           # slice-sampling chains in lock step
        lockstep_opt = options.get("lockstep_opt", _SCIPY_LBFGSB_REENTRANT)   # batch the L-BFGS-B runs

        X, y, s2 = self._convert_shapes(X, y, s2)
        if X is not None:
            self.X = X
        if y is not None:
            self.y = y
        if s2 is not None:
            self.s2 = s2
        cov_n, noise_n, _ = self._counts()
        info = [self.covariance.get_bounds_info(self.X, self.y),
                self.noise.get_bounds_info(self.X, self.y),
                self.mean.get_bounds_info(self.X, self.y)]
        self.hyper_priors["df"][np.isnan(self.hyper_priors["df"])] = df_base

        current = (isinstance(lower_bounds, str) and lower_bounds == "current"
                   and isinstance(upper_bounds, str) and upper_bounds == "current")
        if current and (np.any(np.isnan(self.lower_bounds)) or np.any(np.isnan(self.upper_bounds))):
            self.set_bounds(self.get_recommended_bounds(self.lower_bounds, self.upper_bounds))
        else:
            self.set_bounds(self.get_recommended_bounds(lower_bounds, upper_bounds))
        LB, UB = self.lower_bounds, self.upper_bounds
        PLB = np.concatenate([i["PLB"] for i in info])
        PUB = np.concatenate([i["PUB"] for i in info])
        PLB = np.minimum(np.maximum(PLB, LB), UB)
        PUB = np.maximum(np.minimum(PUB, UB), LB)

        if hyp0 is None:
            if self.posteriors is not None:
                hyp0 = self.get_hyperparameters(as_array=True)
            else:
                hyp0 = np.reshape(np.minimum(np.maximum((PLB + PUB) / 2, LB), UB), (1, -1))
        elif isinstance(hyp0, dict):
            hyp0 = self.hyperparameters_from_dict(hyp0)
        hyp0 = np.atleast_2d(hyp0)
        use_prior = self.no_prior is not True

        def design_objective(H):                     # whole design in one GPU batch
            return self._nlz_batch(H, False, use_prior)
        design_objective.batched = True

        tol = tol_opt_mcmc if (s_N > 0 and sampler_name != "laplace") else tol_opt
        if init_N > 0:
            X0, y0 = f_min_fill(design_objective, hyp0, LB, UB, PLB, PUB, self.hyper_priors,
                                init_N, init_method)
            hyp = X0[0:np.maximum(opts_N, 1), :]
            if noise_n > 0 and 1 < opts_N < init_N:
                # second start: best point among the 20% lowest-noise design rows (:1112-1125)
                rest, rest_y = X0[opts_N:, :], y0[opts_N:]
                order = np.argsort(rest[:, cov_n])
                rest, rest_y = rest[order, :], rest_y[order]
                hyp[1, :] = rest[np.argmin(rest_y[0:math.ceil(0.2 * np.size(rest_y))]), :]
            widths_default = np.std(X0, axis=0, ddof=1) if init_N > 1 else np.zeros(shape=PLB.shape)
        else:
            nll = design_objective(hyp0)
            hyp = hyp0[np.argsort(nll), :]
            widths_default = PUB - PLB
        zero = widths_default == 0
        if np.any(zero):
            if np.shape(hyp)[0] > 1:
                widths_default[zero] = np.std(hyp, axis=0, ddof=1)[zero]
                zero = widths_default == 0
            if np.any(zero):
                widths_default[zero] = np.minimum(1, UB[zero] - LB[zero])

        # keep the starts strictly inside the box (:1159-1166)
        lo, hi = np.reshape(LB.copy(), (1, -1)), np.reshape(UB.copy(), (1, -1))
        free = lo != hi
        lo_i, hi_i = free & np.isfinite(lo), free & np.isfinite(hi)
        lo[lo_i] = np.nextafter(lo[lo_i], np.inf)
        hi[hi_i] = np.nextafter(hi[hi_i], -np.inf)
        hyp = np.minimum(hi, np.maximum(lo, hyp))

        def opt_objective(h):
            return self.__gp_obj_fun(h, True, False)

        nll = np.full((np.maximum(opts_N, 1),), np.inf)
        results = []
        opts_N = np.minimum(opts_N, hyp.shape[0])
        if lockstep_opt and opts_N > 1:
            # the opts_N runs are independent: advance them together, one batched nlZ+gradient
            # evaluation per round (same iterates as running them one after another)
            results = minimize_lockstep(lambda H: self._nlz_batch(H, True, use_prior), hyp[:opts_N, :],
                                        list(zip(LB, UB)), tol)
            for i, res in enumerate(results):
                hyp[i, :] = res.x
                nll[i] = res.fun
        else:
            for i in range(opts_N):
                res = sp.optimize.minimize(fun=opt_objective, x0=hyp[i, :], jac=True,
                                           bounds=list(zip(LB, UB)), tol=tol)
                results.append(res)
                hyp[i, :] = res.x
                nll[i] = res.fun
        if opts_N > 0:
            optimize_result = results[np.argmin(nll)]
            hyp_start = hyp[np.argmin(nll), :].copy()
        else:
            optimize_result = None
            hyp_start = hyp[0, :].copy()
        if s_N == 0:
            hyp_start = np.reshape(hyp_start, (1, -1))
            self.update(hyp=hyp_start)
            return hyp_start, optimize_result, None

        if sampler_name != "slicesample":
            raise ValueError("Unknown sampler!")
        widths = widths_default if widths is None else np.minimum(widths, widths_default)
        if n_chains > 1:
            # K independent chains from the optimum, advanced in lock step (batch of K per round)
            per_chain = -(-s_N // n_chains)
            mc = MultiChainSliceSampler(lambda H: -self._nlz_batch(H, False, use_prior), hyp_start, widths,
                                        LB, UB, n_chains)
            sampling_result = mc.sample(per_chain * thin, burn=burn_in)
            kept = sampling_result["samples"][:, thin - 1::thin, :]            # (K, per_chain, P)
            hyp = np.transpose(kept, (1, 0, 2)).reshape(-1, kept.shape[2])[:s_N]
            self.update(hyp=hyp)
            return hyp, optimize_result, sampling_result
        # speculative shrinking: the next few proposals of a coordinate update go to the GPU as one
        # batch (same chain as the sequential sampler, fewer and better-filled calls)
        slicer = SliceSampler(lambda h: self.__gp_obj_fun(h, False, True), hyp_start, widths, LB, UB,
                              {"display": "off", "diagnostics": False,
                               "log_f_batch": lambda H: -self._nlz_batch(H, False, use_prior),
                               "speculate": self._speculation_depths(options.get("speculate"))})
        sampling_result = slicer.sample(s_N * thin, burn=burn_in)
        hyp = sampling_result["samples"][thin - 1::thin, :]
        self.update(hyp=hyp)
        return hyp, optimize_result, sampling_result

    def _speculation_depths(self, spec):
        """Proposals per batched call and coordinate for the slice sampler.  A move along a MEAN
        hyperparameter re-uses the cached factor (O(N^2), latency-bound: a deeper batch costs the
        same), one along a covariance / noise hyperparameter refactors every row of the batch."""
        cov_N, noise_N, mean_N = self._counts()
        if spec is None:
            spec = (2, 4)
        if np.ndim(spec) == 0:
            return int(spec)
        spec = np.asarray(spec, dtype=int)
        if spec.size == 2:
            return np.concatenate((np.full(cov_N + noise_N, spec[0]), np.full(mean_N, spec[1])))
        return spec

    # ------------------------------------------------------------------ not in this round
    def _not_built(self, name):
        raise NotImplementedError(
            f"GP.{name} is outside the hot path built so far (SURVEY.md 8f 'next' rows)")

    def quad(self, mu, sigma, compute_var=False, separate_samples=False):
        """Bayesian quadrature of the GP against Gaussian measures N(mu, diag(sigma^2))
        (gaussian_process.py:1818-1981; squared-exponential kernel only): one GPU call for
        all measures and all hyperparameter samples."""
        from .covariance_functions import SquaredExponential
        if not isinstance(self.covariance, SquaredExponential):
            raise ValueError("Bayesian quadrature only supports the squared exponential kernel.")
        D = self.D
        mu = np.tile(mu, (1, D)) if np.size(mu) == 1 else np.atleast_2d(np.asarray(mu, dtype=float))
        sigma = np.tile(sigma, (1, D)) if np.size(sigma) == 1 else np.atleast_2d(np.asarray(sigma, dtype=float))
        sigma = np.ascontiguousarray(np.broadcast_to(sigma, mu.shape))
        batch = self._device_batch()
        if batch is None:
            raise RuntimeError("GP.quad: the posteriors hold no factors; call "
                               "update(compute_posterior=True) first")
        return self.engine.quad(batch, mu, sigma, compute_var=compute_var, separate=separate_samples)

    def predict_full(self, x_star, y_star=None, s2_star=None, add_noise=False):
        """Posterior mean (M, Ns) and full covariance (M, M, Ns) at x_star for every
        hyperparameter sample (gaussian_process.py:1561-1661)."""
        x_star, y_star, s2_star = self._convert_shapes(x_star, y_star, s2_star)
        if self.y is None:
            return self._predict_full_prior(x_star, y_star, s2_star, add_noise)
        batch = self._device_batch()
        if batch is None:
            raise RuntimeError("GP.predict_full: the posteriors hold no factors; call "
                               "update(compute_posterior=True) first")
        return self.engine.predict_full(batch, x_star, None if y_star is None else y_star.reshape(-1),
                                        None if s2_star is None else s2_star.reshape(-1),
                                        add_noise=add_noise)

    def _predict_full_prior(self, x_star, y_star, s2_star, add_noise):
        """GP without training data: prior mean and covariance through the plugin kernels
        (gaussian_process.py:1621-1624, :1649-1659)."""
        cov_n, noise_n, mean_n = self._counts()
        s_N, M = self.posteriors.size, x_star.shape[0]
        mu, cov = np.zeros((M, s_N)), np.zeros((s_N, M, M))
        for s, post in enumerate(self.posteriors):
            h = np.asarray(post.hyp, dtype=float)
            mu[:, s] = np.reshape(self.mean.compute(h[cov_n + noise_n:cov_n + noise_n + mean_n], x_star), -1)
            C = self.covariance.compute(h[:cov_n], x_star)
            cov[s] = (C + C.T) / 2
            if add_noise:
                mult = post.sn2_mult if post.sn2_mult is not None else 1
                sn2 = self.noise.compute(h[cov_n:cov_n + noise_n], x_star, y_star, s2_star)
                cov[s] += np.dot(np.eye(M), sn2) * mult
        return mu, cov.transpose(1, 2, 0)

    def random_function(self, X_star, add_noise=False):
        """Draw one function from the GP (prior if there is no data, else posterior) at
        X_star (gaussian_process.py:2241-2329).  The mean and covariance come from the GPU
        path; the (M, M) factorisation for the draw is a small host operation.  Consumes the
        global NumPy RNG in the reference's order."""
        X_star = np.atleast_2d(np.asarray(X_star, dtype=float))
        M = X_star.shape[0]
        cov_n, noise_n, mean_n = self._counts()
        s = np.random.randint(0, np.size(self.posteriors))
        hyp = np.asarray(self.posteriors[s].hyp, dtype=float)
        if self.y is None:
            f_mu = np.reshape(self.mean.compute(hyp[cov_n + noise_n:cov_n + noise_n + mean_n], X_star), (-1, 1))
            C = self.covariance.compute(hyp[:cov_n], X_star) + np.spacing(1) * np.eye(M)
        else:
            mu, cov = self.predict_full(X_star)
            f_mu, C = mu[:, s:s + 1], cov[:, :, s]
        C = (C + C.T) / 2
        Tm = _robust_cholesky(C)
        f_star = np.dot(Tm.T, np.random.standard_normal((Tm.shape[0], 1))) + f_mu
        if not add_noise:
            return f_star
        sn2 = self.noise.compute(hyp[cov_n:cov_n + noise_n], X_star, None, None)
        mult = self.posteriors[s].sn2_mult
        mult = 1 if mult is None else mult
        return f_star + np.sqrt(sn2 * mult) * np.random.standard_normal(size=f_mu.shape)

    def plot(self, *a, **k):
        self._not_built("plot")
