"""Coordinate-wise slice sampling with bounds and burn-in width adaptation, with the
interface of the reference's ``SliceSampler`` (gpyreg/slice_sample.py).  One chain is
strictly sequential (every proposal depends on the last accepted point), so this stays a
host-side driver; each log-density evaluation is one B=1 call into the GPU path.

The global NumPy RNG is consumed in the same order as the reference (one shuffle per
sweep; per coordinate one draw for the slice level, one for the bracket position, one per
shrink proposal), so a fixed ``np.random.seed`` gives the same chain whenever the
log-density values agree.

Speculative shrinking (``options["log_f_batch"]``, ``options["speculate"] = k``): the next
proposals of the shrink loop are a deterministic function of the random numbers still to be
drawn (if proposal 1 is rejected the bracket shrinks to a known side, proposal 2 follows, ...).
With a batched log density the sampler evaluates the next k proposals in ONE call (k may differ
per coordinate) and accepts the first that clears the slice level; the RNG state is rewound so that exactly as many draws
are consumed as the sequential algorithm would have used.  The chain is identical, bit for
bit; only the number of (batched) calls drops, from ~2 per coordinate to ~1.
"""
import ctypes
import logging

import numpy as np


class _RngRewind:
    """Save / restore the global NumPy RNG (legacy ``np.random``) around a speculative draw.
    ``np.random.get_state`` + ``set_state`` cost ~100 us per round trip, more than a small GP
    evaluation; copying the 2.5 KB Mersenne-Twister state of the global bit generator directly takes
    ~3 us.  Only uniform draws happen between save and restore, so the Gaussian cache of the
    legacy generator is not involved.  The raw path is self-checked once against the public API
    and silently replaced by it if NumPy's internals ever differ."""

    _SIZE = 624 * 4 + 4          # mt19937_state: uint32 key[624]; int pos

    def __init__(self):
        self._raw = self._probe()
        self._buf = ctypes.create_string_buffer(self._SIZE)
        self._state = None

    @staticmethod
    def _address():
        bg = np.random.mtrand._rand._bit_generator
        if type(bg).__name__ != "MT19937":
            raise TypeError("global bit generator is not MT19937")
        return bg.ctypes.state_address

    def _probe(self):
        try:
            public = np.random.get_state()
            try:
                buf = ctypes.create_string_buffer(self._SIZE)
                ctypes.memmove(buf, self._address(), self._SIZE)
                a = np.random.rand(700)                          # crosses a state regeneration
                ctypes.memmove(self._address(), buf, self._SIZE)
                b = np.random.rand(700)
                np.random.set_state(public)
                c = np.random.rand(700)
                return bool(np.array_equal(a, b) and np.array_equal(a, c))
            finally:
                np.random.set_state(public)
        except Exception:
            return False

    def save(self):
        if self._raw:
            try:
                ctypes.memmove(self._buf, self._address(), self._SIZE)
                return
            except Exception:
                self._raw = False
        self._state = np.random.get_state()

    def restore(self):
        if self._raw:
            ctypes.memmove(self._address(), self._buf, self._SIZE)
        else:
            np.random.set_state(self._state)


class SliceSampler:
    def __init__(self, log_f, x0, widths=None, LB=None, UB=None, options=None):
        self.log_f = log_f
        self.x0 = np.atleast_1d(np.asarray(x0, dtype=float)).copy() if np.ndim(x0) <= 1 else np.asarray(x0)
        if np.ndim(self.x0) > 1:
            raise ValueError("The initial point x0 needs to be a scalar or a 1D array")
        D = self.x0.size

        def as_bound(v, default):
            if v is None:
                return np.full((D,), default)
            v = np.asarray(v, dtype=float)
            return np.tile(v, D) if v.size == 1 else v.copy()

        self.LB = as_bound(LB, -np.inf)
        self.UB = as_bound(UB, np.inf)
        self.LB_out = np.nextafter(self.LB, -np.inf)
        self.UB_out = np.nextafter(self.UB, np.inf)
        if widths is None:
            self.widths = ((self.UB - self.LB) / 2).copy()
            self.base_widths = None
        else:
            widths = np.asarray(widths)
            if not np.iscomplexobj(widths):          # complex widths are rejected below, like the reference
                widths = widths.astype(float)
            self.widths = np.tile(widths, D) if widths.size == 1 else widths.copy()
            self.base_widths = self.widths.copy()
        self.widths[np.isinf(self.widths)] = 10
        self.widths[self.LB == self.UB] = 1          # irrelevant for pinned coordinates
        if np.shape(self.LB) != np.shape(self.x0) or np.shape(self.UB) != np.shape(self.x0):
            raise ValueError("LB and UB need to be None, scalars, or 1D arrays of "
                             "the same size as X0.")
        if not np.all(self.UB >= self.LB):
            raise ValueError("All upper bounds UB need to be equal or greater than "
                             "lower bounds LB.")
        if np.any(self.widths <= 0) or np.any(~np.isfinite(self.widths)) or np.any(~np.isreal(self.widths)):
            raise ValueError("The widths vector needs to be all positive real numbers.")
        if np.any(self.x0 < self.LB) or np.any(self.x0 > self.UB):
            raise ValueError("The initial starting point X0 is outside the bounds.")
        self.func_count = 0
        options = options or {}
        self.step_out = options.get("step_out", False)
        self.display = options.get("display", "full")
        self.adaptive = options.get("adaptive", True)
        self.log_prior = options.get("log_prior", None)
        self.diagnostics = options.get("diagnostics", True)
        self.log_f_batch = options.get("log_f_batch", None)
        # proposals evaluated per batched call: one number, or one per coordinate (a caller that
        # knows which coordinates are cheap to evaluate in a batch can speculate deeper there)
        spec = options.get("speculate", 3 if self.log_f_batch is not None else 1)
        self.speculate = np.broadcast_to(np.asarray(spec, dtype=int), (D,)).copy()
        if np.any(self.speculate < 1):
            raise ValueError("The speculate option needs to be a positive integer (or one per coordinate).")
        self.batch_calls = 0
        self._rewind = _RngRewind()
        self.shrink_counts = np.zeros((D, 65), dtype=np.int64)   # [coordinate][proposals needed]
        self.logger = logging.getLogger("SliceSampler")
        self.logger.setLevel({"off": logging.WARN, "summary": logging.INFO}.get(self.display, logging.DEBUG))

    # -- target density with bounds ------------------------------------------------
    def _log_density(self, x):
        """log p(x) (+ log prior), -inf outside the bounds or where p is NaN/-inf
        (slice_sample.py:653-683)."""
        if np.any(x < self.LB) or np.any(x > self.UB):
            return -np.inf, np.nan, -np.inf
        lp = 0.0
        if self.log_prior is not None:
            lp = self.log_prior(x)
            if np.isnan(lp):
                self.logger.warning("Prior density function returned NaN. Trying to continue.")
                return -np.inf, np.nan, -np.inf
            if not np.isfinite(lp):
                return -np.inf, np.nan, -np.inf
        f_val = self.log_f(x)
        self.func_count += 1
        f_sum = np.sum(f_val)
        if np.any(np.isnan(f_val)):
            self.logger.warning("Target density function returned NaN. Trying to continue.")
            return -np.inf, f_val, lp
        return f_sum + lp, f_val, lp

    def _log_density_many(self, pts, d=None, inside=None):
        """Batched log density of the rows of pts (no prior): -inf outside the bounds / for NaN.
        When only coordinate d differs between the rows the caller passes the bounds mask it
        already has (the other coordinates are inside by construction)."""
        if inside is None:
            inside = np.all((pts >= self.LB) & (pts <= self.UB), axis=1)
        n = pts.shape[0]
        n_in = int(np.count_nonzero(inside))
        if n_in == 0:
            return np.full((n,), -np.inf), inside
        vals = np.asarray(self.log_f_batch(pts if n_in == n else pts[inside]), dtype=float).reshape(-1)
        self.batch_calls += 1
        nan = vals != vals
        if nan.any():
            self.logger.warning("Target density function returned NaN. Trying to continue.")
            vals = np.where(nan, -np.inf, vals)
        if n_in == n:
            return vals, inside
        out = np.full((n,), -np.inf)
        out[inside] = vals
        return out, inside

    # -- sampling --------------------------------------------------------------------
    def sample(self, N, thin=1, burn=None):
        xx = self.x0.astype(float).copy()
        D = xx.size
        if burn is None:
            burn = 0 if self.func_count > 0 else round(N / 3)
        if not np.isscalar(thin) or thin <= 0:
            raise ValueError("The thinning factor option needs to be a positive integer.")
        if not np.isscalar(burn) or burn < 0:
            raise ValueError("The burn-in samples option needs to be a non-negative integer.")
        n_sweeps = N + (N - 1) * (thin - 1) + burn
        samples = np.zeros((N, D))
        log_Px, f_val, log_prior = self._log_density(xx)
        if np.any(~np.isfinite(log_Px)):
            raise ValueError("The initial starting point X0 needs to evaluate to a "
                             "real number (not Inf or NaN).")
        f_vals = np.zeros((N, np.size(f_val)))
        log_priors = np.zeros((N,))
        s1, s2 = np.zeros((D,)), np.zeros((D,))     # running moments over the 2nd half of burn-in
        perm = np.arange(D)
        for it in range(n_sweeps):
            lo, hi, prop = xx.copy(), xx.copy(), xx.copy()
            np.random.shuffle(perm)
            for d in perm:
                if self.LB[d] == self.UB[d]:
                    continue
                level = log_Px + np.log(np.random.rand())          # slice height
                u = np.random.rand()                               # random bracket placement
                lo[d] = np.fmax(lo[d] - u * self.widths[d], self.LB_out[d])
                hi[d] = np.fmin(hi[d] + (1 - u) * self.widths[d], self.UB_out[d])
                if self.step_out:
                    while self._log_density(lo)[0] > level:
                        lo[d] -= self.widths[d]
                    while self._log_density(hi)[0] > level:
                        hi[d] += self.widths[d]
                n_shrink = 0
                k_spec = int(self.speculate[d])
                speculative = (self.log_f_batch is not None and k_spec > 1
                               and self.log_prior is None)
                while True:                                        # shrink until accepted
                    if speculative:
                        # the next k proposals, assuming each one before is rejected
                        self._rewind.save()
                        us = np.random.rand(k_spec)
                        self._rewind.restore()
                        l_, h_, cands = lo[d], hi[d], []
                        for u in us:
                            c = u * (h_ - l_) + l_
                            cands.append(c)
                            if c > xx[d]:
                                h_ = c
                            elif c < xx[d]:
                                l_ = c
                            else:
                                break
                        pts = np.empty((len(cands), D))
                        pts[:] = prop
                        pts[:, d] = cands
                        # prop is inside the bounds in every other coordinate: check d only
                        lb_d, ub_d = self.LB[d], self.UB[d]
                        inside = np.array([lb_d <= c <= ub_d for c in cands])
                        vals, inside = self._log_density_many(pts, d, inside)
                        stop = False
                        for c, v, ins in zip(cands, vals, inside):
                            n_shrink += 1
                            np.random.rand()                       # the draw the sequential sampler makes
                            prop[d] = c
                            log_Px, f_val, log_prior = v, (v if ins else np.nan), (0.0 if ins else -np.inf)
                            self.func_count += int(ins)
                            if v > level:
                                stop = True
                                break
                            if c > xx[d]:
                                hi[d] = c
                            elif c < xx[d]:
                                lo[d] = c
                            else:
                                self.logger.warning("WARNING: Shrunk to current position and still "
                                                    " not acceptable!")
                                stop = True
                                break
                        if stop:
                            break
                        continue
                    n_shrink += 1
                    prop[d] = np.random.rand() * (hi[d] - lo[d]) + lo[d]
                    log_Px, f_val, log_prior = self._log_density(prop)
                    if log_Px > level:
                        break
                    if prop[d] > xx[d]:
                        hi[d] = prop[d]
                    elif prop[d] < xx[d]:
                        lo[d] = prop[d]
                    else:
                        self.logger.warning("WARNING: Shrunk to current position and still "
                                            " not acceptable!")
                        break
                self.shrink_counts[d, min(n_shrink, 64)] += 1
                if it < burn and self.adaptive:                    # width adaptation
                    span = self.UB[d] - self.LB[d]
                    if n_shrink > 3:
                        floor = np.abs(np.spacing(span)) if np.isfinite(span) else np.spacing(1)
                        self.widths[d] = np.maximum(self.widths[d] / 1.1, floor)
                    elif n_shrink < 2:
                        self.widths[d] = np.minimum(self.widths[d] * 1.2, span)
                xx[d] = prop[d]
                lo[d] = hi[d] = xx[d]
            if it >= burn and (it - burn) % thin == 0:
                k = (it - burn) // thin
                samples[k, :] = xx
                f_vals[k, :] = f_val
                log_priors[k] = log_prior
            if burn / 2 <= it < burn:
                s1 += xx
                s2 += xx ** 2
                if it == burn - 1 and self.adaptive:
                    n_stored = np.floor(burn / 2)
                    new_w = np.fmin(5 * np.sqrt(np.maximum(s2 / n_stored - (s1 / n_stored) ** 2, 0)),
                                    self.UB_out - self.LB_out)
                    if self.base_widths is None:
                        self.widths = new_w
                    else:
                        self.widths = np.maximum(new_w, np.sqrt(new_w * self.base_widths))
        self.x0 = xx
        R = eff_N = None
        exit_flag = 0
        if self.diagnostics:
            exit_flag, R, eff_N = _diagnose(samples)
        return {"samples": samples, "f_vals": f_vals, "exit_flag": exit_flag,
                "log_priors": log_priors, "R": R, "eff_N": eff_N}


def _diagnose(samples):
    """Split-chain potential scale reduction and a crude effective sample size."""
    N, D = samples.shape
    if N < 8:
        return 0, None, None
    h = N // 2
    a, b = samples[:h], samples[h:2 * h]
    W = 0.5 * (a.var(0, ddof=1) + b.var(0, ddof=1))
    Bv = h * np.var(np.stack([a.mean(0), b.mean(0)]), axis=0, ddof=1)
    with np.errstate(all="ignore"):
        R = np.sqrt(((h - 1) / h * W + Bv / h) / W)
        x = samples - samples.mean(0)
        rho1 = np.sum(x[1:] * x[:-1], 0) / np.sum(x * x, 0)
        eff_N = N * (1 - rho1) / (1 + rho1)
    flag = 1
    if np.any(R > 1.5):
        flag = -3
    elif np.any(R > 1.1):
        flag = -2
    elif np.any(eff_N < N / 10.0):
        flag = -1
    return flag, R, eff_N
