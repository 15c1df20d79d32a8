"""Batched inference drivers around the GPU path.

The reference drives the numerical core one hyperparameter vector at a time: opts_N
independent L-BFGS-B runs one after another (gaussian_process.py:1177-1187) and ONE slice
sampling chain whose proposals are strictly sequential (slice_sample.py:437-457).  Both
leave the GPU at batch size 1.  Here

* :func:`minimize_lockstep` runs the independent L-BFGS-B runs concurrently and gathers the
  objective requests they have pending into one batched evaluation per round;
* :class:`MultiChainSliceSampler` advances K independent slice-sampling chains in lock step:
  every round each chain has exactly one pending proposal, so a round is one batch of K.

Neither changes what a single run / chain computes: with one start or one chain they reduce
to the reference's sequential algorithm.
"""
import threading

import numpy as np
import scipy.optimize


def minimize_lockstep(fun_batch, x0s, bounds, tol):
    """L-BFGS-B from every row of x0s (scipy.optimize.minimize, jac=True, same arguments as
    gaussian_process.py:1178-1184), evaluated in lock step.

    fun_batch(H (b, P)) -> (f (b,), g (b, P)).  Returns the list of OptimizeResult."""
    x0s = np.atleast_2d(np.asarray(x0s, dtype=float))
    n = x0s.shape[0]
    if n == 0:
        return []
    lock = threading.Condition()
    pending = {}           # run index -> x awaiting evaluation
    answers = {}           # run index -> (f, g)
    state = {"active": n, "error": None}
    results = [None] * n

    def objective(i):
        def f(x):
            with lock:
                pending[i] = np.array(x, dtype=float)
                lock.notify_all()
                while i not in answers and state["error"] is None:
                    lock.wait()
                if state["error"] is not None:
                    raise state["error"]
                return answers.pop(i)
        return f

    def worker(i):
        try:
            results[i] = scipy.optimize.minimize(fun=objective(i), x0=x0s[i], jac=True, bounds=bounds,
                                                 tol=tol)
        except BaseException as e:          # propagate to the coordinator
            with lock:
                if state["error"] is None:
                    state["error"] = e
        finally:
            with lock:
                state["active"] -= 1
                lock.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    while True:
        with lock:
            while state["active"] > 0 and len(pending) < state["active"] and state["error"] is None:
                lock.wait()
            if state["error"] is not None or state["active"] == 0:
                break
            idx = sorted(pending)
            H = np.stack([pending.pop(i) for i in idx])
        try:
            f, g = fun_batch(H)
        except BaseException as e:
            with lock:
                state["error"] = e
                lock.notify_all()
            break
        with lock:
            for r, i in enumerate(idx):
                answers[i] = (float(f[r]), np.array(g[r], dtype=float))
            lock.notify_all()
    for t in threads:
        t.join()
    if state["error"] is not None:
        raise state["error"]
    return results


class MultiChainSliceSampler:
    """K independent coordinate-wise slice samplers (same algorithm and adaptation rules as
    :class:`gpyreg_b200.slice_sample.SliceSampler` / slice_sample.py:232-602) advanced in lock
    step.  ``log_f_batch(X (k, D)) -> (k,)`` evaluates the log density of k points at once.
    Every chain has its own ``numpy.random.Generator``; their seeds are drawn from the global
    NumPy RNG, so ``np.random.seed`` makes the whole run reproducible."""

    def __init__(self, log_f_batch, x0, widths, LB, UB, n_chains, adaptive=True):
        x0 = np.asarray(x0, dtype=float)
        self.K = int(n_chains)
        self.x = np.tile(x0, (self.K, 1)) if x0.ndim == 1 else x0.copy()
        if self.x.shape[0] != self.K:
            raise ValueError("x0 must be (D,) or (n_chains, D)")
        D = self.x.shape[1]
        self.LB = np.full((D,), -np.inf) if LB is None else np.asarray(LB, dtype=float).copy()
        self.UB = np.full((D,), np.inf) if UB is None else np.asarray(UB, dtype=float).copy()
        self.LB_out, self.UB_out = np.nextafter(self.LB, -np.inf), np.nextafter(self.UB, np.inf)
        w = (self.UB - self.LB) / 2 if widths is None else np.asarray(widths, dtype=float)
        w = np.tile(w, D) if w.size == 1 else w.copy()
        self.base_widths = None if widths is None else w.copy()
        w[np.isinf(w)] = 10
        w[self.LB == self.UB] = 1
        if np.any(w <= 0) or np.any(~np.isfinite(w)):
            raise ValueError("The widths vector needs to be all positive real numbers.")
        if np.any(self.x < self.LB) or np.any(self.x > self.UB):
            raise ValueError("The initial starting point X0 is outside the bounds.")
        self.widths = np.tile(w, (self.K, 1))
        self.log_f_batch = log_f_batch
        self.adaptive = adaptive
        self.func_count = 0
        self.rounds = 0
        seeds = np.random.randint(0, 2 ** 31 - 1, size=self.K)
        self.rng = [np.random.default_rng(int(s)) for s in seeds]

    def _evaluate(self, pts):
        """log density of every chain's pending point, -inf outside the bounds / for NaN."""
        inside = np.all((pts >= self.LB) & (pts <= self.UB), axis=1)
        out = np.full((pts.shape[0],), -np.inf)
        if np.any(inside):
            vals = np.asarray(self.log_f_batch(pts[inside]), dtype=float).reshape(-1)
            self.func_count += int(inside.sum())
            vals = np.where(np.isnan(vals), -np.inf, vals)
            out[inside] = vals
        self.rounds += 1
        return out

    def sample(self, N, thin=1, burn=0):
        """N recorded samples per chain -> dict with samples (K, N, D), f_vals (K, N)."""
        K, D = self.x.shape
        n_sweeps = N + (N - 1) * (thin - 1) + burn
        samples = np.zeros((K, N, D))
        f_vals = np.zeros((K, N))
        logp = self._evaluate(self.x.copy())
        if np.any(~np.isfinite(logp)):
            raise ValueError("The initial starting point X0 needs to evaluate to a "
                             "real number (not Inf or NaN).")
        free = np.nonzero(self.LB != self.UB)[0]
        sweep = np.zeros(K, dtype=int)                 # current sweep of each chain
        order = [None] * K                             # this sweep's coordinate order
        pos = np.zeros(K, dtype=int)                   # position in `order`
        lo, hi = self.x.copy(), self.x.copy()
        level = np.zeros(K)
        n_shrink = np.zeros(K, dtype=int)
        s1, s2 = np.zeros((K, D)), np.zeros((K, D))
        done = np.zeros(K, dtype=bool) | (n_sweeps == 0) | (free.size == 0)
        fresh = np.ones(K, dtype=bool)                 # needs a new coordinate / bracket

        def open_bracket(c):
            if order[c] is None:
                order[c] = self.rng[c].permutation(free)
                pos[c] = 0
            d = order[c][pos[c]]
            level[c] = logp[c] + np.log(self.rng[c].random())
            u = self.rng[c].random()
            lo[c, d] = max(self.x[c, d] - u * self.widths[c, d], self.LB_out[d])
            hi[c, d] = min(self.x[c, d] + (1 - u) * self.widths[c, d], self.UB_out[d])
            n_shrink[c] = 0
            fresh[c] = False

        def finish_coordinate(c, d):
            """Width adaptation, then move on; end-of-sweep bookkeeping."""
            it = sweep[c]
            if it < burn and self.adaptive:
                span = self.UB[d] - self.LB[d]
                if n_shrink[c] > 3:
                    floor = abs(np.spacing(span)) if np.isfinite(span) else np.spacing(1)
                    self.widths[c, d] = max(self.widths[c, d] / 1.1, floor)
                elif n_shrink[c] < 2:
                    self.widths[c, d] = min(self.widths[c, d] * 1.2, span)
            lo[c, d] = hi[c, d] = self.x[c, d]
            pos[c] += 1
            fresh[c] = True
            if pos[c] < len(order[c]):
                return
            order[c] = None                             # sweep complete
            if it >= burn and (it - burn) % thin == 0:
                k = (it - burn) // thin
                samples[c, k] = self.x[c]
                f_vals[c, k] = logp[c]
            if burn / 2 <= it < burn:
                s1[c] += self.x[c]
                s2[c] += self.x[c] ** 2
                if it == burn - 1 and self.adaptive:
                    m = np.floor(burn / 2)
                    new_w = np.fmin(5 * np.sqrt(np.maximum(s2[c] / m - (s1[c] / m) ** 2, 0)),
                                    self.UB_out - self.LB_out)
                    if self.base_widths is None:
                        self.widths[c] = new_w
                    else:
                        self.widths[c] = np.maximum(new_w, np.sqrt(new_w * self.base_widths))
            sweep[c] += 1
            if sweep[c] >= n_sweeps:
                done[c] = True

        while not np.all(done):
            act = np.nonzero(~done)[0]
            props = self.x[act].copy()
            dims = np.zeros(act.size, dtype=int)
            for r, c in enumerate(act):
                if fresh[c]:
                    open_bracket(c)
                d = order[c][pos[c]]
                dims[r] = d
                n_shrink[c] += 1
                props[r, d] = self.rng[c].random() * (hi[c, d] - lo[c, d]) + lo[c, d]
            vals = self._evaluate(props)
            for r, c in enumerate(act):
                d = dims[r]
                if vals[r] > level[c]:                 # accepted
                    self.x[c, d] = props[r, d]
                    logp[c] = vals[r]
                    finish_coordinate(c, d)
                elif props[r, d] > self.x[c, d]:
                    hi[c, d] = props[r, d]
                elif props[r, d] < self.x[c, d]:
                    lo[c, d] = props[r, d]
                else:                                  # shrunk to the current point
                    finish_coordinate(c, d)
        return {"samples": samples, "f_vals": f_vals, "widths": self.widths.copy(),
                "func_count": self.func_count, "rounds": self.rounds}
