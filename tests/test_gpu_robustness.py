"""State-handling regressions of the device layer (factor cache, data upload, context and
posterior lifetimes, several GP objects on one GPU), all through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import gp_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(N=300, D=3, B=12, mean_kind=2):
    from bench import benign_hyp, synth_data
    spec = orc.ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=mean_kind)
    X, y = synth_data(N, D, seed=0)
    return spec, X, y, benign_hyp(spec, B, y, seed=1)


def _engine(spec, X, y):
    from gpyreg_b200 import Engine
    e = Engine(0)
    e.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    e.set_data(X, y, None)
    return e


_NOCACHE = r"""
import sys, json
import numpy as np
sys.path.insert(0, ".")
from gpyreg_b200 import Engine
d = np.load(sys.argv[1])
e = Engine(0)
e.set_model(*[int(v) for v in d["model"][:4]], tuple(int(v) for v in d["model"][4:]))
e.set_data(d["X"], d["y"], None)
print(json.dumps([v.hex() for v in e.nlz_batch(d["hyp"])[0]]))
"""


def test_factor_cache_survives_invalidation(tmp_path):
    """after the cache is invalidated (gradient call = workspace re-allocation
    with W, new data, new model) a SMALLER cacheable call must not revive the entries of the
    earlier, larger batch.  Reference values: a process with GPB_NLZ_CACHE=0."""
    spec, X, y, hyp = _problem()
    path = str(tmp_path / "p.npz")
    np.savez(path, X=X, y=y, hyp=hyp,
             model=np.array([spec.cov_kind, spec.degree, int(spec.ard), spec.mean_kind, *spec.noise_params]))
    env = dict(os.environ, GPB_NLZ_CACHE="0")
    out = subprocess.run([sys.executable, "-c", _NOCACHE, path], cwd=ROOT, env=env, check=True,
                         capture_output=True, text=True, timeout=600).stdout.strip().splitlines()[-1]
    import json
    ref = np.array([float.fromhex(v) for v in json.loads(out)])

    e = _engine(spec, X, y)
    np.testing.assert_array_equal(e.nlz_batch(hyp)[0], ref)              # 12 rows cached
    e.nlz_batch(hyp[5:6], want_grad=True)                                # re-allocates (W), overwrites slot 0
    np.testing.assert_array_equal(e.nlz_batch(hyp[5:6])[0], ref[5:6])    # miss, re-keys slot 0 only
    for r in (7, 0, 11, 5):                                              # old rows: must NOT hit dead slots
        np.testing.assert_array_equal(e.nlz_batch(hyp[r:r + 1])[0], ref[r:r + 1])
    # same N, different data: nothing of the old batch may survive
    np.testing.assert_array_equal(e.nlz_batch(hyp)[0], ref)
    e.set_data(X[::-1].copy(), y[::-1].copy(), None)
    np.testing.assert_array_equal(e.nlz_batch(hyp[:1])[0], e.nlz_batch(hyp[:1], want_grad=True)[0])
    got = e.nlz_batch(hyp[3:4])[0]
    assert rel_err(got, ref[3:4]) <= 1e-11                               # permutation invariance, fresh factor
    # mean-only design (all covariance/noise hyperparameters equal): every row has the same key
    same = np.repeat(hyp[:1], 6, axis=0)
    same[:, spec.cov_n + spec.noise_n] += np.arange(6) * 0.01
    e.set_data(X, y, None)
    a = e.nlz_batch(same)[0]
    e.nlz_batch(same[:2], want_grad=True)
    # slots 0 and 1 now hold inverses, not factors; a one-row call re-validates the cache, and the
    # batched calls after it must find the factor in slot 0, not in their own (dead) slots
    b = np.concatenate([e.nlz_batch(same[:1])[0], e.nlz_batch(same[1:4])[0], e.nlz_batch(same[4:6])[0]])
    np.testing.assert_array_equal(a, b)
    e.close()


def test_inplace_data_edit_is_seen():
    """the reference reads gp.X / gp.y afresh on every evaluation, so an
    in-place edit of a single element must reach the GPU."""
    import gpyreg_b200 as g
    from gpyreg_b200.covariance_functions import Matern
    from gpyreg_b200.mean_functions import ConstantMean
    from gpyreg_b200.noise_functions import GaussianNoise
    spec, X, y, hyp = _problem(N=200, B=2, mean_kind=1)
    gp = g.GP(3, Matern(5), ConstantMean(), GaussianNoise(constant_add=True))
    gp.X, gp.y = X.copy(), y.copy()
    before = gp._GP__compute_nlZ(hyp[0], False, False)
    assert rel_err(before, orc.nlz_batch(spec, hyp[:1], X, y, None, False)) <= 1e-9
    gp.y[137, 0] += 0.5                         # one element, not at any strided probe position
    after = gp._GP__compute_nlZ(hyp[0], False, False)
    assert rel_err(after, orc.nlz_batch(spec, hyp[:1], gp.X, gp.y, None, False)) <= 1e-9
    assert after != before
    gp.X[3, 1] = np.nextafter(gp.X[3, 1], 10.0)  # a one-ulp edit
    key = gp._data_key
    gp._GP__compute_nlZ(hyp[0], False, False)
    assert gp._data_key != key


def test_several_gps_share_one_context():
    """GP objects share the device's engine (no per-GP workspace hoarding) and
    switching between them re-uploads the right data."""
    import gpyreg_b200 as g
    from gpyreg_b200.covariance_functions import Matern
    from gpyreg_b200.mean_functions import ConstantMean
    from gpyreg_b200.noise_functions import GaussianNoise
    spec, X, y, hyp = _problem(N=260, B=2, mean_kind=1)
    gps = []
    for k in range(3):
        gp = g.GP(3, Matern(5), ConstantMean(), GaussianNoise(constant_add=True))
        gp.update(X_new=X[k * 40:k * 40 + 140], y_new=y[k * 40:k * 40 + 140], hyp=hyp)
        gps.append(gp)
    assert gps[0].engine is gps[1].engine is gps[2].engine
    Xs = np.random.default_rng(0).uniform(-3, 3, (20, 3))
    for k in (2, 0, 1, 0):
        gp = gps[k]
        Xk, yk = X[k * 40:k * 40 + 140], y[k * 40:k * 40 + 140]
        nlz = gp._GP__compute_nlZ(hyp[1], False, False)
        assert rel_err(nlz, orc.nlz_batch(spec, hyp[1:2], Xk, yk, None, False)) <= 1e-9
        mu, s2 = gp.predict(Xs)
        rmu, rs2 = orc.predict(spec, orc.posterior_batch(spec, hyp, Xk, yk, None), Xk, yk, Xs)
        assert np.max(np.abs(mu - rmu)) <= 1e-8 * (1 + np.max(np.abs(rmu)))
        assert np.max(np.abs(s2 - rs2)) <= 1e-8 * np.max(np.abs(rs2))


def test_posterior_handles_die_with_their_context():
    """closing an engine releases its posterior batches; stale Python handles
    neither crash nor double-free."""
    spec, X, y, hyp = _problem(N=150, B=2)
    e = _engine(spec, X, y)
    p1, p2 = e.posterior_batch(hyp), e.posterior_batch(hyp[:1])
    p1.free()
    e.close()
    assert p2._h is None
    p2.free()
    p1.free()
    del p1, p2, e


def test_posterior_rebuild_of_selected_samples():
    """gaussian_process.py:864-868: after a rank-one update only the samples whose update was unstable are
    recomputed (on the extended data); the others keep their rank-one factors.  Natural instability is rare,
    so the recompute is exercised on a chosen sample: it must equal a fresh posterior of that row on N+1
    points bit for bit, the untouched sample must equal the rank-one result bit for bit, and predictions
    through the mixed batch must match the oracle."""
    from gpyreg_b200 import Engine
    spec, X, y, hyp = _problem(N=201, B=3, mean_kind=1)
    Xn, yn, X0, y0 = X[-1], y[-1], X[:-1], y[:-1]
    e = _engine(spec, X0, y0)
    post = e.posterior_batch(hyp)
    e.predict(post, X0[:4])                               # W = L^-1 exists before the append
    st = e.posterior_append(post, Xn, float(yn[0]))
    assert st is not None and not st.any() and post.N == 201
    kept = [post.fetch(b, "alpha").copy() for b in range(3)]
    e.set_data(X, y, None)                                # the context now holds the N+1 points
    e.posterior_rebuild(post, [1])
    fresh = e.posterior_batch(hyp)
    np.testing.assert_array_equal(post.fetch(1, "alpha"), fresh.fetch(1, "alpha"))
    np.testing.assert_array_equal(post.fetch(1, "L"), fresh.fetch(1, "L"))
    for b in (0, 2):
        np.testing.assert_array_equal(post.fetch(b, "alpha"), kept[b])
    Xs = np.random.default_rng(1).uniform(-3, 3, (40, 3))
    mu, s2 = e.predict(post, Xs, add_noise=True, separate=True)
    rmu, rs2 = orc.predict(spec, orc.posterior_batch(spec, hyp, X, y, None), X, y, Xs, add_noise=True,
                           separate_samples=True)
    assert np.max(np.abs(mu - rmu)) <= 1e-8 * (1 + np.max(np.abs(rmu)))
    assert np.max(np.abs(s2 - rs2)) <= 1e-8 * np.max(np.abs(rs2))
    with pytest.raises(Exception, match="data the posteriors cover"):
        e.set_data(X0, y0, None)
        e.posterior_rebuild(post, [0])
    post.free()
    fresh.free()
    e.close()


def test_gp_deepcopy_and_pickle_rebuild_their_factors():
    """PyVBMC deep-copies and pickles its GPs.  A copy carries host records (alpha, sW, flags; the (N, N)
    factor stays behind), shares nothing with the original, predicts identically after re-creating its
    factors on the device, and still serves ``posteriors[i].L``."""
    import copy
    import pickle
    import gpyreg_b200 as g
    from gpyreg_b200.covariance_functions import Matern
    from gpyreg_b200.mean_functions import ConstantMean
    from gpyreg_b200.noise_functions import GaussianNoise
    spec, X, y, hyp = _problem(N=180, B=3, mean_kind=1)
    gp = g.GP(3, Matern(5), ConstantMean(), GaussianNoise(constant_add=True))
    gp.update(X_new=X, y_new=y, hyp=hyp)
    Xs = np.random.default_rng(0).uniform(-3, 3, (25, 3))
    ref = gp.predict(Xs, add_noise=True)
    L0 = gp.posteriors[1].L.copy()
    for other in (copy.deepcopy(gp), pickle.loads(pickle.dumps(gp))):
        assert other._post_batch is None and other.posteriors[0]._batch is None
        assert not other.posteriors[0]._have["L"] and other.posteriors[0]._have["alpha"]
        np.testing.assert_array_equal(other.posteriors[2].alpha, gp.posteriors[2].alpha)
        got = other.predict(Xs, add_noise=True)
        np.testing.assert_array_equal(got[0], ref[0])
        np.testing.assert_array_equal(got[1], ref[1])
        np.testing.assert_array_equal(other.posteriors[1].L, L0)
        other.update(X_new=Xs[:1], y_new=np.zeros((1, 1)))          # rank-one update on the copy only
        assert other.X.shape[0] == 181 and gp.X.shape[0] == 180
    np.testing.assert_array_equal(gp.predict(Xs, add_noise=True)[0], ref[0])
    # a bare copy of the records is complete (what the reference's test_cleaning does)
    posts = copy.deepcopy(gp.posteriors)
    np.testing.assert_array_equal(posts[1].L, L0)
    assert posts[1]._batch is None


def test_gradient_after_large_nlz_batch_adds_w_beside_the_arena():
    """W = L^-1 has its own allocation and capacity: a gradient call after a large nlZ-only batch allocates a small W
    next to the slots that are there (it used to release and re-partition the whole workspace: ~1 s of a cfg3 fit), a
    larger gradient call grows W, a workspace limit that leaves no room beside the arena re-partitions it.  Every
    result is bit-identical to a fresh context's."""
    spec, X, y, hyp = _problem(N=300, D=3, B=40)
    fresh = _engine(spec, X, y)
    ref_nlz = fresh.nlz_batch(hyp)[0]
    ref_g = fresh.nlz_batch(hyp, want_grad=True)
    fresh.close()

    e = _engine(spec, X, y)
    np.testing.assert_array_equal(e.nlz_batch(hyp)[0], ref_nlz)                 # 40 slots, no W
    for rows in (slice(0, 3), slice(5, 7), slice(0, 17), slice(0, 40), slice(2, 3)):   # W: 3 slots -> 17 -> 40
        got = e.nlz_batch(hyp[rows], want_grad=True)
        np.testing.assert_array_equal(got[0], ref_g[0][rows])
        np.testing.assert_array_equal(got[1], ref_g[1][rows])
        np.testing.assert_array_equal(e.nlz_batch(hyp[:9])[0], ref_nlz[:9])     # nlZ-only calls in between
    e.close()

    # a limit with room for ~6 slots without W: the gradient call must re-partition (3 slots with W) and chunk
    e = _engine(spec, X, y)
    per_noW = (384 * 384 + 2 * 3 * 128 * 128) * 8 + 64 * 1024
    e.set_workspace_limit(6 * per_noW)
    np.testing.assert_array_equal(e.nlz_batch(hyp)[0], ref_nlz)
    got = e.nlz_batch(hyp, want_grad=True)
    np.testing.assert_array_equal(got[0], ref_g[0])
    np.testing.assert_array_equal(got[1], ref_g[1])
    np.testing.assert_array_equal(e.nlz_batch(hyp)[0], ref_nlz)
    e.close()
