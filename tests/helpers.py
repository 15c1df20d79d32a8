"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle.gp_oracle import ModelSpec

COV_TAGS = {
    "se_ard": (0, 0, 1), "mat1_ard": (1, 1, 1), "mat3_ard": (1, 3, 1),
    "mat5_ard": (1, 5, 1), "rq_ard": (2, 0, 1), "se_iso": (0, 0, 0),
    "mat1_iso": (1, 1, 0), "mat3_iso": (1, 3, 0), "mat5_iso": (1, 5, 0),
}

CORE_TAGS = ["cfg2_se_const", "cfg3_mat5_negquad", "cfg4_rq_const",
             "cfg5_mat3iso_const", "ex1_mat3_negquad_user", "se_zero_allnoise",
             "mat1_zero", "seiso_negquad", "mat5iso_zero_user", "multi_tile_se"]


def spec_from_array(a):
    D, ck, deg, ard, mk, p0, p1, p2 = (int(v) for v in a)
    return ModelSpec(D=D, cov_kind=ck, degree=deg, ard=bool(ard), mean_kind=mk,
                     noise_params=(p0, p1, p2))


def case(g, tag):
    """Pull one golden case out of the flat npz dict."""
    pre = tag + "."
    return {k[len(pre):]: v for k, v in g.items() if k.startswith(pre)}


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def grad_err(a, b):
    """max|da| / max|grad| per vector, NaN positions must match."""
    a = np.atleast_2d(np.asarray(a, dtype=float))
    b = np.atleast_2d(np.asarray(b, dtype=float))
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    worst = 0.0
    for ra, rb in zip(a, b):
        ok = ~np.isnan(rb)
        if ok.any():
            worst = max(worst, np.max(np.abs(ra[ok] - rb[ok])) / np.max(np.abs(rb[ok])))
    return worst
