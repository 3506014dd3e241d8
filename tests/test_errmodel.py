"""Error-model refit: dada2_b200/errmodel.py against what can be pinned without R -- see the module
header: `loess` itself is PARITY UNPINNED (R's stats::loess is neither in /root/reference nor runnable here)."""
import numpy as np
import pytest

from dada2_b200 import errmodel
from tests import cases


def test_accumulate_trans_pads_to_the_widest_matrix():
    a, b = np.ones((16, 3)), 2 * np.ones((16, 5))
    r = errmodel.accumulateTrans([a, b])
    assert r.shape == (16, 5) and np.all(r[:, :3] == 3) and np.all(r[:, 3:] == 2)


def test_loess_reproduces_quadratics_and_ignores_weight_scale():
    x = np.arange(41.0)
    y = 0.5 - 0.07 * x + 0.002 * x * x
    w = np.linspace(1, 5, 41)
    f1 = errmodel.loess_direct(x, y, w, x)
    f2 = errmodel.loess_direct(x, y, 1000.0 * w, x)
    assert np.allclose(f1, y, atol=1e-9) and np.allclose(f1, f2, atol=1e-9)
    assert np.isnan(errmodel.loess_direct(x[5:30], y[5:30], w[5:30], x)[[0, 4, 30, 40]]).all()      # no extrapolation


def _counts(rng, err, depth):
    trans = np.zeros((16, err.shape[1]))
    for i in range(4):
        for q in range(err.shape[1]):
            trans[4 * i:4 * i + 4, q] = rng.multinomial(int(depth[q]), err[4 * i:4 * i + 4, q] / err[4 * i:4 * i + 4, q].sum())
    return trans


def test_loessErrfun_structure_and_recovery():
    rng = np.random.default_rng(3)
    truth = cases.tperr1()
    depth = np.full(41, 4e6)
    depth[:2] = 0                                        # no data at q = 0, 1: NA rows, filled from the first fitted value
    err = errmodel.loessErrfun(_counts(rng, truth, depth))
    assert err.shape == (16, 41)
    for i in range(4):
        assert np.allclose(err[4 * i:4 * i + 4].sum(axis=0), 1.0, atol=1e-12)
    off = np.array([k for k in range(16) if k % 5])
    assert err[off].max() <= errmodel.MAX_ERROR_RATE and err[off].min() >= errmodel.MIN_ERROR_RATE
    assert np.array_equal(err[:, 0], err[:, 2]) and np.array_equal(err[:, 1], err[:, 2])
    # a smooth truth is recovered to a few per cent in the well-covered quality range
    smooth = np.array(truth)
    for r in off:
        smooth[r] = 10 ** np.polyval(np.polyfit(np.arange(41), np.log10(truth[r]), 2), np.arange(41))
    for i in range(4):
        smooth[5 * i] = 1 - (smooth[4 * i:4 * i + 4].sum(axis=0) - smooth[5 * i])
    fit = errmodel.loessErrfun(_counts(rng, smooth, np.full(41, 3e7)))
    assert np.max(np.abs(np.log10(fit[off, 5:36] / smooth[off, 5:36]))) < 0.05


def test_loessErrfun_rejects_starved_input():
    t = np.zeros((16, 41)); t[:, 30] = 5
    with pytest.raises(ValueError):
        errmodel.loessErrfun(t)


def test_selfconsist_loop_on_the_cpu_oracle():
    """The loop of R/dada.R:256-391 around the CPU oracle: pass 0 with all-ones err and MAX_CLUST = 1, refits until the matrix repeats."""
    from oracle import port
    seqs, ab, pri, err0, q, opts = cases.build_case("syn2000_default")
    log = []

    def runner(e, max_clust):
        log.append((np.array(e), max_clust))
        return port.dada_uniques(seqs, ab, None, e, q, max_clust=max_clust)
    out = errmodel.learnErrors(runner)
    assert log[0][1] == 1 and np.all(log[0][0] == 1.0) and all(mc == 0 for _, mc in log[1:])
    assert np.all(log[1][0][[0, 5, 10, 15]] == 1.0)        # self-transitions of the initial estimate forced to 1
    assert 2 <= out["passes"] <= 11
    e = out["err_out"]
    assert e.shape == (16, 41) and np.all((e > 0) & (e <= 1))
    if out["passes"] < 11:                                 # converged: the final matrix was seen before
        assert any(np.array_equal(e, x) for x in out["err_in"])
