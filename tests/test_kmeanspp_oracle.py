"""The restatement of kmeanspp() (R/utils.R:351-375: k-means++ seeding + stats::kmeans Hartigan-Wong) that checks
scrc_kmeanspp.  R is not available, so the oracle itself is anchored on what can be verified independently."""
import numpy as np
import pytest

from oracle import kmeanspp_oracle as ko


def _blobs(rng, n=90, p=5, k=3, sep=4.0):
    c = rng.standard_normal((k, p)) * sep
    lab = rng.integers(0, k, n)
    return c[lab] + rng.standard_normal((n, p)), lab


def test_revsort_sorts_descending_and_carries_the_index(rng):
    a = list(rng.random(37)); a[5] = a[9]                      # a tie
    ib = list(range(1, 38))
    ref = sorted(a, reverse=True)
    orig = list(a)
    ko.revsort(a, ib)
    assert a == ref and all(orig[i - 1] == v for i, v in zip(ib, a))


def test_sampling_walks_the_decreasing_cumulative_mass():
    ps = [0.1, 0.6, 0.3]
    assert ko.sample_int_prob_1(ps, 0.05) == 2                  # largest probability first
    assert ko.sample_int_prob_1(ps, 0.6) == 2                   # rT <= mass is inclusive
    assert ko.sample_int_prob_1(ps, 0.61) == 3
    assert ko.sample_int_prob_1(ps, 0.95) == 1
    draws = [ko.sample_int_prob_1(ps, u) for u in (np.arange(2000) + 0.5) / 2000]
    assert np.allclose(np.bincount(draws, minlength=4)[1:] / 2000, ps, atol=2e-3)


def test_seeding_uses_d2_weights(rng):
    """Arthur & Vassilvitskii: the next centre is drawn with probability proportional to the squared distance to the
    nearest centre chosen so far -- checked by sweeping the uniform over (0, 1)."""
    x, _ = _blobs(rng, n=40, p=3)
    first = 7
    d2 = ((x - x[first - 1]) ** 2).sum(axis=1)
    us = (np.arange(4000) + 0.5) / 4000
    picks = np.array([ko.kmeanspp_init(x, 2, first, [u])[1] for u in us])
    freq = np.bincount(picks, minlength=41)[1:] / us.size
    assert freq[first - 1] == 0.0
    assert np.allclose(freq, d2 / d2.sum(), atol=1.5e-3)


def test_hartigan_wong_ends_in_a_point_transfer_optimum(rng):
    """AS 136 stops when no single point can move to another cluster and lower the total within-cluster sum of squares
    (R2 = n_l/(n_l+1) d(i,l)^2 against R1 = n_1/(n_1-1) d(i,1)^2); and it is never worse than Lloyd from the same start."""
    from sklearn.cluster import KMeans
    for seed in range(4):
        r = np.random.default_rng(seed)
        x, _ = _blobs(r, n=120, p=6, k=4, sep=1.5)
        init = r.choice(120, 4, replace=False)
        cl, c, wss, ifault = ko.kmeans_hartigan_wong(x, x[init], 50)
        assert ifault == 0
        cl0 = cl - 1
        nc = np.bincount(cl0, minlength=4)
        cen = np.stack([x[cl0 == l].mean(axis=0) for l in range(4)])
        assert np.allclose(cen, c, atol=1e-12) and np.isclose(wss.sum(), ((x - cen[cl0]) ** 2).sum())
        d2 = ((x[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        for i in range(120):
            l1 = cl0[i]
            if nc[l1] == 1:
                continue
            r1 = nc[l1] / (nc[l1] - 1.0) * d2[i, l1]
            for l in range(4):
                if l != l1:
                    assert nc[l] / (nc[l] + 1.0) * d2[i, l] >= r1 - 1e-9
        lloyd = KMeans(n_clusters=4, init=x[init], n_init=1, algorithm="lloyd", max_iter=300, tol=0).fit(x)
        assert wss.sum() <= lloyd.inertia_ + 1e-9


def test_kmeanspp_picks_the_first_best_initialisation(rng):
    x, lab = _blobs(rng, n=80, p=4, k=3, sep=6.0)
    firsts = [3, 17, 40]
    us = rng.random((3, 2))
    cl, tot, faults = ko.kmeanspp(x, 3, firsts, us, n_max_iter=10)
    assert tot.shape == (3,) and np.all(faults == 0)
    best = int(np.argmin(tot))
    again, _, _ = ko.kmeanspp(x, 3, firsts[best:best + 1], us[best:best + 1], n_max_iter=10)
    assert np.array_equal(cl, again)
    # well separated blobs are recovered up to a relabelling
    from oracle.scregclust_oracle import adjusted_rand_index
    assert adjusted_rand_index(cl, lab + 1) == 1.0
