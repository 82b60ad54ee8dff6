"""scrc_kmeanspp (csrc/kmeans.cu) against the restatement of kmeanspp() / stats::kmeans in oracle/kmeanspp_oracle.py:
same seeds, same Hartigan-Wong transfers, hence identical labels and objectives (distances are sequential sums in the
same order on both sides)."""
import numpy as np
import pytest

from oracle import kmeanspp_oracle as ko

pytestmark = pytest.mark.gpu


def _data(seed, n, p, k, sep):
    r = np.random.default_rng(seed)
    c = r.standard_normal((k, p)) * sep
    return c[r.integers(0, k, n)] + r.standard_normal((n, p)), r


@pytest.mark.parametrize("seed,n,p,k,sep,n_init,iters", [(0, 150, 12, 4, 2.0, 3, 10), (1, 97, 5, 2, 1.0, 2, 10),
                                                        (2, 200, 33, 7, 0.8, 4, 100), (3, 64, 3, 5, 0.5, 3, 2)])
def test_kmeanspp_matches_oracle(seed, n, p, k, sep, n_init, iters):
    from scregclust_b200 import api
    x, r = _data(seed, n, p, k, sep)
    firsts = r.integers(1, n + 1, n_init)
    us = r.random((n_init, k - 1))
    want_cl, want_tot, want_fault = ko.kmeanspp(x, k, firsts, us, n_max_iter=iters)
    cl, tot, fault = api.kmeanspp(x, k, firsts, us, n_max_iter=iters, return_details=True)
    assert np.array_equal(fault, want_fault)
    assert np.allclose(tot, want_tot, rtol=1e-13, atol=0)
    assert np.array_equal(cl, want_cl)


def test_session_kmeanspp_clusters_cross_corr1():
    """k_start_cl <- kmeanspp(cross_corr1, n_modules, n_initializations, 100L) (R/scregclust.R:1131-1136)."""
    from scregclust_b200.driver import Session
    from scregclust_b200.synthetic import make_workload
    from oracle import scregclust_oracle as orc
    w = make_workload("tiny", seed=5)
    r = np.random.default_rng(11)
    T = w.z1_target.shape[1]
    firsts = r.integers(1, T + 1, 3); us = r.random((3, w.n_modules - 1))
    with Session(w.z1_reg, w.z2_reg, w.z1_target, w.z2_target, w.z1_target_sds, w.n_modules) as s:
        cl, tot, fault = s.kmeanspp(firsts, us, n_max_iter=100, return_details=True)
        cc = s.cross_corr()
    # the oracle runs on the library's own cross_corr1 (fast_cor agrees with the oracle's to 1e-12, not bit for bit)
    want_cl, want_tot, want_fault = ko.kmeanspp(cc, w.n_modules, firsts, us, n_max_iter=100)
    assert np.allclose(cc, orc.fast_cor(w.z1_target, w.z1_reg), atol=1e-12)
    assert np.array_equal(fault, want_fault) and np.allclose(tot, want_tot, rtol=1e-13) and np.array_equal(cl, want_cl)
