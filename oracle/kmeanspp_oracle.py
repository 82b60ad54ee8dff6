"""CPU restatement of the initial clustering of scregclust(): ``kmeanspp()`` (R/utils.R:351-375) with
``kmeanspp_init()`` (R/utils.R:275-322) and ``stats::kmeans`` (Hartigan-Wong, the default algorithm).

TEST INFRASTRUCTURE ONLY: imported by tests/ as the checker of scrc_kmeanspp; the product path never does.

PARITY UNPINNED.  The reference delegates to base R, which is not in this image (no R, no R sources):
  * ``dist()``            -- R 4.x src/library/stats/src/distance.c, ``R_euclidean``: sequential sum of squared
                             differences over the columns, then ``sqrt``;
  * ``sample.int(n, 1, prob = ps)`` -- src/main/random.c: ``FixupProb`` (plain double sum, divide), then
                             ``ProbSampleNoReplace``: ``revsort`` (src/main/sort.c, a heap sort into DEcreasing order
                             that carries the index vector along), ``rT = unif_rand()``, walk the cumulative mass;
  * ``sum()``             -- long double accumulation (src/main/summary.c, ``rsum``);
  * ``stats::kmeans``     -- Hartigan & Wong (1979), Algorithm AS 136, Applied Statistics 28, 100-108
                             (src/library/stats/src/kmns.c is a translation of the published Fortran): initial
                             two-closest-centres assignment, then alternating optimal-transfer (OPTRA) and
                             quick-transfer (QTRAN) stages; ``iter.max`` bounds the OPTRA/QTRAN rounds, 50 m the
                             QTRAN steps.
All restated here from the published algorithm / from memory of those sources; nothing in the container can check
them against R.  What the tests can and do check independently: the seeding probabilities are the D^2 weights of
Arthur & Vassilvitskii; the result of the Hartigan-Wong restatement is a fixed point of single-point transfers (no
transfer lowers the within-cluster sum of squares), and its objective is never worse than Lloyd's from the same start.

R's RNG stays with the caller: ``first_centers[i]`` is the value of ``sample(n, size = 1L)`` and ``uniforms[i, :]`` the
K - 1 ``unif_rand()`` draws that the K - 1 ``sample.int(prob =)`` calls of initialisation i consume, in that order.
"""
from __future__ import annotations

import math

import numpy as np

BIG = 1.0e30


def revsort(a, ib):
    """R's ``revsort``: sorts ``a`` into descending order by heap sort, permuting ``ib`` alongside (in place)."""
    n = len(a)
    if n <= 1:
        return
    a1 = [0.0] + list(a)
    b1 = [0] + list(ib)
    l = (n >> 1) + 1
    ir = n
    while True:
        if l > 1:
            l -= 1
            ra, ii = a1[l], b1[l]
        else:
            ra, ii = a1[ir], b1[ir]
            a1[ir], b1[ir] = a1[1], b1[1]
            ir -= 1
            if ir == 1:
                a1[1], b1[1] = ra, ii
                break
        i = l
        j = l << 1
        while j <= ir:
            if j < ir and a1[j] > a1[j + 1]:
                j += 1
            if ra > a1[j]:
                a1[i], b1[i] = a1[j], b1[j]
                i = j
                j += i
            else:
                j = ir + 1
        a1[i], b1[i] = ra, ii
    a[:] = a1[1:]
    ib[:] = b1[1:]


def sample_int_prob_1(ps, u):
    """``sample.int(length(ps), size = 1, prob = ps)`` with ``unif_rand() == u``; returns a 1-based index."""
    p = [float(v) for v in ps]
    s = 0.0
    for v in p:                       # FixupProb
        if v > 0.0:
            s += v
    p = [v / s for v in p]
    perm = list(range(1, len(p) + 1))
    revsort(p, perm)                  # ProbSampleNoReplace
    rT = 1.0 * u
    mass = 0.0
    n1 = len(p) - 1
    j = 0
    while j < n1:
        mass += p[j]
        if rT <= mass:
            break
        j += 1
    return perm[j]


def _euclid(x, i, c):
    """R_euclidean between rows i and c (0-based) of x."""
    d = 0.0
    for j in range(x.shape[1]):
        dev = x[i, j] - x[c, j]
        d += dev * dev
    return math.sqrt(d)


def kmeanspp_init(x, n_cluster, first_center, uniforms):
    """R/utils.R:275-322.  ``first_center`` 1-based; returns the 1-based row indices of the initial centres."""
    n = x.shape[0]
    centers = [int(first_center)]
    mind = np.full(n, np.inf)          # min over the centres chosen so far of dist(i, centre)
    for r in range(1, n_cluster):
        c = centers[-1] - 1
        for i in range(n):
            d = _euclid(x, i, c)
            if d < mind[i]:
                mind[i] = d
        remaining = [i for i in range(1, n + 1) if i not in centers]           # setdiff(seq_len(n), centers)
        log_ws_sq = [math.log(mind[i - 1] * mind[i - 1]) if mind[i - 1] > 0.0 else -math.inf for i in remaining]
        mx = max(log_ws_sq)
        e = [math.exp(v - mx) for v in log_ws_sq]
        tot = np.longdouble(0.0)
        for v in e:                    # R's sum(): long double accumulation
            tot += np.longdouble(v)
        tot = float(tot)
        ps = [v / tot for v in e]
        centers.append(remaining[sample_int_prob_1(ps, float(uniforms[r - 1])) - 1])
    return centers


def _d2(a, c, i, l):
    s = 0.0
    for j in range(a.shape[1]):
        t = a[i, j] - c[l, j]
        s += t * t
    return s


def kmeans_hartigan_wong(a, centers, iter_max):
    """AS 136 as called by stats::kmeans(x, centers, iter.max).  ``a``: m x n data, ``centers``: k x n initial centres.
    Returns (cluster 1-based, centres, wss, ifault)  -- ifault 0 ok, 1 empty cluster, 2 not converged, 4 QTRAN cap."""
    a = np.asarray(a, dtype=np.float64)
    c = np.array(centers, dtype=np.float64, copy=True)
    m, n = a.shape
    k = c.shape[0]
    ic1 = np.zeros(m, dtype=np.int64); ic2 = np.zeros(m, dtype=np.int64)
    nc = np.zeros(k, dtype=np.int64)
    an1 = np.zeros(k); an2 = np.zeros(k); ncp = np.zeros(k, dtype=np.int64); d = np.zeros(m)
    itran = np.zeros(k, dtype=np.int64); live = np.zeros(k, dtype=np.int64)
    imaxqtr = min(2 ** 31 - 1, 50 * m)
    # ---- two closest centres of every point
    for i in range(m):
        i1, i2 = 0, 1
        dt = [_d2(a, c, i, 0), _d2(a, c, i, 1)]
        if dt[0] > dt[1]:
            i1, i2 = 1, 0
            dt = [dt[1], dt[0]]
        for l in range(2, k):
            db = _d2(a, c, i, l)
            if db >= dt[1]:
                continue
            if db >= dt[0]:
                dt[1] = db; i2 = l
            else:
                dt[1] = dt[0]; i2 = i1; dt[0] = db; i1 = l
        ic1[i], ic2[i] = i1, i2
    # ---- centres = means of their points
    c[:] = 0.0
    for i in range(m):
        l = ic1[i]
        nc[l] += 1
        for j in range(n):
            c[l, j] += a[i, j]
    if np.any(nc == 0):
        return ic1 + 1, c, np.zeros(k), 1
    for l in range(k):
        aa = float(nc[l])
        for j in range(n):
            c[l, j] /= aa
        an2[l] = aa / (aa + 1.0)
        an1[l] = aa / (aa - 1.0) if aa > 1.0 else BIG
        itran[l] = 1
        ncp[l] = -1
    indx = 0
    ifault = 2
    for _ in range(iter_max):
        # ================= OPTRA
        for l in range(k):
            if itran[l] == 1:
                live[l] = m
        stop = False
        for i in range(m):
            indx += 1
            l1 = ic1[i]
            if nc[l1] != 1:
                if ncp[l1] != 0:
                    d[i] = _d2(a, c, i, l1) * an1[l1]
                l2 = ic2[i]; ll = l2
                r2 = _d2(a, c, i, l2) * an2[l2]
                for l in range(k):
                    if ((i + 1 >= live[l1] and i + 1 >= live[l]) or l == l1 or l == ll):
                        continue
                    rr = r2 / an2[l]
                    dc = _d2(a, c, i, l)
                    if dc >= rr:
                        continue
                    r2 = dc * an2[l]; l2 = l
                if r2 >= d[i]:
                    ic2[i] = l2
                else:
                    indx = 0
                    live[l1] = m + i + 1; live[l2] = m + i + 1
                    ncp[l1] = i + 1; ncp[l2] = i + 1
                    al1 = float(nc[l1]); alw = al1 - 1.0; al2 = float(nc[l2]); alt = al2 + 1.0
                    for j in range(n):
                        c[l1, j] = (c[l1, j] * al1 - a[i, j]) / alw
                        c[l2, j] = (c[l2, j] * al2 + a[i, j]) / alt
                    nc[l1] -= 1; nc[l2] += 1
                    an2[l1] = alw / al1
                    an1[l1] = alw / (alw - 1.0) if alw > 1.0 else BIG
                    an1[l2] = alt / al2
                    an2[l2] = alt / (alt + 1.0)
                    ic1[i] = l2; ic2[i] = l1
            if indx == m:
                stop = True
                break
        if not stop:
            for l in range(k):
                itran[l] = 0
                live[l] -= m
        if indx == m:
            ifault = 0
            break
        # ================= QTRAN
        icoun = 0; istep = 0
        capped = False
        done = False
        while not done:
            for i in range(m):
                icoun += 1; istep += 1
                if istep >= imaxqtr:
                    capped = True; done = True
                    break
                l1 = ic1[i]; l2 = ic2[i]
                if nc[l1] != 1:
                    if istep <= ncp[l1]:
                        d[i] = _d2(a, c, i, l1) * an1[l1]
                    if istep < ncp[l1] or istep < ncp[l2]:
                        r2 = d[i] / an2[l2]
                        dd = _d2(a, c, i, l2)
                        if dd < r2:
                            icoun = 0; indx = 0
                            itran[l1] = 1; itran[l2] = 1
                            ncp[l1] = istep + m; ncp[l2] = istep + m
                            al1 = float(nc[l1]); alw = al1 - 1.0; al2 = float(nc[l2]); alt = al2 + 1.0
                            for j in range(n):
                                c[l1, j] = (c[l1, j] * al1 - a[i, j]) / alw
                                c[l2, j] = (c[l2, j] * al2 + a[i, j]) / alt
                            nc[l1] -= 1; nc[l2] += 1
                            an2[l1] = alw / al1
                            an1[l1] = alw / (alw - 1.0) if alw > 1.0 else BIG
                            an1[l2] = alt / al2
                            an2[l2] = alt / (alt + 1.0)
                            ic1[i] = l2; ic2[i] = l1
                if icoun == m:
                    done = True
                    break
        if capped:
            ifault = 4
            break
        if k == 2:
            ifault = 0
            break
        ncp[:] = 0
    # ---- within-cluster sums of squares (centres recomputed as means)
    wss = np.zeros(k)
    c[:] = 0.0
    for i in range(m):
        for j in range(n):
            c[ic1[i], j] += a[i, j]
    for j in range(n):
        for l in range(k):
            c[l, j] /= float(nc[l])
        for i in range(m):
            da = a[i, j] - c[ic1[i], j]
            wss[ic1[i]] += da * da
    return ic1 + 1, c, wss, ifault


def kmeanspp(x, n_cluster, first_centers, uniforms, n_max_iter=10):
    """R/utils.R:351-375: one Hartigan-Wong run per initialisation, the clustering with the smallest tot.withinss
    (first minimum, ``which.min``).  Returns (cluster 1-based, tot_withinss per initialisation, ifault per initialisation)."""
    x = np.asarray(x, dtype=np.float64)
    n_init = len(first_centers)
    runs, tot, faults = [], [], []
    for r in range(n_init):
        idx = kmeanspp_init(x, n_cluster, first_centers[r], np.asarray(uniforms)[r])
        cl, _, wss, ifault = kmeans_hartigan_wong(x, x[np.array(idx) - 1], n_max_iter)
        t = np.longdouble(0.0)
        for v in wss:
            t += np.longdouble(v)
        runs.append(cl); tot.append(float(t)); faults.append(ifault)
    best = int(np.argmin(tot))
    return runs[best], np.array(tot), np.array(faults)
