"""Seeded synthetic workloads of the BASELINE.json shapes.

Planted-module data: regulators ``Zr ~ N(0,1)`` (optionally AR(1)-correlated
columns), ``K`` true modules with 8 regulators each, coefficients shared per
module plus a small per-target perturbation, ``Zt = Zr B + N(0,1)``, 5 % pure
noise targets.  The first half of the cells is split 1 (stands in for
``split_indices`` of R/scregclust.R:880-894) and the initial module labels are
uniform random (stands in for ``initial_target_modules``, R/scregclust.R:1145-1147),
so neither R's RNG nor k-means enters the comparison.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (n_cells, n_targets, n_regulators, n_modules, penalties)
    "tiny":    (400, 120, 24, 3, np.array([0.1, 0.2])),
    "small":   (2000, 400, 100, 5, np.array([0.1, 0.2, 0.3])),
    "config1": (2700, 5500, 500, 10, np.array([0.1])),
    "config2": (5000, 2000, 500, 10, np.geomspace(0.05, 0.5, 10)),
    "config3": (20000, 5000, 1600, 20, np.geomspace(0.05, 0.5, 20)),
    "config4": (100000, 10000, 1600, 40, np.geomspace(0.05, 0.5, 8)),
}


@dataclass
class Workload:
    z1_reg: np.ndarray       # n1 x P, scaled with split-1 statistics (R/scregclust.R:1000)
    z2_reg: np.ndarray       # n2 x P (R/scregclust.R:1007-1009)
    z1_target: np.ndarray    # n1 x T, centred (R/scregclust.R:998)
    z2_target: np.ndarray    # n2 x T, centred with split-1 means (R/scregclust.R:1004-1006)
    z1_target_sds: np.ndarray
    k_init: np.ndarray       # T, 1-based module labels
    k_true: np.ndarray       # T, planted labels (-1 noise)
    penalization: np.ndarray
    n_modules: int
    name: str = ""


def make_raw(n_cells, n_targets, n_reg, n_modules, seed=0, ar1=0.0, regs_per_module=8,
             noise_frac=0.05):
    rng = np.random.default_rng(seed)
    zr = rng.standard_normal((n_cells, n_reg))
    if ar1 > 0.0:
        c = np.sqrt(1.0 - ar1 * ar1)
        for j in range(1, n_reg):
            zr[:, j] = ar1 * zr[:, j - 1] + c * zr[:, j]
    k_true = rng.integers(1, n_modules + 1, size=n_targets)
    n_noise = int(round(noise_frac * n_targets))
    if n_noise > 0:
        k_true[rng.choice(n_targets, n_noise, replace=False)] = -1
    rpm = min(regs_per_module, n_reg)
    zt = rng.standard_normal((n_cells, n_targets))
    for j in range(1, n_modules + 1):
        regs = rng.choice(n_reg, rpm, replace=False)
        shared = 0.4 * rng.standard_normal(rpm)
        cols = np.flatnonzero(k_true == j)
        coef = shared[:, None] + 0.1 * rng.standard_normal((rpm, cols.size))
        zt[:, cols] += zr[:, regs] @ coef          # only the module's regulators are non-zero
    return zr, zt, k_true


def scale_split(zr, zt):
    """Split into halves and scale as R/scregclust.R:998-1009 does."""
    n = zr.shape[0]
    n1 = n // 2
    z1r, z2r, z1t, z2t = zr[:n1], zr[n1:], zt[:n1], zt[n1:]
    m1t = z1t.mean(axis=0)
    sds_t = z1t.std(axis=0, ddof=1)
    m1r = z1r.mean(axis=0)
    s1r = z1r.std(axis=0, ddof=1)
    f = np.asfortranarray
    return (f((z1r - m1r) / s1r), f((z2r - m1r) / s1r), f(z1t - m1t), f(z2t - m1t), sds_t)


def make_workload(name="tiny", seed=0, ar1=0.0, shape=None, penalization=None) -> Workload:
    if shape is None:
        n_cells, n_targets, n_reg, n_modules, pens = CONFIGS[name]
    else:
        n_cells, n_targets, n_reg, n_modules = shape
        pens = np.asarray(penalization, dtype=np.float64)
    if penalization is not None:
        pens = np.asarray(penalization, dtype=np.float64)
    zr, zt, k_true = make_raw(n_cells, n_targets, n_reg, n_modules, seed=seed, ar1=ar1)
    z1r, z2r, z1t, z2t, sds = scale_split(zr, zt)
    k_init = np.random.default_rng(seed + 100).integers(1, n_modules + 1, size=n_targets).astype(np.int32)
    return Workload(z1r, z2r, z1t, z2t, sds, k_init, k_true.astype(np.int32),
                    np.asarray(pens, dtype=np.float64), int(n_modules), name)
