"""CPU-only checks of the drop-in boundary: the library builds for sm_100a, loads without a GPU,
exports every symbol include/scregclust_b200.h declares, and the argument checks that run before
any device work return the reference's error texts (src/optim.cpp:76-122,408-410; utils.cpp:20,52)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from scregclust_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _capi.load()


def test_header_symbols_exported(lib):
    from scregclust_b200 import _capi
    header = open(os.path.join(ROOT, "include", "scregclust_b200.h")).read()
    declared = set(re.findall(r"\b(scrc_[a-z_0-9]+)\s*\(", header))
    declared -= {"scrc_ctx", "scrc_options", "scrc_cycle_out"}
    assert len(declared) >= 25
    raw = C.CDLL(_capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(_capi.SIGNATURES), (declared ^ set(_capi.SIGNATURES))


def test_version_and_defaults(lib):
    from scregclust_b200 import _capi
    assert lib.scrc_version() == 201
    o = _capi.Options()
    lib.scrc_default_options(C.byref(o))
    # R/scregclust.R:193-196 and src/optim.cpp:67-69
    assert (o.max_optim_iter, o.tol_coop_rel, o.tol_coop_abs, o.tol_nnls) == (10000, 1e-8, 1e-12, 1e-4)
    assert (o.rho0, o.alpha0, o.n_update, o.eps_corr) == (0.2, 1.5, 2, 0.2)


def test_argument_errors_match_reference_texts(lib):
    from scregclust_b200 import api
    from scregclust_b200._capi import ScrcError
    y = np.zeros((4, 2)); w = np.ones(3)
    with pytest.raises(ScrcError, match="same number of rows"):
        api.coop_lasso(y, np.zeros((5, 3)), 0.1, w)
    with pytest.raises(ScrcError, match="COOP LASSO: Matrix dimensions of y and x need to be positive"):
        api.coop_lasso(np.zeros((4, 0)), np.zeros((4, 3)), 0.1, w)
    x = np.ones((4, 3))
    with pytest.raises(ScrcError, match="beta_0 needs to be of size p x m"):
        api.coop_lasso(y, x, 0.1, w, beta_0=np.zeros((2, 2)))
    with pytest.raises(ScrcError, match="rho_0 > 0 needs to hold"):
        api.coop_lasso(y, x, 0.1, w, rho_0=-1.0)
    with pytest.raises(ScrcError, match="0 < alpha_0 < 2 needs to hold"):
        api.coop_lasso(y, x, 0.1, w, alpha_0=0.0)
    with pytest.raises(ScrcError, match="NNLS: Matrix dimensions of y and x need to be positive"):
        api.coef_nnls(np.zeros((5, 0)), np.zeros((5, 2)))


def test_host_side_array_helpers(lib):
    # alloc_array / reset_array are host-side in the library (no device needed)
    from scregclust_b200 import api
    from scregclust_b200._capi import ScrcError
    z = np.arange(12.0).reshape(4, 3)
    arr = api.alloc_array(z, 5)
    assert arr.shape == (5, 4, 3) and all(np.array_equal(arr[q], z) for q in range(5))
    arr[2] = -1.0
    api.reset_array(arr, z)
    assert np.array_equal(arr[2], z)
    with pytest.raises(ScrcError, match="reset_array: input has wrong dimensions"):
        api.reset_array(arr, z[:, :2])


def test_no_cpu_fallback_without_device(lib):
    """Without a usable CUDA device the compute entry points fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from scregclust_b200 import api
    from scregclust_b200._capi import ScrcError
    with pytest.raises(ScrcError) as ei:
        api.fast_cor(np.random.rand(10, 3), np.random.rand(10, 2))
    assert ei.value.code <= -100


def build_c_client(tmp_path):
    """Compiles tests/c/abi_smoke.c as strict C99 against the public header and the shared library."""
    import subprocess
    from scregclust_b200 import _capi
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lscregclust_b200", "-lm",
           f"-Wl,-rpath,{libdir}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_plain_c_client_links_and_reports_errors(lib, tmp_path):
    """The boundary is usable from plain C (what cgo / JNI / .Call glue binds): header is valid C99,
    argument errors carry the reference's texts, and without a device nothing computes."""
    import subprocess
    exe = build_c_client(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "C ABI smoke: ok" in out.stdout, out.stdout + out.stderr


def test_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """tests/c/abi_layout.c prints sizeof / offsetof(last member) of every struct of the public header as a C compiler sees
    them; the ctypes mirrors of scregclust_b200/_capi.py must agree (a member added on one side only shifts the rest)."""
    import ctypes as C
    import shutil
    import subprocess
    from scregclust_b200 import _capi
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_layout")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_layout.c"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    mirrors = {"scrc_options": (_capi.Options, "eps_corr"), "scrc_cycle_out": (_capi.CycleOut, "admm_sweeps"),
               "scrc_prior": (_capi.Prior, "aggregate"), "scrc_quick": (_capi.Quick, "percent"),
               "scrc_grid_params": (_capi.GridParams, "cancel"), "scrc_grid_info": (_capi.GridInfo, "device_calls"),
               "scrc_grid_penalty_info": (_capi.GridPenaltyInfo, "n_records"),
               "scrc_grid_final": (_capi.GridFinal, "importance"), "scrc_grid_backend": (_capi.GridBackend, "coeffs_fit")}
    seen = set()
    for line in out.splitlines():
        name, size, last = line.split()
        cls, member = mirrors[name]
        assert C.sizeof(cls) == int(size), name
        assert getattr(cls, member).offset == int(last), (name, member)
        seen.add(name)
    assert seen == set(mirrors)
