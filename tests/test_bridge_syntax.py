"""The Rcpp bridge a maintainer drops into the R package (bridge/bridge.cpp) must type-check against the public header:
R and Rcpp are not installed in the build container, so the bridge is compiled with `g++ -fsyntax-only` against a
minimal mock of the Rcpp / R interfaces it uses (bridge/mock/).  This pins every C-ABI call of the bridge -- argument
order, pointer types, struct fields -- to include/scregclust_b200.h."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bridge_type_checks_against_the_header():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=return-type",
           "-I", os.path.join(ROOT, "bridge", "mock"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "bridge", "bridge.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]


def test_bridge_keeps_the_reference_call_symbols():
    """Rcpp::compileAttributes() derives the .Call symbols from the exported function names: the five routines of
    src/RcppExports.cpp:104-112 that R/RcppExports.R binds must be exported under the same names and arities."""
    src = open(os.path.join(ROOT, "bridge", "bridge.cpp")).read()
    exported = re.findall(r"// \[\[Rcpp::export\]\]\s*\n\s*[\w:<>]+\s+(\w+)\s*\(", src)
    for name in ("coop_lasso", "coef_nnls", "allocate_into_modules", "alloc_array", "reset_array"):
        assert name in exported, name

    def arity(fn):
        body = src[src.index(f" {fn}("):]
        depth, n, i = 0, 1, body.index("(")
        for ch in body[i:]:
            depth += ch in "(<"
            depth -= ch in ")>"
            n += (ch == "," and depth == 1)
            if depth == 0:
                break
        return n
    assert [arity(f) for f in ("coop_lasso", "coef_nnls", "allocate_into_modules", "alloc_array", "reset_array")] == [13, 4, 7, 2, 2]
