/* Plain C99 client of include/scregclust_b200.h: proves that the boundary is a C ABI (no C++ or
 * torch types), checks the argument errors that must be reported before any device work with the
 * reference's message texts (src/optim.cpp:77,81,116,121,409), and -- when a device is present --
 * runs one tiny coop_lasso / coef_nnls through the drop-in entry points.
 * Built and run by tests/test_capi_cpu.py (CPU: error paths) and tests/test_gpu_kernels.py (GPU). */
#include "scregclust_b200.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

static int expect(int rc, int want, const char* needle) {
    const char* msg = scrc_last_error();
    if (rc != want || (needle && !strstr(msg, needle))) {
        printf("FAIL rc=%d (want %d) msg='%s' (want '%s')\n", rc, want, msg, needle ? needle : "");
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    int bad = 0, it = 0, iters[2];
    double y[8] = {1.0, -0.5, 0.25, 2.0, 0.5, 1.5, -1.0, 0.75};
    double x[8] = {1.0, 0.0, 0.5, -1.0, 0.0, 1.0, 0.5, 0.25};
    double w[2] = {1.0, 1.0}, beta[4];
    scrc_options o;
    scrc_default_options(&o);
    if (scrc_version() <= 0 || o.max_optim_iter != 10000 || o.n_update != 2) { printf("FAIL defaults\n"); bad = 1; }

    /* argument checks happen on the host, before the device is touched */
    bad |= expect(scrc_coop_lasso(y, 0, 2, x, 0, 2, 0.1, w, NULL, 0, 0, 0.2, 1.5, 2, 0.2, 100, 1e-8, 1e-12, beta, &it),
                  SCRC_ERR_COOP_DIMS, "Matrix dimensions of y and x need to be positive");
    bad |= expect(scrc_coop_lasso(y, 4, 2, x, 3, 2, 0.1, w, NULL, 0, 0, 0.2, 1.5, 2, 0.2, 100, 1e-8, 1e-12, beta, &it),
                  SCRC_ERR_COOP_ROWS, "same number of rows");
    bad |= expect(scrc_coop_lasso(y, 4, 2, x, 4, 2, 0.1, w, NULL, 0, 0, 0.0, 1.5, 2, 0.2, 100, 1e-8, 1e-12, beta, &it),
                  SCRC_ERR_COOP_RHO, "rho_0 > 0");
    bad |= expect(scrc_coop_lasso(y, 4, 2, x, 4, 2, 0.1, w, NULL, 0, 0, 0.2, 2.5, 2, 0.2, 100, 1e-8, 1e-12, beta, &it),
                  SCRC_ERR_COOP_ALPHA, "alpha_0");
    bad |= expect(scrc_coef_nnls(x, 0, 2, y, 0, 2, 1e-12, 100, beta, iters), SCRC_ERR_NNLS_DIMS, "NNLS: Matrix dimensions");

    if (argc > 1 && strcmp(argv[1], "--gpu") == 0) {
        int rc = scrc_coop_lasso(y, 4, 2, x, 4, 2, 0.05, w, NULL, 0, 0, 0.2, 1.5, 2, 0.2, 1000, 1e-8, 1e-12, beta, &it);
        if (rc != SCRC_OK || it <= 0 || !isfinite(beta[0])) { printf("FAIL coop_lasso rc=%d it=%d %s\n", rc, it, scrc_last_error()); bad = 1; }
        rc = scrc_coef_nnls(x, 4, 2, y, 4, 2, 1e-12, 1000, beta, iters);
        if (rc != SCRC_OK || beta[0] < 0.0 || beta[1] < 0.0 || beta[2] < 0.0 || beta[3] < 0.0) { printf("FAIL nnls rc=%d %s\n", rc, scrc_last_error()); bad = 1; }
        if (scrc_launch_count() == 0ull) { printf("FAIL no kernels launched\n"); bad = 1; }
    } else {
        /* without a device every compute entry point must fail loudly: no CPU fallback */
        int rc = scrc_coop_lasso(y, 4, 2, x, 4, 2, 0.05, w, NULL, 0, 0, 0.2, 1.5, 2, 0.2, 1000, 1e-8, 1e-12, beta, &it);
        if (rc == SCRC_OK) { printf("FAIL: compute succeeded without a device?\n"); bad = 1; }
    }
    printf(bad ? "C ABI smoke: FAILED\n" : "C ABI smoke: ok\n");
    return bad;
}
