/* Prints sizeof and the offset of the last member of every struct of the public header; tests/test_capi_cpu.py compares
 * them with the ctypes mirrors in scregclust_b200/_capi.py (a field added on one side only would shift every later one). */
#include <stddef.h>
#include <stdio.h>

#include "scregclust_b200.h"

#define ROW(T, last) printf(#T " %zu %zu\n", sizeof(T), offsetof(T, last))

int main(void) {
    ROW(scrc_options, eps_corr);
    ROW(scrc_cycle_out, admm_sweeps);
    ROW(scrc_prior, aggregate);
    ROW(scrc_quick, percent);
    ROW(scrc_grid_params, cancel);
    ROW(scrc_grid_info, device_calls);
    ROW(scrc_grid_penalty_info, n_records);
    ROW(scrc_grid_final, importance);
    ROW(scrc_grid_backend, coeffs_fit);
    return 0;
}
