/* Plain-C consumer of include/jgb200.h, linked against libjgb200.so without ctypes: catches header / export drift at
 * link time (every declared entry point is referenced below) and, on a box with a GPU, runs a 3-bus Newton-Raphson power
 * flow through the same calls the Julia shim makes (jgb_nr_setup ... jgb_nr_run).
 *   gcc -std=c99 -I include tests/abi_c_test.c -L juliagrid.jl_b200 -ljgb200 -lm -o abi_c_test
 * prints "symbols N" and either "no device (rc -5)" or "nr ok ...". Exit code 0 on success. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "jgb200.h"

#define SYM(f) { #f, (void (*)(void))f }
static const struct { const char* name; void (*fn)(void); } table[] = {
    SYM(jgb_create), SYM(jgb_destroy), SYM(jgb_last_error), SYM(jgb_abi_version), SYM(jgb_synchronize),
    SYM(jgb_nr_setup), SYM(jgb_nr_dims), SYM(jgb_nr_pattern), SYM(jgb_nr_set_injection), SYM(jgb_nr_set_state),
    SYM(jgb_nr_get_state), SYM(jgb_nr_update_y), SYM(jgb_nr_mismatch), SYM(jgb_nr_solve), SYM(jgb_nr_get_vectors),
    SYM(jgb_nr_run), SYM(jgb_nr_batch), SYM(jgb_nr_batch_dev), SYM(jgb_nr_set_branches), SYM(jgb_nr_power),
    SYM(jgb_wls_setup), SYM(jgb_wls_dims), SYM(jgb_wls_gain_pattern), SYM(jgb_wls_set_mean), SYM(jgb_wls_set_state),
    SYM(jgb_wls_get_state), SYM(jgb_wls_increment), SYM(jgb_wls_solve), SYM(jgb_wls_get_vectors), SYM(jgb_wls_run),
    SYM(jgb_wls_batch), SYM(jgb_wls_batch_dev), SYM(jgb_wls_update_rows), SYM(jgb_wls_update_y),
    SYM(jgb_wls_update_branch), SYM(jgb_wls_residual_test), SYM(jgb_wls_remove_row),
    SYM(jgb_lin_setup), SYM(jgb_lin_refactor), SYM(jgb_lin_projection), SYM(jgb_lin_solve),
    SYM(jgb_lin_solve_projected), SYM(jgb_lin_solve_dev), SYM(jgb_lin_dims),
    SYM(jgb_fnr_setup), SYM(jgb_fnr_set_injection), SYM(jgb_fnr_set_state), SYM(jgb_fnr_get_state),
    SYM(jgb_fnr_mismatch), SYM(jgb_fnr_solve), SYM(jgb_fnr_run), SYM(jgb_fnr_batch),
    SYM(jgb_comm_unique_id), SYM(jgb_comm_init), SYM(jgb_allgather_states), SYM(jgb_comm_wait), SYM(jgb_comm_size),
    SYM(jgb_stat), SYM(jgb_profile), SYM(jgb_selfcheck_symbolic), SYM(jgb_selfcheck_tree), SYM(jgb_selfcheck_tasks),
};

int main(void) {
    const int nsym = (int)(sizeof(table) / sizeof(table[0]));
    for (int i = 0; i < nsym; ++i)
        if (!table[i].fn) { printf("missing %s\n", table[i].name); return 2; }
    printf("symbols %d abi %d\n", nsym, jgb_abi_version());
    int32_t rc = 0;
    jgb_ctx* ctx = jgb_create(0, NULL, &rc);
    if (!ctx) {
        printf("no device (rc %d): %s\n", rc, jgb_last_error(NULL));
        return rc == -5 ? 0 : 3;
    }
    /* 3 buses (1 slack, 2 and 3 PQ), three identical lines z = 0.01 + 0.1i between every pair: Ybus is full */
    const double zr = 0.01, zx = 0.1, d = zr * zr + zx * zx;
    const double yr = zr / d, yi = -zx / d;
    int64_t colptr[4] = {1, 4, 7, 10}, rowval[9] = {1, 2, 3, 1, 2, 3, 1, 2, 3};
    double y[18], yt[18];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            const int q = 3 * c + r;
            y[2 * q] = (r == c) ? 2 * yr : -yr;
            y[2 * q + 1] = (r == c) ? 2 * yi : -yi;
        }
    memcpy(yt, y, sizeof(y));                      /* symmetric network */
    int8_t type[3] = {3, 1, 1};
    if (jgb_nr_setup(ctx, 3, colptr, rowval, y, yt, type, 1) != 0) { printf("setup: %s\n", jgb_last_error(ctx)); return 4; }
    int64_t dim = 0, nnz = 0;
    jgb_nr_dims(ctx, &dim, &nnz);
    if (dim != 4 || nnz != 16) { printf("dims %lld %lld\n", (long long)dim, (long long)nnz); return 5; }
    double ps[3] = {0, 0, 0}, qs[3] = {0, 0, 0}, pd[3] = {0, 0.5, 0.3}, qd[3] = {0, 0.2, 0.1};
    double vm[3] = {1.02, 1, 1}, va[3] = {0, 0, 0};
    if (jgb_nr_set_injection(ctx, ps, qs, pd, qd) || jgb_nr_set_state(ctx, vm, va)) return 6;
    int64_t it = 0;
    double sp = 0, sq = 0;
    rc = jgb_nr_run(ctx, 20, 1e-10, &it, &sp, &sq);
    if (rc != 0 || it < 2 || it > 6 || !(sp < 1e-10) || !(sq < 1e-10)) { printf("run rc %d it %lld\n", rc, (long long)it); return 7; }
    jgb_nr_get_state(ctx, vm, va);
    /* check bus 2 against the power-flow equations evaluated here */
    double p2 = 0, q2 = 0;
    for (int j = 0; j < 3; ++j) {
        const double g = (j == 1) ? 2 * yr : -yr, b = (j == 1) ? 2 * yi : -yi, t = va[1] - va[j];
        p2 += vm[1] * vm[j] * (g * cos(t) + b * sin(t));
        q2 += vm[1] * vm[j] * (g * sin(t) - b * cos(t));
    }
    if (fabs(p2 + 0.5) > 1e-9 || fabs(q2 + 0.2) > 1e-9) { printf("mismatch %g %g\n", p2 + 0.5, q2 + 0.2); return 8; }
    /* an out-of-range argument is an error code with a message, never a crash */
    if (jgb_nr_update_y(ctx, 1, &colptr[3], y, yt) != -1 || !strlen(jgb_last_error(ctx))) return 9;
    printf("nr ok: %lld iterations, V2 = %.6f /_ %.6f rad\n", (long long)it, vm[1], va[1]);
    jgb_destroy(ctx);
    return 0;
}
