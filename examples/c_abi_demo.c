/* Plain-C caller of the ABI (include/topay_b200.h): what a cgo / JNI / ctypes binding sees.
 * Build:  gcc -std=c99 -I include examples/c_abi_demo.c -L topay_b200 -ltopay_b200 -Wl,-rpath,$PWD/topay_b200 -o c_abi_demo
 * Without a CUDA device every compute entry point refuses with TOPAY_ERR_NO_DEVICE (there is no CPU path);
 * with one, the program builds a small field, solves one candidate and applies the success gate. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "topay_b200.h"

int main(void) {
    printf("%s\n", topay_version());
    topay_robot_params rp;
    topay_opt_params opt;
    topay_robot_params_default(&rp);
    topay_opt_params_default(&opt);
    printf("dof %d, int_K %d, lbfgs mem %d, chassis radius %.3f\n", TOPAY_DOF, opt.int_K, opt.s2_lbfgs.mem_size,
           rp.chassis_colli_radius);

    /* host-only entry points work anywhere */
    int32_t ok[4] = {0, 1, 1, 1};
    double dur[4] = {1.0, 5.0, 3.0, 3.0};
    printf("select_shortest -> %d (expects 2)\n", topay_select_shortest(ok, dur, 4));

    topay_grid_desc gd = {{10.0, 10.0, 1.6}, 0.1, 0.4, 0.155};
    topay_field* field = NULL;
    int rc = topay_field_create(&gd, 0, &field);
    if (rc == TOPAY_ERR_NO_DEVICE) {
        printf("no CUDA device: %s\n", topay_strerror(rc));
        return 0;
    }
    if (rc != TOPAY_OK) {
        printf("field_create failed: %s\n", topay_last_error());
        return 1;
    }
    /* one post next to a straight 4 m path */
    float post[3 * 8];
    for (int i = 0; i < 8; i++) {
        post[3 * i] = 0.0f;
        post[3 * i + 1] = 1.2f;
        post[3 * i + 2] = 0.05f + 0.1f * (float)i;
    }
    topay_field_clear(field, 1);
    topay_field_rasterize_points(field, post, 8);
    rc = topay_field_rebuild(field);
    if (rc != TOPAY_OK) return 1;

    topay_solver* solver = NULL;
    rc = topay_solver_create(&opt, &rp, field, 1, 16, &solver);
    if (rc != TOPAY_OK) {
        printf("solver_create failed: %s\n", topay_last_error());
        return 1;
    }
    enum { LEN = 9 };
    double path[LEN * 10];
    memset(path, 0, sizeof(path));
    for (int i = 0; i < LEN; i++) {
        path[10 * i] = -2.0 + 0.5 * i;          /* x */
        path[10 * i + 1] = 0.2 * sin(0.7 * i);  /* y */
    }
    int32_t len = LEN, status = 0, pieces = 0, evals = 0, best_d = -1, best_c = -1;
    double bvel[20] = {0}, bacc[20] = {0}, cost = 0, duration = 0;
    topay_result_batch res;
    memset(&res, 0, sizeof(res));
    res.status = &status;
    res.piece_num = &pieces;
    res.evals = &evals;
    res.cost = &cost;
    res.duration = &duration;
    rc = topay_solver_solve_batch(solver, 1, &len, path, bvel, bacc, &res, &best_d, &best_c);
    if (rc != TOPAY_OK) {
        printf("solve failed: %s\n", topay_last_error());
        return 1;
    }
    int32_t feasible = 0, feasible_print = 0, winner = -1;
    topay_feasibility fz;
    memset(&fz, 0, sizeof(fz));
    fz.feasible = &feasible;
    fz.feasible_print = &feasible_print;
    rc = topay_solver_check_feasible(solver, &fz, &winner);
    printf("status %d, pieces %d, evals %d, cost %.3f, duration %.3f s, feasible %d, winner %d\n", status, pieces, evals,
           cost, duration, feasible, winner);
    topay_solver_destroy(solver);
    topay_field_destroy(field);
    return rc == TOPAY_OK ? 0 : 1;
}
