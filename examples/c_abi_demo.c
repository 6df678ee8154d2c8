/* c_abi_demo.c — a plain-C consumer of libdynoed_cuda.so (what a Julia `ccall` / cgo / JNI binding
 * does): Lotka-Volterra fishing OED (reference model /root/reference/test/references/
 * LotkaVolterra.jl:10-36), FisherD criterion, RK4 with 4 substeps on the merged grid t_k = k/10.
 * Evaluates the reference test's optimum (u = [0.62, 0, ...], w1 = w2 = last 4 of 30 on) and a
 * small batch, prints objective, F and a few gradient entries.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo \
 *       -Ldynamicoed.jl_b200 -ldynoed_cuda -Wl,-rpath,$PWD/dynamicoed.jl_b200 && /tmp/c_abi_demo
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dynoed.h"

#define N 30      /* merged intervals */
#define NVAR 71   /* 10 u | 30 w1 | 30 w2 | tau   (src/discretize.jl:110-137, sorted keys) */

int main(void) {
    double tgrid[N + 1];
    int32_t ind[3 * N];                 /* rows: u, w1, w2; 0-based interval of each variable */
    int j, k;
    for (j = 0; j <= N; ++j) tgrid[j] = j / 10.0;          /* Julia ranges give the double nearest k/10 */
    for (j = 0; j < N; ++j) { ind[j] = j / 3; ind[N + j] = j; ind[2 * N + j] = j; }
    int32_t ctrl_off[1] = {0}, w_off[2] = {10, 40};

    dynoed_problem_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    d.model = "lv_fishing";
    d.criterion = DYNOED_FISHER_D;
    d.method = DYNOED_RK4;
    d.n_intervals = N;
    d.tgrid = tgrid;
    d.substeps = 4;
    d.indicators = ind;
    d.ctrl_offsets = ctrl_off;
    d.w_offsets = w_off;
    d.ic_offset = 0;
    d.tau_offset = 70;
    d.n_var = NVAR;

    dynoed_ctx* ctx = NULL;
    if (dynoed_create(&d, 0, &ctx) != DYNOED_OK) {
        fprintf(stderr, "dynoed_create: %s\n", dynoed_last_error(NULL));
        return 2;
    }
    enum { B = 4 };
    double des[B * NVAR], crit[B], grad[B * NVAR], fim[B * 3];
    memset(des, 0, sizeof des);
    des[0] = 0.62;                                           /* design 0: the reference optimum */
    for (k = 26; k < 30; ++k) { des[10 + k] = 1.0; des[40 + k] = 1.0; }
    des[70] = 1.0;
    for (j = 1; j < B; ++j) {                                /* designs 1..: simple deterministic fill */
        for (k = 0; k < NVAR; ++k) des[j * NVAR + k] = ((j * 37 + k * 11) % 17) / 17.0;
        des[j * NVAR + 70] = 0.1 * j;
    }
    if (dynoed_eval_batch(ctx, des, B, crit, grad, fim, 0) != DYNOED_OK) {
        fprintf(stderr, "dynoed_eval_batch: %s\n", dynoed_last_error(ctx));
        return 3;
    }
    int64_t best; double bv;
    dynoed_argmin(ctx, &best, &bv);
    for (j = 0; j < B; ++j)
        printf("design %d objective %.17g F %.17g %.17g %.17g grad[0] %.17g grad[39] %.17g grad[70] %.17g\n", j,
               crit[j], fim[3 * j], fim[3 * j + 1], fim[3 * j + 2], grad[j * NVAR], grad[j * NVAR + 39],
               grad[j * NVAR + 70]);
    printf("argmin %lld %.17g\n", (long long)best, bv);
    dynoed_destroy(ctx);
    return 0;
}
