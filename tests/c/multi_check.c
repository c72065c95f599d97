/* multi_check.c — drives the multi-GPU entry of the C ABI from plain C, through include/en234gpu.h only (no Python, no
 * torch, no MPI): one linear static step (assemble -> boundary conditions -> PCG solve -> re-assemble -> boundary
 * conditions, src/staticstep.f90:44-70) of a jittered hex8 block on ONE device through en234gpu_create, then on several
 * parts through en234gpu_create_multi, and compares displacements, reaction forces, norms and iteration counts.
 *
 *   multi_check nx ny nz n_parts dev0[,dev1,...] [method]      devices may repeat (parts co-scheduled on one GPU)
 * exit code 0: the partitioned result equals the one-device result (1e-12 relative on u and rforce). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "en234gpu.h"

static double jitter(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) & 0xffff) / 65535.0 - 0.5; }

#define CK(call) do { if (call) { fprintf(stderr, "multi_check: %s failed: %s\n", #call, en234gpu_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: multi_check nx ny nz n_parts dev0[,dev1,...] [method]\n"); return 2; }
    const int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]), n_parts = atoi(argv[4]);
    int devices[16], nd = 0;
    for (char* t = strtok(argv[5], ","); t && nd < 16; t = strtok(NULL, ",")) devices[nd++] = atoi(t);
    while (nd < n_parts) { devices[nd] = devices[nd - 1]; ++nd; }
    const int method = argc > 6 ? atoi(argv[6]) : EN234_SOLVE_PCG_BSR;
    const int N = (nx + 1) * (ny + 1) * (nz + 1), E = nx * ny * nz, ndof = 3 * N;
    double* xyz = malloc(sizeof(double) * 3 * N);
    int* conn = malloc(sizeof(int) * 8 * E);
    int* flag = malloc(sizeof(int) * E);
    int* pidx = calloc(E, sizeof(int));
    const double props[2] = {100.0, 0.3};
    unsigned seed = 12345u;
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                const int n = (k * (ny + 1) + j) * (nx + 1) + i;
                const int inner = i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz;
                xyz[3 * n] = i + (inner ? 0.3 * jitter(&seed) : 0.0);
                xyz[3 * n + 1] = j + (inner ? 0.3 * jitter(&seed) : 0.0);
                xyz[3 * n + 2] = k + (inner ? 0.3 * jitter(&seed) : 0.0);
            }
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int e = (k * ny + j) * nx + i, n0 = (k * (ny + 1) + j) * (nx + 1) + i + 1, up = (nx + 1) * (ny + 1);
                int* c = conn + 8 * e;                       /* ABAQUS C3D8 order, 1-based */
                c[0] = n0; c[1] = n0 + 1; c[2] = n0 + 1 + nx + 1; c[3] = n0 + nx + 1;
                c[4] = c[0] + up; c[5] = c[1] + up; c[6] = c[2] + up; c[7] = c[3] + up;
                flag[e] = EN234_ID_LINELAST;
            }
    /* face x = 0 clamped, nodal forces (0, 0, -0.01) on the face x = nx */
    int nfix = 0;
    int* fixdof = malloc(sizeof(int) * 3 * (ny + 1) * (nz + 1));
    double* fixval = calloc(3 * (ny + 1) * (nz + 1), sizeof(double));
    double* fext = calloc(ndof, sizeof(double));
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j) {
            const int n0 = (k * (ny + 1) + j) * (nx + 1), n1 = n0 + nx;
            for (int c = 0; c < 3; ++c) fixdof[nfix++] = 3 * n0 + c;
            fext[3 * n1 + 2] = -0.01;
        }
    const double tol = 1e-24;
    const int maxit = 20000;
    double *u[2], *rf[2], nrm[2][4];
    int its[2];
    for (int run = 0; run < 2; ++run) {
        double* ut = calloc(ndof, sizeof(double));
        double* du = calloc(ndof, sizeof(double));
        double* r = calloc(ndof, sizeof(double));
        double* corr = calloc(nfix, sizeof(double));
        int fail = 0, sfail = 0;
        if (run == 0) {
            en234gpu_handle h;
            CK(en234gpu_create(&h, devices[0], N, E, xyz, conn, flag, pidx, props, 2));
            CK(en234gpu_set_operator_mode(h, method));
            CK(en234gpu_assemble(h, ut, du, NULL, r, &nrm[run][0], &fail));
            CK(en234gpu_apply_boundaryconditions(h, fext, nfix, fixdof, corr, &nrm[run][1]));
            CK(en234gpu_solve(h, method, tol, maxit, du, &nrm[run][2], &its[run], &sfail));
            CK(en234gpu_assemble(h, ut, du, NULL, r, &nrm[run][0], &fail));
            for (int k = 0; k < nfix; ++k) corr[k] = fixval[k] - du[fixdof[k]];
            CK(en234gpu_apply_boundaryconditions(h, fext, nfix, fixdof, corr, &nrm[run][3]));
            CK(en234gpu_destroy(h));
        } else {
            en234gpu_multi_handle h;
            CK(en234gpu_create_multi(&h, n_parts, devices, N, E, xyz, conn, flag, pidx, props, 2));
            CK(en234gpu_multi_set_operator_mode(h, method));
            CK(en234gpu_multi_assemble(h, ut, du, NULL, r, &nrm[run][0], &fail));
            CK(en234gpu_multi_apply_boundaryconditions(h, fext, nfix, fixdof, corr, &nrm[run][1]));
            CK(en234gpu_multi_solve(h, method, tol, maxit, du, &nrm[run][2], &its[run], &sfail));
            CK(en234gpu_multi_assemble(h, ut, du, NULL, r, &nrm[run][0], &fail));
            for (int k = 0; k < nfix; ++k) corr[k] = fixval[k] - du[fixdof[k]];
            CK(en234gpu_multi_apply_boundaryconditions(h, fext, nfix, fixdof, corr, &nrm[run][3]));
            struct en234gpu_stats st;
            CK(en234gpu_multi_get_stats(h, &st));
            printf("multi_check: %d parts, solve %.3f ms, %lld kernel launches\n", n_parts, st.solve_ms, st.kernel_launches);
            CK(en234gpu_multi_destroy(h));
        }
        if (fail || sfail) { fprintf(stderr, "multi_check: run %d reported fail (assembly %d, solve %d)\n", run, fail, sfail); return 1; }
        u[run] = du; rf[run] = r;
        free(ut); free(corr);
    }
    double eu = 0, su = 0, er = 0, sr = 0;
    for (int d = 0; d < ndof; ++d) {
        eu = fmax(eu, fabs(u[0][d] - u[1][d])); su = fmax(su, fabs(u[0][d]));
        er = fmax(er, fabs(rf[0][d] - rf[1][d])); sr = fmax(sr, fabs(rf[0][d]));
    }
    const double final_ratio = nrm[1][3] / nrm[1][0];
    printf("multi_check: %dx%dx%d, method %d, %d parts vs one device: max rel u %.3e, max rel rforce %.3e, CG iterations %d vs %d, "
           "correction norm %.15e vs %.15e, out-of-balance ratio %.2e\n",
           nx, ny, nz, method, n_parts, eu / su, er / sr, its[1], its[0], nrm[1][2], nrm[0][2], final_ratio);
    const int ok = eu / su < 1e-12 && er / sr < 1e-10 && abs(its[0] - its[1]) <= 2 + its[0] / 25 && final_ratio < 1e-9 &&
                   fabs(nrm[0][2] - nrm[1][2]) < 1e-11 * nrm[0][2];
    printf("multi_check: %s\n", ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}
