// en234_run — command-line static analysis on the GPU path:  en234_run <file.in> [--matrix-free] [--cg-tol x]
//             [--cg-maxit n] [--host-api] [--log file.out] [--device d] [--state-print file | --no-state-print]
// Plays the role of `program en234fea` (src/main.f90:127-143) for the static step: read the input file, run
// compute_static_step, report the convergence history.  (The reference hard-codes its input path, main.f90:11-35.)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/en234host.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: en234_run <file.in> [--matrix-free] [--cg-tol x] [--cg-maxit n] [--host-api] [--log f] [--device d] [--state-print f | --no-state-print]\n"); return 2; }
    int method = EN234_SOLVE_PCG_BSR, host_api = 0, device = 0, maxit = 0;
    double tol = 0.0;
    const char* log = nullptr;
    const char* state_print = nullptr;
    bool no_state_print = false;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--matrix-free")) method = EN234_SOLVE_PCG_MATRIXFREE;
        else if (!std::strcmp(argv[i], "--host-api")) host_api = 1;
        else if (!std::strcmp(argv[i], "--cg-tol") && i + 1 < argc) tol = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--cg-maxit") && i + 1 < argc) maxit = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--log") && i + 1 < argc) log = argv[++i];
        else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--state-print") && i + 1 < argc) state_print = argv[++i];
        else if (!std::strcmp(argv[i], "--no-state-print")) no_state_print = true;
    }
    en234host_model_t m = en234host_model_read(argv[1]);
    if (!m) { std::fprintf(stderr, "input error: %s\n", en234host_last_error()); return 1; }
    const int* p; long n;
    en234host_model_int(m, "n_nodes", &p, &n); const int nn = p[0];
    en234host_model_int(m, "n_elements", &p, &n); const int ne = p[0];
    std::printf(" Total no. elements %d\n Total no. nodes %d\n Total no. DOF %d\n", ne, nn, 3 * nn);
    en234gpu_handle h;
    if (en234host_gpu_create(m, device, &h)) { std::fprintf(stderr, "device error: %s\n", en234host_last_error()); return 1; }
    if (en234gpu_set_operator_mode(h, method)) { std::fprintf(stderr, "%s\n", en234gpu_last_error()); return 1; }
    en234host_model_int(m, "checkstiffness_flag", &p, &n);
    if (p[0] != 0) {                                    // CHECK STIFFNESS key: main.f90:136, before the static step
        double err = 0;
        int nu = 0, nnu = 0;
        if (en234host_check_stiffness(m, p[0], nullptr, nullptr, log ? (std::string(log) + ".checkstiffness").c_str() : nullptr, &err, &nu, &nnu))
            std::fprintf(stderr, "check stiffness: %s\n", en234host_last_error());
        else
            std::printf(" STIFFNESS CHECK element %d: normalized difference %.3e (%s), unsymmetric entries %d / %d\n", p[0], err,
                        err > 1e-6 ? "NOT consistent" : "consistent with residual", nu, nnu);
    }
    {   // PRINT STATE block: written to the file the input names (print_state, src/printing.f90) unless overridden
        int flags = 0, steps = 1;
        double scale = 1.0;
        char fname[512] = {0};
        en234host_model_get_state_print(m, &flags, &steps, &scale, nullptr, 0, fname, (int)sizeof fname);
        if ((flags & 1) && !no_state_print) {
            const char* path = state_print ? state_print : fname;
            if (en234host_model_set_state_print(m, flags, steps, scale, nullptr, path)) { std::fprintf(stderr, "%s\n", en234host_last_error()); return 1; }
            std::printf(" state prints -> %s\n", path);
        }
    }
    en234host_model_int(m, "explicitdynamicstep", &p, &n);
    if (p[0]) {                                         // EXPLICIT DYNAMIC STEP (main.f90:139 explicit_dynamic_step)
        std::vector<double> u(3 * (size_t)nn), v(3 * (size_t)nn);
        long long dinfo[8] = {0};
        double ms = 0.0;
        if (en234host_run_dynamic(m, h, 0, nullptr, u.data(), v.data(), nullptr, nullptr, nullptr, dinfo, &ms)) {
            std::fprintf(stderr, "explicit dynamic step failed: %s\n", en234host_last_error());
            return 1;
        }
        double umax = 0.0, vmax = 0.0;
        for (size_t d = 0; d < u.size(); ++d) { umax = std::fmax(umax, std::fabs(u[d])); vmax = std::fmax(vmax, std::fabs(v[d])); }
        std::printf(" explicit dynamic steps %lld, state prints %lld, kernels launched %lld, device ms %.3f\n max |u| %.6e  max |v| %.6e\n",
                    dinfo[0], dinfo[1], dinfo[2], ms, umax, vmax);
        en234gpu_destroy(h);
        en234host_model_free(m);
        return 0;
    }
    std::vector<double> u(3 * (size_t)nn), rf(3 * (size_t)nn), newton(8 * 4096);
    std::vector<int> cg(4096);
    long long info[8] = {0};
    int rc = en234host_run_static(m, h, method, tol, maxit, 1, host_api, u.data(), rf.data(), newton.data(), 4096, cg.data(), 4096, info, log);
    if (rc) { std::fprintf(stderr, "static step failed: %s\n", en234host_last_error()); return 1; }
    std::printf(" steps completed %lld, cutbacks %lld, state prints %lld\n", info[0], info[1], info[4]);
    for (int i = 0; i < info[2]; ++i)
        std::printf(" step %3d it %2d  correction %12.5e  force %12.5e  unbalanced %12.5e  ratio %12.5e  %s\n",
                    (int)newton[8 * i], (int)newton[8 * i + 1], newton[8 * i + 2], newton[8 * i + 3], newton[8 * i + 4],
                    newton[8 * i + 5], newton[8 * i + 7] != 0 ? "converged" : "");
    for (int i = 0; i < info[3]; ++i) std::printf(" CG solve %d: %d iterations\n", i + 1, cg[i]);
    en234gpu_stats st;
    en234gpu_get_stats(h, &st);
    std::printf(" device ms: assemble %.3f  bc %.3f  solve %.3f  kernels launched %lld\n", st.assemble_ms, st.bc_ms, st.solve_ms, st.kernel_launches);
    en234gpu_destroy(h);
    en234host_model_free(m);
    return 0;
}
