// Static step driver of the product: the loop of src/staticstep.f90:1-162 (time stepping, Newton iterations,
// cutbacks) and compute_static_time_increment (:164-246), with the reference's assemble / apply-BC / solve
// calls replaced by the en234gpu C ABI (include/en234gpu.h).  No numerical work happens here.
#include <cmath>
#include <cstdio>

#include "print_state.h"
#include "staticstep.h"

namespace en234 {
namespace {

// user_codes/src/convergencecheck.f90:1-25
bool convergencecheck(const Model& m, int step, int iteration, double correction, double force, double unbalanced,
                      StepResult& res, FILE* log) {
    NewtonRecord r;
    r.step = step; r.iteration = iteration;
    r.correction_norm = correction; r.nodal_force_norm = force; r.unbalanced_force_norm = unbalanced;
    r.ratio = unbalanced / (force + correction);
    r.tolerance = m.newtonraphson_tolerance;
    bool converged = false;
    if (force + correction > 0.0) { if (r.ratio < m.newtonraphson_tolerance) converged = true; }
    else if (unbalanced == 0.0) converged = true;
    r.converged = converged;
    res.newton.push_back(r);
    if (log) {
        // same six lines as format 99042 (D15.6 written here as E15.6)
        std::fprintf(log, "\n\n  Newton Raphson iteration   %5d\n  Correction norm            %15.6E\n"
                          "  Generalized Force norm     %15.6E\n  Out of balance force norm  %15.6E\n"
                          "  Ratio                      %15.6E\n  Tolerance                  %15.6E\n",
                     iteration, correction, force, unbalanced, r.ratio, r.tolerance);
    }
    return converged;
}

struct TimeCtl { double TIME = 0, DTIME = 0; int current_step_number = 1; double abq_PNEWDT = 1.0; };

// src/staticstep.f90:164-246; returns false where the reference executes `stop`
bool compute_static_time_increment(const Model& m, const TimeCtl& t, int iteration, bool converged, bool& cont,
                                   double& newdt, FILE* log) {
    cont = true;
    newdt = t.DTIME;
    if (m.abaqusformat && t.abq_PNEWDT < 1) { newdt = t.abq_PNEWDT * t.DTIME; return true; }
    if (!converged) {
        newdt = t.DTIME / 2.0;
        if (newdt < m.timestep_min) {
            if (log) std::fprintf(log, " Time step has been reduced to the minimum allowable value without convergence \n Analysis has been terminated \n");
            return false;
        }
        if (log) std::fprintf(log, "\n\n +++ Newton-Raphson iterations have not converged +++ \n     Timestep has been reduced to: %13.5G\n", newdt);
        return true;
    }
    if (iteration < m.max_newton_iterations / 5) newdt = 1.25 * t.DTIME;
    if (m.abaqusformat && t.abq_PNEWDT < 2.0) newdt = t.abq_PNEWDT * t.DTIME;
    if (newdt > m.timestep_max) newdt = m.timestep_max;
    if (t.TIME + t.DTIME + newdt >= m.max_total_time) {
        newdt = m.max_total_time - t.TIME - t.DTIME;
        if (newdt <= 0.0) cont = false;
    }
    if (t.current_step_number == m.max_no_steps) cont = false;
    if (log)
        std::fprintf(log, "\n\n    Step number %5d has completed successfully\n    Elapsed time:                   %12.5G\n"
                          "    Time step has been adjusted to: %12.5G\n", t.current_step_number, t.TIME + t.DTIME, newdt);
    return true;
}

}  // namespace

// One device pass of  assemble -> apply BCs  for the current (dof_total, dof_increment); device-resident or
// host-buffer API.  Returns fail.
static int assemble_and_bc(en234gpu_handle h, const Model& m, const StepOptions& o, const StepLoads& L,
                           const std::vector<double>& dof_total, std::vector<double>& dof_increment,
                           std::vector<double>& rforce, double& force_norm, double& unbalanced, std::string& err) {
    int fail = 0;
    if (o.use_host_api) {
        // corrections must be formed against the current increment (solver_direct.f90:705-709)
        std::vector<double> corr(L.fixed_dof.size());
        for (size_t i = 0; i < corr.size(); ++i)
            corr[i] = L.fixed_value[i] - (dof_total[L.fixed_dof[i]] + dof_increment[L.fixed_dof[i]]);
        if (en234gpu_assemble(h, dof_total.data(), dof_increment.data(), L.any_element_ext ? L.element_ext.data() : nullptr,
                              rforce.data(), &force_norm, &fail)) { err = en234gpu_last_error(); return -1; }
        if (fail) return 1;
        if (en234gpu_apply_boundaryconditions(h, L.any_ext ? L.ext.data() : nullptr, (int)corr.size(), L.fixed_dof.data(),
                                              corr.data(), &unbalanced)) { err = en234gpu_last_error(); return -1; }
    } else {
        if (en234gpu_assemble_dev(h, &force_norm, &fail)) { err = en234gpu_last_error(); return -1; }
        if (fail) return 1;
        if (en234gpu_apply_boundaryconditions_dev(h, &unbalanced)) { err = en234gpu_last_error(); return -1; }
    }
    (void)m;
    return 0;
}

int compute_static_step(const Model& m, en234gpu_handle h, const StepOptions& opt, StepResult& res, FILE* log,
                        const std::vector<int>* g2l, size_t n_local) {
    // partitioned runs: every rank executes this same loop on its rank-local DOF arrays; the norms returned by the
    // device entry points are global (all-reduced), so all ranks take identical step/cutback decisions
    const size_t ndof = g2l ? n_local : (size_t)m.ndof();
    res.dof_total.assign(ndof, 0.0);
    res.rforce.assign(ndof, 0.0);
    std::vector<double> dof_increment(ndof, 0.0);
    // every analysis starts from zero Lagrange multipliers (read_input_file.f90:139); the handle keeps them between calls
    if (!m.constraints.empty() && en234gpu_reset_lagrange_multipliers(h)) { res.error = en234gpu_last_error(); return -1; }
    TimeCtl t;
    t.DTIME = m.timestep_initial;
    bool cont = true, converged = false;
    double newdt = 0.0;
    // SOLVER, DIRECT inputs are solved by PCG converged far below the Newton tolerance; CONJUGATE GRADIENT inputs
    // use the reference's rule and constants unless overridden at run time.
    // UNSYMMETRIC stiffness (built-in continuum elements): Jacobi-BiCGSTAB on the assembled matrix
    const int method = (m.unsymmetric_stiffness && opt.method == EN234_SOLVE_PCG_BSR) ? EN234_SOLVE_BICGSTAB_BSR : opt.method;
    const bool bicg = method == EN234_SOLVE_BICGSTAB_BSR;
    const double cg_tol = opt.cg_tolerance > 0 ? opt.cg_tolerance : (m.solvertype == 1 ? (bicg ? 1e-30 : 1e-28) : m.cg_tolerance);
    const int cg_maxit = opt.max_cg_iterations > 0 ? opt.max_cg_iterations : (m.solvertype == 1 ? 100000 : m.max_cg_iterations);
    const bool newton_loop = m.solvertype == 1 || m.nonlinear || opt.check_linear_cg;   // staticstep.f90:60-70 vs :120-132
    StepLoads L;
    int guard = 0;
    while (cont) {
        if (++guard > 100000) { res.error = "time stepping did not terminate"; return 2; }
        m.current_step_number = t.current_step_number;
        evaluate_loads(m, t.TIME, t.DTIME, res.dof_total.data(), dof_increment.data(), L, g2l, n_local);
        if (!opt.use_host_api &&
            en234gpu_upload_state(h, res.dof_total.data(), dof_increment.data(), L.any_element_ext ? L.element_ext.data() : nullptr,
                                  L.any_ext ? L.ext.data() : nullptr, (int)L.fixed_dof.size(), L.fixed_dof.data(),
                                  L.fixed_value.data())) { res.error = en234gpu_last_error(); return -1; }
        double force_norm = 0, unbalanced = 0, correction = 0;
        int fail = assemble_and_bc(h, m, opt, L, res.dof_total, dof_increment, res.rforce, force_norm, unbalanced, res.error);
        if (fail < 0) return -1;
        auto cutback = [&](int iteration) -> bool {
            if (!compute_static_time_increment(m, t, iteration, false, cont, newdt, log)) return false;
            t.DTIME = newdt;
            std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
            res.n_cutbacks++;
            return true;
        };
        if (fail) { if (!cutback(0)) { res.error = "time step cut below minimum"; return 3; } continue; }
        int its = 0, sfail = 0;
        auto solve = [&]() -> int {
            int rc = opt.use_host_api
                         ? en234gpu_solve(h, method, cg_tol, cg_maxit, dof_increment.data(), &correction, &its, &sfail)
                         : en234gpu_solve_dev(h, method, cg_tol, cg_maxit, &correction, &its, &sfail);
            res.cg_iterations.push_back(its);
            if (rc) { res.error = en234gpu_last_error(); return -1; }
            return sfail;
        };
        fail = solve();
        if (fail < 0) return -1;
        if (fail) { if (!cutback(0)) { res.error = "time step cut below minimum"; return 3; } continue; }
        converged = true;
        int iteration = 0;
        if (newton_loop) {
            for (iteration = 1; iteration <= m.max_newton_iterations; ++iteration) {
                fail = assemble_and_bc(h, m, opt, L, res.dof_total, dof_increment, res.rforce, force_norm, unbalanced, res.error);
                if (fail < 0) return -1;
                if (fail) { converged = false; break; }
                converged = convergencecheck(m, t.current_step_number, iteration, correction, force_norm, unbalanced, res, log);
                if (converged) break;
                fail = solve();
                if (fail < 0) return -1;
                if (fail) { converged = false; break; }
            }
        }
        if (!compute_static_time_increment(m, t, iteration, converged, cont, newdt, log)) { res.error = "time step cut below minimum"; return 3; }
        if (!converged) {
            t.DTIME = newdt;
            std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
            res.n_cutbacks++;
            continue;
        }
        if (!opt.use_host_api &&
            en234gpu_download_state(h, dof_increment.data(), res.rforce.data())) { res.error = en234gpu_last_error(); return -1; }
        // state print: first step, every state_print_steps-th step, last step (staticstep.f90:227-231), before the
        // increment is added to dof_total (:81-86); one device only
        const bool print_now = t.current_step_number == 1 || (m.state_print_steps > 0 && t.current_step_number % m.state_print_steps == 0) || !cont;
        if (m.stateprint && opt.state_print_collective && g2l && print_now) {
            // partitioned run: collective over the ranks (projection on the parts, fields gathered over the communicator);
            // opt.state_print is the open file on the writing rank and null on the others
            if (print_state_partitioned(m, h, t.TIME + t.DTIME, res.dof_total.data(), dof_increment.data(), *g2l, ndof, opt.state_print,
                                        res.error)) return -1;
            res.n_state_prints++;
        }
        // NB a linear solve that fails inside the Newton loop above forces a cutback here; the reference ignores the flag at
        // that point (staticstep.f90:130) — deliberate deviation, documented in INTEGRATION.md
        if (m.stateprint && opt.state_print && !g2l && print_now) {
            if (print_state(m, h, t.TIME + t.DTIME, res.dof_total.data(), dof_increment.data(), opt.state_print, res.error)) return -1;
            res.n_state_prints++;
        }
        t.current_step_number++;
        for (size_t d = 0; d < ndof; ++d) res.dof_total[d] += dof_increment[d];
        if (en234gpu_commit_state(h)) { res.error = en234gpu_last_error(); return -1; }   // staticstep.f90:87
        // NB: like the reference (staticstep.f90:81-89) dof_increment is NOT reset here — the next step starts
        // from the previous increment as its predictor.
        res.step_time.push_back(t.TIME + t.DTIME);
        t.TIME += t.DTIME;
        t.DTIME = newdt;
        res.n_steps_completed++;
    }
    return 0;
}

}  // namespace en234
