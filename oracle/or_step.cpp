// ORACLE — test infrastructure only (see or_model.h).  Static step driver: restates
// src/staticstep.f90 (compute_static_step :1-162, compute_static_time_increment :164-246) and
// user_codes/src/convergencecheck.f90:1-25.
#include <cmath>
#include <cstdio>

#include "or_model.h"

namespace oracle {

static bool convergencecheck(const Model& m, int step, int iteration, const AsmOut& a, StepLog& log) {
    NewtonRecord r;
    r.step = step;
    r.iteration = iteration;
    r.correction_norm = a.correction_norm;
    r.nodal_force_norm = a.nodal_force_norm;
    r.unbalanced_force_norm = a.unbalanced_force_norm;
    r.ratio = a.unbalanced_force_norm / (a.nodal_force_norm + a.correction_norm);
    r.tolerance = m.newtonraphson_tolerance;
    bool converged = false;
    if (a.nodal_force_norm + a.correction_norm > 0.0) {
        if (r.ratio < m.newtonraphson_tolerance) converged = true;
    } else if (a.unbalanced_force_norm == 0.0) converged = true;
    r.converged = converged;
    log.newton.push_back(r);
    return converged;
}

struct TimeCtl {
    double TIME = 0, DTIME = 0;
    int current_step_number = 1;
    double abq_PNEWDT = 1.0;
};

// compute_static_time_increment (staticstep.f90:164-246).  Returns false if the analysis must stop
// (the reference executes `stop` when the step falls below timestep_min).
static bool time_increment(const Model& m, TimeCtl& t, int iteration, bool converged, bool& cont, double& newdt) {
    cont = true;
    newdt = t.DTIME;
    if (m.abaqusformat && t.abq_PNEWDT < 1) { newdt = t.abq_PNEWDT * t.DTIME; return true; }
    if (!converged) {
        newdt = t.DTIME / 2.0;
        if (newdt < m.timestep_min) return false;
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
    return true;
}

int compute_static_step(const Model& m, std::vector<double>& dof_total, std::vector<double>& rforce,
                        std::vector<double>& svars, StepLog& log) {
    const int ndof = m.ndof();
    std::vector<double> dof_increment(ndof, 0.0);
    long long nsv = 0;
    for (int e = 0; e < m.n_elements; ++e) nsv += m.element_n_states[e];
    std::vector<double> updated(nsv, 0.0);
    if ((long long)svars.size() != nsv) svars.assign(nsv, 0.0);
    TimeCtl t;
    t.DTIME = m.timestep_initial;
    AsmOut a;
    bool cont = true, converged = false;
    double newdt = 0;
    int guard = 0;
    if (m.solvertype != 1 && m.n_constraints() > 0) return 4;       // constraints: skyline path only in this restatement
    if (m.solvertype == 1) {                                           // direct :31-91
        Skyline s;
        std::vector<int> order(m.n_nodes);
        if ((int)m.node_order.size() == m.n_nodes || (int)m.node_order.size() == m.n_nodes + m.n_constraints()) order = m.node_order;
        else for (int i = 0; i < m.n_nodes; ++i) order[i] = i;
        compute_profile(m, order, s, &log);
        while (cont) {
            if (++guard > 100000) return 2;
            m.current_step_number = t.current_step_number;
            updated = svars;
            int fail = assemble_direct_stiffness(m, s, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, &updated, a);
            if (fail) {
                if (!time_increment(m, t, 0, false, cont, newdt)) return 3;
                t.DTIME = newdt;
                std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
                log.n_cutbacks++;
                continue;
            }
            apply_direct_boundaryconditions(m, s, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, a);
            solve_direct(m, s, dof_increment.data(), a, &log);
            converged = true;
            int iteration;
            for (iteration = 1; iteration <= m.max_newton_iterations; ++iteration) {
                updated = svars;
                fail = assemble_direct_stiffness(m, s, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, &updated, a);
                if (fail) break;
                apply_direct_boundaryconditions(m, s, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, a);
                converged = convergencecheck(m, t.current_step_number, iteration, a, log);
                if (converged) break;
                solve_direct(m, s, dof_increment.data(), a, &log);
            }
            if (!time_increment(m, t, iteration, converged, cont, newdt)) return 3;
            if (!converged) {
                std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
                t.DTIME = newdt;
                log.n_cutbacks++;
                continue;
            }
            t.current_step_number++;
            for (int d = 0; d < ndof; ++d) dof_total[d] += dof_increment[d];
            svars = updated;
            log.step_time.push_back(t.TIME + t.DTIME);
            log.step_dtime.push_back(t.DTIME);
            t.TIME += t.DTIME;
            t.DTIME = newdt;
            log.n_steps_completed++;
        }
        log.lagrange_multipliers = s.lagrange_multipliers;
    } else {                                                           // conjugate gradient :93-160
        CooUpper c;
        initialize_cg_solver(m, c);
        while (cont) {
            if (++guard > 100000) return 2;
            m.current_step_number = t.current_step_number;
            updated = svars;
            int fail = assemble_conjugate_gradient_stiffness(m, c, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, &updated, a);
            if (fail) {
                if (!time_increment(m, t, 0, false, cont, newdt)) return 3;
                std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
                t.DTIME = newdt;   // the reference omits this at :104 (a latent infinite loop); applied here
                log.n_cutbacks++;
                continue;
            }
            apply_conjugategradient_boundaryconditions(m, c, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, a, m.cg_apply_nodal_forces);
            int its = 0;
            fail = solve_conjugategradient(m, c, dof_increment.data(), a, &its);
            log.cg_iterations.push_back(its);
            if (fail) {
                if (!time_increment(m, t, 0, false, cont, newdt)) return 3;
                t.DTIME = newdt;
                std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
                log.n_cutbacks++;
                continue;
            }
            converged = true;
            int iteration = 0;
            if (m.nonlinear || m.cg_check_linear) {                    // :120-132 (+ Q6 option)
                for (iteration = 1; iteration <= m.max_newton_iterations; ++iteration) {
                    updated = svars;
                    fail = assemble_conjugate_gradient_stiffness(m, c, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, &updated, a);
                    if (fail) { converged = false; break; }
                    apply_conjugategradient_boundaryconditions(m, c, dof_total.data(), dof_increment.data(), t.TIME, t.DTIME, a, m.cg_apply_nodal_forces);
                    converged = convergencecheck(m, t.current_step_number, iteration, a, log);
                    if (converged) break;
                    fail = solve_conjugategradient(m, c, dof_increment.data(), a, &its);
                    log.cg_iterations.push_back(its);
                }
            }
            if (!time_increment(m, t, iteration, converged, cont, newdt)) return 3;
            if (!converged) {
                t.DTIME = newdt;
                std::fill(dof_increment.begin(), dof_increment.end(), 0.0);
                log.n_cutbacks++;
                continue;
            }
            t.current_step_number++;
            for (int d = 0; d < ndof; ++d) dof_total[d] += dof_increment[d];
            svars = updated;
            log.step_time.push_back(t.TIME + t.DTIME);
            log.step_dtime.push_back(t.DTIME);
            t.TIME += t.DTIME;
            t.DTIME = newdt;
            log.n_steps_completed++;
        }
    }
    rforce = a.rforce;
    return 0;
}

}  // namespace oracle
