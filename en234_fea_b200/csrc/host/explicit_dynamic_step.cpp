// Explicit dynamic step driver of the product: the loop of src/explicit_dynamic_step.f90:1-83 with assemble_lumped_mass,
// apply_dynamic_boundaryconditions and assemble_dynamic_force replaced by the en234gpu_dynamic_* entry points
// (include/en234gpu.h).  No numerical work happens here: the host builds the time-independent load tables once and decides
// which steps print.
#include <algorithm>
#include <cstdio>

#include "explicit_dynamic_step.h"
#include "loads.h"
#include "print_state.h"

namespace en234 {

int explicit_dynamic_step(const Model& m, en234gpu_handle h, int n_steps, const double* initial_velocity, FILE* state_print,
                          DynamicResult& res, const std::vector<int>* g2l, size_t n_local, const std::vector<int>* local_elements) {
    if (!m.explicitdynamicstep) { res.error = "the input has no EXPLICIT DYNAMIC STEP block"; return 1; }
    DynamicLoads L;
    if (!dynamic_load_tables(m, L, g2l, local_elements)) { res.error = L.error; return 1; }
    auto ip = [](const std::vector<int>& v) { return v.empty() ? nullptr : v.data(); };
    auto dp = [](const std::vector<double>& v) { return v.empty() ? nullptr : v.data(); };
    if (en234gpu_dynamic_setup(h, L.element_density.data(), (int)L.history_ptr.size() - 1, L.history_ptr.data(), dp(L.history_time),
                               dp(L.history_value), (int)L.load_history.size(), L.load_ptr.data(), ip(L.load_dof), dp(L.load_val),
                               ip(L.load_history), dp(L.load_constant), (int)L.pdof_dof.size(), ip(L.pdof_dof), ip(L.pdof_history),
                               dp(L.pdof_value), ip(L.pdof_mode))) { res.error = en234gpu_last_error(); return -1; }
    if (initial_velocity && en234gpu_dynamic_set_state(h, nullptr, initial_velocity, nullptr, 0.0)) { res.error = en234gpu_last_error(); return -1; }
    const int total = n_steps > 0 ? n_steps : m.no_dynamic_steps;
    const double dt = m.dynamic_timestep;
    const size_t nd = g2l ? n_local : (size_t)m.ndof();
    const bool printing = m.stateprint && state_print != nullptr && m.state_print_steps > 0 && !g2l;
    std::vector<double> ut, du;
    int done = 0, fail = 0;
    double ms = 0.0;
    res.device_ms = 0.0;
    while (done < total) {
        // whole steps up to the next printing step run as one batch on the device
        int batch = total - done;
        if (printing) batch = std::min(batch, m.state_print_steps - 1 - done % m.state_print_steps);
        if (batch > 0) {
            if (en234gpu_dynamic_steps(h, dt, batch, &fail, &ms)) { res.error = en234gpu_last_error(); return -1; }
            res.device_ms += ms;
            done += batch;
            if (fail) { res.error = "non-finite acceleration or non-positive Jacobian in step batch ending at step " + std::to_string(done); return 2; }
            if (done >= total) break;
        }
        if (printing) {                                               // step done + 1 prints: mod(istep, state_print_steps) == 0
            if (en234gpu_dynamic_half_step(h, dt, 1, &fail)) { res.error = en234gpu_last_error(); return -1; }
            ut.resize(nd); du.resize(nd);
            double t = 0.0;
            if (en234gpu_dynamic_get_state(h, ut.data(), du.data(), nullptr, nullptr, nullptr, nullptr, &t)) { res.error = en234gpu_last_error(); return -1; }
            // the reference's header shows TIME + DTIME with DTIME left at the dynamic time step (explicit_dynamic_step.f90:25)
            if (print_state(m, h, t + dt, ut.data(), du.data(), state_print, res.error)) return -1;
            res.n_state_prints++;
            if (en234gpu_dynamic_half_step(h, dt, 2, &fail)) { res.error = en234gpu_last_error(); return -1; }
            done += 1;
            if (fail) { res.error = "non-finite acceleration or non-positive Jacobian in step " + std::to_string(done); return 2; }
        }
    }
    res.steps = done;
    res.dof_total.resize(nd); res.velocity.resize(nd); res.acceleration.resize(nd); res.rforce.resize(nd); res.lumped_mass.resize(nd);
    if (en234gpu_dynamic_get_state(h, res.dof_total.data(), nullptr, res.velocity.data(), res.acceleration.data(), res.rforce.data(),
                                   res.lumped_mass.data(), &res.time)) { res.error = en234gpu_last_error(); return -1; }
    return 0;
}

}  // namespace en234
