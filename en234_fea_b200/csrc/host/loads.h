#pragma once
#include <string>
#include <vector>

#include "model.h"

namespace en234 {

struct StepLoads {
    std::vector<double> element_ext;        // added to the element residuals (pre-BC rhs, counted in rforce)
    std::vector<double> ext;                // added by the BC routine
    std::vector<int> fixed_dof;             // 0-based DOF indices, application order
    std::vector<double> fixed_value;        // prescribed value of dof_total + dof_increment
    std::vector<double> fixed_correction;   // value - (dof_total + dof_increment)
    bool any_element_ext = false, any_ext = false;
};

// g2l != null (partitioned runs): map global DOF -> rank-local DOF (-1 = not on this rank); outputs and the DOF
// arrays are then rank-local with n_local entries
void evaluate_loads(const Model& m, double time, double dtime, const double* dof_total, const double* dof_increment,
                    StepLoads& out, const std::vector<int>* g2l = nullptr, size_t n_local = 0);

// Time-independent load data of an explicit dynamic step (apply_dynamic_boundaryconditions, explicit_dynamic_step.f90:
// 720-1128) in the form en234gpu_dynamic_setup takes: every distributed load / prescribed force becomes the nodal forces of a
// unit magnitude (sparse, distinct DOFs) times a scale the device evaluates each step (a history at TIME+DTIME, or a
// constant); prescribed DOFs become (dof, history | value, mode) with the last entry winning where a DOF is listed twice.
struct DynamicLoads {
    std::vector<int> load_ptr{0}, load_dof, load_history;
    std::vector<double> load_val, load_constant;
    std::vector<int> pdof_dof, pdof_history, pdof_mode;     // mode 0: du = value - u; 1: du = DTIME*value; 2: du = value
    std::vector<double> pdof_value;
    std::vector<int> history_ptr{0};
    std::vector<double> history_time, history_value;
    std::vector<double> element_density;                    // DENSITY, or the third property of a VUEL element
    std::string error;
};
// partitioned runs: g2l maps a global DOF to the rank-local DOF (-1: not on this rank), local_elements lists the rank's
// elements (0-based global numbers); loads and prescribed DOFs are then restricted to the rank's DOFs with their FULL values
// (every copy of an interface DOF receives the whole external force, like the element load vector of the static step)
bool dynamic_load_tables(const Model& m, DynamicLoads& out, const std::vector<int>* g2l = nullptr,
                         const std::vector<int>* local_elements = nullptr);

}  // namespace en234
