#pragma once
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

}  // namespace en234
