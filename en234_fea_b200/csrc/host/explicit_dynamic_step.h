#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../../../include/en234gpu.h"
#include "model.h"

namespace en234 {

struct DynamicResult {
    std::vector<double> dof_total, velocity, acceleration, rforce, lumped_mass;
    double time = 0.0, device_ms = 0.0;
    int steps = 0, n_state_prints = 0;
    std::string error;
};

// src/explicit_dynamic_step.f90:1-83 on the device; n_steps <= 0: the input's NUMBER OF STEPS
// partitioned runs (g2l / local_elements as in dynamic_load_tables): every rank runs this loop on its rank-local arrays of
// n_local DOFs; state prints are single-device and skipped
int explicit_dynamic_step(const Model& m, en234gpu_handle h, int n_steps, const double* initial_velocity, FILE* state_print,
                          DynamicResult& res, const std::vector<int>* g2l = nullptr, size_t n_local = 0,
                          const std::vector<int>* local_elements = nullptr);

}  // namespace en234
