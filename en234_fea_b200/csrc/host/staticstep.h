#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../../../include/en234gpu.h"
#include "loads.h"
#include "model.h"

namespace en234 {

struct NewtonRecord {          // one print of user_codes/src/convergencecheck.f90:10-18
    int step, iteration;
    double correction_norm, nodal_force_norm, unbalanced_force_norm, ratio, tolerance;
    int converged;
};

struct StepOptions {
    int method = EN234_SOLVE_PCG_BSR;
    double cg_tolerance = 0.0;        // <= 0: reference value for CONJUGATE GRADIENT inputs, 1e-28 for DIRECT inputs
    int max_cg_iterations = 0;        // <= 0: reference value / 100000
    bool check_linear_cg = false;     // run the re-assemble + convergencecheck loop for LINEAR + CG too
    bool use_host_api = false;        // true: host-buffer entry points (what a Fortran driver binds) each call
    FILE* state_print = nullptr;      // open Tecplot file for the PRINT STATE block (null: state prints are skipped)
    bool state_print_collective = false;   // partitioned runs: every rank takes part in the state prints, one rank holds the file
};

struct StepResult {
    std::vector<double> dof_total, rforce;
    std::vector<NewtonRecord> newton;
    std::vector<int> cg_iterations;
    std::vector<double> step_time;
    int n_steps_completed = 0, n_cutbacks = 0, n_state_prints = 0;
    std::string error;
};

int compute_static_step(const Model& m, en234gpu_handle h, const StepOptions& opt, StepResult& res, FILE* log,
                        const std::vector<int>* g2l = nullptr, size_t n_local = 0);

}  // namespace en234
