// State print of the static step: nodal projection of integration-point fields on the device + Tecplot FEPOINT
// writer (src/printing.f90:249-1093), 3-D hex8 zones.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../../../include/en234gpu.h"
#include "model.h"

namespace en234 {

// FIELD VARIABLES names -> en234gpu_project_fields codes, with the reference's prefix matching (src/parse.f90:155-176)
// and the per-element-type lists of el_linelast_3Dbasic.f90:392-408 / abaqus_tecplot_interface.f90:163-179 /
// continuum_elements.f90:1291-1345.  Names no element routine recognises map to EN234_FIELD_NONE (a zero field).
std::vector<int> field_codes(const Model& m, const std::vector<std::string>& names);

// fields[k * n_nodes + n]; dof arrays may be null (device-resident state)
int project_fields(const Model& m, en234gpu_handle h, const std::vector<std::string>& names, bool lumped,
                   const double* dof_total, const double* dof_increment, std::vector<double>& fields, std::string& err);

// one ZONE of the Tecplot file, appended to `f` (the reference keeps the unit open over the whole analysis)
int print_state(const Model& m, en234gpu_handle h, double time_end, const double* dof_total, const double* dof_increment,
                FILE* f, std::string& err);

// the writer alone: whole-mesh DOF arrays and fields[k * n_nodes + n]
void write_state_zone(const Model& m, double time_end, const double* dof_total, const double* dof_increment,
                      const std::vector<double>& fields, FILE* f);
// element partitions (one process per GPU): collective over the ranks; `f` is the open file on the writing rank, null elsewhere
int print_state_partitioned(const Model& m, en234gpu_handle h, double time_end, const double* dof_total_local,
                            const double* dof_increment_local, const std::vector<int>& g2l, size_t n_local_dof, FILE* f,
                            std::string& err);

std::string normalise_path(const std::string& input_name);   // Fileparameters.f90:46-54: backslashes of the .in files

}  // namespace en234
