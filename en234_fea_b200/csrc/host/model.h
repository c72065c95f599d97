// Host-side data model of the product: the subset of the reference's module-global arrays that the
// static-step hot path reads (modules/Mesh.f90:51-107, modules/Boundaryconditions.f90:4-76,
// modules/Staticstepparameters.f90:5-18), as plain vectors.  Filled by read_input_file() or user_mesh().
#pragma once
#include <string>
#include <vector>

namespace en234 {

struct History { std::string name; std::vector<double> t, v; };
struct NamedSet { std::string name; std::vector<int> items; };   // 1-based node / element numbers

struct PrescribedDof {        // prescribeddof_list / prescribedforce_list (Boundaryconditions.f90:34-55)
    int flag = 1;             // 1 VALUE, 2 HISTORY, 3 SUBROUTINE (user_prescribeddof)
    int subroutine_parameter_number = 0;   // 1-based SUBROUTINE PARAMETERS block (flag 3)
    int dof = 1;              // 1-based
    int node_number = 0;      // single node, or
    int node_set = 0;         // 1-based node set (0 = none)
    int history_number = 0;   // 1-based
    double value = 0.0;
    int rate_flag = 0;
};

struct SubroutineParameters { std::string name; std::vector<double> values; };   // SUBROUTINE PARAMETERS block

struct DistributedLoad {      // distributedload_list (Boundaryconditions.f90:57-66)
    int flag = 0;             // native: 1 VALUE, 2 HISTORY, 3 NORMAL, 4 SUBROUTINE; ABAQUS UEL: face number n of "Un"
    int element_set = 0;      // 1-based
    int face = 0;             // native loads only
    int history_number = 0;
    std::vector<double> values;
    bool uel_format = false;
};

struct Constraint {           // constraint_list (Boundaryconditions.f90; read_input_file.f90:1429-1541)
    int flag = 1;             // 1 CONSTRAIN NODE PAIRS, 2 TIE NODES (two-node ties); 3 CONSTRAIN NODE SET (user_constraint code)
    int node1 = 0, dof1 = 0, node2 = 0, dof2 = 0;   // 1-based; flag 3: node1 / node2 are the node set and the DOF-list set
    int index_parameters = 0; // 1-based CONSTRAINT PARAMETERS block (0 = none)
};

struct ElementGroup {         // one PARAMETERS/PROPERTIES/CONNECTIVITY block of ELEMENTS, USER
    int flag = 0, n_nodes = 8, n_states = 0;
    int prop_index = 0, n_props = 0;
    int first_element = 0, n_elements = 0;
    std::string zone;
};

struct Model {
    // mesh
    int n_nodes = 0, n_elements = 0;
    int n_coords = 3, n_dof = 3;
    std::vector<double> coords;            // 3 per node
    int nodes_per_element = 8;             // 8 (hex8), or 4 / 10 / 20: meshes of tet4 / tet10 / hex20 elements of type 1001
    std::vector<int> connectivity;         // nodes_per_element per element, 1-based
    std::vector<int> element_flag, element_prop_index, element_n_props, element_n_states;
    std::vector<double> element_properties;
    std::vector<ElementGroup> groups;
    std::vector<double> element_density;   // DENSITY of the ELEMENTS block in force when the element was read (0: none)
    // MATERIAL blocks + ELEMENTS, INTERNAL (C3D8 + UMAT, element flag 10003): parsed so that the reference's golden
    // input can be read, but not part of the GPU hot path — en234gpu_create rejects flag 10003.
    struct MaterialDef { std::string name; int n_states = 0, prop_index = 0, n_props = 0; };
    std::vector<MaterialDef> materials;
    std::vector<double> material_properties;
    std::vector<int> element_material;     // 1-based material per element (0 for user elements)
    bool abaqusformat = false;             // set when any element id is "Un"/"VUn" (read_input_file.f90:546-555)
    // boundary conditions
    std::vector<History> histories;
    std::vector<NamedSet> nodesets, elementsets;
    std::vector<PrescribedDof> prescribed_dofs, prescribed_forces;
    std::vector<SubroutineParameters> subroutine_parameters;
    std::vector<Constraint> constraints;                        // CONSTRAINTS block
    std::vector<SubroutineParameters> constraint_parameters;    // CONSTRAINT PARAMETERS blocks
    // current_step_number of module Staticstepparameters: the reference's user hooks read it as a global
    // (user_codes/src/user_prescribeddof.f90:7); set by compute_static_step before the loads are evaluated
    mutable int current_step_number = 1;
    std::vector<DistributedLoad> distributed_loads;
    // static step (defaults read_input_file.f90:120-129, :1735-1751)
    bool staticstep = false;
    double timestep_initial = 0, timestep_min = 0, timestep_max = 0, max_total_time = 0;
    int max_no_steps = 1;
    int solvertype = 1;                    // 1 direct, 2 conjugate gradient
    bool nonlinear = true;
    double newtonraphson_tolerance = 1e-8;
    int max_newton_iterations = 1;
    bool unsymmetric_stiffness = false;
    int checkstiffness_flag = 0;           // element flag named by CHECK STIFFNESS (0 = none)
    // PRINT STATE block of the static step (read_input_file.f90:1772-1866; defaults :136-149, modules/Printparameters.f90)
    bool stateprint = false;
    std::string state_print_filename;      // as written in the input (backslashes kept; normalised when opened)
    int state_print_steps = 1;             // STATE PRINT STEPS (:1715-1726)
    bool print_dof = false, print_displacedmesh = false, use_lumped_projection_matrix = false;
    double displacementscalefactor = 1.0;
    std::vector<std::string> field_variable_names;
    std::string state_print_path;          // run-time: where run_static writes the state prints ("" = state prints off)
    // EXPLICIT DYNAMIC STEP (read_input_file.f90:1920-1966, modules/Dynamicstepparameters.f90)
    bool explicitdynamicstep = false;
    double dynamic_timestep = 0.0;
    int no_dynamic_steps = 0;
    // run-time solver options that are compile-time parameters in modules/Stiffness.f90:35-37
    double cg_tolerance = 1e-17;
    int max_cg_iterations = 1000;
    std::string error;

    int ndof() const { return 3 * n_nodes; }
};

// src/read_input_file.f90 + src/parse.f90 — subset listed in SURVEY.md section 5
bool read_input_file(const std::string& path, Model& m);
bool read_input_text(const std::string& text, Model& m);
// user_codes/src/user_mesh.f90 hook (MESH / USER SUBROUTINE): synthetic structured hex8 block
bool user_mesh(const std::vector<double>& params, Model& m);
// writes the same mesh/BCs as literal .in text (parser round-trip tests, n <= 16)
std::string write_input_text(const Model& m);

double interpolate_history_table(const History& h, double time);   // Boundaryconditions.f90:161-221

}  // namespace en234
