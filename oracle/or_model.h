// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product (en234_fea_b200/) may include, link or
// call anything in this directory; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// CPU restatement (C++17, IEEE double, built with -O2 -ffp-contract=off) of the EN234_FEA hot path:
// hex8 element stiffness/residual -> global assembly -> linear solve inside each static Newton step.
// All citations are file:line relative to /root/reference/EN234_FEA/.
//
// PARITY PINNING: the reference is serial Fortran and no Fortran compiler exists in this image, so it
// cannot be executed here.  The oracle is pinned by (1) the analytic solution of
// input_files/Abaqus_uel_linear_elastic_3d.in, (2) the reference's own CHECK STIFFNESS criterion
// (src/checkstiffness.f90:481-532), (3) direct-vs-CG agreement on input_files/Abaqus_uel_holeplate_3d.in,
// (4) the golden log Output_Files/Abaqus_umat_holeplate_3d.out (C3D8 B-bar + UMAT; pins the
// driver/assembly/BC/skyline-LU/Newton logic and the GPS profile statistics).
// Absolute hex8 K_e values of the U1/1001 elements have no golden vector in the reference:
// for those "parity unpinned" by reference outputs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

// Data model: mirrors the flat 1-based arrays of modules/Mesh.f90:51-107 and
// modules/Boundaryconditions.f90:4-76, restricted to 3-coordinate / 3-DOF nodes and 8-node elements.
struct Model {
    int n_nodes = 0, n_elements = 0;
    int nodes_per_element = 8;              // 8 (hex8: every element type of the path), or 4 / 10 / 20 for meshes of
                                            // tet4 / tet10 / hex20 elements of type 1001 (el_linelast_3Dbasic.f90:82-85)
    std::vector<double> coords;             // 3*n_nodes, xyz per node            (Mesh.f90:84)
    std::vector<int> connectivity;          // nodes_per_element*n_elements, 1-based node numbers (Mesh.f90:69)
    std::vector<int> element_flag;          // element_list%flag                  (Mesh.f90:16-29)
    std::vector<int> element_prop_index;    // 0-based offset into element_properties
    std::vector<int> element_n_props;
    std::vector<double> element_properties;
    std::vector<int> element_n_states;      // svars per element (48 for the shipped inputs)
    std::vector<int> element_material;      // 1-based material number for flag 10003 (C3D8), else 0
    std::vector<int> material_prop_index, material_n_props;
    std::vector<double> material_properties;

    std::vector<double> element_density;    // densities(element%density_index), 0 if none (explicit dynamics, native elements)
    int explicitdynamicstep = 0, no_dynamic_steps = 0;   // modules/Dynamicstepparameters.f90
    double dynamic_timestep = 0.0;

    // Boundary conditions (Boundaryconditions.f90:34-66)
    std::vector<int> history_ptr;           // nh+1
    std::vector<double> history_time, history_value;
    std::vector<int> nodeset_ptr, nodeset_nodes;          // 1-based node numbers
    std::vector<int> elementset_ptr, elementset_elements; // 1-based element numbers
    std::vector<int> pdof_flag, pdof_dof, pdof_node, pdof_nodeset, pdof_history, pdof_rate;
    std::vector<double> pdof_value;
    // SUBROUTINE-type prescribed DOFs (flag 3): parameter block per DOF entry (1-based, 0 = none), CSR of the blocks
    std::vector<int> pdof_subparam, subparam_ptr;
    std::vector<double> subparam_values;
    mutable int current_step_number = 1;   // Staticstepparameters%current_step_number, read by user_prescribeddof
    std::vector<int> pforce_flag, pforce_dof, pforce_node, pforce_nodeset, pforce_history;
    std::vector<double> pforce_value;
    std::vector<int> dload_flag, dload_elset, dload_face, dload_history, dload_val_ptr;
    std::vector<double> dload_values;
    // Two-node tie constraints (constraint_list with flag 1 / 2, Boundaryconditions.f90; CONSTRAIN NODE PAIRS / TIE NODES of the
    // CONSTRAINTS block, read_input_file.f90:1429-1498): dof1 of node1 equals dof2 of node2, enforced by one Lagrange
    // multiplier each (solver_direct.f90:294-328).  1-based nodes and DOFs.  General constraints (flag 3) call user code
    // (user_constraint.f90) and are not restated.
    std::vector<int> con_node1, con_dof1, con_node2, con_dof2;
    int n_constraints() const { return (int)con_node1.size(); }

    // Static step parameters (Staticstepparameters.f90:5-18, defaults read_input_file.f90:120-129)
    double timestep_initial = 0, timestep_min = 0, timestep_max = 0, max_total_time = 0;
    double newtonraphson_tolerance = 1e-8;
    int max_no_steps = 1, solvertype = 1, nonlinear = 1, max_newton_iterations = 1, unsymmetric = 0;
    int abaqusformat = 0;
    // CG run-time options (compile-time parameters in modules/Stiffness.f90:35-37)
    double cg_tolerance = 1e-17;
    int max_cg_iterations = 1000;
    // Oracle options (deviations from the literal reference are opt-in and documented):
    int gps_renumber = 0;            // reserved (Gibbs-Poole-Stockmeyer order, staticstep.f90:38, is not restated)
    std::vector<int> node_order;     // optional bandwidth-reducing node order for the skyline (0-based nodes, position ->
                                     // node); the tests pass a reverse Cuthill-McKee order from scipy.  Empty = identity.
    int cg_apply_nodal_forces = 1;   // 0: literal CG BC routine, which drops FORCES
    int cg_check_linear = 0;         // 1: run the re-assemble + convergencecheck loop for LINEAR+CG too (Q6)

    int ndof() const { return 3 * n_nodes; }
};

// One line of the Newton log (user_codes/src/convergencecheck.f90:10-18)
struct NewtonRecord {
    int step, iteration;
    double correction_norm, nodal_force_norm, unbalanced_force_norm, ratio, tolerance;
    int converged;
};

struct StepLog {
    std::vector<NewtonRecord> newton;
    std::vector<int> cg_iterations;        // one entry per CG solve
    std::vector<double> step_time, step_dtime;
    int n_steps_completed = 0;
    int n_cutbacks = 0;
    // direct-solver diagnostics (solver_direct.f90:1381-1386, 1154-1159)
    long long profile_neq = 0, profile_bwmax = 0, profile_bwmean = 0, profile_bwrms = 0, profile_size = 0;
    std::vector<double> lu_dimx, lu_dimn;
    std::vector<int> lu_ifig;
    std::vector<double> lagrange_multipliers;   // final values (two-node tie constraints)
};

// ---- elements (or_elements.cpp) -------------------------------------------------------------
// K is column-major 24x24 (K[row + 24*col]), matching element_stiffness(ROW,COL).
void el_linelast_3dbasic(const double* coords24, const double* du24, const double* utot24,
                         const double* props, double* K, double* R);
// the same routine for n_nodes = 4, 10, 8, 20 with the reference's integration rules and shape functions
// (Element_Utilities.f90:586-689, :863-1025) as written there; K column-major (3 n_nodes)^2.  Returns 1 when det == 0.
int el_linelast_3d_general(int n_nodes, const double* coords, const double* du, const double* utot, const double* props,
                           double* K, double* R);
int integration_points_3d(int n_nodes, double xi[][3], double* w);                       // number of points
void shapefunctions_3d(int n_nodes, const double* xi, double* f, double dfdxi[][3]);
void uel_linelastic_3d(const double* coords24, const double* U24, const double* props, int nsvars,
                       int ndload, const int* jdltyp, const double* adlmag, double* K, double* R,
                       double* svars, double* energy8, double* pnewdt);
void el_neohooke_3d(const double* coords24, const double* U24, const double* props, double* K,
                    double* R, double* svars, int* fail);
void c3d8_bbar_static(const double* coords24, const double* du24, const double* utot24,
                      const double* matprops, int nprops, double dtime, const double* svars_in, double* K, double* R,
                      double* svars_out, int nsv);
// Dispatch on element flag exactly as solver_direct.f90:129-258 does.  Returns fail.
int element_static(const Model& m, int lmn /*0-based*/, const double* dof_total,
                   const double* dof_increment, double time, double dtime, double* K, double* R,
                   double* svars /*n_states or null*/);
void traction_boundarycondition_static_3d(int flag, const double* face_coords12,
                                          const double* traction, int ntract, double* R12);
void traction_boundarycondition_dynamic_3d(int flag, const double* face_coords12, const double* traction3, int ntract, double* R12);
void vuel_pressure_3d(const double* coords24, int face, double adlmag, double* R24);
void facenodes_hex8(int face, int list[4]);
double interpolate_history_table(const double* t, const double* v, int n, double time);
// Element ids understood by the oracle
enum { ID_LINELAST = 1001, ID_NEOHOOKE = 1002, FLAG_UEL_BASE = 99999, FLAG_C3D8 = 10003 };

// ---- direct (skyline) solver (or_direct.cpp) ------------------------------------------------
struct Skyline {
    int neq = 0;
    std::vector<int> ieqs;       // dof (0-based) -> equation (0-based)
    std::vector<long long> jp;   // jp[j] = number of stored entries in columns 0..j (cumulative heights)
    std::vector<double> aupp, diag, rhs;
    std::vector<double> alow;    // lower triangle stored row-wise (unsymmetric_stiffness only, Stiffness.f90:17-28)
    bool unsym = false;
    std::vector<double> lagrange_multipliers;   // one per constraint (module Boundaryconditions), equation ieqs[ndof + nc]
};
void compute_profile(const Model& m, const std::vector<int>& node_order, Skyline& s, StepLog* log);
void gps_node_order(const Model& m, std::vector<int>& node_order);   // bandwidth_minimizer.f90
struct StepLog;
void projection_mass_lu(const Model& m, const std::vector<int>& node_order, StepLog* log);   // printing.f90:1214-1383, :475-488
void projection_mass_factor(const Model& m, const std::vector<int>& node_order, Skyline& s, StepLog* log);
void skyline_backsubstitution(Skyline& s);   // solver_direct.f90:976-1032 on s.rhs
struct AsmOut {
    std::vector<double> rforce;
    double nodal_force_norm = 0, unbalanced_force_norm = 0, correction_norm = 0;
};
int assemble_direct_stiffness(const Model& m, Skyline& s, const double* dof_total,
                              const double* dof_increment, double time, double dtime,
                              std::vector<double>* svars, AsmOut& out);
void apply_direct_boundaryconditions(const Model& m, Skyline& s, const double* dof_total,
                                     const double* dof_increment, double time, double dtime, AsmOut& out);
void solve_direct(const Model& m, Skyline& s, double* dof_increment, AsmOut& out, StepLog* log);

// ---- conjugate gradient solver (or_cg.cpp) --------------------------------------------------
struct CooUpper {
    int neq = 0;
    std::vector<int> rows, cols;           // strictly-upper entries, first-touch order
    std::vector<double> aupp, diag, rhs, sol;
    std::vector<std::vector<std::pair<int, int>>> rowbins;  // per row: (col, index)
};
void initialize_cg_solver(const Model& m, CooUpper& c);
int assemble_conjugate_gradient_stiffness(const Model& m, CooUpper& c, const double* dof_total,
                                          const double* dof_increment, double time, double dtime,
                                          std::vector<double>* svars, AsmOut& out);
void apply_conjugategradient_boundaryconditions(const Model& m, CooUpper& c, const double* dof_total,
                                                const double* dof_increment, double time, double dtime,
                                                AsmOut& out, bool apply_nodal_forces);
int cg_iteration_loop(CooUpper& c, double tol, int maxit, int* iters, int fixed_iters = -1);
int solve_conjugategradient(const Model& m, CooUpper& c, double* dof_increment, AsmOut& out, int* iters);

// ---- static step driver (or_step.cpp) -------------------------------------------------------
int compute_static_step(const Model& m, std::vector<double>& dof_total, std::vector<double>& rforce,
                        std::vector<double>& svars, StepLog& log);


// ---- explicit dynamics (or_dynamic.cpp): src/explicit_dynamic_step.f90 -------------------------------------------
struct DynamicState {
    std::vector<double> dof_total, dof_increment, velocity, acceleration, rforce, lumped_mass;
    double time = 0.0;
    int steps_done = 0;
};
void assemble_lumped_mass(const Model& m, std::vector<double>& lumped_mass);                         // :85-392
// runs steps [s.steps_done, s.steps_done + n_steps); the first call initialises velocity = dof_increment, acceleration = 0
int explicit_dynamic_steps(const Model& m, DynamicState& s, int n_steps);

// ---- state print (or_print.cpp): nodal projection of Gauss-point fields + Tecplot file, src/printing.f90:249-1093 ----
struct PrintOptions {
    std::vector<std::string> field_names;   // FIELD VARIABLES of the PRINT STATE block (read_input_file.f90:1822-1834)
    bool print_dof = false, displaced_mesh = false, lumped = false;
    double displacement_scale_factor = 1.0;
};
// field[k + nf*n]: int_V f_k N^n dV summed over the elements in element order (printing.f90:1386-1668)
int assemble_field_projection(const Model& m, const double* dof_total, const double* dof_increment, const double* svars,
                              const std::vector<std::string>& names, std::vector<double>& field);
void lumped_projection_mass(const Model& m, std::vector<double>& lump);                       // printing.f90:1096-1212
int project_fields(const Model& m, const double* dof_total, const double* dof_increment, const double* svars,
                   const std::vector<std::string>& names, bool lumped, std::vector<double>& field);
int print_state(const Model& m, const PrintOptions& o, double time_end, const double* dof_total,
                const double* dof_increment, const double* svars, std::string& out);

}  // namespace oracle
