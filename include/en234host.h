/* en234host.h — C ABI of the host-side driver that sits above en234gpu.h: the `.in` parser
 * (reference: src/read_input_file.f90 + src/parse.f90), the MESH/USER SUBROUTINE hook
 * (user_codes/src/user_mesh.f90), boundary-condition evaluation (src/solver_direct.f90:501-734,798-912) and the
 * static step loop (src/staticstep.f90).  Written in C++ because no Fortran compiler exists in this image; the
 * Fortran binding of the same device entry points is shipped as source in fortran/ (see INTEGRATION.md).
 * Every function returns 0 on success unless stated otherwise. */
#ifndef EN234HOST_H
#define EN234HOST_H

#include "en234gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct en234host_model* en234host_model_t;

/* parse an input file / literal text; on failure returns NULL and en234host_last_error() explains */
en234host_model_t en234host_model_read(const char* path);
en234host_model_t en234host_model_from_text(const char* text);
void en234host_model_free(en234host_model_t m);
const char* en234host_last_error(void);

/* Flat views of the model (names as in modules/Mesh.f90 / Boundaryconditions.f90 where they exist):
 * int arrays:  connectivity element_flag element_prop_index element_n_props element_n_states element_material
 *              material_prop_index material_n_props history_ptr nodeset_ptr
 *              nodeset_nodes elementset_ptr elementset_elements pdof_flag pdof_dof pdof_node pdof_nodeset pdof_history
 *              pdof_rate pdof_subparam subparam_ptr pforce_flag pforce_dof pforce_node pforce_nodeset pforce_history dload_flag dload_elset
 *              dload_face dload_history dload_val_ptr
 * int scalars (length 1): n_nodes n_elements max_no_steps solvertype nonlinear max_newton_iterations unsymmetric
 *              abaqusformat max_cg_iterations checkstiffness_flag
 * double arrays: coords element_properties material_properties history_time history_value pdof_value pforce_value dload_values
 *              subparam_values
 * double scalars: timestep_initial timestep_min timestep_max max_total_time newtonraphson_tolerance cg_tolerance
 * The pointers stay valid until the model is freed or modified. */
int en234host_model_int(en234host_model_t m, const char* name, const int** p, long* n);
int en234host_model_double(en234host_model_t m, const char* name, const double** p, long* n);
/* names of node/element sets: returns the 1-based index or 0 */
int en234host_model_find_set(en234host_model_t m, int element_set, const char* name);
/* run-time overrides: cg_tolerance max_cg_iterations solvertype nonlinear newtonraphson_tolerance
 * max_newton_iterations max_no_steps timestep_initial timestep_min timestep_max max_total_time */
int en234host_model_set(en234host_model_t m, const char* name, double value);
/* literal `.in` text of the model; returns the number of bytes needed (call with cap = 0 to size the buffer) */
long en234host_model_write_text(en234host_model_t m, char* buf, long cap);

/* boundary-condition vectors at TIME, DTIME (see en234gpu_assemble / en234gpu_apply_boundaryconditions) */
int en234host_evaluate_loads(en234host_model_t m, double time, double dtime, const double* dof_total,
                             const double* dof_increment, double* f_element_ext, double* f_ext, int* n_fixed,
                             int* fixed_dof /*cap = sum of set sizes*/, double* fixed_value, double* fixed_correction,
                             int fixed_cap);

/* device handle for the whole mesh of a model */
int en234host_gpu_create(en234host_model_t m, int device, en234gpu_handle* h);

/* compute_static_step (src/staticstep.f90:1-162) on the device path.
 *   newton[8*max_rows]: step, iteration, correction, force, unbalanced, ratio, tolerance, converged per row
 *   info[8]: n_steps_completed, n_cutbacks, n_newton_rows, n_cg_solves, n_state_prints, 0...
 *   log_path: optional file receiving the reference-format convergence lines (NULL = none) */
int en234host_run_static(en234host_model_t m, en234gpu_handle h, int method, double cg_tolerance,
                         int max_cg_iterations, int check_linear_cg, int use_host_api, double* dof_total,
                         double* rforce, double* newton, int max_rows, int* cg_iterations, int max_cg_rows,
                         long long* info, const char* log_path);

/* Element partition of a model for multi-GPU runs (1-D slabs of whole elements along z for structured blocks,
 * contiguous element ranges otherwise).  Produces the rank-local mesh and the halo description consumed by
 * en234gpu_set_partition().  Query sizes first with the *_sizes call. */
typedef struct en234host_part* en234host_part_t;
en234host_part_t en234host_partition(en234host_model_t m, int n_ranks, int rank);
void en234host_part_free(en234host_part_t p);
/* int arrays: local_to_global_node (0-based global) local_elements (0-based global) connectivity (1-based LOCAL)
 *             neighbour_rank shared_ptr shared_nodes ; unsigned char array `owned` via en234host_part_owned
 * int scalars: n_nodes n_elements n_neighbours */
int en234host_part_int(en234host_part_t p, const char* name, const int** ptr, long* n);
const unsigned char* en234host_part_owned(en234host_part_t p);
/* device handle for one partition (local mesh + halo already registered; call en234gpu_comm_init next) */
int en234host_part_gpu_create(en234host_model_t m, en234host_part_t p, int device, en234gpu_handle* h);
/* en234host_run_static on one rank of a partitioned run (all ranks call it collectively): dof_total / rforce are the
 * rank-local arrays (3 * part n_nodes, local node order); norms, Newton records and CG counts are global */
int en234host_run_static_part(en234host_model_t m, en234host_part_t p, en234gpu_handle h, int method,
                              double cg_tolerance, int max_cg_iterations, int check_linear_cg, int use_host_api,
                              double* dof_total_local, double* rforce_local, double* newton, int max_rows,
                              int* cg_iterations, int max_cg_rows, long long* info, const char* log_path);

/* ---- state print: the step after the hot path (print_state, src/printing.f90:249-1093; 3-D hex8 zones).
 * The PRINT STATE block of the static step is parsed (read_input_file.f90:1772-1866); en234host_run_static writes a
 * Tecplot FEPOINT zone at the first, every STATE PRINT STEPS-th and the last step (staticstep.f90:227-231) once a path
 * has been set with en234host_model_set_state_print (library default: off, so that running an input never writes files
 * the caller did not ask for; the en234_run driver turns it on with the input's own file name).
 * flags: bit 0 PRINT STATE present / enabled, bit 1 DEGREES OF FREEDOM, bit 2 DISPLACED MESH, bit 3 lumped projection
 * matrix (printing.f90:464-471; the static step default is the consistent matrix, read_input_file.f90:144). */
int en234host_model_get_state_print(en234host_model_t m, int* flags, int* state_print_steps, double* scale_factor,
                                    char* field_names_csv, int cap_names, char* file_name, int cap_file);
int en234host_model_set_state_print(en234host_model_t m, int flags, int state_print_steps, double scale_factor,
                                    const char* field_names_csv /*NULL: keep*/, const char* path /*NULL: off*/);
/* Nodal projection of the named integration-point fields on the device (en234gpu_project_fields): S11 S22 S33 S12 S13 S23
 * SMISES for elements 1001 / U1 (el_linelast_3Dbasic.f90:392-408, abaqus_tecplot_interface.f90:163-179), S.., E11..E23,
 * U<n> for C3D8 (continuum_elements.f90:1291-1345); names are matched by prefix like the reference's strcmp, unknown
 * names give zero fields.  fields[k * n_nodes + n].  dof arrays: host, or NULL for the device-resident state. */
int en234host_project_fields(en234host_model_t m, en234gpu_handle h, const char* field_names_csv, int lumped,
                             const double* dof_total, const double* dof_increment, double* fields);
/* One zone of the Tecplot file for the given state: VARIABLES line, ZONE header (A10,D12.4,A15,I5,A3,I5,A9), one
 * (n(1X,E18.10)) line per node in order of first appearance in the element list, (8(1x,i5)) connectivity.  Node / element
 * counts above 99999 are written in full (the reference's I5 prints asterisks). */
int en234host_print_state(en234host_model_t m, en234gpu_handle h, double time_end, const double* dof_total,
                          const double* dof_increment /*NULL: zero*/, const char* path, int append);

/* explicit_dynamic_step (src/explicit_dynamic_step.f90:1-83) for an input with an EXPLICIT DYNAMIC STEP block (TIME STEP,
 * NUMBER OF STEPS, STATE PRINT STEPS, PRINT STATE; read_input_file.f90:1920-2103): builds the load tables
 * (apply_dynamic_boundaryconditions), runs no_dynamic_steps steps on the device and writes a state print every
 * STATE PRINT STEPS-th step at the point of the step where the reference prints (after the boundary conditions, before the
 * element forces) when a path has been set with en234host_model_set_state_print.  n_steps <= 0: the input's
 * NUMBER OF STEPS.  initial_velocity may be NULL (the reference starts from velocity = dof_increment = 0).  Outputs may be
 * NULL.  info[8]: steps done, state prints, kernels launched, 0...; device_ms: CUDA-event time of the steps. */
int en234host_run_dynamic(en234host_model_t m, en234gpu_handle h, int n_steps, const double* initial_velocity,
                          double* dof_total, double* velocity, double* acceleration, double* rforce, double* lumped_mass,
                          long long* info, double* device_ms);

/* one rank of a partitioned explicit dynamic run (collective; one halo sum of the element forces per step); arrays are
 * rank-local, 3 * (part n_nodes); no state prints */
int en234host_run_dynamic_part(en234host_model_t m, en234host_part_t p, en234gpu_handle h, int n_steps,
                               const double* initial_velocity_local, double* dof_total_local, double* velocity_local,
                               double* acceleration_local, double* rforce_local, double* lumped_mass_local, long long* info,
                               double* device_ms);

/* inspection of the load tables en234host_run_dynamic hands to en234gpu_dynamic_setup (tests): load `load_index` as a dense
 * vector of nodal forces for a unit magnitude (distributed loads first, then prescribed forces) with its history id /
 * constant scale, the prescribed-DOF list, the element densities; p == NULL: whole mesh, else rank-local numbering.
 * Returns the number of loads, -1 on error. */
int en234host_dynamic_tables(en234host_model_t m, en234host_part_t p, int load_index, double* dense, int* history,
                             double* constant, int* n_pdof, int* pdof_dof, int* pdof_history, double* pdof_value, int* pdof_mode,
                             int cap_pdof, double* element_density);

/* 0 when every element of the model is one the device implements; else 1 with the reason in en234host_last_error — e.g. a
 * "U1" element whose property count fits neither the shipped linear-elastic UEL (E, nu) nor the VUEL (E, nu, density): an
 * input file names element routines by id only, so a model written for another UEL source would otherwise run with the
 * wrong physics.  en234host_gpu_create / en234host_part_gpu_create apply the same test. */
int en234host_model_device_support(en234host_model_t m);

/* check_stiffness (src/checkstiffness.f90): the reference's self-test of an element routine, run through the device twins
 * below — coded stiffness against the numerical derivative of the residual (h = 1e-7), err = sum (K_num - K)^2 / sum K^2
 * (consistent when <= 1e-6), plus the counts of entries the symmetry report flags.  element_flag: as parsed from the
 * CHECK STIFFNESS key (model int scalar checkstiffness_flag), dof arrays may be NULL (zero), log_path receives the
 * reference-format report (NULL = none). */
int en234host_check_stiffness(en234host_model_t m, int element_flag, const double* dof_total, const double* dof_increment,
                              const char* log_path, double* err, int* n_unsymmetric, int* n_numerical_unsymmetric);

/* ---- the reference's element plugin entry points with their own argument lists and gfortran linkage (all arguments
 * by reference; default integer / logical = int32): user_codes/src/user_element.f90:5-49 and
 * user_codes/src/abaqus_uel_linelastic_3d.for:16-28.  A Fortran driver links these in place of the CPU routines;
 * each call evaluates one hex8 element on device 0 (diagnostics / CHECK STIFFNESS use; the timed path evaluates
 * elements inside the assembly kernel).  element ids 1001, 1002 / JTYPE 1, 2 only: anything else sets fail (PNEWDT
 * = 0.5 for UEL) instead of stopping the process. */
void user_element_static_(const int* lmn, const int* element_identifier, const int* n_nodes,
                          const int* node_property_list, const int* n_properties, const double* element_properties,
                          const int* n_int_properties, const int* int_element_properties,
                          const double* element_coords, const int* length_coord_array, const double* dof_increment,
                          const double* dof_total, const int* length_dof_array, const int* n_state_variables,
                          const double* initial_state_variables, double* updated_state_variables,
                          double* element_stiffness, double* element_residual, int* fail);
void uel_(double* RHS, double* AMATRX, double* SVARS, double* ENERGY, const int* NDOFEL, const int* NRHS,
          const int* NSVARS, const double* PROPS, const int* NPROPS, const double* COORDS, const int* MCRD,
          const int* NNODE, const double* U, const double* DU, const double* V, const double* A, const int* JTYPE,
          const double* TIME, const double* DTIME, const int* KSTEP, const int* KINC, const int* JELEM,
          const double* PARAMS, const int* NDLOAD, const int* JDLTYP, const double* ADLMAG, const double* PREDEF,
          const int* NPREDF, const int* LFLAGS, const int* MLVARX, const double* DDLMAG, const int* MDLOAD,
          double* PNEWDT, const int* JPROPS, const int* NJPROP, const double* PERIOD);

#ifdef __cplusplus
}
#endif
#endif
