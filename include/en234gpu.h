/* en234gpu.h — C ABI of the B200 (sm_100a) implementation of the EN234_FEA static-step hot path.
 *
 * Drop-in boundary: every entry point replaces one argument-less module-global subroutine call of the
 * reference's static step driver (citations relative to /root/reference/EN234_FEA/).  All arguments are
 * plain pointers and sizes in the reference's own array layouts (modules/Mesh.f90:51-107): a Fortran
 * driver binds them with ISO_C_BINDING (fortran/en234_gpu_iface.f90).
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via
 * en234gpu_last_error()); the library never calls exit().  Recoverable numerical failures (NaN in
 * the residual, non-positive Jacobian, CG breakdown / non-convergence) are reported through *fail = 1 so
 * the driver can cut the time step back exactly as src/staticstep.f90:46-54,109-117 does.
 * Host arrays stay caller-owned and are only touched during a call.  One host thread per handle.
 * There is no CPU fallback: without a CUDA device en234gpu_create() fails.
 */
#ifndef EN234GPU_H
#define EN234GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct en234gpu_ctx* en234gpu_handle;

/* Element identifiers understood on the device (element_list%flag, modules/Mesh.f90:16-29):
 *   1001            native linear elastic hex8   (user_codes/src/el_linelast_3Dbasic.f90, user_element.f90:62)
 *   99999 + 1 (U1)  ABAQUS UEL linear elastic hex8 (user_codes/src/abaqus_uel_linelastic_3d.for; flag = n+99999,
 *                   src/read_input_file.f90:546-550)
 *   1002 / U2       compressible neo-Hookean hex8 (new; props mu, K as input_files/Abaqus_uel_hyperelastic.in:44) */
#define EN234_ID_LINELAST 1001
#define EN234_ID_NEOHOOKE 1002
#define EN234_FLAG_UEL_BASE 99999
/*   10003           built-in C3D8 continuum element (finite-strain B-bar hex8, src/continuum_elements.f90:368-761) with
 *                   the linear-elastic ABAQUS UMAT (user_codes/src/abaqus_umat_elastic.for); props E, nu of its
 *                   MATERIAL; 15 state variables per integration point; UNSYMMETRIC tangent -> EN234_SOLVE_BICGSTAB_BSR.
 *                   A mesh is either all-C3D8 or all user elements. */
#define EN234_FLAG_C3D8 10003

/* Solver methods for en234gpu_solve */
#define EN234_SOLVE_PCG_BSR 0        /* Jacobi-PCG on the assembled 3x3-block CSR matrix               */
#define EN234_SOLVE_PCG_MATRIXFREE 1 /* Jacobi-PCG with the element-by-element (matrix-free) operator  */
#define EN234_SOLVE_BICGSTAB_BSR 2   /* Jacobi-BiCGSTAB on the assembled matrix (unsymmetric stiffness; 1 GPU) */

const char* en234gpu_last_error(void);
int en234gpu_device_count(void);

/* Replaces initialize_cg_solver (src/solver_conjugate_gradient.f90:1039-1244; called at src/staticstep.f90:95)
 * and, for the direct branch, generate_node_numbers/compute_profile/allocate_direct_stiffness
 * (src/staticstep.f90:38-40): builds the sparsity pattern, the element-to-CSR slot maps and all device
 * storage once per analysis.
 *   coords[3*n_nodes]          xyz per node           (Mesh.f90:84 `coords`)
 *   connectivity[8*n_elements] 1-based node numbers   (Mesh.f90:69 `connectivity`)
 *   element_flag[n_elements]   element_list%flag
 *   element_prop_index[n_elements] 0-based offset of the element's first property in element_properties
 *   element_properties[n_element_properties]          (Mesh.f90:76 `element_properties`)
 * Partitioned runs (one handle per GPU/process) pass the rank-local mesh and then call
 * en234gpu_set_partition() before the first assemble. */
int en234gpu_create(en234gpu_handle* h, int device, int n_nodes, int n_elements, const double* coords,
                    const int* connectivity, const int* element_flag, const int* element_prop_index,
                    const double* element_properties, int n_element_properties);
/* The same for meshes of 4-node / 10-node tetrahedra or 20-node bricks of element type 1001 (el_linelast_3Dbasic.f90:82-85:
 * 1, 4, 27 integration points; shape functions Element_Utilities.f90:863-1025 as written there); connectivity holds
 * nodes_per_element entries per element; nodes_per_element = 8 is en234gpu_create.  These meshes run the assembled-operator
 * path on one device (assemble / apply_boundaryconditions / solve, host-buffer and device-resident forms); the matrix-free
 * operators, state prints, explicit dynamics and partitions are hex8 paths. */
int en234gpu_create_nodes(en234gpu_handle* h, int device, int nodes_per_element, int n_nodes, int n_elements,
                          const double* coords, const int* connectivity, const int* element_flag,
                          const int* element_prop_index, const double* element_properties, int n_element_properties);
int en234gpu_destroy(en234gpu_handle h);

/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's
 * current stream, so that caller-side CUDA events bracket it. NULL = the handle's own stream. */
int en234gpu_set_stream(en234gpu_handle h, void* cuda_stream);

/* Replaces assemble_conjugate_gradient_stiffness (src/solver_conjugate_gradient.f90:1-414; call sites
 * src/staticstep.f90:98,122) / assemble_direct_stiffness (src/solver_direct.f90:5-439; :44,:61):
 * element loop (gather :104-121, element routine :129-258) + deterministic assembly (:264-287 semantics,
 * contributions summed in ascending element order) + NaN scan (:402-415) + rforce = -rhs (:417-421) +
 * nodal_force_norm = ||rhs||_2 (:423).
 *   dof_total, dof_increment [3*n_nodes]   (Mesh.f90:85-86)
 *   f_element_ext [3*n_nodes] or NULL      load vectors that the reference's element routines add to their
 *                                          own residual (ABAQUS UEL DLOAD faces, abaqus_uel_linelastic_3d.for:207-238);
 *                                          evaluated on the host (surface terms) and added before rforce/norm.
 *   rforce [3*n_nodes] out or NULL         (Mesh.f90:91) */
int en234gpu_assemble(en234gpu_handle h, const double* dof_total, const double* dof_increment,
                      const double* f_element_ext, double* rforce, double* nodal_force_norm, int* fail);

/* Replaces apply_cojugategradient_boundaryconditions (src/solver_conjugate_gradient.f90:417-671; :107,:127) /
 * apply_direct_boundaryconditions (src/solver_direct.f90:444-738; :55,:65):
 *   f_ext [3*n_nodes] or NULL   native-element tractions (:501-568) + prescribed nodal forces (:646-682), host-evaluated
 *   fixed_dof[n_fixed]          0-based DOF indices (3*(node-1)+dof-1) of the prescribed DOFs, in application order
 *   fixed_correction[n_fixed]   value - (dof_total+dof_increment)   (:709,:729)
 * fixdof_* semantics (:742-792 / CG :673-713): rhs_j -= K_jn c_n, row and column n zeroed, diag_n = 1,
 * rhs_n = c_n; a DOF listed twice keeps its last correction.  unbalanced_force_norm = ||rhs||_2 (:736). */
int en234gpu_apply_boundaryconditions(en234gpu_handle h, const double* f_ext, int n_fixed, const int* fixed_dof,
                                      const double* fixed_correction, double* unbalanced_force_norm);

/* Replaces solve_conjugategradient + cg_iteration_loop (src/solver_conjugate_gradient.f90:717-912; :108,:130)
 * and, for large meshes, solve_direct (src/solver_direct.f90:916-973; :56,:68).
 * Jacobi-PCG with the reference's recurrences and stopping rule: x0 = 0, stop when
 * (r.r)/(b.b) < cg_tolerance (reference value 1e-17, modules/Stiffness.f90:35), *fail = 1 after
 * max_cg_iterations (reference 1000, Stiffness.f90:37); zero right-hand side returns x = 0 (:805-814).
 *   dof_increment [3*n_nodes] inout: += solution (:760-765);  correction_norm = ||solution||_2 (:771) */
int en234gpu_solve(en234gpu_handle h, int method, double cg_tolerance, int max_cg_iterations, double* dof_increment,
                   double* correction_norm, int* cg_iterations, int* fail);

/* Select the operator used by assemble / apply_boundaryconditions / solve: EN234_SOLVE_PCG_BSR (default) forms the
 * block-CSR matrix; EN234_SOLVE_PCG_MATRIXFREE never forms it (assemble computes residuals only, the BC routine
 * evaluates K[:,C]c and the Jacobi diagonal element by element).  Call before assemble. */
int en234gpu_set_operator_mode(en234gpu_handle h, int method);

/* ---- device-resident variants: same operations on the handle's device copies (no host<->device copies) */
int en234gpu_upload_state(en234gpu_handle h, const double* dof_total, const double* dof_increment,
                          const double* f_element_ext, const double* f_ext, int n_fixed, const int* fixed_dof,
                          const double* fixed_value /* prescribed VALUES (not corrections) */);
int en234gpu_assemble_dev(en234gpu_handle h, double* nodal_force_norm, int* fail);
/* corrections are formed on the device as fixed_value - (dof_total + dof_increment) */
int en234gpu_apply_boundaryconditions_dev(en234gpu_handle h, double* unbalanced_force_norm);
int en234gpu_solve_dev(en234gpu_handle h, int method, double cg_tolerance, int max_cg_iterations,
                       double* correction_norm, int* cg_iterations, int* fail);
int en234gpu_download_state(en234gpu_handle h, double* dof_increment, double* rforce);
int en234gpu_zero_increment_dev(en234gpu_handle h);

/* ---- element state variables (initial_state_variables / updated_state_variables, modules/Mesh.f90:92-93): owned by the
 * handle, zero at creation; every assembly recomputes the updated set from the initial one; the driver commits at the
 * end of a converged step (initial = updated, src/staticstep.f90:87).  n_state_variables = 120 per C3D8 element. */
long long en234gpu_n_state_variables(en234gpu_handle h);
int en234gpu_commit_state(en234gpu_handle h);
int en234gpu_get_state_variables(en234gpu_handle h, int updated /*0 initial, 1 updated*/, double* out);
int en234gpu_set_state_variables(en234gpu_handle h, const double* initial);

/* ---- element plugin surface.  Batched device evaluation of the element routines for a list of elements:
 * the device twin of user_element_static (user_codes/src/user_element.f90:5-49) / UEL
 * (user_codes/src/abaqus_uel_linelastic_3d.for:16-28).  element_stiffness is returned per element as the
 * reference's column-major (ROW,COL) 24x24 array, element_residual as 24 values.  Used by CHECK STIFFNESS
 * style tests; not part of the timed path. */
int en234gpu_element_batch(en234gpu_handle h, int n, const int* element_numbers /*1-based*/, const double* dof_total,
                           const double* dof_increment, double* element_stiffness /*n*576*/,
                           double* element_residual /*n*24*/, int* fail);
/* One element, host arrays in and out (24 coords xyz per node, 24+24 DOFs): what the Fortran-signature twins
 * user_element_static_ / uel_ (include/en234host.h) call.  svars48 / energy may be NULL. */
int en234gpu_element_single(int element_flag, const double* coords24, const double* props, int n_props,
                            const double* dof_total24, const double* dof_increment24, double* element_stiffness,
                            double* element_residual, double* svars48, double* energy, int* fail);
/* Gauss-point state: Cauchy stress s11,s22,s33,s12,s13,s23 at the 8 integration points of every element
 * (the 48 SVARS of abaqus_uel_linelastic_3d.for:193-195), for dof_total + dof_increment. */
int en234gpu_get_svars(en234gpu_handle h, const double* dof_total, const double* dof_increment,
                       double* svars /*48*n_elements*/);

/* Field projection of print_state (src/printing.f90:457-490; the step after the hot path): nodal least-squares fit of
 * integration-point fields.  Right-hand sides int f N dV per node in the reference's element order (assemble_field_projection,
 * :1386-1668), the consistent projection matrix int N N dV with the 27-point rule (:1214-1383) solved by Jacobi-PCG on the
 * device, three fields per pass (the reference LU-factorises a skyline copy: agreement to the solver tolerance), or the
 * row-sum lumped matrix (:1096-1212) when lumped != 0.
 * field_codes[k]: EN234_FIELD_S11..S23 (0..5), EN234_FIELD_SMISES (6; elements 1001 / U1), EN234_FIELD_E11..E23 (7..12;
 * C3D8), EN234_FIELD_STATE0 + i (state variable i of the point; C3D8); anything else projects to a zero field, as an
 * unknown name does in the reference.  dof_total / dof_increment: host arrays, or NULL to use the device-resident state.
 * fields: host, [k * n_nodes + n].  tolerance <= 0: 1e-28 on r.r / b.b; max_iterations <= 0: 500.
 * Overwrites the assembled stiffness values (a new en234gpu_assemble* must precede the next solve).
 * Partitioned contexts: collective over the ranks; arrays are rank-local; right-hand sides, matrix diagonal and lumped entries
 * of interface nodes are halo-summed, the solve is the partitioned PCG loop of the static step. */
#define EN234_FIELD_S11 0
#define EN234_FIELD_SMISES 6
#define EN234_FIELD_E11 7
#define EN234_FIELD_STATE0 100
#define EN234_FIELD_NONE (-1)
#define EN234_MAX_FIELDS 24
int en234gpu_project_fields(en234gpu_handle h, int n_fields, const int* field_codes, int lumped, double tolerance,
                            int max_iterations, const double* dof_total, const double* dof_increment, double* fields,
                            int* iterations);

/* ---- explicit dynamic step (src/explicit_dynamic_step.f90:1-83): central differences with the row-sum lumped mass matrix.
 * The element loop of assemble_dynamic_force (:394-718) is the residual half of the static element loop, so the device path
 * reuses the residual-only element kernel; time, displacement, velocity, acceleration stay on the device and a chunk of
 * steps runs without a host round trip (16 steps per CUDA graph).  Linear elastic hex8 elements (1001 with a DENSITY, or the
 * VUEL of user_codes/src/abaqus_vuel.for whose third property is the density); one device.
 * en234gpu_dynamic_setup replaces assemble_lumped_mass (:85-392; computed once: it depends on geometry and density only)
 * and takes the time-independent part of apply_dynamic_boundaryconditions (:720-1128):
 *   element_density[n_elements];
 *   history tables as CSR (history_ptr[n_histories + 1], time / value pairs; Boundaryconditions.f90:4-20), 1-based ids;
 *   loads: CSR of (0-based DOF, nodal force of a unit magnitude) with DISTINCT DOFs inside one load, applied in order as
 *          f[dof] += scale * value with scale = history(TIME + DTIME) when load_history[l] > 0, else load_constant[l];
 *   prescribed DOFs (distinct): value = history(TIME + DTIME) when pdof_history[k] > 0 else pdof_value[k];
 *          pdof_mode 0: dof_increment = value - dof_total (:1086, :1102), 1: dof_increment = DTIME * value (:1104),
 *          2: dof_increment = value (rate flag on a single node, :1080-1086).
 * The state starts at zero; en234gpu_dynamic_set_state overrides it (NULL = zeros; velocity is what the reference takes from
 * dof_increment at :30). */
int en234gpu_dynamic_setup(en234gpu_handle h, const double* element_density, int n_histories, const int* history_ptr,
                           const double* history_time, const double* history_value, int n_loads, const int* load_ptr,
                           const int* load_dof, const double* load_val, const int* load_history, const double* load_constant,
                           int n_pdof, const int* pdof_dof, const int* pdof_history, const double* pdof_value,
                           const int* pdof_mode);
int en234gpu_dynamic_set_state(en234gpu_handle h, const double* dof_total, const double* velocity, const double* acceleration,
                               double time);
/* n_steps whole steps (:36-70 without the prints); fail = 1: non-finite acceleration or non-positive Jacobian;
 * device_ms (may be NULL): CUDA-event time of the steps */
int en234gpu_dynamic_steps(en234gpu_handle h, double dt, int n_steps, int* fail, double* device_ms);
/* one step in two halves — phase 1: predictor + boundary conditions (:38-43), phase 2: element forces + update (:64-70) —
 * so that the state can be read where the reference prints it (:45-53) */
int en234gpu_dynamic_half_step(en234gpu_handle h, double dt, int phase, int* fail);
/* any pointer may be NULL; rforce is the last step's total nodal force (external + internal) */
int en234gpu_dynamic_get_state(en234gpu_handle h, double* dof_total, double* dof_increment, double* velocity,
                               double* acceleration, double* rforce, double* lumped_mass, double* time);

/* ---- inspection (parity tests, reports) */
long long en234gpu_nnz_blocks(en234gpu_handle h);
/* Scalar CSR view of the current system matrix: rowptr[3N+1] (64-bit), col[9*nnzb] (0-based), val[9*nnzb] */
int en234gpu_get_csr(en234gpu_handle h, long long* rowptr, int* col, double* val);
int en234gpu_get_rhs(en234gpu_handle h, double* rhs /*3N*/, double* diag /*3N or NULL*/);
/* y = A x with the assembled (method 0) or matrix-free (method 1) operator, for tests */
int en234gpu_apply_operator(en234gpu_handle h, int method, const double* x, double* y);

/* ---- two-node tie constraints (CONSTRAINTS block: CONSTRAIN NODE PAIRS / TIE NODES, constraint_list flag 1 / 2)
 * Replaces the Lagrange-multiplier rows the reference appends to the equation system (src/solver_direct.f90:292-328; same
 * block in src/solver_conjugate_gradient.f90:263-296) and their update after a solve (:967-969).  dof1 / dof2: 0-based DOF
 * numbers, constraint k demands u[dof1[k]] = u[dof2[k]].  After this call, assemble / apply_boundaryconditions / solve treat
 * the system exactly as the reference does — rows of tied DOFs carry -/+ lambda, rforce = -rhs, the three norms of the Newton
 * log include the multiplier rows and the multiplier corrections — but the solve eliminates every tied group onto one unknown
 * inside the PCG loop instead of factorising the indefinite saddle-point matrix (en234_fea_b200/csrc/en234gpu_ties.cuh).
 * One device, assembled operator (PCG, or BiCGSTAB for an unsymmetric stiffness: the same two wrapper kernels around its
 * operator); a group may contain one prescribed DOF (the others follow it).
 * Returns non-zero for loops of constraints or two prescribed DOFs tied together.  n_ties = 0 removes the constraints. */
int en234gpu_set_ties(en234gpu_handle h, int n_ties, const int* dof1, const int* dof2);
/* The elimination plan of a set of ties for a given prescribed-DOF list, without a device (host logic of en234gpu_set_ties;
 * used by the CPU tests): sizes[4] = groups, members, touched DOFs, constraints; group_ptr has groups + 1 entries; peel_tie /
 * peel_dof give, in recovery order, each constraint and the DOF whose equilibrium row yields its multiplier.  Output arrays
 * may be NULL.  Non-zero (en234gpu_last_error) for loops and for two prescribed DOFs in one group. */
int en234gpu_tie_plan(int n_ties, const int* dof1, const int* dof2, int n_fixed, const int* fixed_dof, int* sizes,
                      int* group_root, int* group_ptr, int* group_fixed, int* member_dof, int* peel_tie, int* peel_dof);
/* lagrange_multipliers = 0 (read_input_file.f90:139): call before a new analysis on a handle that has already run one */
int en234gpu_reset_lagrange_multipliers(en234gpu_handle h);
/* lagrange_multipliers(1:n_constraints) of module Boundaryconditions after the last solve */
int en234gpu_get_lagrange_multipliers(en234gpu_handle h, int n_ties, double* out);

/* ---- multi-GPU (one handle per process/GPU; element partitions with shared-node halo sums)
 * neighbours: for each neighbour rank its shared nodes as LOCAL 0-based node indices, listed in the same
 * (global-node-number) order on both sides.  owned[n_nodes]: 1 if this rank owns the node (exactly one
 * owner per global node) — dot products and norms count owned DOFs only. */
int en234gpu_get_unique_id(char* id128);
int en234gpu_comm_init(en234gpu_handle h, int n_ranks, int rank, const char* id128);
/* buf[n] (host) becomes its sum over all ranks of the communicator, on every rank (post-processing helper: the Tecplot
 * writer of a partitioned run gathers the rank-local nodal fields with it).  No-op on an unpartitioned handle. */
int en234gpu_comm_allreduce_host(en234gpu_handle h, double* buf, long long n);
int en234gpu_set_partition(en234gpu_handle h, int n_neighbours, const int* neighbour_rank, const int* shared_ptr,
                           const int* shared_nodes, const unsigned char* owned);
/* Optional peer-memory path for the collectives inside the PCG loop (CUDA IPC over NVLink): halo sums become direct
 * stores into the neighbours' buffers overlapped with the interior rows of the SpMV, the dot-product all-reduces one
 * single-CTA kernel each.  Call after comm_init + set_partition on every rank:
 *   en234gpu_comm_arena(h, max over ranks of the largest shared-node list, handle64 out)
 *   exchange the 64-byte handles (all-gather), then en234gpu_comm_connect(h, handles[n_ranks*64]).
 * Without these calls (or with EN234_COMM=nccl) the loop uses ncclSend/ncclRecv/ncclAllReduce. */
int en234gpu_comm_arena(en234gpu_handle h, int max_shared_nodes, char* handle64);
int en234gpu_comm_connect(en234gpu_handle h, const char* handles);

/* ---- several GPUs behind one handle, in one process (no MPI / torchrun / NCCL needed by the caller).
 * The reference's driver is one serial program (src/main.f90 -> src/staticstep.f90:93-160); this is the entry a Fortran
 * driver binds to use all GPUs of a node without changing its structure: en234gpu_create_multi takes the WHOLE mesh (same
 * arrays as en234gpu_create), splits the elements into n_parts contiguous ranges (whole layers of a structured block),
 * creates one context per part on devices[part] (NULL: part p on device p mod device count) — each driven by its own host
 * thread inside the library — and connects them through peer memory (cudaDeviceEnablePeerAccess).  The en234gpu_multi_*
 * calls below have the argument lists of their single-GPU namesakes, with whole-mesh arrays in the reference's layouts;
 * the norms, iteration counts and fail flags they return are the global ones.  The collectives of the PCG loop (halo sums
 * of A p on shared nodes, the three dot products) are folded into the kernels (peer-memory stores + flags).
 * devices[] may repeat a device (tests on a one-GPU box; the parts' kernels are then co-scheduled on it).
 * Element types: the symmetric ones (1001, U1, 1002, U2); methods EN234_SOLVE_PCG_BSR / _MATRIXFREE. */
typedef struct en234gpu_multi_ctx* en234gpu_multi_handle;
int en234gpu_create_multi(en234gpu_multi_handle* h, int n_parts, const int* devices, int n_nodes, int n_elements,
                          const double* coords, const int* connectivity, const int* element_flag,
                          const int* element_prop_index, const double* element_properties, int n_element_properties);
int en234gpu_multi_destroy(en234gpu_multi_handle h);
int en234gpu_multi_set_operator_mode(en234gpu_multi_handle h, int method);
int en234gpu_multi_assemble(en234gpu_multi_handle h, const double* dof_total, const double* dof_increment,
                            const double* f_element_ext, double* rforce, double* nodal_force_norm, int* fail);
int en234gpu_multi_apply_boundaryconditions(en234gpu_multi_handle h, const double* f_ext, int n_fixed, const int* fixed_dof,
                                            const double* fixed_correction, double* unbalanced_force_norm);
int en234gpu_multi_solve(en234gpu_multi_handle h, int method, double cg_tolerance, int max_cg_iterations,
                         double* dof_increment, double* correction_norm, int* cg_iterations, int* fail);
/* device-resident variants (see en234gpu_upload_state ... above) */
int en234gpu_multi_upload_state(en234gpu_multi_handle h, const double* dof_total, const double* dof_increment,
                                const double* f_element_ext, const double* f_ext, int n_fixed, const int* fixed_dof,
                                const double* fixed_value);
int en234gpu_multi_assemble_dev(en234gpu_multi_handle h, double* nodal_force_norm, int* fail);
int en234gpu_multi_apply_boundaryconditions_dev(en234gpu_multi_handle h, double* unbalanced_force_norm);
int en234gpu_multi_solve_dev(en234gpu_multi_handle h, int method, double cg_tolerance, int max_cg_iterations,
                             double* correction_norm, int* cg_iterations, int* fail);
int en234gpu_multi_zero_increment_dev(en234gpu_multi_handle h);
int en234gpu_multi_download_state(en234gpu_multi_handle h, double* dof_increment, double* rforce);
/* inspection */
/* The element partition en234gpu_create_multi builds (host code, no device needed): contiguous element ranges (whole layers
 * of a layer-numbered block), local nodes = interface nodes first then interior, each in ascending global order.
 * sizes[6] = local nodes, local elements, first element (0-based), interface nodes, neighbours, total shared entries;
 * arrays: local_to_global[nodes] (0-based), local_connectivity[8*elements] (1-based local), neighbour_rank[neighbours],
 * shared_ptr[neighbours+1], shared_nodes[shared] (local, same order on both sides), owned[nodes]. */
int en234gpu_partition_sizes(int n_parts, int part, int n_nodes, int n_elements, const int* connectivity, int* sizes);
int en234gpu_partition_arrays(int n_parts, int part, int n_nodes, int n_elements, const int* connectivity,
                              int* local_to_global, int* local_connectivity, int* neighbour_rank, int* shared_ptr,
                              int* shared_nodes, unsigned char* owned);
int en234gpu_multi_n_parts(en234gpu_multi_handle h);
en234gpu_handle en234gpu_multi_part_handle(en234gpu_multi_handle h, int part);
int en234gpu_multi_part_info(en234gpu_multi_handle h, int part, int* device, int* n_nodes, int* n_elements,
                             int* n_interface, int* n_neighbours);

/* ---- measurement: device time (ms, CUDA events on the handle's stream) of the last call of each phase,
 * kernel-launch counter, and per-iteration SpMV time of the last solve */
struct en234gpu_stats {
    double assemble_ms, bc_ms, solve_ms, spmv_ms_avg;
    long long kernel_launches;   /* cumulative number of kernels launched by this handle */
    int last_cg_iterations;
    double last_cg_err;          /* (r.r)/(b.b) at exit */
    int mf_congruent;            /* 1: every element is a translate of element 0 (bitwise) with one material, and the
                                    matrix-free operator multiplies by that single K_e */
};
int en234gpu_get_stats(en234gpu_handle h, struct en234gpu_stats* s);
/* slowest part's phase times, launch counters summed over the parts */
int en234gpu_multi_get_stats(en234gpu_multi_handle h, struct en234gpu_stats* s);
/* average device time (ms, CUDA events on the handle's stream) of `reps` back-to-back launches of the operator
 * kernel y = A p (method 0: block-CSR SpMV; 1: matrix-free apply) after `warm` untimed ones — the per-launch
 * duration the roofline figures are computed from */
int en234gpu_bench_operator(en234gpu_handle h, int method, int warm, int reps, double* ms_per_launch);
/* measurement helper, collective over the ranks: microseconds per call of which = 0 the scalar all-reduce kernel of the PCG
 * loop, 1 a peer-memory halo sum of a DOF vector, 2 an empty kernel (launch floor); 64 calls per captured graph x reps */
int en234gpu_bench_collective(en234gpu_handle h, int which, int reps, double* us_per_call);
/* measured FP64 FMA throughput of the handle's device in TFLOP/s (an FMA micro-kernel, best of 5): the roof of the
 * FP64-bound kernels (element kernel, general-mesh matrix-free operator), reported beside the HBM roof */
int en234gpu_measure_fp64_peak(en234gpu_handle h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
