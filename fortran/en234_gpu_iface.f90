!  ISO_C_BINDING interface to liben234gpu.so (include/en234gpu.h) for the EN234_FEA Fortran driver.
!
!  SOURCE ONLY: no Fortran compiler exists in the build image or on the GPU box, so this module and the patched
!  static step below are shipped uncompiled; the functionally identical C++ driver (en234_fea_b200/csrc/host)
!  carries the tests and measurements.  See INTEGRATION.md for how a maintainer adds these two files to the
!  reference build (they replace nothing: solver type 3 is new, types 1 and 2 keep the reference's CPU paths).
!
!  Argument layouts are the reference's own module arrays (modules/Mesh.f90:51-107):
!    coords(length_coords), connectivity(length_connectivity), dof_total/dof_increment/rforce(length_dofs)
!
module en234_gpu_iface
    use, intrinsic :: iso_c_binding
    implicit none

    integer(c_int), parameter :: EN234_SOLVE_PCG_BSR = 0
    integer(c_int), parameter :: EN234_SOLVE_PCG_MATRIXFREE = 1
    integer(c_int), parameter :: EN234_SOLVE_BICGSTAB_BSR = 2      ! unsymmetric stiffness (SOLVER, ..., UNSYMMETRIC)

    type(c_ptr), save :: gpu_handle = c_null_ptr
    ! several GPUs behind one handle (en234gpu_create_multi): the driver stays one serial program, the library partitions
    ! the mesh, runs one host thread per GPU and takes / returns the whole-mesh module arrays
    type(c_ptr), save :: gpu_multi_handle = c_null_ptr
    integer, save     :: gpu_count = 1

    interface
        ! replaces initialize_cg_solver (src/solver_conjugate_gradient.f90:1039-1244; call site src/staticstep.f90:95)
        function en234gpu_create(h, device, n_nodes, n_elements, coords, connectivity, element_flag, &
                                 element_prop_index, element_properties, n_element_properties) bind(C, name='en234gpu_create')
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out)          :: h
            integer(c_int), value             :: device, n_nodes, n_elements, n_element_properties
            real(c_double), intent(in)        :: coords(*), element_properties(*)
            integer(c_int), intent(in)        :: connectivity(*), element_flag(*), element_prop_index(*)
            integer(c_int)                    :: en234gpu_create
        end function
        function en234gpu_destroy(h) bind(C, name='en234gpu_destroy')
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int)     :: en234gpu_destroy
        end function
        ! replaces assemble_conjugate_gradient_stiffness (src/staticstep.f90:98,122) / assemble_direct_stiffness (:44,:61)
        function en234gpu_assemble(h, dof_total, dof_increment, f_element_ext, rforce, nodal_force_norm, fail) &
                bind(C, name='en234gpu_assemble')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            real(c_double), intent(in)  :: dof_total(*), dof_increment(*)
            type(c_ptr), value          :: f_element_ext          ! c_loc(array) or c_null_ptr
            real(c_double), intent(out) :: rforce(*), nodal_force_norm
            integer(c_int), intent(out) :: fail
            integer(c_int)              :: en234gpu_assemble
        end function
        ! replaces apply_cojugategradient_boundaryconditions (:107,:127) / apply_direct_boundaryconditions (:55,:65)
        function en234gpu_apply_boundaryconditions(h, f_ext, n_fixed, fixed_dof, fixed_correction, &
                                                   unbalanced_force_norm) bind(C, name='en234gpu_apply_boundaryconditions')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            type(c_ptr), value          :: f_ext                  ! c_loc(array) or c_null_ptr
            integer(c_int), value       :: n_fixed
            integer(c_int), intent(in)  :: fixed_dof(*)           ! 0-based DOF indices
            real(c_double), intent(in)  :: fixed_correction(*)
            real(c_double), intent(out) :: unbalanced_force_norm
            integer(c_int)              :: en234gpu_apply_boundaryconditions
        end function
        ! replaces solve_conjugategradient (:108,:130) / solve_direct (:56,:68)
        function en234gpu_solve(h, method, cg_tolerance, max_cg_iterations, dof_increment, correction_norm, &
                                cg_iterations, fail) bind(C, name='en234gpu_solve')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value            :: h
            integer(c_int), value         :: method, max_cg_iterations
            real(c_double), value         :: cg_tolerance
            real(c_double), intent(inout) :: dof_increment(*)
            real(c_double), intent(out)   :: correction_norm
            integer(c_int), intent(out)   :: cg_iterations, fail
            integer(c_int)                :: en234gpu_solve
        end function
        ! initial_state_variables = updated_state_variables on the device (src/staticstep.f90:87); built-in C3D8 elements
        function en234gpu_commit_state(h) bind(C, name='en234gpu_commit_state')
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int)     :: en234gpu_commit_state
        end function
        ! print_state's field projection on the device (src/printing.f90:457-490): replaces assemble_field_projection +
        ! assemble_projection_mass_matrix + LU_decomposition + backsubstitution (or the lumped division);
        ! fields(n_nodes, n_fields) in Fortran order = fields[k*n_nodes + n] (transpose of field_variables(k,n));
        ! field_codes: 0..5 S11 S22 S33 S12 S13 S23, 6 SMISES, 7..12 E11..E23, -1 for a name no element routine knows
        function en234gpu_project_fields(h, n_fields, field_codes, lumped, tolerance, max_iterations, dof_total, &
                                         dof_increment, fields, iterations) bind(C, name='en234gpu_project_fields')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value            :: h
            integer(c_int), value         :: n_fields, lumped, max_iterations
            integer(c_int), intent(in)    :: field_codes(*)
            real(c_double), value         :: tolerance
            real(c_double), intent(in)    :: dof_total(*), dof_increment(*)
            real(c_double), intent(out)   :: fields(*)
            integer(c_int), intent(out)   :: iterations
            integer(c_int)                :: en234gpu_project_fields
        end function
        ! explicit dynamics (src/explicit_dynamic_step.f90:36-70): n_steps whole central-difference steps on the device after
        ! en234gpu_dynamic_setup; replaces assemble_lumped_mass / apply_dynamic_boundaryconditions / assemble_dynamic_force
        function en234gpu_dynamic_steps(h, dt, n_steps, fail, device_ms) bind(C, name='en234gpu_dynamic_steps')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            real(c_double), value       :: dt
            integer(c_int), value       :: n_steps
            integer(c_int), intent(out) :: fail
            real(c_double), intent(out) :: device_ms
            integer(c_int)              :: en234gpu_dynamic_steps
        end function
        function en234gpu_dynamic_get_state(h, dof_total, dof_increment, velocity, acceleration, rforce, lumped_mass, time) &
                bind(C, name='en234gpu_dynamic_get_state')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            real(c_double), intent(out) :: dof_total(*), dof_increment(*), velocity(*), acceleration(*), rforce(*), lumped_mass(*)
            real(c_double), intent(out) :: time
            integer(c_int)              :: en234gpu_dynamic_get_state
        end function
        ! time-independent part of assemble_lumped_mass / apply_dynamic_boundaryconditions (src/explicit_dynamic_step.f90:85-392,
        ! :720-1128): densities, history tables (CSR), unit nodal forces of every load, prescribed DOFs (include/en234gpu.h)
        function en234gpu_dynamic_setup(h, element_density, n_histories, history_ptr, history_time, history_value, n_loads, &
                                        load_ptr, load_dof, load_val, load_history, load_constant, n_pdof, pdof_dof, &
                                        pdof_history, pdof_value, pdof_mode) bind(C, name='en234gpu_dynamic_setup')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            integer(c_int), value       :: n_histories, n_loads, n_pdof
            real(c_double), intent(in)  :: element_density(*), history_time(*), history_value(*), load_val(*), load_constant(*), &
                                           pdof_value(*)
            integer(c_int), intent(in)  :: history_ptr(*), load_ptr(*), load_dof(*), load_history(*), pdof_dof(*), &
                                           pdof_history(*), pdof_mode(*)
            integer(c_int)              :: en234gpu_dynamic_setup
        end function
        function en234gpu_device_count() bind(C, name='en234gpu_device_count')
            import :: c_int
            integer(c_int) :: en234gpu_device_count
        end function
        ! ---- two-node tie constraints (constraint_list%flag < 3): replaces the Lagrange-multiplier block of
        ! assemble_direct_stiffness / assemble_conjugate_gradient_stiffness (solver_direct.f90:292-328) and the multiplier update
        ! of the solve (:967-969).  dof1 / dof2: 0-based positions in dof_total.
        function en234gpu_set_ties(h, n_ties, dof1, dof2) bind(C, name='en234gpu_set_ties')
            import :: c_ptr, c_int
            type(c_ptr), value         :: h
            integer(c_int), value      :: n_ties
            integer(c_int), intent(in) :: dof1(*), dof2(*)
            integer(c_int)             :: en234gpu_set_ties
        end function
        function en234gpu_reset_lagrange_multipliers(h) bind(C, name='en234gpu_reset_lagrange_multipliers')
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int)     :: en234gpu_reset_lagrange_multipliers
        end function
        function en234gpu_get_lagrange_multipliers(h, n_ties, lagrange_multipliers) bind(C, name='en234gpu_get_lagrange_multipliers')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            integer(c_int), value       :: n_ties
            real(c_double), intent(out) :: lagrange_multipliers(*)
            integer(c_int)              :: en234gpu_get_lagrange_multipliers
        end function
        ! ---- several GPUs, one process: same argument lists as the single-GPU calls, whole-mesh arrays
        function en234gpu_create_multi(h, n_parts, devices, n_nodes, n_elements, coords, connectivity, element_flag, &
                                       element_prop_index, element_properties, n_element_properties) &
                bind(C, name='en234gpu_create_multi')
            import :: c_ptr, c_int, c_double
            type(c_ptr), intent(out)          :: h
            integer(c_int), value             :: n_parts, n_nodes, n_elements, n_element_properties
            type(c_ptr), value                :: devices              ! c_loc(integer array) or c_null_ptr: part p on GPU p
            real(c_double), intent(in)        :: coords(*), element_properties(*)
            integer(c_int), intent(in)        :: connectivity(*), element_flag(*), element_prop_index(*)
            integer(c_int)                    :: en234gpu_create_multi
        end function
        function en234gpu_multi_destroy(h) bind(C, name='en234gpu_multi_destroy')
            import :: c_ptr, c_int
            type(c_ptr), value :: h
            integer(c_int)     :: en234gpu_multi_destroy
        end function
        function en234gpu_multi_assemble(h, dof_total, dof_increment, f_element_ext, rforce, nodal_force_norm, fail) &
                bind(C, name='en234gpu_multi_assemble')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            real(c_double), intent(in)  :: dof_total(*), dof_increment(*)
            type(c_ptr), value          :: f_element_ext
            real(c_double), intent(out) :: rforce(*), nodal_force_norm
            integer(c_int), intent(out) :: fail
            integer(c_int)              :: en234gpu_multi_assemble
        end function
        function en234gpu_multi_apply_boundaryconditions(h, f_ext, n_fixed, fixed_dof, fixed_correction, &
                                                         unbalanced_force_norm) &
                bind(C, name='en234gpu_multi_apply_boundaryconditions')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value          :: h
            type(c_ptr), value          :: f_ext
            integer(c_int), value       :: n_fixed
            integer(c_int), intent(in)  :: fixed_dof(*)
            real(c_double), intent(in)  :: fixed_correction(*)
            real(c_double), intent(out) :: unbalanced_force_norm
            integer(c_int)              :: en234gpu_multi_apply_boundaryconditions
        end function
        function en234gpu_multi_solve(h, method, cg_tolerance, max_cg_iterations, dof_increment, correction_norm, &
                                      cg_iterations, fail) bind(C, name='en234gpu_multi_solve')
            import :: c_ptr, c_int, c_double
            type(c_ptr), value            :: h
            integer(c_int), value         :: method, max_cg_iterations
            real(c_double), value         :: cg_tolerance
            real(c_double), intent(inout) :: dof_increment(*)
            real(c_double), intent(out)   :: correction_norm
            integer(c_int), intent(out)   :: cg_iterations, fail
            integer(c_int)                :: en234gpu_multi_solve
        end function
        function en234gpu_last_error() bind(C, name='en234gpu_last_error')
            import :: c_ptr
            type(c_ptr) :: en234gpu_last_error
        end function
    end interface

contains

    ! Builds the element arrays the C ABI expects from module Mesh and creates the device handle.
    ! Called once where the reference calls initialize_cg_solver (src/staticstep.f90:95).
    subroutine initialize_gpu_solver
        use Types
        use ParamIO
        use Mesh
        use Boundaryconditions, only : n_constraints, constraint_list
        implicit none
        integer(c_int), allocatable :: flags(:), propidx(:), tie1(:), tie2(:)
        integer :: lmn, status, nc
        allocate(flags(n_elements), propidx(n_elements))
        do lmn = 1, n_elements
            flags(lmn) = element_list(lmn)%flag
            propidx(lmn) = element_list(lmn)%element_property_index - 1      ! 0-based offset
        end do
        ! all GPUs of the node behind one handle when there are several (the static step then calls the en234gpu_multi_*
        ! twins with the same module arrays), else the single-GPU handle
        gpu_count = en234gpu_device_count()
        if (n_constraints > 0) gpu_count = 1        ! tie constraints: one device
        if (gpu_count > 1) then
            status = en234gpu_create_multi(gpu_multi_handle, int(gpu_count, c_int), c_null_ptr, int(n_nodes, c_int), &
                                           int(n_elements, c_int), coords, connectivity, flags, propidx, element_properties, &
                                           int(length_element_properties, c_int))
        else
            status = en234gpu_create(gpu_handle, 0_c_int, int(n_nodes, c_int), int(n_elements, c_int), coords, connectivity, &
                                     flags, propidx, element_properties, int(length_element_properties, c_int))
        end if
        if (status /= 0) then
            write(IOW, *) ' *** Error creating the GPU solver (en234gpu_create) *** '
            stop
        end if
        deallocate(flags, propidx)
        if (n_constraints > 0) then
            ! CONSTRAIN NODE PAIRS / TIE NODES (flag 1 / 2); general constraints (flag 3, user_constraint) stay on the CPU path
            allocate(tie1(n_constraints), tie2(n_constraints))
            do nc = 1, n_constraints
                if (constraint_list(nc)%flag > 2) then
                    write(IOW, *) ' *** CONSTRAIN NODE SET constraints are not available on the GPU solver *** '
                    stop
                end if
                tie1(nc) = node_list(constraint_list(nc)%node1)%dof_index + constraint_list(nc)%dof1 - 2
                tie2(nc) = node_list(constraint_list(nc)%node2)%dof_index + constraint_list(nc)%dof2 - 2
            end do
            status = en234gpu_set_ties(gpu_handle, int(n_constraints, c_int), tie1, tie2)
            if (status /= 0) then
                write(IOW, *) ' *** Error in the constraint definition (en234gpu_set_ties) *** '
                stop
            end if
            deallocate(tie1, tie2)
        end if
    end subroutine initialize_gpu_solver

    ! The three calls of the static step, on one GPU or on all GPUs of the node (same arguments: whole-mesh module arrays)
    function gpu_assemble(dof_total, dof_increment, f_element_ext, rforce, nodal_force_norm, fail) result(status)
        real(c_double), intent(in)  :: dof_total(*), dof_increment(*)
        type(c_ptr), value          :: f_element_ext
        real(c_double), intent(out) :: rforce(*), nodal_force_norm
        integer(c_int), intent(out) :: fail
        integer(c_int)              :: status
        if (gpu_count > 1) then
            status = en234gpu_multi_assemble(gpu_multi_handle, dof_total, dof_increment, f_element_ext, rforce, nodal_force_norm, fail)
        else
            status = en234gpu_assemble(gpu_handle, dof_total, dof_increment, f_element_ext, rforce, nodal_force_norm, fail)
        end if
    end function gpu_assemble

    function gpu_apply_boundaryconditions(f_ext, n_fixed, fixed_dof, fixed_correction, unbalanced_force_norm) result(status)
        type(c_ptr), value          :: f_ext
        integer(c_int), value       :: n_fixed
        integer(c_int), intent(in)  :: fixed_dof(*)
        real(c_double), intent(in)  :: fixed_correction(*)
        real(c_double), intent(out) :: unbalanced_force_norm
        integer(c_int)              :: status
        if (gpu_count > 1) then
            status = en234gpu_multi_apply_boundaryconditions(gpu_multi_handle, f_ext, n_fixed, fixed_dof, fixed_correction, &
                                                             unbalanced_force_norm)
        else
            status = en234gpu_apply_boundaryconditions(gpu_handle, f_ext, n_fixed, fixed_dof, fixed_correction, &
                                                       unbalanced_force_norm)
        end if
    end function gpu_apply_boundaryconditions

    function gpu_solve(method, cg_tolerance, max_cg_iterations, dof_increment, correction_norm, cg_iterations, fail) result(status)
        use Boundaryconditions, only : n_constraints, lagrange_multipliers
        integer(c_int), value         :: method, max_cg_iterations
        real(c_double), value         :: cg_tolerance
        real(c_double), intent(inout) :: dof_increment(*)
        real(c_double), intent(out)   :: correction_norm
        integer(c_int), intent(out)   :: cg_iterations, fail
        integer(c_int)                :: status
        if (gpu_count > 1) then
            status = en234gpu_multi_solve(gpu_multi_handle, method, cg_tolerance, max_cg_iterations, dof_increment, &
                                          correction_norm, cg_iterations, fail)
        else
            status = en234gpu_solve(gpu_handle, method, cg_tolerance, max_cg_iterations, dof_increment, correction_norm, &
                                    cg_iterations, fail)
            ! module Boundaryconditions keeps the multipliers (solver_direct.f90:967-969); the library owns them between calls
            if (status == 0 .and. n_constraints > 0) &
                status = en234gpu_get_lagrange_multipliers(gpu_handle, int(n_constraints, c_int), lagrange_multipliers)
        end if
    end function gpu_solve

end module en234_gpu_iface
