!  GPU branch of the static step for EN234_FEA — SOURCE ONLY (uncompiled here: no Fortran compiler in the image).
!
!  compute_static_step_gpu is the CG branch of compute_static_step (src/staticstep.f90:93-160) with the three
!  module-global calls replaced by the C ABI of liben234gpu.so:
!      assemble_conjugate_gradient_stiffness(fail)      -> en234gpu_assemble   (gpu_assemble etc. below dispatch to the
!                                                          en234gpu_multi_* twins when the node has several GPUs)
!      apply_cojugategradient_boundaryconditions        -> gpu_collect_boundaryconditions + en234gpu_apply_boundaryconditions
!      solve_conjugategradient(fail)                    -> en234gpu_solve
!  convergencecheck, compute_static_time_increment, print_state and the state update are the reference's own routines,
!  called unchanged.  To enable it add to src/staticstep.f90 after the solvertype==2 branch:
!      else if (solvertype==3) then
!         call compute_static_step_gpu
!  and to the SOLVER parser (src/read_input_file.f90:1739-1748):
!      else if (strcmp(strpar(2),'GPU',3) ) then
!         solvertype = 3
!
subroutine compute_static_step_gpu
    use, intrinsic :: iso_c_binding
    use Types
    use ParamIO
    use Globals
    use Mesh
    use Boundaryconditions
    use Staticstepparameters
    use Stiffness, only : cg_tolerance, max_cg_iterations
    use en234_gpu_iface
    implicit none

    logical :: fail, continue_timesteps, converged, activatestateprint, activateuserprint
    integer :: iteration, status, n_fixed
    integer(c_int) :: cfail, cg_its
    real (prec) :: new_time_increment
    real (prec), allocatable, target :: f_ext(:), f_element_ext(:)
    integer(c_int), allocatable :: fixed_dof(:)
    real (prec), allocatable :: fixed_correction(:)

    allocate(f_ext(length_dofs), f_element_ext(length_dofs), fixed_dof(length_dofs), fixed_correction(length_dofs))

    DTIME = timestep_initial
    TIME = 0.d0
    continue_timesteps = .true.
    current_step_number = 1
    call initialize_gpu_solver

    do while (continue_timesteps)
        call gpu_collect_boundaryconditions(f_ext, f_element_ext, n_fixed, fixed_dof, fixed_correction)
        status = gpu_assemble(dof_total, dof_increment, c_loc(f_element_ext), rforce, &
                                   nodal_force_norm, cfail)
        fail = (status /= 0) .or. (cfail /= 0)
        if (.not.fail) then
            status = gpu_apply_boundaryconditions(c_loc(f_ext), int(n_fixed, c_int), fixed_dof, &
                                                       fixed_correction, unbalanced_force_norm)
            status = gpu_solve(EN234_SOLVE_PCG_BSR, cg_tolerance, int(max_cg_iterations, c_int), &
                                    dof_increment, correction_norm, cg_its, cfail)
            fail = (status /= 0) .or. (cfail /= 0)
        endif
        if (fail) then                          ! Force a timestep cutback (staticstep.f90:99-117)
            converged = .false.
            iteration = 0
            call compute_static_time_increment(iteration, converged, continue_timesteps, &
                activatestateprint, activateuserprint, new_time_increment)
            DTIME = new_time_increment
            dof_increment = 0.d0
            cycle
        endif
        converged = .true.

        do iteration = 1, max_newton_iterations   ! also run for LINEAR problems, like the direct branch (:60-70)
            call gpu_collect_boundaryconditions(f_ext, f_element_ext, n_fixed, fixed_dof, fixed_correction)
            status = gpu_assemble(dof_total, dof_increment, c_loc(f_element_ext), rforce, &
                                       nodal_force_norm, cfail)
            if (status /= 0 .or. cfail /= 0) then
                converged = .false.
                exit
            endif
            status = gpu_apply_boundaryconditions(c_loc(f_ext), int(n_fixed, c_int), fixed_dof, &
                                                       fixed_correction, unbalanced_force_norm)
            call convergencecheck(iteration, converged)
            if (converged) exit
            status = gpu_solve(EN234_SOLVE_PCG_BSR, cg_tolerance, int(max_cg_iterations, c_int), &
                                    dof_increment, correction_norm, cg_its, cfail)
        end do

        call compute_static_time_increment(iteration, converged, continue_timesteps, &
            activatestateprint, activateuserprint, new_time_increment)
        if (.not.converged) then
            DTIME = new_time_increment
            dof_increment = 0.d0
            cycle
        endif

        if (activatestateprint) call print_state
        if (activateuserprint) call user_print(current_step_number)

        current_step_number = current_step_number + 1
        dof_total = dof_total + dof_increment
        initial_state_variables = updated_state_variables
        status = en234gpu_commit_state(gpu_handle)       ! the device copy of the element state (C3D8 elements)
        TIME = TIME + DTIME
        DTIME = new_time_increment
    end do

    status = en234gpu_destroy(gpu_handle)
    deallocate(f_ext, f_element_ext, fixed_dof, fixed_correction)

end subroutine compute_static_step_gpu

!  Evaluates on the host what the reference's BC routine adds to the right-hand side
!  (src/solver_conjugate_gradient.f90:417-671 / src/solver_direct.f90:501-734): native-element tractions and
!  prescribed nodal forces into f_ext, the list of prescribed DOFs with their corrections.  ABAQUS-UEL distributed
!  loads (which the reference's UEL adds to its own residual, abaqus_uel_linelastic_3d.for:207-238) go to f_element_ext.
subroutine gpu_collect_boundaryconditions(f_ext, f_element_ext, n_fixed, fixed_dof, fixed_correction)
    use, intrinsic :: iso_c_binding
    use Types
    use Globals, only : TIME, DTIME
    use Mesh
    use Boundaryconditions
    use Element_Utilities, only : facenodes
    implicit none

    real (prec), intent(out)    :: f_ext(*), f_element_ext(*), fixed_correction(*)
    integer, intent(out)        :: n_fixed
    integer(c_int), intent(out) :: fixed_dof(*)

    integer :: k, i, j, n, idof, iof, nhist, nnodes, load, flag, elset, ifac, nel, lmn, ntract, nfacenodes, ix, iu, nsr
    integer :: list(8)
    real (prec) :: dofvalue, ucur, dloadvalue, force_value, traction(3)
    real (prec) :: face_coords(24), face_du(24), face_u(24), face_k(24,24), face_r(24)

    f_ext(1:length_dofs) = 0.d0
    f_element_ext(1:length_dofs) = 0.d0

    ! distributed loads on native elements (solver_direct.f90:501-568)
    do load = 1, n_distributedloads
        flag = distributedload_list(load)%flag
        elset = distributedload_list(load)%element_set
        ifac = distributedload_list(load)%face
        nel = elementset_list(elset)%n_elements
        iof = elementset_list(elset)%index
        if (flag >= 4) cycle
        do k = 1, nel
            lmn = element_lists(iof+k-1)
            if (element_list(lmn)%flag > 99999) cycle        ! UEL faces: see gpu_uel_dload below
            traction = 0.d0
            if (flag < 3) then
                ntract = distributedload_list(load)%n_dload_values
                traction(1:ntract) = dload_values(distributedload_list(load)%index_dload_values: &
                                                  distributedload_list(load)%index_dload_values+ntract-1)
            else
                ntract = 1
                traction(1) = 1.d0
            endif
            if (flag == 2) traction = traction/dsqrt(dot_product(traction,traction))
            if (flag > 1) then
                j = history_list(distributedload_list(load)%history_number)%index
                nhist = history_list(distributedload_list(load)%history_number)%n_timevals
                call interpolate_history_table(history_data(1,j), nhist, TIME+DTIME, dloadvalue)
                traction = traction*dloadvalue
            endif
            call facenodes(3, element_list(lmn)%n_nodes, ifac, list, nfacenodes)
            ix = 0
            do j = 1, nfacenodes
                n = connectivity(element_list(lmn)%connect_index + list(j) - 1)
                face_coords(ix+1:ix+3) = coords(node_list(n)%coord_index:node_list(n)%coord_index+2)
                ix = ix + 3
            end do
            iu = 3*nfacenodes
            face_du(1:iu) = 0.d0
            face_u(1:iu) = 0.d0
            call traction_boundarycondition_static(flag, 3, 3, nfacenodes, face_coords(1:ix), ix, face_du(1:iu), &
                face_u(1:iu), iu, traction(1:ntract), ntract, face_k(1:iu,1:iu), face_r(1:iu))
            nsr = 0
            do j = 1, nfacenodes
                n = connectivity(element_list(lmn)%connect_index + list(j) - 1)
                do i = 1, 3
                    nsr = nsr + 1
                    f_ext(node_list(n)%dof_index + i - 1) = f_ext(node_list(n)%dof_index + i - 1) + face_r(nsr)
                end do
            end do
        end do
    end do

    ! prescribed nodal forces (solver_direct.f90:646-682; applied on the GPU path for both solver types, SURVEY Q3)
    do k = 1, n_prescribedforces
        if (prescribedforce_list(k)%flag == 1) then
            force_value = dof_values(prescribedforce_list(k)%index_dof_values)
        else
            iof = history_list(prescribedforce_list(k)%history_number)%index
            nhist = history_list(prescribedforce_list(k)%history_number)%n_timevals
            call interpolate_history_table(history_data(1,iof), nhist, TIME+DTIME, force_value)
        endif
        idof = prescribedforce_list(k)%dof
        if (prescribedforce_list(k)%node_set == 0) then
            n = prescribedforce_list(k)%node_number
            f_ext(idof + node_list(n)%dof_index - 1) = f_ext(idof + node_list(n)%dof_index - 1) + force_value
        else
            nnodes = nodeset_list(prescribedforce_list(k)%node_set)%n_nodes
            iof = nodeset_list(prescribedforce_list(k)%node_set)%index
            do i = 1, nnodes
                n = node_lists(iof+i-1)
                f_ext(idof + node_list(n)%dof_index - 1) = f_ext(idof + node_list(n)%dof_index - 1) + force_value
            end do
        endif
    end do

    ! prescribed DOFs -> (0-based DOF index, correction) in application order (solver_direct.f90:685-734)
    n_fixed = 0
    do k = 1, n_prescribeddof
        if (prescribeddof_list(k)%flag == 1) then
            dofvalue = dof_values(prescribeddof_list(k)%index_dof_values)
        else
            iof = history_list(prescribeddof_list(k)%history_number)%index
            nhist = history_list(prescribeddof_list(k)%history_number)%n_timevals
            call interpolate_history_table(history_data(1,iof), nhist, TIME+DTIME, dofvalue)
        endif
        idof = prescribeddof_list(k)%dof
        if (prescribeddof_list(k)%node_set == 0) then
            nnodes = 1
        else
            nnodes = nodeset_list(prescribeddof_list(k)%node_set)%n_nodes
            iof = nodeset_list(prescribeddof_list(k)%node_set)%index
        endif
        do i = 1, nnodes
            if (prescribeddof_list(k)%node_set == 0) then
                n = prescribeddof_list(k)%node_number
            else
                n = node_lists(iof+i-1)
            endif
            j = idof + node_list(n)%dof_index - 1
            if (prescribeddof_list(k)%rate_flag == 0) then
                ucur = dof_increment(j) + dof_total(j)
            else
                ucur = dof_increment(j)
            endif
            n_fixed = n_fixed + 1
            fixed_dof(n_fixed) = j - 1
            fixed_correction(n_fixed) = dofvalue - ucur
        end do
    end do

    ! ABAQUS-UEL distributed loads: evaluate the face integrals of abaqus_uel_linelastic_3d.for:207-238 for every
    ! loaded UEL face into f_element_ext (the device element kernels do not see DLOADs)
    call gpu_uel_dload(f_element_ext)

end subroutine gpu_collect_boundaryconditions

!  Face load vectors of the ABAQUS UEL elements: calls the reference's own UEL with zero displacements and
!  LFLAGS(3)=5 semantics are not available in EN234_FEA, so the residual of a zero-displacement call (which is
!  exactly the DLOAD contribution, the internal force being zero) is used.
subroutine gpu_uel_dload(f_element_ext)
    use Types
    use Globals, only : TIME, DTIME
    use Mesh
    use Boundaryconditions
    use User_Subroutine_Storage
    use Staticstepparameters, only : current_step_number, max_total_time
    implicit none
    real (prec), intent(inout) :: f_element_ext(*)

    integer :: lmn, j, k, n, ix, nd, abq_MDLOAD, abq_LFLAGS(5), jprops(1)
    real (prec) :: element_coords(24), zero24(24), rhs24(24,1), amatrx(24,24), svars(48), energy(8)
    real (prec) :: abq_time(2), params(3), predef(2,1,8), pnewdt
    real (prec), allocatable :: adlmag(:), ddlmag(:)
    integer, allocatable :: jdltyp(:)

    call generate_abaqus_dloads                        ! src/solver_direct.f90:798-912
    if (length_abq_dlmag_array == 0) return
    allocate(adlmag(length_abq_dlmag_array), ddlmag(length_abq_dlmag_array), jdltyp(length_abq_dlmag_array))
    zero24 = 0.d0
    abq_LFLAGS = 0
    abq_LFLAGS(3) = 1
    abq_time = TIME
    params = 0.d0
    predef = 0.d0
    jprops = 0
    do lmn = 1, n_elements
        if (element_list(lmn)%flag <= 99999) cycle
        abq_MDLOAD = abq_uel_bc_list(lmn)%mdload
        if (abq_MDLOAD == 0) cycle
        k = abq_uel_bc_list(lmn)%mag_index
        adlmag(1:abq_MDLOAD) = abq_uel_bc_mag(k:k+abq_MDLOAD-1)
        jdltyp(1:abq_MDLOAD) = abq_uel_bc_typ(k:k+abq_MDLOAD-1)
        ddlmag(1:abq_MDLOAD) = abq_uel_bc_dmag(k:k+abq_MDLOAD-1)
        ix = 0
        do j = 1, 8
            n = connectivity(element_list(lmn)%connect_index + j - 1)
            element_coords(ix+1:ix+3) = coords(node_list(n)%coord_index:node_list(n)%coord_index+2)
            ix = ix + 3
        end do
        call UEL(rhs24, amatrx, svars, energy, 24, 1, 48, &
            element_properties(element_list(lmn)%element_property_index), element_list(lmn)%n_element_properties, &
            element_coords, 3, 8, zero24, zero24, zero24, zero24, element_list(lmn)%flag-99999, abq_time, DTIME, &
            1, current_step_number, lmn, params, abq_MDLOAD, jdltyp, adlmag, predef, 1, abq_LFLAGS, 24, ddlmag, &
            abq_MDLOAD, pnewdt, jprops, 0, max_total_time)
        nd = 0
        do j = 1, 8
            n = connectivity(element_list(lmn)%connect_index + j - 1)
            do k = 1, 3
                nd = nd + 1
                f_element_ext(node_list(n)%dof_index + k - 1) = f_element_ext(node_list(n)%dof_index + k - 1) + rhs24(nd,1)
            end do
        end do
    end do
    deallocate(adlmag, ddlmag, jdltyp)
end subroutine gpu_uel_dload
