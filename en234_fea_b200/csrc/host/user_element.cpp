// Host-callable twins of the reference's element plugin entry points, with the reference's own argument lists and
// the gfortran calling convention (every argument by reference, explicit-shape arrays as bare pointers, default
// integer/logical = 4 bytes), so that a Fortran driver links them in place of
//   user_codes/src/user_element.f90:5-49        subroutine user_element_static  -> symbol user_element_static_
//   user_codes/src/abaqus_uel_linelastic_3d.for:16-28   SUBROUTINE UEL          -> symbol uel_
// Both evaluate ONE hex8 element on the device (en234gpu_element_single): diagnostics / CHECK STIFFNESS use only,
// the timed path evaluates elements inside the assembly kernel.  Unsupported element types set fail / PNEWDT and
// return (the library never stops the process; the reference writes a message and stops, user_element.f90:83-86).
#include <cstdio>
#include <cstring>

#include "../../../include/en234host.h"

namespace en234 {
void uel_face_load(const double* coords24, int face, double mag, double* R24);
}

extern "C" {

void user_element_static_(const int* lmn, const int* element_identifier, const int* n_nodes,
                          const int* node_property_list /* type(node): 7 int32 per node, Mesh.f90:5-14 */,
                          const int* n_properties, const double* element_properties, const int* n_int_properties,
                          const int* int_element_properties, const double* element_coords,
                          const int* length_coord_array, const double* dof_increment, const double* dof_total,
                          const int* length_dof_array, const int* n_state_variables,
                          const double* initial_state_variables, double* updated_state_variables,
                          double* element_stiffness, double* element_residual, int* fail) {
    (void)lmn; (void)node_property_list; (void)n_int_properties; (void)int_element_properties;
    const int nd = *length_dof_array;
    std::memset(element_stiffness, 0, sizeof(double) * (size_t)nd * nd);          // user_element.f90:55-56
    std::memset(element_residual, 0, sizeof(double) * nd);
    *fail = 0;                                                                     // :58
    if (updated_state_variables != initial_state_variables)                        // :60
        std::memcpy(updated_state_variables, initial_state_variables, sizeof(double) * (size_t)*n_state_variables);
    const int id = *element_identifier;
    if ((id != 1001 && id != 1002) || *n_nodes != 8 || *length_coord_array != 24 || nd != 24 || *n_properties < 2) {
        std::fprintf(stderr, " **** user_element_static (GPU twin): element type %d with %d nodes is not available on "
                             "the device path (hex8 ids 1001, 1002 only)\n", id, *n_nodes);
        *fail = 1;
        return;
    }
    int f = 0;
    if (en234gpu_element_single(id, element_coords, element_properties, *n_properties, dof_total, dof_increment,
                                element_stiffness, element_residual, nullptr, nullptr, &f)) {
        std::fprintf(stderr, " **** user_element_static (GPU twin): %s\n", en234gpu_last_error());
        *fail = 1;
        return;
    }
    *fail = f;
}

void uel_(double* RHS, double* AMATRX, double* SVARS, double* ENERGY, const int* NDOFEL, const int* NRHS,
          const int* NSVARS, const double* PROPS, const int* NPROPS, const double* COORDS, const int* MCRD,
          const int* NNODE, const double* U, const double* DU, const double* V, const double* A, const int* JTYPE,
          const double* TIME, const double* DTIME, const int* KSTEP, const int* KINC, const int* JELEM,
          const double* PARAMS, const int* NDLOAD, const int* JDLTYP, const double* ADLMAG, const double* PREDEF,
          const int* NPREDF, const int* LFLAGS, const int* MLVARX, const double* DDLMAG, const int* MDLOAD,
          double* PNEWDT, const int* JPROPS, const int* NJPROP, const double* PERIOD) {
    (void)NRHS; (void)DU; (void)V; (void)A; (void)TIME; (void)DTIME; (void)KSTEP; (void)KINC; (void)JELEM;
    (void)PARAMS; (void)PREDEF; (void)NPREDF; (void)LFLAGS; (void)DDLMAG; (void)JPROPS; (void)NJPROP; (void)PERIOD;
    const int nd = *NDOFEL, ml = *MLVARX;
    for (int i = 0; i < ml; ++i) RHS[i] = 0.0;                                     // abaqus_uel_linelastic_3d.for:142
    std::memset(AMATRX, 0, sizeof(double) * (size_t)nd * nd);                      // :143
    for (int i = 0; i < 8; ++i) ENERGY[i] = 0.0;                                   // :160
    if (*NNODE != 8 || *MCRD != 3 || nd != 24 || ml < 24 || *NPROPS < 2 || (*JTYPE != 1 && *JTYPE != 2)) {
        std::fprintf(stderr, " **** UEL (GPU twin): U%d with %d nodes is not available on the device path "
                             "(hex8 U1, U2 only)\n", *JTYPE, *NNODE);
        *PNEWDT = 0.5;                                                             // forces a cutback in the caller
        return;
    }
    // U holds the DOFs at the end of the increment (solver_direct.f90:234); the element needs only that sum
    const double zero[24] = {0};
    double sv[48], en = 0.0, K[576];
    int f = 0;
    if (en234gpu_element_single(99999 + *JTYPE, COORDS, PROPS, *NPROPS, U, zero, K, RHS, sv, &en, &f) || f) {
        if (!f) std::fprintf(stderr, " **** UEL (GPU twin): %s\n", en234gpu_last_error());
        *PNEWDT = 0.5;
        return;
    }
    std::memcpy(AMATRX, K, sizeof(K));
    ENERGY[1] = en;                                                                // :190-191
    if (*NSVARS >= 48) std::memcpy(SVARS, sv, sizeof(sv));                         // :193-195
    *PNEWDT = 1.0;                                                                 // :199
    const int md = *MDLOAD;
    (void)md;
    for (int j = 0; j < *NDLOAD; ++j) {                                            // :207-238, column 1 of (MDLOAD,*)
        int face = JDLTYP[j] < 0 ? -JDLTYP[j] : JDLTYP[j];
        if (face < 1 || face > 6) continue;
        en234::uel_face_load(COORDS, face, ADLMAG[j], RHS);
    }
}

}  // extern "C"
