// check_stiffness (src/checkstiffness.f90:1-571), the reference's own self-test of an element routine: the coded stiffness
// is compared with the numerical derivative of the residual, column by column (h = 1e-7, :427-463 / :481-513), the
// criterion is the squared Frobenius ratio err = sum (K_num - K)^2 / sum K^2 <= 1e-6 (:515-532), followed by the symmetry
// report of both matrices (:534-547).  The element routines are the device twins with the reference's signatures
// (user_element.cpp): user_element_static_ for native identifiers, uel_ for ABAQUS "Un" elements.  Activated like the
// reference by the CHECK STIFFNESS key (main.f90:136) through en234host_check_stiffness.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../../include/en234host.h"
#include "model.h"

namespace en234 {

int check_stiffness(const Model& m, int element_flag, const double* dof_total, const double* dof_increment, FILE* log,
                    double* err_out, int* n_unsymmetric, int* n_numerical_unsymmetric, std::string& error) {
    int lmn = 0;                                                       // find an element of type element_flag (:107-115)
    while (lmn < m.n_elements && m.element_flag[lmn] != element_flag) ++lmn;
    if (lmn >= m.n_elements) { error = "check_stiffness: no element with identifier " + std::to_string(element_flag) + " was found in mesh"; return 1; }
    if (element_flag == 10003) { error = "check_stiffness: the C3D8 + UMAT element uses an approximate stiffness (CHECK MATERIAL TANGENT is the reference's tool for it)"; return 1; }
    const int nn = 8, iu = 24, ix = 24;
    std::vector<double> xc(ix), du(iu), ut(iu), K(iu * iu), K1(iu * iu), Kn(iu * iu, 0.0), r0(iu), r1(iu);
    std::vector<int> nodes(nn), node_props(7 * nn, 0);
    for (int a = 0; a < nn; ++a) {
        nodes[a] = m.connectivity[8 * (size_t)lmn + a];
        for (int i = 0; i < 3; ++i) {
            xc[3 * a + i] = m.coords[3 * (size_t)(nodes[a] - 1) + i];
            du[3 * a + i] = dof_increment ? dof_increment[3 * (size_t)(nodes[a] - 1) + i] : 0.0;
            ut[3 * a + i] = dof_total ? dof_total[3 * (size_t)(nodes[a] - 1) + i] : 0.0;
        }
        int* np = &node_props[7 * a];                                 // type(node), Mesh.f90:5-14
        np[1] = 3 * a + 1; np[2] = 3; np[3] = 3 * a + 1; np[4] = 3;
    }
    const double* props = m.element_properties.data() + m.element_prop_index[lmn];
    const int nprops = m.element_n_props[lmn], nstate = 1, one = 1, zero = 0, lm1 = lmn + 1;
    double sv0[1] = {0.0}, sv1[1] = {0.0};
    const bool uel = element_flag > 99999;
    auto evaluate = [&](std::vector<double>& Kout, std::vector<double>& Rout) -> bool {
        if (uel) {                                                    // :400-413: U = u + du, DU = du, NRHS = 1, MLVARX = NDOFEL
            std::vector<double> U(iu), V(iu, 0.0), SV(48), EN(8), predef(2 * nn, 0.0);
            for (int d = 0; d < iu; ++d) U[d] = du[d] + ut[d];
            const int jtype = element_flag - 99999, mcrd = 3, nsv = 48, lflags[5] = {1, m.nonlinear ? 1 : 0, 1, 0, 0}, jdl = 0;
            const double time2[2] = {0.0, 0.0}, dtime = m.timestep_initial, params[3] = {0, 0, 0}, adl = 0.0, period = m.max_total_time;
            double pnewdt = 1.0;
            uel_(Rout.data(), Kout.data(), SV.data(), EN.data(), &iu, &one, &nsv, props, &nprops, xc.data(), &mcrd, &nn, U.data(),
                 du.data(), V.data(), V.data(), &jtype, time2, &dtime, &one, &one, &lm1, params, &zero, &jdl, &adl, predef.data(),
                 &one, lflags, &iu, &adl, &one, &pnewdt, &zero, &zero, &period);
            return pnewdt >= 1.0;
        }
        int fail = 0;
        user_element_static_(&lm1, &element_flag, &nn, node_props.data(), &nprops, props, &zero, &zero, xc.data(), &ix, du.data(),
                             ut.data(), &iu, &nstate, sv0, sv1, Kout.data(), Rout.data(), &fail);
        return fail == 0;
    };
    if (log) std::fprintf(log, "  \n  \n  ===================== STIFFNESS CHECK ============================\n");
    if (!evaluate(K, r0)) { error = "check_stiffness: the element routine reported failure"; return 1; }
    int icount = 0;
    for (int j = 0; j < nn; ++j)
        for (int i = 0; i < 3; ++i) {
            du[icount] += 1.0e-7;
            if (!evaluate(K1, r1)) { error = "check_stiffness: the element routine reported failure"; return 1; }
            for (int r = 0; r < iu; ++r) Kn[r + iu * icount] = -(r1[r] - r0[r]) / 1.0e-7;      // :447 / :499
            if (log) {
                std::fprintf(log, "\n  Column %11d  node %11d  DOF %11d\n", icount + 1, nodes[j], i + 1);
                for (int r = 0; r < iu; ++r)
                    std::fprintf(log, " Row %4d node %4d DOF %4d Stiffness %15.5E Numerical deriv %15.5E\n", r + 1, j + 1, r % 3 + 1,
                                 K[r + iu * icount], Kn[r + iu * icount]);
            }
            du[icount] -= 1.0e-7;
            ++icount;
        }
    double num = 0.0, den = 0.0;                                      // :515-516
    for (int q = 0; q < iu * iu; ++q) { num += (Kn[q] - K[q]) * (Kn[q] - K[q]); den += K[q] * K[q]; }
    const double err = num / den;
    if (log) {
        if (err > 1.0e-6)
            std::fprintf(log, "\n  The normalized difference between the numerical and exact stiffness matrix is %24.16E\n\n"
                              "  If you are testing a user element, check the printed exact and numerical stiffness \n"
                              "  for errors.   Small differences may be caused by rounding error.   Large ones \n"
                              "  (more than a few %%) are likely to be caused by an error in your coded stiffness \n \n", err);
        else std::fprintf(log, "\n\n Stiffness matrix is consistent with residual \n\n\n");
    }
    int nu = 0, nnu = 0;                                              // :534-547
    for (int i = 0; i < iu; ++i)
        for (int j = 0; j < iu; ++j) {
            double e = K[i + iu * j] - K[j + iu * i], s = K[i + iu * j] + K[j + iu * i];
            if (e * e > 1.0e-6 * s * s) { ++nu; if (log) std::fprintf(log, "  Stiffness is unsymmetric at col, row %11d %11d\n", i + 1, j + 1); }
            e = Kn[i + iu * j] - Kn[j + iu * i]; s = Kn[i + iu * j] + Kn[j + iu * i];
            if (e * e > 1.0e-6 * s * s) { ++nnu; if (log) std::fprintf(log, "  Numerical stiffness is unsymmetric at col, row %11d %11d\n", i + 1, j + 1); }
        }
    if (err_out) *err_out = err;
    if (n_unsymmetric) *n_unsymmetric = nu;
    if (n_numerical_unsymmetric) *n_numerical_unsymmetric = nnu;
    return 0;
}

}  // namespace en234
