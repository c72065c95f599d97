// ORACLE — test infrastructure only (see or_model.h).  Conjugate-gradient path: restates
// src/solver_conjugate_gradient.f90 (assembly :1-414, BCs :417-671, fixdof :673-713, solve :717-773,
// cg_iteration_loop :776-912, initialise :1039-1244).  Storage keeps the reference's semantics: diagonal
// separate, strictly-upper COO entries in first-touch order (getindex :916-1037 appends a new entry the
// first time a (row,col) pair is seen).  The O(n_nodes^2) set-up (:1094-1189) and the O(nnz)-per-DOF
// fixdof scan (:687-708) are replaced by equivalent-result O(nnz) code.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "or_model.h"

namespace oracle {

void oracle_bc_loads_and_corrections(const Model& m, const double* dof_total, const double* dof_increment,
                                     double time, double dtime, bool nodal_forces, const int* ieqs,
                                     std::vector<double>& rhs, std::vector<std::pair<int, double>>& fix);

// initialize_cg_solver: identity equation numbering (:1078-1091)
void initialize_cg_solver(const Model& m, CooUpper& c) {
    c.neq = m.ndof();
    c.rows.clear(); c.cols.clear(); c.aupp.clear();
    c.diag.assign(c.neq, 0.0);
    c.rhs.assign(c.neq, 0.0);
    c.sol.assign(c.neq, 0.0);
    c.rowbins.assign(c.neq, {});
}

static inline int getindex(CooUpper& c, int irow, int icol) {
    auto& bin = c.rowbins[irow];
    for (auto& p : bin)
        if (p.first == icol) return p.second;
    int idx = (int)c.aupp.size();
    bin.emplace_back(icol, idx);
    c.rows.push_back(irow);
    c.cols.push_back(icol);
    c.aupp.push_back(0.0);
    return idx;
}

// assemble_conjugate_gradient_stiffness (:100-399); thresh = 0 (Stiffness.f90:36)
int assemble_conjugate_gradient_stiffness(const Model& m, CooUpper& c, const double* dof_total,
                                          const double* dof_increment, double time, double dtime,
                                          std::vector<double>* svars, AsmOut& out) {
    std::fill(c.aupp.begin(), c.aupp.end(), 0.0);
    std::fill(c.rhs.begin(), c.rhs.end(), 0.0);
    std::fill(c.diag.begin(), c.diag.end(), 0.0);
    const int npe = m.nodes_per_element, nd = 3 * npe;
    std::vector<double> Kbuf((size_t)nd * nd), Rbuf(nd);
    double *K = Kbuf.data(), *R = Rbuf.data();
    const double thresh = 0.0;
    long long sv_off = 0;
    for (int lmn = 0; lmn < m.n_elements; ++lmn) {
        double* sv = nullptr;
        if (svars && m.element_n_states[lmn] > 0) sv = svars->data() + sv_off;
        sv_off += m.element_n_states[lmn];
        if (element_static(m, lmn, dof_total, dof_increment, time, dtime, K, R, sv)) return 1;
        for (int i1 = 0; i1 < npe; ++i1) {                      // :233-258
            int node1 = m.connectivity[(size_t)npe * lmn + i1] - 1;
            for (int j1 = 0; j1 < 3; ++j1) {
                int irow = 3 * node1 + j1, nsr = 3 * i1 + j1;
                c.rhs[irow] += R[nsr];
                for (int i2 = 0; i2 < npe; ++i2) {
                    int node2 = m.connectivity[(size_t)npe * lmn + i2] - 1;
                    for (int j2 = 0; j2 < 3; ++j2) {
                        int icol = 3 * node2 + j2, nsc = 3 * i2 + j2;
                        if (icol == irow) c.diag[irow] += K[nsr + nd * nsc];
                        else if (icol > irow) {
                            if (std::fabs(K[nsr + nd * nsc]) + std::fabs(K[nsc + nd * nsr]) > thresh)
                                c.aupp[getindex(c, irow, icol)] += K[nsr + nd * nsc];
                        }
                    }
                }
            }
        }
    }
    for (int n = 0; n < c.neq; ++n)                              // :378-391
        if (c.rhs[n] == c.rhs[n] + 1.0 || std::isnan(c.rhs[n])) return 1;
    out.rforce.assign(c.neq, 0.0);                                // :393-399
    double nn = 0.0;
    for (int n = 0; n < c.neq; ++n) { out.rforce[n] = -c.rhs[n]; nn += c.rhs[n] * c.rhs[n]; }
    out.nodal_force_norm = std::sqrt(nn);
    return 0;
}

// apply_cojugategradient_boundaryconditions (:417-671).  The reference CG routine has no prescribed
// nodal-force block; `apply_nodal_forces` selects the literal behaviour (false) or the
// direct solver's (true).  fixdof_conjugategradient (:673-713) is applied for each prescribed DOF in
// order; the per-DOF scan over all entries is replaced by one pass per DOF over an index of the
// entries touching it — identical arithmetic in identical order for each rhs entry.
void apply_conjugategradient_boundaryconditions(const Model& m, CooUpper& c, const double* dof_total,
                                                const double* dof_increment, double time, double dtime,
                                                AsmOut& out, bool apply_nodal_forces) {
    std::vector<int> ident(c.neq);
    for (int i = 0; i < c.neq; ++i) ident[i] = i;
    std::vector<std::pair<int, double>> fix;
    oracle_bc_loads_and_corrections(m, dof_total, dof_increment, time, dtime, apply_nodal_forces, ident.data(),
                                    c.rhs, fix);
    if (!fix.empty()) {
        // index: for each dof, the COO entries having it as row or column, ascending entry index
        std::vector<int> cnt(c.neq + 1, 0);
        const int nz = (int)c.aupp.size();
        for (int i = 0; i < nz; ++i) { cnt[c.rows[i] + 1]++; cnt[c.cols[i] + 1]++; }
        for (int i = 0; i < c.neq; ++i) cnt[i + 1] += cnt[i];
        std::vector<int> pos(cnt.begin(), cnt.end() - 1), idx(2 * (size_t)nz);
        for (int i = 0; i < nz; ++i) { idx[pos[c.rows[i]]++] = i; idx[pos[c.cols[i]]++] = i; }
        for (auto& f : fix) {
            const int n = f.first;
            const double v = f.second;
            for (int q = cnt[n]; q < cnt[n + 1]; ++q) {
                int i = idx[q];
                if (c.rows[i] == n) { c.rhs[c.cols[i]] -= v * c.aupp[i]; c.aupp[i] = 0.0; }
                if (c.cols[i] == n) { c.rhs[c.rows[i]] -= v * c.aupp[i]; c.aupp[i] = 0.0; }
            }
            c.diag[n] = 1.0;
            c.rhs[n] = v;
        }
    }
    double nn = 0.0;
    for (int n = 0; n < c.neq; ++n) nn += c.rhs[n] * c.rhs[n];
    out.unbalanced_force_norm = std::sqrt(nn);
}

// cg_iteration_loop (:776-912): Jacobi-PCG, x0 = 0, err = (r.r)/(b.b), stop when err < tol.
// fixed_iters >= 0 runs exactly that many iterations (CPU-baseline timing window) ignoring tol.
int cg_iteration_loop(CooUpper& c, double tol, int maxit, int* iters, int fixed_iters) {
    const int neq = c.neq;
    const int size = (int)c.aupp.size();
    std::vector<double> p(neq, 0.0), r(neq), z(neq), precon(neq);
    double* sol = c.sol.data();
    const double* rhs = c.rhs.data();
    const double* diag = c.diag.data();
    const double* aupp = c.aupp.data();
    const int* irow = c.rows.data();
    const int* icol = c.cols.data();
    *iters = 0;
    std::fill(c.sol.begin(), c.sol.end(), 0.0);
    int is = 0;
    while (is < neq && rhs[is] == 0.0) ++is;
    if (is == neq) return 0;                                     // zero RHS :805-814
    for (int i = 0; i < neq; ++i) precon[i] = 1.0 / diag[i];     // :817-820
    std::fill(r.begin(), r.end(), 0.0);                          // r = rhs - A*sol, sol = 0 (:825-832)
    for (int i = 0; i < size; ++i) {
        r[irow[i]] += aupp[i] * sol[icol[i]];
        r[icol[i]] += aupp[i] * sol[irow[i]];
    }
    for (int i = 0; i < neq; ++i) r[i] = r[i] + diag[i] * sol[i];
    for (int i = 0; i < neq; ++i) r[i] = rhs[i] - r[i];
    double bnrm = 0.0, err = 0.0;
    for (int i = 0; i < neq; ++i) bnrm += rhs[i] * rhs[i];
    for (int i = 0; i < neq; ++i) err += r[i] * r[i];
    err = err / bnrm;
    if (fixed_iters < 0 && err < tol) return 0;
    for (int i = 0; i < neq; ++i) z[i] = r[i] * precon[i];
    int iter = 0;
    double bkden = 0.0;
    while (fixed_iters >= 0 ? iter < fixed_iters : err > tol) {
        ++iter;
        if (fixed_iters < 0 && iter > maxit) { *iters = iter; return 1; }     // :854-860
        double bknum = 0.0;
        for (int i = 0; i < neq; ++i) bknum += z[i] * r[i];
        if (iter == 1) for (int i = 0; i < neq; ++i) p[i] = z[i];
        else { double bk = bknum / bkden; for (int i = 0; i < neq; ++i) p[i] = bk * p[i] + z[i]; }
        bkden = bknum;
        std::fill(z.begin(), z.end(), 0.0);                      // z = A p :877-883
        for (int i = 0; i < size; ++i) {
            z[irow[i]] += aupp[i] * p[icol[i]];
            z[icol[i]] += aupp[i] * p[irow[i]];
        }
        for (int i = 0; i < neq; ++i) z[i] = z[i] + diag[i] * p[i];
        double akden = 0.0;
        for (int i = 0; i < neq; ++i) akden += z[i] * p[i];
        double ak = bknum / akden;
        for (int i = 0; i < neq; ++i) sol[i] = sol[i] + ak * p[i];
        for (int i = 0; i < neq; ++i) r[i] = r[i] - ak * z[i];
        for (int i = 0; i < neq; ++i) z[i] = r[i] * precon[i];
        err = 0.0;
        for (int i = 0; i < neq; ++i) err += r[i] * r[i];
        err = err / bnrm;
    }
    *iters = iter;
    return 0;
}

// solve_conjugategradient (:717-773)
int solve_conjugategradient(const Model& m, CooUpper& c, double* dof_increment, AsmOut& out, int* iters) {
    int fail = cg_iteration_loop(c, m.cg_tolerance, m.max_cg_iterations, iters);
    if (fail) return 1;
    double nn = 0.0;
    for (int d = 0; d < c.neq; ++d) { dof_increment[d] += c.sol[d]; nn += c.sol[d] * c.sol[d]; }
    out.correction_norm = std::sqrt(nn);
    return 0;
}

}  // namespace oracle
