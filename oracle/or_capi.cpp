// ORACLE — test infrastructure only (see or_model.h).  Plain C API for ctypes (tests/, smoke(), bench.py).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "or_model.h"

using namespace oracle;

namespace {
struct System {
    const Model* m = nullptr;
    int kind = 0;          // 1 direct, 2 cg
    Skyline sky;
    CooUpper coo;
    AsmOut out;
    StepLog log;
};
template <class T> void assign(std::vector<T>& v, const T* p, long n) { v.assign(p, p + n); }
std::vector<double> g_last_lagrange;
}  // namespace

extern "C" {

void* or_model_new() { return new Model(); }
void or_model_free(void* h) { delete (Model*)h; }

int or_model_set_int(void* h, const char* name, const int* p, long n) {
    Model& m = *(Model*)h;
    std::string k(name);
#define VI(f) if (k == #f) { assign(m.f, p, n); return 0; }
#define SI(f) if (k == #f) { m.f = p[0]; return 0; }
    VI(connectivity) VI(element_flag) VI(element_prop_index) VI(element_n_props) VI(element_n_states)
    VI(element_material) VI(material_prop_index) VI(material_n_props)
    VI(history_ptr) VI(nodeset_ptr) VI(nodeset_nodes) VI(elementset_ptr) VI(elementset_elements)
    VI(pdof_flag) VI(pdof_dof) VI(pdof_node) VI(pdof_nodeset) VI(pdof_history) VI(pdof_rate) VI(pdof_subparam) VI(subparam_ptr)
    VI(pforce_flag) VI(pforce_dof) VI(pforce_node) VI(pforce_nodeset) VI(pforce_history)
    VI(con_node1) VI(con_dof1) VI(con_node2) VI(con_dof2)
    VI(node_order) VI(dload_flag) VI(dload_elset) VI(dload_face) VI(dload_history) VI(dload_val_ptr)
    SI(n_nodes) SI(n_elements) SI(nodes_per_element) SI(max_no_steps) SI(solvertype) SI(nonlinear) SI(max_newton_iterations)
    SI(unsymmetric) SI(abaqusformat) SI(max_cg_iterations) SI(gps_renumber) SI(cg_apply_nodal_forces)
    SI(cg_check_linear) SI(explicitdynamicstep) SI(no_dynamic_steps)
#undef VI
#undef SI
    return 1;
}

int or_model_set_double(void* h, const char* name, const double* p, long n) {
    Model& m = *(Model*)h;
    std::string k(name);
#define VD(f) if (k == #f) { assign(m.f, p, n); return 0; }
#define SD(f) if (k == #f) { m.f = p[0]; return 0; }
    VD(coords) VD(element_properties) VD(material_properties) VD(history_time) VD(history_value)
    VD(pdof_value) VD(pforce_value) VD(dload_values) VD(subparam_values) VD(element_density)
    SD(timestep_initial) SD(timestep_min) SD(timestep_max) SD(max_total_time) SD(newtonraphson_tolerance)
    SD(cg_tolerance) SD(dynamic_timestep)
#undef VD
#undef SD
    return 1;
}

// Fill defaults for optional arrays and validate sizes.  Returns 0 if the model is usable.
int or_model_finalize(void* h) {
    Model& m = *(Model*)h;
    const int E = m.n_elements, N = m.n_nodes;
    if ((long)m.coords.size() != 3L * N || (long)m.connectivity.size() != (long)m.nodes_per_element * E) return 1;
    if (m.nodes_per_element != 8 && m.nodes_per_element != 4 && m.nodes_per_element != 10 && m.nodes_per_element != 20) return 4;
    if (m.nodes_per_element != 8) {                       // other element shapes: element 1001 only, no face loads
        for (int f : m.element_flag) if (f != ID_LINELAST) return 5;
        if (!m.dload_flag.empty()) return 6;
    }
    if ((int)m.element_flag.size() != E) return 2;
    if (m.element_prop_index.empty()) m.element_prop_index.assign(E, 0);
    if (m.element_n_props.empty()) m.element_n_props.assign(E, (int)m.element_properties.size());
    if (m.element_n_states.empty()) m.element_n_states.assign(E, 0);
    if (m.element_material.empty()) m.element_material.assign(E, 0);
    if (m.element_density.empty()) m.element_density.assign(E, 0.0);
    if (m.history_ptr.empty()) m.history_ptr.assign(1, 0);
    if (m.nodeset_ptr.empty()) m.nodeset_ptr.assign(1, 0);
    if (m.elementset_ptr.empty()) m.elementset_ptr.assign(1, 0);
    if (m.dload_val_ptr.empty()) m.dload_val_ptr.assign(m.dload_flag.size() + 1, 0);
    for (int c : m.connectivity)
        if (c < 1 || c > N) return 3;
    const size_t nc = m.con_node1.size();
    if (m.con_dof1.size() != nc || m.con_node2.size() != nc || m.con_dof2.size() != nc) return 7;
    for (size_t k = 0; k < nc; ++k)
        if (m.con_node1[k] < 1 || m.con_node1[k] > N || m.con_node2[k] < 1 || m.con_node2[k] > N || m.con_dof1[k] < 1 ||
            m.con_dof1[k] > 3 || m.con_dof2[k] < 1 || m.con_dof2[k] > 3) return 7;
    return 0;
}

// ---- single element ---------------------------------------------------------------------------
// flag: 1001 native linear elastic, 1002 neo-Hookean, 100000 (=U1) UEL linear elastic, 100001 (=U2) UEL neo-Hookean
int or_element(int flag, const double* coords24, const double* dof_increment24, const double* dof_total24,
               const double* props, int ndload, const int* jdltyp, const double* adlmag, double* K, double* R,
               double* svars48, double* energy8) {
    double U[24];
    for (int i = 0; i < 24; ++i) U[i] = dof_increment24[i] + dof_total24[i];
    int fail = 0;
    if (flag == ID_LINELAST) el_linelast_3dbasic(coords24, dof_increment24, dof_total24, props, K, R);
    else if (flag == FLAG_UEL_BASE + 1) {
        double pn;
        uel_linelastic_3d(coords24, U, props, 48, ndload, jdltyp, adlmag, K, R, svars48, energy8, &pn);
    } else if (flag == ID_NEOHOOKE || flag == FLAG_UEL_BASE + 2) el_neohooke_3d(coords24, U, props, K, R, svars48, &fail);
    else return -1;
    return fail;
}

// ---- the other 3-D element shapes of element 1001 (tet4, tet10, hex20): shape-function table, integration rule, element
int or_shapefunctions_3d(int n_nodes, const double* xi, double* f, double* df /*n_nodes x 3*/) {
    double d[20][3];
    if (n_nodes != 4 && n_nodes != 10 && n_nodes != 8 && n_nodes != 20) return 1;
    shapefunctions_3d(n_nodes, xi, f, d);
    for (int a = 0; a < n_nodes; ++a) for (int j = 0; j < 3; ++j) df[3 * a + j] = d[a][j];
    return 0;
}
int or_integration_points_3d(int n_nodes, double* xi /*27 x 3*/, double* w /*27*/) {
    double x[27][3];
    const int n = integration_points_3d(n_nodes, x, w);
    for (int k = 0; k < n; ++k) for (int j = 0; j < 3; ++j) xi[3 * k + j] = x[k][j];
    return n;
}
int or_element_general(int n_nodes, const double* coords, const double* dof_increment, const double* dof_total,
                       const double* props, double* K, double* R) {
    return el_linelast_3d_general(n_nodes, coords, dof_increment, dof_total, props, K, R);
}

// CHECK STIFFNESS (src/checkstiffness.f90:481-532): forward difference of the residual, h = 1e-7,
// column by column; returns sum((Knum-K)^2)/sum(K^2) and the max |K_ij - K_ji|.
int or_check_stiffness(int flag, const double* coords24, const double* dof_increment24, const double* dof_total24,
                       const double* props, double* err_ratio, double* max_asym) {
    double K[576], R0[24], K1[576], R1[24], du[24];
    std::memcpy(du, dof_increment24, sizeof(du));
    if (or_element(flag, coords24, du, dof_total24, props, 0, nullptr, nullptr, K, R0, nullptr, nullptr)) return 1;
    const double h = 1.0e-7;
    double num = 0.0, den = 0.0;
    for (int j = 0; j < 24; ++j) {
        du[j] = dof_increment24[j] + h;
        if (or_element(flag, coords24, du, dof_total24, props, 0, nullptr, nullptr, K1, R1, nullptr, nullptr)) return 1;
        du[j] = dof_increment24[j];
        for (int i = 0; i < 24; ++i) {
            double kn = -(R1[i] - R0[i]) / h;
            num += (kn - K[i + 24 * j]) * (kn - K[i + 24 * j]);
            den += K[i + 24 * j] * K[i + 24 * j];
        }
    }
    *err_ratio = num / den;
    double a = 0.0;
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) a = std::fmax(a, std::fabs(K[i + 24 * j] - K[j + 24 * i]));
    *max_asym = a;
    return 0;
}

// Gibbs-Poole-Stockmeyer node order of the model (bandwidth_minimizer.f90): order[i] = 0-based node numbered i+1
int or_gps_node_order(void* model, int* order) {
    std::vector<int> o;
    gps_node_order(*(Model*)model, o);
    for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
    return 0;
}

// conditioning line of the field-projection mass matrix LU (print_state): out = {max diag, min diag, digits lost}
int or_projection_mass_lu(void* model, const int* order, double* out) {
    Model* m = (Model*)model;
    std::vector<int> o(order, order + m->n_nodes);
    StepLog log;
    projection_mass_lu(*m, o, &log);
    out[0] = log.lu_dimx[0]; out[1] = log.lu_dimn[0]; out[2] = log.lu_ifig[0];
    return 0;
}

// ---- state print (or_print.cpp) ----------------------------------------------------------------
static std::vector<std::string> split_names(const char* csv) {
    std::vector<std::string> v;
    std::string cur;
    for (const char* p = csv ? csv : ""; ; ++p) {
        if (*p == ',' || *p == 0) { if (!cur.empty()) v.push_back(cur); cur.clear(); if (!*p) break; }
        else cur += *p;
    }
    return v;
}
// raw = 1: the assembled right-hand sides int f N dV only; else the projected nodal fields.  out[k + nf*n]
int or_project_fields(void* model, const double* dof_total, const double* dof_increment, const double* svars,
                      const char* names_csv, int lumped, int raw, double* out) {
    const Model& m = *(Model*)model;
    std::vector<std::string> names = split_names(names_csv);
    std::vector<double> f;
    int rc = raw ? assemble_field_projection(m, dof_total, dof_increment, svars, names, f)
                 : project_fields(m, dof_total, dof_increment, svars, names, lumped != 0, f);
    if (rc) return rc;
    std::memcpy(out, f.data(), sizeof(double) * f.size());
    return 0;
}
int or_lumped_projection_mass(void* model, double* out) {
    std::vector<double> l;
    lumped_projection_mass(*(Model*)model, l);
    std::memcpy(out, l.data(), sizeof(double) * l.size());
    return 0;
}
// flags: bit 0 DEGREES OF FREEDOM, bit 1 DISPLACED MESH, bit 2 lumped projection matrix
int or_print_state(void* model, const double* dof_total, const double* dof_increment, const double* svars,
                   const char* names_csv, int flags, double scale_factor, double time_end, const char* path) {
    PrintOptions o;
    o.field_names = split_names(names_csv);
    o.print_dof = flags & 1; o.displaced_mesh = flags & 2; o.lumped = flags & 4;
    o.displacement_scale_factor = scale_factor;
    std::string text;
    if (int rc = print_state(*(Model*)model, o, time_end, dof_total, dof_increment, svars, text)) return rc;
    FILE* f = std::fopen(path, "w");
    if (!f) return -1;
    std::fwrite(text.data(), 1, text.size(), f);
    std::fclose(f);
    return 0;
}

// ---- explicit dynamics (or_dynamic.cpp) -----------------------------------------------------------
namespace { struct Dyn { const Model* m; DynamicState s; }; }
void* or_dynamic_new(void* model, const double* dof_total0, const double* velocity0) {
    Dyn* d = new Dyn();
    d->m = (Model*)model;
    const size_t nd = (size_t)d->m->ndof();
    d->s.dof_total.assign(nd, 0.0);
    d->s.dof_increment.assign(nd, 0.0);
    if (dof_total0) d->s.dof_total.assign(dof_total0, dof_total0 + nd);
    if (velocity0) d->s.dof_increment.assign(velocity0, velocity0 + nd);     // velocity(1:neq) = dof_increment(1:neq), :30
    return d;
}
void or_dynamic_free(void* h) { delete (Dyn*)h; }
int or_dynamic_steps(void* h, int n) { Dyn* d = (Dyn*)h; return explicit_dynamic_steps(*d->m, d->s, n); }
// any pointer may be NULL; time_steps = {TIME, steps done}
void or_dynamic_get(void* h, double* dof_total, double* dof_increment, double* velocity, double* acceleration, double* rforce,
                    double* lumped_mass, double* time_steps) {
    Dyn* d = (Dyn*)h;
    auto cp = [](double* dst, const std::vector<double>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double)); };
    cp(dof_total, d->s.dof_total); cp(dof_increment, d->s.dof_increment); cp(velocity, d->s.velocity);
    cp(acceleration, d->s.acceleration); cp(rforce, d->s.rforce); cp(lumped_mass, d->s.lumped_mass);
    if (time_steps) { time_steps[0] = d->s.time; time_steps[1] = d->s.steps_done; }
}
int or_lumped_mass(void* model, double* out) {
    std::vector<double> l;
    assemble_lumped_mass(*(Model*)model, l);
    std::memcpy(out, l.data(), l.size() * sizeof(double));
    return 0;
}

// ---- assembled systems ------------------------------------------------------------------------
void* or_system_new(void* model, int kind) {
    System* s = new System();
    s->m = (Model*)model;
    s->kind = kind;
    if (kind == 1) {
        std::vector<int> order(s->m->n_nodes);
        if ((int)s->m->node_order.size() == s->m->n_nodes) order = s->m->node_order;
        else for (int i = 0; i < s->m->n_nodes; ++i) order[i] = i;
        compute_profile(*s->m, order, s->sky, &s->log);
    } else initialize_cg_solver(*s->m, s->coo);
    return s;
}
void or_system_free(void* h) { delete (System*)h; }

// assemble: returns fail; rforce[ndof]; norms[0] = nodal_force_norm
int or_system_assemble(void* h, const double* dof_total, const double* dof_increment, double time, double dtime,
                       double* rforce, double* norms, double* svars) {
    System& s = *(System*)h;
    std::vector<double> sv;
    long nsv = 0;
    for (int e = 0; e < s.m->n_elements; ++e) nsv += s.m->element_n_states[e];
    if (svars) sv.assign(svars, svars + nsv);
    int fail = s.kind == 1
                   ? assemble_direct_stiffness(*s.m, s.sky, dof_total, dof_increment, time, dtime, svars ? &sv : nullptr, s.out)
                   : assemble_conjugate_gradient_stiffness(*s.m, s.coo, dof_total, dof_increment, time, dtime, svars ? &sv : nullptr, s.out);
    if (fail) return fail;
    if (svars) std::memcpy(svars, sv.data(), sizeof(double) * nsv);
    if (rforce) std::memcpy(rforce, s.out.rforce.data(), sizeof(double) * s.m->ndof());
    if (norms) norms[0] = s.out.nodal_force_norm;
    return 0;
}
// apply BCs: norms[1] = unbalanced_force_norm
void or_system_apply_bc(void* h, const double* dof_total, const double* dof_increment, double time, double dtime,
                        double* norms) {
    System& s = *(System*)h;
    if (s.kind == 1) apply_direct_boundaryconditions(*s.m, s.sky, dof_total, dof_increment, time, dtime, s.out);
    else apply_conjugategradient_boundaryconditions(*s.m, s.coo, dof_total, dof_increment, time, dtime, s.out,
                                                    s.m->cg_apply_nodal_forces);
    if (norms) norms[1] = s.out.unbalanced_force_norm;
}
long or_system_nnz_upper(void* h) {
    System& s = *(System*)h;
    return s.kind == 1 ? (long)s.sky.aupp.size() : (long)s.coo.aupp.size();
}
// CG system export: strictly-upper COO (0-based dof indices), diagonal and rhs
void or_system_get_coo(void* h, int* rows, int* cols, double* aupp, double* diag, double* rhs) {
    System& s = *(System*)h;
    if (s.kind == 2) {
        long nz = (long)s.coo.aupp.size();
        if (rows) std::memcpy(rows, s.coo.rows.data(), sizeof(int) * nz);
        if (cols) std::memcpy(cols, s.coo.cols.data(), sizeof(int) * nz);
        if (aupp) std::memcpy(aupp, s.coo.aupp.data(), sizeof(double) * nz);
        if (diag) std::memcpy(diag, s.coo.diag.data(), sizeof(double) * s.coo.neq);
        if (rhs) std::memcpy(rhs, s.coo.rhs.data(), sizeof(double) * s.coo.neq);
    } else {
        // rhs/diag in dof order
        for (int d = 0; d < s.m->ndof(); ++d) {
            if (diag) diag[d] = s.sky.diag[s.sky.ieqs[d]];
            if (rhs) rhs[d] = s.sky.rhs[s.sky.ieqs[d]];
        }
    }
}
// Solve the current system; adds the solution to dof_increment.  CG: tol/maxit; fixed_iters >= 0 runs a
// fixed iteration window (timing).  norms[2] = correction_norm.  Returns fail.
int or_system_solve(void* h, double tol, int maxit, int fixed_iters, double* dof_increment, double* norms,
                    int* iters) {
    System& s = *(System*)h;
    if (s.kind == 1) {
        solve_direct(*s.m, s.sky, dof_increment, s.out, &s.log);
        if (iters) *iters = 0;
    } else {
        int it = 0;
        int fail = cg_iteration_loop(s.coo, tol, maxit, &it, fixed_iters);
        if (iters) *iters = it;
        if (fail) return 1;
        double nn = 0.0;
        for (int d = 0; d < s.coo.neq; ++d) { dof_increment[d] += s.coo.sol[d]; nn += s.coo.sol[d] * s.coo.sol[d]; }
        s.out.correction_norm = std::sqrt(nn);
    }
    if (norms) norms[2] = s.out.correction_norm;
    return 0;
}
void or_system_profile(void* h, long long* stats5) {
    System& s = *(System*)h;
    stats5[0] = s.log.profile_neq; stats5[1] = s.log.profile_bwmax; stats5[2] = s.log.profile_bwmean;
    stats5[3] = s.log.profile_bwrms; stats5[4] = s.log.profile_size;
}

// ---- whole static analysis --------------------------------------------------------------------
// newton: rows of 8 doubles (step, iteration, correction, force, unbalanced, ratio, tol, converged)
// info: [n_steps_completed, n_cutbacks, n_newton_records, n_cg_solves, neq, bwmax, bwmean, bwrms, size, n_lu]
int or_run_static(void* model, double* dof_total, double* rforce, double* svars, double* newton, int max_newton_rows,
                  int* cg_iters, int max_cg_rows, long long* info, double* lu_info /*3 per LU*/, int max_lu) {
    const Model& m = *(Model*)model;
    std::vector<double> ut(dof_total, dof_total + m.ndof()), rf, sv;
    long nsv = 0;
    for (int e = 0; e < m.n_elements; ++e) nsv += m.element_n_states[e];
    if (svars) sv.assign(svars, svars + nsv);
    StepLog log;
    int rc = compute_static_step(m, ut, rf, sv, log);
    std::memcpy(dof_total, ut.data(), sizeof(double) * m.ndof());
    if (rforce && !rf.empty()) std::memcpy(rforce, rf.data(), sizeof(double) * m.ndof());
    if (svars && !sv.empty()) std::memcpy(svars, sv.data(), sizeof(double) * nsv);
    int nr = 0;
    for (auto& r : log.newton) {
        if (nr >= max_newton_rows) break;
        double* q = newton + 8 * nr++;
        q[0] = r.step; q[1] = r.iteration; q[2] = r.correction_norm; q[3] = r.nodal_force_norm;
        q[4] = r.unbalanced_force_norm; q[5] = r.ratio; q[6] = r.tolerance; q[7] = r.converged;
    }
    int nc = 0;
    for (int it : log.cg_iterations) { if (nc >= max_cg_rows) break; cg_iters[nc++] = it; }
    int nl = 0;
    for (size_t i = 0; i < log.lu_dimx.size() && nl < max_lu; ++i, ++nl) {
        lu_info[3 * nl] = log.lu_dimx[i]; lu_info[3 * nl + 1] = log.lu_dimn[i]; lu_info[3 * nl + 2] = log.lu_ifig[i];
    }
    info[0] = log.n_steps_completed; info[1] = log.n_cutbacks; info[2] = nr; info[3] = nc;
    info[4] = log.profile_neq; info[5] = log.profile_bwmax; info[6] = log.profile_bwmean;
    info[7] = log.profile_bwrms; info[8] = log.profile_size; info[9] = nl;
    g_last_lagrange = log.lagrange_multipliers;
    return rc;
}
// Lagrange multipliers of the two-node tie constraints after the last or_run_static (module Boundaryconditions keeps them)
int or_last_lagrange(double* out, int n) {
    if ((int)g_last_lagrange.size() != n) return 1;
    for (int i = 0; i < n; ++i) out[i] = g_last_lagrange[i];
    return 0;
}

}  // extern "C"
