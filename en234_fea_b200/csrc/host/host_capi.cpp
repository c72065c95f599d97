// C ABI of the host driver (include/en234host.h).
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

#include "../../../include/en234host.h"
#include "loads.h"
#include "model.h"
#include "explicit_dynamic_step.h"
#include "print_state.h"
#include "staticstep.h"

using namespace en234;

namespace en234 {
int check_stiffness(const Model& m, int element_flag, const double* dof_total, const double* dof_increment, FILE* log,
                    double* err_out, int* n_unsymmetric, int* n_numerical_unsymmetric, std::string& error);
}


struct en234host_model {
    Model m;
    std::map<std::string, std::vector<int>> I;
    std::map<std::string, std::vector<double>> D;
    bool flat = false;
};

struct en234host_part {
    int n_ranks = 1, rank = 0;
    std::map<std::string, std::vector<int>> I;
    std::vector<unsigned char> owned;
    std::vector<double> coords;
};

static thread_local std::string g_err;

static void flatten(en234host_model& w) {
    Model& m = w.m;
    auto& I = w.I;
    auto& D = w.D;
    I.clear(); D.clear();
    I["n_nodes"] = {m.n_nodes}; I["n_elements"] = {m.n_elements}; I["nodes_per_element"] = {m.nodes_per_element};
    I["connectivity"] = m.connectivity; I["element_flag"] = m.element_flag;
    I["element_prop_index"] = m.element_prop_index; I["element_n_props"] = m.element_n_props;
    I["element_n_states"] = m.element_n_states;
    D["coords"] = m.coords; D["element_properties"] = m.element_properties;
    I["element_material"] = m.element_material;
    I["element_material"].resize(m.n_elements, 0);
    I["material_prop_index"]; I["material_n_props"];
    for (auto& md : m.materials) { I["material_prop_index"].push_back(md.prop_index); I["material_n_props"].push_back(md.n_props); }
    D["material_properties"] = m.material_properties;
    auto& hp = I["history_ptr"]; hp = {0};
    for (auto& h : m.histories) {
        D["history_time"].insert(D["history_time"].end(), h.t.begin(), h.t.end());
        D["history_value"].insert(D["history_value"].end(), h.v.begin(), h.v.end());
        hp.push_back((int)D["history_time"].size());
    }
    D["history_time"]; D["history_value"];
    auto sets = [&](const std::vector<NamedSet>& v, const char* pn, const char* ln) {
        auto& p = I[pn]; p = {0};
        auto& l = I[ln];
        for (auto& s : v) { l.insert(l.end(), s.items.begin(), s.items.end()); p.push_back((int)l.size()); }
    };
    sets(m.nodesets, "nodeset_ptr", "nodeset_nodes");
    sets(m.elementsets, "elementset_ptr", "elementset_elements");
    auto dofs = [&](const std::vector<PrescribedDof>& v, const std::string& pre, bool rate) {
        for (const char* k : {"_flag", "_dof", "_node", "_nodeset", "_history"}) I[pre + k];
        if (rate) I[pre + "_rate"];
        D[pre + "_value"];
        for (auto& p : v) {
            I[pre + "_flag"].push_back(p.flag); I[pre + "_dof"].push_back(p.dof); I[pre + "_node"].push_back(p.node_number);
            I[pre + "_nodeset"].push_back(p.node_set); I[pre + "_history"].push_back(p.history_number);
            if (rate) I[pre + "_rate"].push_back(p.rate_flag);
            D[pre + "_value"].push_back(p.value);
        }
    };
    dofs(m.prescribed_dofs, "pdof", true);
    dofs(m.prescribed_forces, "pforce", false);
    I["pdof_subparam"];                                               // SUBROUTINE PARAMETERS block per prescribed DOF (0 = none)
    for (auto& p : m.prescribed_dofs) I["pdof_subparam"].push_back(p.subroutine_parameter_number);
    I["subparam_ptr"] = {0};
    D["subparam_values"];
    for (auto& sp : m.subroutine_parameters) {
        D["subparam_values"].insert(D["subparam_values"].end(), sp.values.begin(), sp.values.end());
        I["subparam_ptr"].push_back((int)D["subparam_values"].size());
    }
    for (const char* k : {"con_node1", "con_dof1", "con_node2", "con_dof2"}) I[k];     // two-node tie constraints
    for (auto& c : m.constraints) {
        I["con_node1"].push_back(c.node1); I["con_dof1"].push_back(c.dof1);
        I["con_node2"].push_back(c.node2); I["con_dof2"].push_back(c.dof2);
    }
    for (const char* k : {"dload_flag", "dload_elset", "dload_face", "dload_history"}) I[k];
    I["dload_val_ptr"] = {0};
    D["dload_values"];
    for (auto& d : m.distributed_loads) {
        I["dload_flag"].push_back(d.flag); I["dload_elset"].push_back(d.element_set); I["dload_face"].push_back(d.face);
        I["dload_history"].push_back(d.history_number);
        D["dload_values"].insert(D["dload_values"].end(), d.values.begin(), d.values.end());
        I["dload_val_ptr"].push_back((int)D["dload_values"].size());
    }
    I["max_no_steps"] = {m.max_no_steps}; I["solvertype"] = {m.solvertype}; I["nonlinear"] = {m.nonlinear ? 1 : 0};
    I["max_newton_iterations"] = {m.max_newton_iterations}; I["unsymmetric"] = {m.unsymmetric_stiffness ? 1 : 0};
    I["abaqusformat"] = {m.abaqusformat ? 1 : 0}; I["max_cg_iterations"] = {m.max_cg_iterations};
    I["checkstiffness_flag"] = {m.checkstiffness_flag};
    D["element_density"] = m.element_density;
    D["element_density"].resize(m.n_elements, 0.0);
    I["explicitdynamicstep"] = {m.explicitdynamicstep ? 1 : 0}; I["no_dynamic_steps"] = {m.no_dynamic_steps};
    D["dynamic_timestep"] = {m.dynamic_timestep};
    D["timestep_initial"] = {m.timestep_initial}; D["timestep_min"] = {m.timestep_min};
    D["timestep_max"] = {m.timestep_max}; D["max_total_time"] = {m.max_total_time};
    D["newtonraphson_tolerance"] = {m.newtonraphson_tolerance}; D["cg_tolerance"] = {m.cg_tolerance};
    w.flat = true;
}

extern "C" {

const char* en234host_last_error(void) { return g_err.c_str(); }

en234host_model_t en234host_model_read(const char* path) {
    auto* w = new en234host_model();
    if (!read_input_file(path, w->m)) { g_err = w->m.error; delete w; return nullptr; }
    return w;
}
en234host_model_t en234host_model_from_text(const char* text) {
    auto* w = new en234host_model();
    if (!read_input_text(text, w->m)) { g_err = w->m.error; delete w; return nullptr; }
    return w;
}
void en234host_model_free(en234host_model_t m) { delete m; }

int en234host_model_int(en234host_model_t w, const char* name, const int** p, long* n) {
    if (!w->flat) flatten(*w);
    auto it = w->I.find(name);
    if (it == w->I.end()) { g_err = std::string("unknown int array ") + name; return 1; }
    *p = it->second.data(); *n = (long)it->second.size();
    return 0;
}
int en234host_model_double(en234host_model_t w, const char* name, const double** p, long* n) {
    if (!w->flat) flatten(*w);
    auto it = w->D.find(name);
    if (it == w->D.end()) { g_err = std::string("unknown double array ") + name; return 1; }
    *p = it->second.data(); *n = (long)it->second.size();
    return 0;
}
int en234host_model_find_set(en234host_model_t w, int element_set, const char* name) {
    const auto& v = element_set ? w->m.elementsets : w->m.nodesets;
    for (size_t i = 0; i < v.size(); ++i) if (v[i].name == name) return (int)i + 1;
    return 0;
}
int en234host_model_set(en234host_model_t w, const char* name, double value) {
    Model& m = w->m;
    std::string k(name);
    w->flat = false;
    if (k == "cg_tolerance") m.cg_tolerance = value;
    else if (k == "max_cg_iterations") m.max_cg_iterations = (int)value;
    else if (k == "solvertype") m.solvertype = (int)value;
    else if (k == "nonlinear") m.nonlinear = value != 0;
    else if (k == "newtonraphson_tolerance") m.newtonraphson_tolerance = value;
    else if (k == "max_newton_iterations") m.max_newton_iterations = (int)value;
    else if (k == "max_no_steps") m.max_no_steps = (int)value;
    else if (k == "timestep_initial") m.timestep_initial = value;
    else if (k == "timestep_min") m.timestep_min = value;
    else if (k == "timestep_max") m.timestep_max = value;
    else if (k == "max_total_time") m.max_total_time = value;
    else { g_err = "unknown option " + k; return 1; }
    return 0;
}
long en234host_model_write_text(en234host_model_t w, char* buf, long cap) {
    std::string s = write_input_text(w->m);
    if (buf && cap > (long)s.size()) std::memcpy(buf, s.c_str(), s.size() + 1);
    return (long)s.size() + 1;
}

int en234host_evaluate_loads(en234host_model_t w, double time, double dtime, const double* dof_total,
                             const double* dof_increment, double* f_element_ext, double* f_ext, int* n_fixed,
                             int* fixed_dof, double* fixed_value, double* fixed_correction, int fixed_cap) {
    StepLoads L;
    evaluate_loads(w->m, time, dtime, dof_total, dof_increment, L);
    const size_t nd = (size_t)w->m.ndof();
    if (f_element_ext) std::memcpy(f_element_ext, L.element_ext.data(), nd * sizeof(double));
    if (f_ext) std::memcpy(f_ext, L.ext.data(), nd * sizeof(double));
    *n_fixed = (int)L.fixed_dof.size();
    if ((int)L.fixed_dof.size() > fixed_cap) { g_err = "fixed_cap too small"; return 1; }
    for (size_t i = 0; i < L.fixed_dof.size(); ++i) {
        if (fixed_dof) fixed_dof[i] = L.fixed_dof[i];
        if (fixed_value) fixed_value[i] = L.fixed_value[i];
        if (fixed_correction) fixed_correction[i] = L.fixed_correction[i];
    }
    return 0;
}

int en234host_check_stiffness(en234host_model_t w, int element_flag, const double* dof_total, const double* dof_increment,
                              const char* log_path, double* err, int* n_unsymmetric, int* n_numerical_unsymmetric) {
    FILE* log = log_path ? std::fopen(log_path, "w") : nullptr;
    std::string e;
    const int rc = check_stiffness(w->m, element_flag, dof_total, dof_increment, log, err, n_unsymmetric, n_numerical_unsymmetric, e);
    if (log) std::fclose(log);
    if (rc) g_err = e;
    return rc;
}

// An input file names element routines only by an id ("U1", a MATERIAL block): which Fortran source stood behind the id is
// decided when the reference is linked.  The device implements the shipped ones and nothing else, so models whose property /
// state counts do not fit those routines are refused instead of being run with the wrong physics.
static bool device_supports(const Model& m, std::string& err) {
    if (!m.constraints.empty() && m.explicitdynamicstep) { err = "CONSTRAINTS are implemented for the static step only"; return false; }
    for (int e = 0; e < m.n_elements; ++e) {
        const int flag = m.element_flag[e], np = m.element_n_props[e];
        if (flag == 100000 && np != 2 && np != 3) {
            // abaqus_uel_linelastic_3d.for: PROPS = E, nu; abaqus_vuel.for: E, nu, density
            err = "element " + std::to_string(e + 1) + " is a U1 / VU1 user element with " + std::to_string(np) +
                  " properties: the device has the shipped linear-elastic UEL (E, nu) and VUEL (E, nu, density) only";
            return false;
        }
        if ((flag == 1001 || flag == 1002 || flag == 100001) && np < 2) {
            err = "element " + std::to_string(e + 1) + " (type " + std::to_string(flag) + ") needs two properties";
            return false;
        }
    }
    return true;
}

int en234host_model_device_support(en234host_model_t w) {
    std::string err;
    if (device_supports(w->m, err)) return 0;
    g_err = err;
    return 1;
}

// Property table handed to en234gpu_create: the element properties followed by the material properties; built-in continuum
// elements (flag 10003) take their properties from their MATERIAL (read_input_file.f90:433-509, :756-960), so their
// property index points into the material's section.  Shared by the whole-mesh and the partitioned creators.
static bool device_property_table(const Model& m, std::vector<double>& props, std::vector<int>& pidx, std::string& err) {
    props = m.element_properties;
    pidx = m.element_prop_index;
    props.insert(props.end(), m.material_properties.begin(), m.material_properties.end());
    for (int e = 0; e < m.n_elements; ++e)
        if (m.element_flag[e] == EN234_FLAG_C3D8) {
            const int mi = e < (int)m.element_material.size() ? m.element_material[e] : 0;
            if (mi < 1 || mi > (int)m.materials.size()) { err = "C3D8 element without a MATERIAL"; return false; }
            // the input file does not say which UMAT the executable was linked with; the device has the linear-elastic one
            // (user_codes/src/abaqus_umat_elastic.for: PROPS = E, nu, no state variables) and nothing else
            if (m.materials[mi - 1].n_props != 2 || m.materials[mi - 1].n_states != 0) {
                err = "material '" + m.materials[mi - 1].name + "' has " + std::to_string(m.materials[mi - 1].n_props) +
                      " properties and " + std::to_string(m.materials[mi - 1].n_states) + " state variables: only the "
                      "linear-elastic UMAT (E, nu; no state variables) is available on the device";
                return false;
            }
            pidx[e] = (int)m.element_properties.size() + m.materials[mi - 1].prop_index;
        }
    return true;
}

int en234host_gpu_create(en234host_model_t w, int device, en234gpu_handle* h) {
    const Model& m = w->m;
    if (!device_supports(m, g_err)) return 1;
    std::vector<double> props;
    std::vector<int> pidx;
    if (!device_property_table(m, props, pidx, g_err)) return 1;
    int rc = en234gpu_create_nodes(h, device, m.nodes_per_element, m.n_nodes, m.n_elements, m.coords.data(), m.connectivity.data(),
                                   m.element_flag.data(), pidx.data(), props.data(), (int)props.size());
    if (rc) { g_err = en234gpu_last_error(); return rc; }
    if (!m.constraints.empty()) {
        // two-node ties (CONSTRAIN NODE PAIRS / TIE NODES): 0-based DOF pairs, u[dof1] = u[dof2] (solver_direct.f90:294-328)
        std::vector<int> d1, d2;
        for (const Constraint& c : m.constraints) {
            d1.push_back(3 * (c.node1 - 1) + c.dof1 - 1);
            d2.push_back(3 * (c.node2 - 1) + c.dof2 - 1);
        }
        rc = en234gpu_set_ties(*h, (int)d1.size(), d1.data(), d2.data());
        if (rc) { g_err = en234gpu_last_error(); en234gpu_destroy(*h); *h = nullptr; }
    }
    return rc;
}

// ---- state print ---------------------------------------------------------------------------------------
static std::vector<std::string> split_csv(const char* csv) {
    std::vector<std::string> v;
    std::string cur;
    for (const char* p = csv ? csv : "";; ++p) {
        if (*p == ',' || *p == 0) { if (!cur.empty()) v.push_back(cur); cur.clear(); if (!*p) break; }
        else if (*p != ' ') cur += *p;
    }
    return v;
}
static void copy_text(const std::string& s, char* dst, int cap) {
    if (!dst || cap <= 0) return;
    std::strncpy(dst, s.c_str(), (size_t)cap - 1);
    dst[cap - 1] = 0;
}

int en234host_model_get_state_print(en234host_model_t w, int* flags, int* state_print_steps, double* scale_factor,
                                    char* field_names_csv, int cap_names, char* file_name, int cap_file) {
    const Model& m = w->m;
    if (flags) *flags = (m.stateprint ? 1 : 0) | (m.print_dof ? 2 : 0) | (m.print_displacedmesh ? 4 : 0) | (m.use_lumped_projection_matrix ? 8 : 0);
    if (state_print_steps) *state_print_steps = m.state_print_steps;
    if (scale_factor) *scale_factor = m.displacementscalefactor;
    std::string names;
    for (size_t i = 0; i < m.field_variable_names.size(); ++i) names += (i ? "," : "") + m.field_variable_names[i];
    copy_text(names, field_names_csv, cap_names);
    copy_text(m.state_print_filename, file_name, cap_file);
    return 0;
}

int en234host_model_set_state_print(en234host_model_t w, int flags, int state_print_steps, double scale_factor,
                                    const char* field_names_csv, const char* path) {
    Model& m = w->m;
    m.stateprint = flags & 1; m.print_dof = flags & 2; m.print_displacedmesh = flags & 4; m.use_lumped_projection_matrix = flags & 8;
    if (state_print_steps > 0) m.state_print_steps = state_print_steps;
    m.displacementscalefactor = scale_factor;
    if (field_names_csv) m.field_variable_names = split_csv(field_names_csv);
    if ((int)m.field_variable_names.size() > EN234_MAX_FIELDS) { g_err = "too many field variables"; return 1; }
    m.state_print_path = path ? normalise_path(path) : std::string();
    return 0;
}

int en234host_project_fields(en234host_model_t w, en234gpu_handle h, const char* field_names_csv, int lumped,
                             const double* dof_total, const double* dof_increment, double* fields) {
    std::vector<double> f;
    std::string err;
    if (project_fields(w->m, h, split_csv(field_names_csv), lumped != 0, dof_total, dof_increment, f, err)) { g_err = err; return 1; }
    if (!f.empty()) std::memcpy(fields, f.data(), f.size() * sizeof(double));
    return 0;
}

int en234host_print_state(en234host_model_t w, en234gpu_handle h, double time_end, const double* dof_total,
                          const double* dof_increment, const char* path, int append) {
    std::vector<double> zero;
    if (!dof_increment) { zero.assign((size_t)w->m.ndof(), 0.0); dof_increment = zero.data(); }
    if (!dof_total) { g_err = "print_state needs dof_total"; return 1; }
    FILE* f = std::fopen(normalise_path(path).c_str(), append ? "a" : "w");
    if (!f) { g_err = std::string("cannot open state print file ") + path; return 1; }
    std::string err;
    const int rc = print_state(w->m, h, time_end, dof_total, dof_increment, f, err);
    std::fclose(f);
    if (rc) g_err = err;
    return rc;
}

int en234host_run_dynamic(en234host_model_t w, en234gpu_handle h, int n_steps, const double* initial_velocity,
                          double* dof_total, double* velocity, double* acceleration, double* rforce, double* lumped_mass,
                          long long* info, double* device_ms) {
    FILE* sp = nullptr;
    if (w->m.stateprint && !w->m.state_print_path.empty()) {
        sp = std::fopen(w->m.state_print_path.c_str(), "w");
        if (!sp) { g_err = "cannot open state print file " + w->m.state_print_path; return -1; }
    }
    en234gpu_stats st0{}, st1{};
    en234gpu_get_stats(h, &st0);
    DynamicResult r;
    const int rc = explicit_dynamic_step(w->m, h, n_steps, initial_velocity, sp, r);
    if (sp) std::fclose(sp);
    if (rc) { g_err = r.error; return rc; }
    en234gpu_get_stats(h, &st1);
    const size_t b = (size_t)w->m.ndof() * sizeof(double);
    if (dof_total) std::memcpy(dof_total, r.dof_total.data(), b);
    if (velocity) std::memcpy(velocity, r.velocity.data(), b);
    if (acceleration) std::memcpy(acceleration, r.acceleration.data(), b);
    if (rforce) std::memcpy(rforce, r.rforce.data(), b);
    if (lumped_mass) std::memcpy(lumped_mass, r.lumped_mass.data(), b);
    if (info) { info[0] = r.steps; info[1] = r.n_state_prints; info[2] = st1.kernel_launches - st0.kernel_launches; }
    if (device_ms) *device_ms = r.device_ms;
    return 0;
}

// the same on one rank of a partitioned run (collective over all ranks): arrays are rank-local (3 * part n_nodes)
int en234host_run_dynamic_part(en234host_model_t w, en234host_part_t p, en234gpu_handle h, int n_steps,
                               const double* initial_velocity_local, double* dof_total_local, double* velocity_local,
                               double* acceleration_local, double* rforce_local, double* lumped_mass_local, long long* info,
                               double* device_ms) {
    const auto& l2g = p->I["local_to_global_node"];
    std::vector<int> g2l((size_t)w->m.ndof(), -1);
    for (size_t i = 0; i < l2g.size(); ++i)
        for (int c = 0; c < 3; ++c) g2l[3 * (size_t)l2g[i] + c] = (int)(3 * i + c);
    const size_t nd = 3 * l2g.size();
    DynamicResult r;
    const int rc = explicit_dynamic_step(w->m, h, n_steps, initial_velocity_local, nullptr, r, &g2l, nd, &p->I["local_elements"]);
    if (rc) { g_err = r.error; return rc; }
    const size_t b = nd * sizeof(double);
    if (dof_total_local) std::memcpy(dof_total_local, r.dof_total.data(), b);
    if (velocity_local) std::memcpy(velocity_local, r.velocity.data(), b);
    if (acceleration_local) std::memcpy(acceleration_local, r.acceleration.data(), b);
    if (rforce_local) std::memcpy(rforce_local, r.rforce.data(), b);
    if (lumped_mass_local) std::memcpy(lumped_mass_local, r.lumped_mass.data(), b);
    if (info) { info[0] = r.steps; info[1] = 0; info[2] = 0; }
    if (device_ms) *device_ms = r.device_ms;
    return 0;
}

// Inspection of the explicit-dynamics load tables (tests): dense nodal forces of a unit magnitude of load `load_index`
// (distributed loads first, then prescribed forces), its history id / constant; prescribed DOFs as (dof, history, value, mode).
// p == NULL: whole mesh; else rank-local numbering.  Returns the number of loads, or -1.
int en234host_dynamic_tables(en234host_model_t w, en234host_part_t p, int load_index, double* dense, int* history,
                             double* constant, int* n_pdof, int* pdof_dof, int* pdof_history, double* pdof_value, int* pdof_mode,
                             int cap_pdof, double* element_density) {
    DynamicLoads L;
    std::vector<int> g2l;
    size_t nd = (size_t)w->m.ndof();
    bool ok;
    if (p) {
        const auto& l2g = p->I["local_to_global_node"];
        g2l.assign((size_t)w->m.ndof(), -1);
        for (size_t i = 0; i < l2g.size(); ++i)
            for (int c = 0; c < 3; ++c) g2l[3 * (size_t)l2g[i] + c] = (int)(3 * i + c);
        nd = 3 * l2g.size();
        ok = dynamic_load_tables(w->m, L, &g2l, &p->I["local_elements"]);
    } else ok = dynamic_load_tables(w->m, L);
    if (!ok) { g_err = L.error; return -1; }
    const int n_loads = (int)L.load_history.size();
    if (load_index >= 0 && load_index < n_loads) {
        if (dense) {
            std::fill(dense, dense + nd, 0.0);
            for (int k = L.load_ptr[load_index]; k < L.load_ptr[load_index + 1]; ++k) dense[L.load_dof[k]] = L.load_val[k];
        }
        if (history) *history = L.load_history[load_index];
        if (constant) *constant = L.load_constant[load_index];
    }
    if (n_pdof) *n_pdof = (int)L.pdof_dof.size();
    for (int k = 0; k < (int)L.pdof_dof.size() && k < cap_pdof; ++k) {
        if (pdof_dof) pdof_dof[k] = L.pdof_dof[k];
        if (pdof_history) pdof_history[k] = L.pdof_history[k];
        if (pdof_value) pdof_value[k] = L.pdof_value[k];
        if (pdof_mode) pdof_mode[k] = L.pdof_mode[k];
    }
    if (element_density) std::copy(L.element_density.begin(), L.element_density.end(), element_density);
    return n_loads;
}

int en234host_run_static(en234host_model_t w, en234gpu_handle h, int method, double cg_tolerance,
                         int max_cg_iterations, int check_linear_cg, int use_host_api, double* dof_total,
                         double* rforce, double* newton, int max_rows, int* cg_iterations, int max_cg_rows,
                         long long* info, const char* log_path) {
    StepOptions o;
    o.method = method; o.cg_tolerance = cg_tolerance; o.max_cg_iterations = max_cg_iterations;
    o.check_linear_cg = check_linear_cg != 0; o.use_host_api = use_host_api != 0;
    StepResult r;
    FILE* log = log_path ? std::fopen(log_path, "w") : nullptr;
    if (w->m.stateprint && !w->m.state_print_path.empty()) {
        o.state_print = std::fopen(w->m.state_print_path.c_str(), "w");
        if (!o.state_print) { if (log) std::fclose(log); g_err = "cannot open state print file " + w->m.state_print_path; return -1; }
    }
    int rc = compute_static_step(w->m, h, o, r, log);
    if (log) std::fclose(log);
    if (o.state_print) std::fclose(o.state_print);
    if (rc) g_err = r.error;
    if (info) info[4] = r.n_state_prints;
    const size_t nd = (size_t)w->m.ndof();
    if (dof_total && r.dof_total.size() == nd) std::memcpy(dof_total, r.dof_total.data(), nd * sizeof(double));
    if (rforce && r.rforce.size() == nd) std::memcpy(rforce, r.rforce.data(), nd * sizeof(double));
    int nr = 0;
    for (auto& q : r.newton) {
        if (nr >= max_rows) break;
        double* p = newton + 8 * nr++;
        p[0] = q.step; p[1] = q.iteration; p[2] = q.correction_norm; p[3] = q.nodal_force_norm;
        p[4] = q.unbalanced_force_norm; p[5] = q.ratio; p[6] = q.tolerance; p[7] = q.converged;
    }
    int nc = 0;
    for (int it : r.cg_iterations) { if (nc >= max_cg_rows) break; cg_iterations[nc++] = it; }
    if (info) { info[0] = r.n_steps_completed; info[1] = r.n_cutbacks; info[2] = nr; info[3] = nc; }
    return rc;
}

// same loop on one rank of a partitioned run: DOF arrays are rank-local (length 3 * part n_nodes)
int en234host_run_static_part(en234host_model_t w, en234host_part_t p, en234gpu_handle h, int method, double cg_tolerance,
                              int max_cg_iterations, int check_linear_cg, int use_host_api, double* dof_total_local,
                              double* rforce_local, double* newton, int max_rows, int* cg_iterations, int max_cg_rows,
                              long long* info, const char* log_path) {
    StepOptions o;
    o.method = method; o.cg_tolerance = cg_tolerance; o.max_cg_iterations = max_cg_iterations;
    o.check_linear_cg = check_linear_cg != 0; o.use_host_api = use_host_api != 0;
    const auto& l2g = p->I["local_to_global_node"];
    std::vector<int> g2l((size_t)w->m.ndof(), -1);
    for (size_t i = 0; i < l2g.size(); ++i)
        for (int c = 0; c < 3; ++c) g2l[3 * (size_t)l2g[i] + c] = (int)(3 * i + c);
    const size_t nd = 3 * l2g.size();
    StepResult r;
    FILE* log = (log_path && p->rank == 0) ? std::fopen(log_path, "w") : nullptr;
    bool state_file_failed = false;
    if (w->m.stateprint && !w->m.state_print_path.empty()) {       // every rank projects, rank 0 writes the Tecplot file
        o.state_print_collective = true;
        if (p->rank == 0) {
            o.state_print = std::fopen(w->m.state_print_path.c_str(), "w");
            state_file_failed = !o.state_print;        // reported after the run: the other ranks are already inside it
        }
    }
    int rc = compute_static_step(w->m, h, o, r, log, &g2l, nd);
    if (log) std::fclose(log);
    if (o.state_print) std::fclose(o.state_print);
    if (rc) g_err = r.error;
    else if (state_file_failed) { rc = -1; g_err = "cannot open state print file " + w->m.state_print_path; }
    if (info) info[4] = r.n_state_prints;
    if (dof_total_local && r.dof_total.size() == nd) std::memcpy(dof_total_local, r.dof_total.data(), nd * sizeof(double));
    if (rforce_local && r.rforce.size() == nd) std::memcpy(rforce_local, r.rforce.data(), nd * sizeof(double));
    int nr = 0;
    for (auto& q : r.newton) {
        if (nr >= max_rows) break;
        double* s = newton + 8 * nr++;
        s[0] = q.step; s[1] = q.iteration; s[2] = q.correction_norm; s[3] = q.nodal_force_norm;
        s[4] = q.unbalanced_force_norm; s[5] = q.ratio; s[6] = q.tolerance; s[7] = q.converged;
    }
    int nc = 0;
    for (int it : r.cg_iterations) { if (nc >= max_cg_rows) break; cg_iterations[nc++] = it; }
    if (info) { info[0] = r.n_steps_completed; info[1] = r.n_cutbacks; info[2] = nr; info[3] = nc; }
    return rc;
}

// ---- element partition --------------------------------------------------------------------------------
// Elements are split into n_ranks contiguous ranges of the element numbering (for the structured blocks of
// user_mesh(), whose elements are numbered x-fastest then y then z, a range of whole z-layers is a slab).
// Each rank keeps the closure of its elements' nodes; a node shared with other ranks is owned by the lowest
// rank that has it.  Shared node lists are sorted by global node number so both sides agree on the order.
en234host_part_t en234host_partition(en234host_model_t w, int n_ranks, int rank) {
    const Model& m = w->m;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { g_err = "bad rank"; return nullptr; }
    if (m.nodes_per_element != 8) { g_err = "element partitions are implemented for 8-node elements only"; return nullptr; }
    auto* p = new en234host_part();
    p->n_ranks = n_ranks; p->rank = rank;
    const long long E = m.n_elements;
    // split on whole z-layers when the model is a single structured block, else on plain ranges
    auto first_of = [&](int r) -> long long {
        long long layer = 1;
        if (m.groups.size() == 1 && m.groups[0].zone == "block") {
            // recover nx*ny from the node set sizes is fragile; use the connectivity stride instead:
            // first element of layer k+1 has first node = first node of layer k + (nx+1)(ny+1)
            const int nxny_nodes = m.connectivity[4] - m.connectivity[0];
            // elements per layer = count of elements whose first node < 1 + nxny_nodes and 5th node <= 2*nxny_nodes
            long long cnt = 0;
            while (cnt < E && m.connectivity[8 * cnt] <= nxny_nodes) ++cnt;
            if (cnt > 0 && E % cnt == 0) layer = cnt;
        }
        long long nl = E / layer;
        long long lf = nl * r / n_ranks;
        return lf * layer;
    };
    // every rank validates the whole split, so an impossible one (more ranks than layers / elements) is refused on all
    // ranks alike instead of leaving one rank with an empty mesh while the others wait in the communicator set-up
    for (int r = 0; r < n_ranks; ++r) {
        const long long a = first_of(r), b = r + 1 == n_ranks ? E : first_of(r + 1);
        if (b <= a) {
            g_err = "partition: rank " + std::to_string(r) + " of " + std::to_string(n_ranks) + " would get no elements (" +
                    std::to_string(E) + " elements; structured blocks are split on whole z-layers)";
            delete p;
            return nullptr;
        }
    }
    const long long e0 = first_of(rank), e1 = rank + 1 == n_ranks ? E : first_of(rank + 1);
    // which ranks touch each node: since ranges are contiguous, compute min/max rank per node
    std::vector<int> rmin(m.n_nodes, n_ranks), rmax(m.n_nodes, -1);
    {
        int r = 0;
        long long next = n_ranks > 1 ? first_of(1) : E;
        for (long long e = 0; e < E; ++e) {
            while (r + 1 < n_ranks && e >= next) { ++r; next = r + 1 < n_ranks ? first_of(r + 1) : E; }
            for (int a = 0; a < 8; ++a) {
                int n = m.connectivity[8 * e + a] - 1;
                rmin[n] = std::min(rmin[n], r);
                rmax[n] = std::max(rmax[n], r);
            }
        }
    }
    // exact rank sets for shared nodes (a node may be shared by non-adjacent ranks on unstructured meshes)
    std::vector<int> g2l(m.n_nodes, -1);
    auto& l2g = p->I["local_to_global_node"];
    auto& lel = p->I["local_elements"];
    auto& conn = p->I["connectivity"];
    for (long long e = e0; e < e1; ++e) {
        lel.push_back((int)e);
        for (int a = 0; a < 8; ++a) {
            int n = m.connectivity[8 * e + a] - 1;
            if (g2l[n] < 0) { g2l[n] = 0; }
        }
    }
    // local numbering: interface nodes (touched by more than one rank) first, then the interior, each in ascending
    // global order (keeps the structured locality).  Interface rows can then be computed first and their halo
    // exchange overlapped with the interior rows.
    for (int n = 0; n < m.n_nodes; ++n) if (g2l[n] == 0 && rmin[n] != rmax[n]) { g2l[n] = -2 - (int)l2g.size(); l2g.push_back(n); }
    p->I["n_interface"] = {(int)l2g.size()};
    for (int n = 0; n < m.n_nodes; ++n) if (g2l[n] == 0) { g2l[n] = -2 - (int)l2g.size(); l2g.push_back(n); }
    for (size_t i = 0; i < l2g.size(); ++i) g2l[l2g[i]] = (int)i;
    for (long long e = e0; e < e1; ++e)
        for (int a = 0; a < 8; ++a) conn.push_back(g2l[m.connectivity[8 * e + a] - 1] + 1);
    // neighbours: ranks r != rank that also touch one of my nodes.  Build per-rank node marks lazily.
    std::map<int, std::vector<int>> shared;   // neighbour rank -> local nodes (ascending global order)
    {
        std::vector<std::vector<int>> touch;  // only for nodes with rmin != rmax
        int r = 0;
        long long next = n_ranks > 1 ? first_of(1) : E;
        std::vector<int> last_rank_seen(m.n_nodes, -1);
        std::vector<std::pair<int, int>> pairs;   // (global node, rank) for multi-rank nodes that I hold
        for (long long e = 0; e < E; ++e) {
            while (r + 1 < n_ranks && e >= next) { ++r; next = r + 1 < n_ranks ? first_of(r + 1) : E; }
            if (r == rank) continue;
            for (int a = 0; a < 8; ++a) {
                int n = m.connectivity[8 * e + a] - 1;
                if (g2l[n] >= 0 && last_rank_seen[n] != r) { last_rank_seen[n] = r; pairs.emplace_back(n, r); }
            }
        }
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        for (auto& pr : pairs) shared[pr.second].push_back(g2l[pr.first]);
    }
    auto& nbr = p->I["neighbour_rank"];
    auto& sp = p->I["shared_ptr"]; sp = {0};
    auto& sn = p->I["shared_nodes"];
    for (auto& kv : shared) {
        nbr.push_back(kv.first);
        sn.insert(sn.end(), kv.second.begin(), kv.second.end());
        sp.push_back((int)sn.size());
    }
    p->owned.assign(l2g.size(), 1);
    for (size_t i = 0; i < l2g.size(); ++i) p->owned[i] = rmin[l2g[i]] == rank ? 1 : 0;
    p->I["n_nodes"] = {(int)l2g.size()};
    p->I["n_elements"] = {(int)lel.size()};
    p->I["n_neighbours"] = {(int)nbr.size()};
    p->coords.resize(3 * l2g.size());
    for (size_t i = 0; i < l2g.size(); ++i)
        for (int c = 0; c < 3; ++c) p->coords[3 * i + c] = m.coords[3 * (size_t)l2g[i] + c];
    (void)rmax;
    return p;
}
void en234host_part_free(en234host_part_t p) { delete p; }
int en234host_part_int(en234host_part_t p, const char* name, const int** ptr, long* n) {
    auto it = p->I.find(name);
    if (it == p->I.end()) { g_err = std::string("unknown partition array ") + name; return 1; }
    *ptr = it->second.data(); *n = (long)it->second.size();
    return 0;
}
const unsigned char* en234host_part_owned(en234host_part_t p) { return p->owned.data(); }

int en234host_part_gpu_create(en234host_model_t w, en234host_part_t p, int device, en234gpu_handle* h) {
    const Model& m = w->m;
    if (!device_supports(m, g_err)) return 1;
    if (!m.constraints.empty()) { g_err = "CONSTRAINTS (tied DOFs) are implemented for one device: run this model unpartitioned"; return 1; }
    const auto& lel = p->I["local_elements"];
    std::vector<double> props;
    std::vector<int> gidx;
    if (!device_property_table(m, props, gidx, g_err)) return 1;
    std::vector<int> flag(lel.size()), pidx(lel.size());
    for (size_t i = 0; i < lel.size(); ++i) { flag[i] = m.element_flag[lel[i]]; pidx[i] = gidx[lel[i]]; }
    int rc = en234gpu_create(h, device, p->I["n_nodes"][0], (int)lel.size(), p->coords.data(), p->I["connectivity"].data(),
                             flag.data(), pidx.data(), props.data(), (int)props.size());
    if (rc) { g_err = en234gpu_last_error(); return rc; }
    rc = en234gpu_set_partition(*h, p->I["n_neighbours"][0], p->I["neighbour_rank"].data(), p->I["shared_ptr"].data(),
                                p->I["shared_nodes"].data(), p->owned.data());
    if (rc) g_err = en234gpu_last_error();
    return rc;
}

}  // extern "C"
