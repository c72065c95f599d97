// MESH / USER SUBROUTINE hook (reference: user_codes/src/user_mesh.f90:1-169, reached from
// src/read_input_file.f90:170-194): the reference lets user code fill the Mesh module arrays from a parameter
// list instead of literal COORDINATES/CONNECTIVITY text.  This implementation generates the synthetic
// structured hex8 blocks of BASELINE.json (SURVEY.md 8d "cube(n, jitter, seed)").
//
// Parameter list:  nx, ny, nz, jitter, seed, element_flag, n_state_vars, prop_1 ... prop_k
//   nodes    id = 1 + i + (nx+1) j + (nx+1)(ny+1) k,  x = (i,j,k) h,  h = 1/nx
//            interior nodes displaced by jitter*h*U(-1,1)^3 drawn from std::mt19937_64(seed)
//   elements id = 1 + i + nx j + nx ny k, ABAQUS C3D8 node order (Element_Utilities.f90:908-916)
//   element_flag: 1001 / 1002 native ids, or 99999+n for the ABAQUS "Un" format
// It also defines the node sets xmin,xmax,ymin,ymax,zmin,zmax and the element sets xmin_el ... zmax_el
// (elements having a face on that boundary; faces 6,4,3,5,1,2 in the table Element_Utilities.f90:1102-1109).
#include <cmath>
#include <cstdint>
#include <random>
#include <sstream>

#include "model.h"

namespace en234 {

bool user_mesh(const std::vector<double>& p, Model& m) {
    if (p.size() < 7) { m.error = "expecting nx, ny, nz, jitter, seed, element flag, n_state_vars, properties..."; return false; }
    const int nx = (int)p[0], ny = (int)p[1], nz = (int)p[2];
    const double jitter = p[3];
    const uint64_t seed = (uint64_t)p[4];
    const int flag = (int)p[5], nsvar = (int)p[6];
    if (nx < 1 || ny < 1 || nz < 1) { m.error = "mesh dimensions must be positive"; return false; }
    if ((long long)(nx + 1) * (ny + 1) * (nz + 1) > 2000000000LL / 3) { m.error = "mesh too large for 32-bit DOF numbers"; return false; }
    if (m.n_nodes != 0 || m.n_elements != 0) { m.error = "USER SUBROUTINE mesh must be the only mesh definition"; return false; }
    const int NX = nx + 1, NY = ny + 1, NZ = nz + 1;
    const double h = 1.0 / nx;
    m.n_nodes = NX * NY * NZ;
    m.n_elements = nx * ny * nz;
    m.coords.resize(3 * (size_t)m.n_nodes);
    std::mt19937_64 rng(seed);
    auto u11 = [&]() { return 2.0 * ((double)(rng() >> 11) * (1.0 / 9007199254740992.0)) - 1.0; };
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                size_t n = (size_t)i + (size_t)NX * j + (size_t)NX * NY * k;
                double x = i * h, y = j * h, z = k * h;
                bool interior = i > 0 && i < nx && j > 0 && j < ny && k > 0 && k < nz;
                if (interior && jitter != 0.0) {
                    x += jitter * h * u11();
                    y += jitter * h * u11();
                    z += jitter * h * u11();
                }
                m.coords[3 * n] = x; m.coords[3 * n + 1] = y; m.coords[3 * n + 2] = z;
            }
    m.connectivity.resize(8 * (size_t)m.n_elements);
    const int prop_index = (int)m.element_properties.size();
    for (size_t q = 7; q < p.size(); ++q) m.element_properties.push_back(p[q]);
    const int nprops = (int)p.size() - 7;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                size_t e = (size_t)i + (size_t)nx * j + (size_t)nx * ny * k;
                int n0 = 1 + i + NX * j + NX * NY * k;
                int* c = &m.connectivity[8 * e];
                c[0] = n0; c[1] = n0 + 1; c[2] = n0 + 1 + NX; c[3] = n0 + NX;
                for (int a = 0; a < 4; ++a) c[4 + a] = c[a] + NX * NY;
            }
    m.element_flag.assign(m.n_elements, flag);
    m.element_prop_index.assign(m.n_elements, prop_index);
    m.element_n_props.assign(m.n_elements, nprops);
    m.element_n_states.assign(m.n_elements, nsvar);
    if (flag > 99999) m.abaqusformat = true;
    ElementGroup g;
    g.flag = flag; g.n_states = nsvar; g.prop_index = prop_index; g.n_props = nprops; g.n_elements = m.n_elements;
    g.zone = "block";
    m.groups.push_back(g);

    const char* nn[6] = {"xmin", "xmax", "ymin", "ymax", "zmin", "zmax"};
    NamedSet ns[6], es[6];
    for (int s = 0; s < 6; ++s) { ns[s].name = nn[s]; es[s].name = std::string(nn[s]) + "_el"; }
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) {
                int n = 1 + i + NX * j + NX * NY * k;
                if (i == 0) ns[0].items.push_back(n);
                if (i == nx) ns[1].items.push_back(n);
                if (j == 0) ns[2].items.push_back(n);
                if (j == ny) ns[3].items.push_back(n);
                if (k == 0) ns[4].items.push_back(n);
                if (k == nz) ns[5].items.push_back(n);
            }
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                int e = 1 + i + nx * j + nx * ny * k;
                if (i == 0) es[0].items.push_back(e);
                if (i == nx - 1) es[1].items.push_back(e);
                if (j == 0) es[2].items.push_back(e);
                if (j == ny - 1) es[3].items.push_back(e);
                if (k == 0) es[4].items.push_back(e);
                if (k == nz - 1) es[5].items.push_back(e);
            }
    for (int s = 0; s < 6; ++s) { m.nodesets.push_back(ns[s]); m.elementsets.push_back(es[s]); }
    return true;
}

// Literal `.in` text of a model (COORDINATES / CONNECTIVITY / sets / BCs / STATIC STEP), 17 significant
// digits so that re-parsing reproduces the doubles bit for bit.  Lines stay within the 100-character limit.
std::string write_input_text(const Model& m) {
    std::ostringstream o;
    o.precision(17);
    o << "MESH\n NODES\n  PARAMETERS, 3, 3, 1\n  DISPLACEMENT DOF, 1, 2, 3\n  COORDINATES\n";
    for (int n = 0; n < m.n_nodes; ++n)
        o << "   " << n + 1 << ", " << m.coords[3 * n] << ", " << m.coords[3 * n + 1] << ", " << m.coords[3 * n + 2] << "\n";
    o << "  END COORDINATES\n END NODES\n";
    for (const auto& md : m.materials) {                              // MATERIAL blocks (read_input_file.f90:433-509)
        o << " MATERIAL, " << md.name << "\n  STATE VARIABLES, " << md.n_states << "\n  PROPERTIES\n";
        for (int k = 0; k < md.n_props; ++k) o << "   " << m.material_properties[md.prop_index + k] << "\n";
        o << "  END PROPERTIES\n END MATERIAL\n";
    }
    bool any_internal = false, any_user = false;
    for (const auto& g : m.groups) { any_internal |= g.flag == 10003; any_user |= g.flag != 10003; }
    if (any_internal) {                                               // ELEMENTS, INTERNAL (read_input_file.f90:756-960)
        o << " ELEMENTS, INTERNAL\n";
        for (const auto& g : m.groups) {
            if (g.flag != 10003) continue;
            o << "  TYPE, C3D8\n  PROPERTIES, " << m.materials[m.element_material[g.first_element] - 1].name
              << "\n  CONNECTIVITY, " << (g.zone.empty() ? "zone" : g.zone) << "\n";
            for (int e = g.first_element; e < g.first_element + g.n_elements; ++e) {
                o << "   " << e + 1;
                for (int a = 0; a < 8; ++a) o << ", " << m.connectivity[8 * (size_t)e + a];
                o << "\n";
            }
            o << "  END CONNECTIVITY\n";
        }
        o << " END ELEMENTS\n";
    }
    if (any_user || !any_internal) o << " ELEMENTS, USER\n";
    for (const auto& g : m.groups) {
        if (g.flag == 10003) continue;
        o << "  PARAMETERS, " << m.nodes_per_element << ", " << g.n_states << ", ";
        if (g.flag > 99999) o << (m.explicitdynamicstep ? "VU" : "U") << g.flag - 99999; else o << g.flag;
        o << "\n";
        if (g.first_element < (int)m.element_density.size() && m.element_density[g.first_element] != 0.0)
            o << "  DENSITY, " << m.element_density[g.first_element] << "\n";
        o << "  PROPERTIES\n";
        for (int k = 0; k < g.n_props; ++k) o << "   " << m.element_properties[g.prop_index + k] << "\n";
        o << "  END PROPERTIES\n  CONNECTIVITY, " << (g.zone.empty() ? "zone" : g.zone) << "\n";
        const int npe = m.nodes_per_element;
        for (int e = g.first_element; e < g.first_element + g.n_elements; ++e) {
            o << "   " << e + 1;
            for (int a = 0; a < npe; ++a) o << ", " << m.connectivity[(size_t)npe * e + a];
            o << "\n";
        }
        o << "  END CONNECTIVITY\n";
    }
    if (any_user || !any_internal) o << " END ELEMENTS\n";
    o << "END MESH\nBOUNDARY CONDITIONS\n";
    for (const auto& h : m.histories) {
        o << " HISTORY, " << h.name << "\n";
        for (size_t i = 0; i < h.t.size(); ++i) o << "  " << h.t[i] << ", " << h.v[i] << "\n";
        o << " END HISTORY\n";
    }
    for (const auto& sp : m.subroutine_parameters) {
        o << " SUBROUTINE PARAMETERS, " << sp.name << "\n";
        for (double v : sp.values) o << "  " << v << "\n";
        o << " END SUBROUTINE PARAMETERS\n";
    }
    auto put_set = [&](const char* key, const char* endkey, const NamedSet& s) {
        o << " " << key << ", " << s.name << "\n";
        for (size_t i = 0; i < s.items.size(); ++i) {
            o << (i % 8 == 0 ? "  " : ", ") << s.items[i];
            if (i % 8 == 7 || i + 1 == s.items.size()) o << "\n";
        }
        o << " " << endkey << "\n";
    };
    for (const auto& s : m.nodesets) put_set("NODESET", "END NODESET", s);
    for (const auto& s : m.elementsets) put_set("ELEMENTSET", "END ELEMENTSET", s);
    auto put_dofs = [&](const char* key, const char* endkey, const std::vector<PrescribedDof>& v) {
        if (v.empty()) return;
        o << " " << key << "\n";
        for (const auto& p : v) {
            o << "  ";
            if (p.node_set) o << m.nodesets[p.node_set - 1].name; else o << p.node_number;
            o << ", " << p.dof << ", ";
            if (p.flag == 1) o << "VALUE, " << p.value;
            else if (p.flag == 3) o << "SUBROUTINE, " << m.subroutine_parameters[p.subroutine_parameter_number - 1].name;
            else o << "HISTORY, " << m.histories[p.history_number - 1].name;
            if (p.rate_flag) o << ", RATE";
            o << "\n";
        }
        o << " " << endkey << "\n";
    };
    put_dofs("DEGREES OF FREEDOM", "END DEGREES OF FREEDOM", m.prescribed_dofs);
    put_dofs("FORCES", "END FORCES", m.prescribed_forces);
    if (!m.distributed_loads.empty()) {
        o << " DISTRIBUTED LOADS\n";
        for (const auto& d : m.distributed_loads) {
            o << "  " << m.elementsets[d.element_set - 1].name << ", ";
            if (d.uel_format) {
                o << "U" << (d.flag < 0 ? -d.flag : d.flag) << (d.flag < 0 ? "NU" : "") << ", ";
                if (d.history_number) o << m.histories[d.history_number - 1].name; else o << d.values[0];
            } else {
                o << d.face << ", ";
                if (d.flag == 1) o << "VALUE";
                else if (d.flag == 2) o << "HISTORY, " << m.histories[d.history_number - 1].name;
                else o << "NORMAL, " << m.histories[d.history_number - 1].name;
                for (double v : d.values) o << ", " << v;
            }
            o << "\n";
        }
        o << " END DISTRIBUTED LOADS\n";
    }
    for (const auto& cp : m.constraint_parameters) {                  // read_input_file.f90:1067-1098
        o << " CONSTRAINT PARAMETERS, " << cp.name << "\n";
        for (double v : cp.values) o << "  " << v << "\n";
        o << " END CONSTRAINT PARAMETERS\n";
    }
    if (!m.constraints.empty()) {                                     // two-node ties, written pair by pair (:1436-1444)
        o << " CONSTRAINTS\n";
        for (const auto& c : m.constraints)
            o << "  CONSTRAIN NODE PAIRS, 1, " << c.node1 << ", " << c.dof1 << ", " << c.node2 << ", " << c.dof2 << "\n";
        o << " END CONSTRAINTS\n";
    }
    o << "END BOUNDARY CONDITIONS\n";
    auto put_print_state = [&]() {
        o << " PRINT STATE, " << m.state_print_filename << "\n";
        if (m.print_dof) o << "  DEGREES OF FREEDOM\n";
        if (!m.field_variable_names.empty()) {
            o << "  FIELD VARIABLES";
            for (const std::string& nm : m.field_variable_names) o << ", " << nm;
            o << "\n";
        }
        if (m.print_displacedmesh) o << "  DISPLACED MESH\n";
        o << "  DISPLACEMENT SCALE FACTOR, " << m.displacementscalefactor << "\n END PRINT STATE\n";
    };
    if (m.staticstep) {
        o << "STATIC STEP\n INITIAL TIME STEP, " << m.timestep_initial << "\n MAX TIME STEP, " << m.timestep_max
          << "\n MIN TIME STEP, " << m.timestep_min << "\n MAX NUMBER OF STEPS, " << m.max_no_steps
          << "\n STOP TIME, " << m.max_total_time << "\n SOLVER, "
          << (m.solvertype == 1 ? "DIRECT" : "CONJUGATE GRADIENT") << ", ";
        if (m.nonlinear) o << "NONLINEAR, " << m.newtonraphson_tolerance << ", " << m.max_newton_iterations;
        else o << "LINEAR";
        if (m.unsymmetric_stiffness) o << ", UNSYMMETRIC";
        o << "\n CG TOLERANCE, " << m.cg_tolerance << "\n CG MAX ITERATIONS, " << m.max_cg_iterations << "\n";
        if (m.stateprint) {
            o << " STATE PRINT STEPS, " << m.state_print_steps << "\n";
            put_print_state();
        }
        o << "END STATIC STEP\n";
    }
    if (m.explicitdynamicstep) {
        o << "EXPLICIT DYNAMIC STEP\n TIME STEP, " << m.dynamic_timestep << "\n NUMBER OF STEPS, " << m.no_dynamic_steps << "\n";
        if (m.stateprint) {
            o << " STATE PRINT STEPS, " << m.state_print_steps << "\n";
            put_print_state();
        }
        o << "END EXPLICIT DYNAMIC STEP\n";
    }
    o << "STOP\n";
    return o.str();
}

}  // namespace en234
