// Keyword parser for EN234_FEA `.in` files — the subset the static-step hot path needs.
// Format rules follow src/parse.f90:4-123 (blanks stripped, comma-separated fields, '%' starts a comment,
// at most 100 characters per line, token typing int/real/string) and src/parse.f90:155-176
// (case-insensitive PREFIX keyword matching).  Keyword structure follows src/read_input_file.f90:160-2135.
// Keywords outside the subset (PRINT ..., CHECK MATERIAL TANGENT, materials, constraints, explicit dynamics)
// are parsed and ignored or rejected with a message — never silently misread.
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "model.h"

namespace en234 {
namespace {

struct Line {
    std::vector<std::string> tok;
    std::vector<int> typ;        // 0 integer, 1 real, 2 string (parse.f90:78-121)
    std::string raw;
    int n() const { return (int)tok.size(); }
};

// keyword test: first `len` characters of token equal `key` ignoring case (parse.f90:155-176)
bool kw(const std::string& t, const char* key) {
    size_t len = std::char_traits<char>::length(key);
    if (t.size() < len) return false;
    for (size_t i = 0; i < len; ++i)
        if (std::toupper((unsigned char)t[i]) != std::toupper((unsigned char)key[i])) return false;
    return true;
}

int classify(const std::string& s) {
    int nd = 0, ne = 0, nz = 0, npm = 0, np = 0, nnum = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        char nx = i + 1 < s.size() ? s[i + 1] : ' ';
        if (c >= '0' && c <= '9') ++nnum;
        else if (c == 'D' || c == 'd') { ++nd; if (nx == '+' || nx == '-') --npm; }
        else if (c == 'E' || c == 'e') { ++ne; if (nx == '+' || nx == '-') --npm; }
        else if (c == '+' || c == '-') ++npm;
        else if (c == '.') ++np;
        else ++nz;
    }
    int t = 0;
    if (np == 1) t = 1;
    if (np > 1 || nd > 1 || ne > 1 || npm > 1) t = 2;
    if (nz > 0) t = 2;
    if (nnum == 0) t = 2;
    return t;
}

// list-directed read of a real: accepts Fortran "1.d0" / "1.D-05" exponents
double to_real(const std::string& s) {
    std::string t = s;
    for (char& c : t)
        if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(t.c_str(), nullptr);
}
int to_int(const std::string& s) { return (int)std::lround(to_real(s)); }

struct Reader {
    std::istream& in;
    Model& m;
    explicit Reader(std::istream& i, Model& mm) : in(i), m(mm) {}
    // next non-blank, non-comment line; false at end of file
    bool next(Line& L) {
        std::string s;
        while (std::getline(in, s)) {
            if (s.size() > 100) s.resize(100);                       // format A100, read_input_file.f90:2487
            L.tok.clear(); L.typ.clear(); L.raw = s;
            // first non-blank character '%' => comment line (parse.f90:49-58)
            size_t k = 0;
            while (k < s.size() && (s[k] == ' ' || s[k] == '\t' || s[k] == '\r')) ++k;
            if (k == s.size() || s[k] == '%') continue;
            std::string cur;
            L.tok.emplace_back();
            for (char c : s) {
                if (c == ' ' || c == '\t' || c == '\r') continue;
                if (c == ',') { L.tok.emplace_back(); continue; }
                if (c == '%') break;
                L.tok.back().push_back(c);
            }
            for (auto& t : L.tok) L.typ.push_back(classify(t));
            return true;
        }
        return false;
    }
    bool fail(const std::string& msg, const Line* L = nullptr) {
        m.error = msg;
        if (L) m.error += " | found: " + L->raw;
        return false;
    }
};

int find_name(const std::vector<History>& v, const std::string& n) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i].name == n) return (int)i + 1;
    return 0;
}
int find_name(const std::vector<NamedSet>& v, const std::string& n) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i].name == n) return (int)i + 1;
    return 0;
}

// skip a block until a line whose first token starts with `endkey`
bool skip_block(Reader& r, const char* endkey) {
    Line L;
    while (r.next(L)) if (kw(L.tok[0], endkey)) return true;
    return r.fail(std::string("unterminated block, expecting ") + endkey);
}

bool parse_nodes(Reader& r) {
    Model& m = r.m;
    Line L;
    while (r.next(L)) {
        if (kw(L.tok[0], "PARAM")) {                                   // :206-213
            if (L.n() < 3) return r.fail("NODES/PARAMETERS needs #coords, #dof", &L);
            m.n_coords = to_int(L.tok[1]);
            m.n_dof = to_int(L.tok[2]);
            if (m.n_coords != 3 || m.n_dof != 3)
                return r.fail("only 3-coordinate / 3-DOF nodes are supported on the GPU hot path", &L);
        } else if (kw(L.tok[0], "DISPLACEMENTDOF")) {                  // :214-232 (used by printing only)
        } else if (kw(L.tok[0], "COOR")) {                             // :233-284
            double scale = L.n() > 1 ? to_real(L.tok[1]) : 1.0;
            while (true) {
                if (!r.next(L)) return r.fail("COORDINATES not terminated");
                if (kw(L.tok[0], "ENDCOOR")) break;
                if (L.typ[0] == 2) return r.fail("COORDINATES key must be terminated by an END COORDINATES", &L);
                if (L.n() > 4 || L.n() < 3) return r.fail("expecting 3 coords of a node", &L);
                for (int k = L.n() - 3; k < L.n(); ++k) m.coords.push_back(to_real(L.tok[k]) * scale);
                m.n_nodes++;
            }
        } else if (kw(L.tok[0], "ENDNODE")) {
            return true;
        } else if (kw(L.tok[0], "INIT") || kw(L.tok[0], "CREATE")) {
            return r.fail("NODES: INITIAL DOF / CREATE NODES are outside the supported subset", &L);
        } else return r.fail("unknown key in NODES block", &L);
    }
    return r.fail("NODES block not terminated");
}

bool parse_elements_user(Reader& r) {
    Model& m = r.m;
    Line L;
    int nnod = 0, nsvar = 0, lmntyp = 0, prop_index = 0, nprops = 0;
    double density = 0.0;
    while (r.next(L)) {
        if (kw(L.tok[0], "PARAM")) {                                   // :530-564
            if (L.n() != 4 || L.typ[1] != 0 || L.typ[2] != 0)
                return r.fail("ELEMENTS/PARAMETERS needs #nodes, #state vars, identifier", &L);
            nnod = to_int(L.tok[1]);
            nsvar = to_int(L.tok[2]);
            if (L.typ[3] == 0) lmntyp = to_int(L.tok[3]);
            else if (kw(L.tok[3], "VU")) { m.abaqusformat = true; lmntyp = to_int(L.tok[3].substr(2)) + 99999; }   // :551-555
            else if (kw(L.tok[3], "U")) { m.abaqusformat = true; lmntyp = to_int(L.tok[3].substr(1)) + 99999; }
            else return r.fail("element identifier must be an integer or Un", &L);
            // el_linelast_3Dbasic.f90:82-85 integrates 4-, 10-, 8- and 20-node elements; a mesh has one shape throughout
            if (nnod != 8 && nnod != 4 && nnod != 10 && nnod != 20)
                return r.fail("3-D elements have 4, 10, 8 or 20 nodes", &L);
            if (nnod != 8 && lmntyp != 1001)
                return r.fail("4-, 10- and 20-node elements are implemented for element type 1001 only", &L);
            if (m.n_elements > 0 && nnod != m.nodes_per_element)
                return r.fail("all elements of a mesh must have the same number of nodes on the device path", &L);
            m.nodes_per_element = nnod;
        } else if (kw(L.tok[0], "DENS")) {                             // :565-574
            if (L.n() < 2 || L.typ[1] > 1) return r.fail("expecting density value following DENSITY key", &L);
            density = to_real(L.tok[1]);
        } else if (kw(L.tok[0], "PROP")) {                             // :575-607
            prop_index = (int)m.element_properties.size();
            nprops = 0;
            while (true) {
                if (!r.next(L)) return r.fail("PROPERTIES not terminated");
                if (kw(L.tok[0], "ENDPROP")) break;
                if (L.typ[0] == 2) return r.fail("PROPERTIES key must be terminated by an END PROPERTIES", &L);
                for (int k = 0; k < L.n(); ++k) {
                    if (L.typ[k] == 2) return r.fail("element properties must be real or integer values", &L);
                    m.element_properties.push_back(to_real(L.tok[k]));
                    ++nprops;
                }
            }
        } else if (kw(L.tok[0], "INTEGERPROP")) {
            if (!skip_block(r, "ENDINTE")) return false;
        } else if (kw(L.tok[0], "INITIALS")) {
            if (!skip_block(r, "ENDINIT")) return false;
        } else if (kw(L.tok[0], "CONN")) {                             // :679-744
            if (nnod == 0) return r.fail("PARAMETERS must be specified for elements before CONNECTIVITY", &L);
            ElementGroup g;
            g.flag = lmntyp; g.n_nodes = nnod; g.n_states = nsvar; g.prop_index = prop_index; g.n_props = nprops;
            g.first_element = m.n_elements;
            if (L.n() > 1 && L.typ[1] == 2) g.zone = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("CONNECTIVITY not terminated");
                if (kw(L.tok[0], "ENDCONN")) break;
                if (L.typ[0] == 2) return r.fail("CONNECTIVITY key must be terminated by an END CONNECTIVITY", &L);
                if (L.n() < nnod || L.n() > nnod + 1) return r.fail("expecting element number (optional) and the element's nodes", &L);
                for (int k = L.n() - nnod; k < L.n(); ++k) m.connectivity.push_back(to_int(L.tok[k]));
                m.element_flag.push_back(lmntyp);
                m.element_prop_index.push_back(prop_index);
                m.element_n_props.push_back(nprops);
                m.element_n_states.push_back(nsvar);
                m.element_density.resize(m.n_elements, 0.0);
                m.element_density.push_back(density);
                m.n_elements++;
            }
            g.n_elements = m.n_elements - g.first_element;
            m.groups.push_back(g);
        } else if (kw(L.tok[0], "ENDELEM")) {
            return true;
        } else return r.fail("unknown key in ELEMENTS, USER block", &L);
    }
    return r.fail("ELEMENTS block not terminated");
}

// MATERIAL, name / STATE VARIABLES, n / PROPERTIES ... END PROPERTIES / END MATERIAL   (read_input_file.f90:433-509)
bool parse_material(Reader& r, const Line& head) {
    Model& m = r.m;
    if (head.n() < 2) return r.fail("MATERIAL key must specify a name for the material", &head);
    Model::MaterialDef md;
    md.name = head.tok[1];
    Line L;
    while (r.next(L)) {
        if (kw(L.tok[0], "STAT")) {
            if (L.n() != 2 || L.typ[1] != 0) return r.fail("STATE VARIABLES key in a MATERIAL must specify no. state variables", &L);
            md.n_states = to_int(L.tok[1]);
        } else if (kw(L.tok[0], "PROP")) {
            md.prop_index = (int)m.material_properties.size();
            md.n_props = 0;
            while (true) {
                if (!r.next(L)) return r.fail("PROPERTIES not terminated");
                if (kw(L.tok[0], "ENDPROP")) break;
                for (int k = 0; k < L.n(); ++k) {
                    if (L.typ[k] == 2) return r.fail("material properties must be real or integer values", &L);
                    m.material_properties.push_back(to_real(L.tok[k]));
                    md.n_props++;
                }
            }
        } else if (kw(L.tok[0], "ENDMATE")) { m.materials.push_back(md); return true; }
        else return r.fail("MATERIAL key must be terminated by an END MATERIAL", &L);
    }
    return r.fail("MATERIAL block not terminated");
}

// ELEMENTS, INTERNAL: TYPE, C3D8 / PROPERTIES, material name / CONNECTIVITY   (read_input_file.f90:756-960)
bool parse_elements_internal(Reader& r) {
    Model& m = r.m;
    Line L;
    int nnod = 0, nsvar = 0, n_int_pts = 0, material = 0;
    while (r.next(L)) {
        if (kw(L.tok[0], "TYPE")) {
            if (L.n() != 2) return r.fail("TYPE key must specify the element type", &L);
            if (kw(L.tok[1], "C3D8")) { nnod = 8; n_int_pts = 8; nsvar = 15; }
            else return r.fail("only TYPE, C3D8 internal continuum elements are recognised", &L);
        } else if (kw(L.tok[0], "DENS")) {
        } else if (kw(L.tok[0], "PROP")) {
            if (L.n() < 2) return r.fail("PROPERTIES of an internal element must name a material", &L);
            material = 0;
            for (size_t i = 0; i < m.materials.size(); ++i) if (m.materials[i].name == L.tok[1]) material = (int)i + 1;
            if (!material) return r.fail("unknown material name", &L);
            nsvar = 15 + m.materials[material - 1].n_states;
        } else if (kw(L.tok[0], "INITIALS")) { if (!skip_block(r, "ENDINIT")) return false; }
        else if (kw(L.tok[0], "CONN")) {
            if (nnod == 0 || material == 0) return r.fail("TYPE and PROPERTIES must precede CONNECTIVITY", &L);
            ElementGroup g;
            g.flag = 10003; g.n_nodes = nnod; g.n_states = nsvar * n_int_pts; g.first_element = m.n_elements;
            if (L.n() > 1 && L.typ[1] == 2) g.zone = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("CONNECTIVITY not terminated");
                if (kw(L.tok[0], "ENDCONN")) break;
                if (L.typ[0] == 2) return r.fail("CONNECTIVITY key must be terminated by an END CONNECTIVITY", &L);
                if (L.n() < nnod || L.n() > nnod + 1) return r.fail("expecting element number (optional) and 8 nodes", &L);
                for (int k = L.n() - nnod; k < L.n(); ++k) m.connectivity.push_back(to_int(L.tok[k]));
                m.element_flag.push_back(10003);
                m.element_prop_index.push_back(0);
                m.element_n_props.push_back(0);
                m.element_n_states.push_back(nsvar * n_int_pts);
                m.element_material.resize(m.n_elements, 0);
                m.element_material.push_back(material);
                m.n_elements++;
            }
            g.n_elements = m.n_elements - g.first_element;
            m.groups.push_back(g);
        } else if (kw(L.tok[0], "ENDELEM")) return true;
        else return r.fail("unknown key in ELEMENTS, INTERNAL block", &L);
    }
    return r.fail("ELEMENTS block not terminated");
}

bool parse_user_mesh(Reader& r) {
    Line L;
    std::vector<double> params;
    while (r.next(L)) {
        if (kw(L.tok[0], "ENDUS")) return user_mesh(params, r.m) ? true : r.fail("user_mesh: " + r.m.error);
        for (int k = 0; k < L.n(); ++k) {
            if (L.typ[k] == 2) return r.fail("expecting a list of parameters defining user-subroutine generated mesh", &L);
            params.push_back(to_real(L.tok[k]));
        }
    }
    return r.fail("USER SUBROUTINE block not terminated");
}

bool parse_mesh(Reader& r) {
    Line L;
    while (r.next(L)) {
        if (kw(L.tok[0], "USERSUB")) { if (!parse_user_mesh(r)) return false; }
        else if (kw(L.tok[0], "NODE")) { if (!parse_nodes(r)) return false; }
        else if (kw(L.tok[0], "MATE")) { if (!parse_material(r, L)) return false; }
        else if (kw(L.tok[0], "ELEM")) {
            if (L.n() < 2) return r.fail("ELEMENT key must be followed by USER or INTERNAL", &L);
            if (kw(L.tok[1], "USER")) { if (!parse_elements_user(r)) return false; }
            else if (kw(L.tok[1], "INTER")) { if (!parse_elements_internal(r)) return false; }
            else return r.fail("ELEMENT key must be followed by USER or INTERNAL", &L);
        } else if (kw(L.tok[0], "ENDMESH")) { r.m.element_material.resize(r.m.n_elements, 0); return true; }
        else return r.fail("expecting an END MESH statement", &L);
    }
    return r.fail("MESH block not terminated");
}

bool parse_dof_like(Reader& r, Line& L, bool is_force, PrescribedDof& p) {
    Model& m = r.m;
    if (L.n() < 4) return r.fail("expecting node/node set, DOF, VALUE/HISTORY, value/name", &L);
    if (L.typ[0] == 0) {
        p.node_number = to_int(L.tok[0]);
        p.node_set = 0;
        if (p.node_number < 1 || p.node_number > m.n_nodes) return r.fail("node number out of range", &L);
    } else if (L.typ[0] == 2) {
        p.node_set = find_name(m.nodesets, L.tok[0]);
        if (!p.node_set) return r.fail("unknown node set name", &L);
    } else return r.fail("expecting a node number or node set name", &L);
    p.dof = to_int(L.tok[1]);
    if (p.dof < 1 || p.dof > 3) return r.fail("attempt to prescribe a DOF the node does not have", &L);
    if (kw(L.tok[2], "VALUE")) { p.flag = 1; p.value = to_real(L.tok[3]); }
    else if (kw(L.tok[2], "HISTO")) {
        p.flag = 2;
        p.history_number = find_name(m.histories, L.tok[3]);
        if (!p.history_number) return r.fail("unknown history name", &L);
    } else if (kw(L.tok[2], "SUBRO")) {                               // read_input_file.f90:1212-1225
        if (is_force) return r.fail("SUBROUTINE-type nodal forces are outside the supported subset", &L);
        p.flag = 3;
        p.subroutine_parameter_number = 0;
        for (size_t i = 0; i < m.subroutine_parameters.size(); ++i)
            if (m.subroutine_parameters[i].name == L.tok[3]) p.subroutine_parameter_number = (int)i + 1;
        if (!p.subroutine_parameter_number) return r.fail("unknown SUBROUTINE PARAMETERS name", &L);
    } else return r.fail("expecting key VALUE, HISTORY, or SUBROUTINE in a dof definition", &L);
    p.rate_flag = (!is_force && kw(L.tok[L.n() - 1], "RATE")) ? 1 : 0;
    return true;
}

bool parse_boundary(Reader& r) {
    Model& m = r.m;
    Line L;
    while (r.next(L)) {
        if (kw(L.tok[0], "HISTORY")) {                                 // :1005-1032
            if (L.n() < 2) return r.fail("HISTORY key was used without specifying a name", &L);
            History h;
            h.name = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("HISTORY not terminated");
                if (kw(L.tok[0], "ENDHIST")) break;
                if (L.n() < 2) return r.fail("expecting time, value in a HISTORY", &L);
                h.t.push_back(to_real(L.tok[0]));
                h.v.push_back(to_real(L.tok[1]));
            }
            m.histories.push_back(h);
        } else if (kw(L.tok[0], "NODESET")) {                          // :1100-1144
            if (L.n() < 2) return r.fail("NODESET needs a name", &L);
            NamedSet s;
            s.name = L.tok[1];
            bool generate = L.n() == 3;
            while (true) {
                if (!r.next(L)) return r.fail("NODESET not terminated");
                if (kw(L.tok[0], "ENDNODE")) break;
                if (generate) {                                        // start, end, increment (:1118-1126)
                    if (L.n() < 3) return r.fail("expecting start, end, increment in a generated NODESET", &L);
                    int a = to_int(L.tok[0]), b = to_int(L.tok[1]), inc = to_int(L.tok[2]);
                    if (inc == 0) inc = 1;
                    for (int k = a; inc > 0 ? k <= b : k >= b; k += inc) s.items.push_back(k);
                } else
                    for (int k = 0; k < L.n(); ++k) s.items.push_back(to_int(L.tok[k]));
            }
            for (int n : s.items) if (n < 1 || n > m.n_nodes) return r.fail("node set " + s.name + ": node out of range");
            m.nodesets.push_back(s);
        } else if (kw(L.tok[0], "ELEMENTSET")) {                       // :1145-1166
            if (L.n() < 2) return r.fail("ELEMENTSET needs a name", &L);
            NamedSet s;
            s.name = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("ELEMENTSET not terminated");
                if (kw(L.tok[0], "ENDELEM")) break;
                for (int k = 0; k < L.n(); ++k) s.items.push_back(to_int(L.tok[k]));
            }
            for (int e : s.items) if (e < 1 || e > m.n_elements) return r.fail("element set " + s.name + ": element out of range");
            m.elementsets.push_back(s);
        } else if (kw(L.tok[0], "DEGREES")) {                          // :1167-1251
            while (true) {
                if (!r.next(L)) return r.fail("DEGREES OF FREEDOM not terminated");
                if (kw(L.tok[0], "ENDDEGR")) break;
                PrescribedDof p;
                if (!parse_dof_like(r, L, false, p)) return false;
                m.prescribed_dofs.push_back(p);
            }
        } else if (kw(L.tok[0], "FORCES")) {                           // :1254-1308
            while (true) {
                if (!r.next(L)) return r.fail("FORCES not terminated");
                if (kw(L.tok[0], "ENDFORC")) break;
                PrescribedDof p;
                if (!parse_dof_like(r, L, true, p)) return false;
                m.prescribed_forces.push_back(p);
            }
        } else if (kw(L.tok[0], "DISTRIB")) {                          // :1310-1428
            while (true) {
                if (!r.next(L)) return r.fail("DISTRIBUTED LOADS not terminated");
                if (kw(L.tok[0], "ENDDISTRIB")) break;
                if (L.n() < 3) return r.fail("expecting element set, face/Un, ...", &L);
                DistributedLoad d;
                d.element_set = find_name(m.elementsets, L.tok[0]);
                if (!d.element_set) return r.fail("unknown element set name", &L);
                if (L.typ[1] > 0) {                                    // ABAQUS UEL: element set, Un, value | history (:1327-1364)
                    if (!kw(L.tok[1], "U")) return r.fail("distributed load must specify a face number or Un", &L);
                    int first = m.elementsets[d.element_set - 1].items.empty() ? 0 : m.elementsets[d.element_set - 1].items[0];
                    if (first && m.element_flag[first - 1] < 99999)
                        return r.fail("an ABAQUS format boundary condition can only be applied to an ABAQUS element", &L);
                    size_t i = 1;
                    while (i < L.tok[1].size() && std::isdigit((unsigned char)L.tok[1][i])) ++i;
                    d.flag = to_int(L.tok[1].substr(1, i - 1));
                    if (i < L.tok[1].size() && L.tok[1][i] == 'N') d.flag = -d.flag;
                    d.uel_format = true;
                    if (L.typ[2] < 2) { d.values.push_back(to_real(L.tok[2])); d.history_number = 0; }
                    else {
                        d.history_number = find_name(m.histories, L.tok[2]);
                        if (!d.history_number) return r.fail("unknown history name", &L);
                        d.values.push_back(1.0);
                    }
                } else {                                               // native: set, face, VALUE|HISTORY|NORMAL, ... (:1366-1425)
                    if (L.n() < 4) return r.fail("expecting element set, face #, VALUE/HISTORY/NORMAL, values", &L);
                    d.face = to_int(L.tok[1]);
                    if (kw(L.tok[2], "VALUE")) {
                        d.flag = 1;
                        for (int k = 3; k < L.n(); ++k) d.values.push_back(to_real(L.tok[k]));
                    } else if (kw(L.tok[2], "HISTO")) {
                        d.flag = 2;
                        d.history_number = find_name(m.histories, L.tok[3]);
                        if (!d.history_number) return r.fail("unknown history name", &L);
                        for (int k = 4; k < L.n(); ++k) d.values.push_back(to_real(L.tok[k]));
                    } else if (kw(L.tok[2], "NORMA")) {
                        d.flag = 3;
                        d.history_number = find_name(m.histories, L.tok[3]);
                        if (!d.history_number) return r.fail("unknown history name", &L);
                    } else return r.fail("SUBROUTINE-type distributed loads are outside the supported subset", &L);
                }
                m.distributed_loads.push_back(d);
            }
        } else if (kw(L.tok[0], "SUBROUTINEP")) {                      // read_input_file.f90:1030-1065
            if (L.n() < 2) return r.fail("SUBROUTINE PARAMETERS key was used without specifying a name for the parameters", &L);
            SubroutineParameters sp;
            sp.name = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("SUBROUTINE PARAMETERS not terminated");
                if (kw(L.tok[0], "ENDSUBR")) break;
                for (int k = 0; k < L.n(); ++k) {
                    if (L.typ[k] == 2) return r.fail("SUBROUTINE PARAMETERS must be terminated with an END SUBROUTINE PARAMETERS", &L);
                    sp.values.push_back(to_real(L.tok[k]));
                }
            }
            m.subroutine_parameters.push_back(sp);
        } else if (kw(L.tok[0], "CONSTRAINTP")) {                      // read_input_file.f90:1067-1098 (tested before CONSTRA)
            if (L.n() < 2) return r.fail("CONSTRAINT PARAMETERS key was used without a name", &L);
            SubroutineParameters cp;
            cp.name = L.tok[1];
            while (true) {
                if (!r.next(L)) return r.fail("CONSTRAINT PARAMETERS not terminated");
                if (kw(L.tok[0], "ENDCONS")) break;
                for (int k = 0; k < L.n(); ++k) {
                    if (L.typ[k] == 2) return r.fail("expecting a list of parameters in a CONSTRAINT PARAMETERS block", &L);
                    cp.values.push_back(to_real(L.tok[k]));
                }
            }
            m.constraint_parameters.push_back(cp);
        } else if (kw(L.tok[0], "CONSTRA")) {                          // :1429-1541
            auto param_index = [&](const Line& Q) -> int {             // optional 7th field: name of a parameter block
                if (Q.n() == 7 && Q.typ[6] == 2)
                    for (size_t i = 0; i < m.constraint_parameters.size(); ++i)
                        if (m.constraint_parameters[i].name == Q.tok[6]) return (int)i + 1;
                return 0;
            };
            while (true) {
                if (!r.next(L)) return r.fail("CONSTRAINTS block was not terminated by an END CONSTRAINTS");
                if (kw(L.tok[0], "ENDCONS")) break;
                if (kw(L.tok[0], "CONSTRAINNODEPA")) {                 // CONSTRAIN NODE PAIRS, n, node1|set1, dof1, node2|set2, dof2
                    if (L.n() < 6) return r.fail("expecting n, node (set), dof, node (set), dof in CONSTRAIN NODE PAIRS", &L);
                    if (L.typ[2] == 0) {
                        Constraint c;
                        c.flag = 1;
                        c.node1 = to_int(L.tok[2]); c.dof1 = to_int(L.tok[3]); c.node2 = to_int(L.tok[4]); c.dof2 = to_int(L.tok[5]);
                        m.constraints.push_back(c);
                    } else {
                        const int s1 = find_name(m.nodesets, L.tok[2]), s2 = find_name(m.nodesets, L.tok[4]);
                        if (!s1 || !s2) return r.fail("unknown node set name in CONSTRAIN NODE PAIRS", &L);
                        const size_t k = (size_t)to_int(L.tok[1]);
                        if (m.nodesets[s1 - 1].items.size() != k || m.nodesets[s2 - 1].items.size() != k)
                            return r.fail("the node sets in a CONSTRAIN NODE PAIRS constraint do not have the correct number of nodes", &L);
                        for (size_t i = 0; i < k; ++i) {
                            Constraint c;
                            c.flag = 1;
                            c.node1 = m.nodesets[s1 - 1].items[i]; c.dof1 = to_int(L.tok[3]);
                            c.node2 = m.nodesets[s2 - 1].items[i]; c.dof2 = to_int(L.tok[5]);
                            c.index_parameters = param_index(L);
                            m.constraints.push_back(c);
                        }
                    }
                } else if (kw(L.tok[0], "TIENO")) {                    // TIE NODES, n, node set, dof1, master node, dof2
                    if (L.n() < 6) return r.fail("expecting n, node set, dof, master node, dof in TIE NODES", &L);
                    const int s1 = find_name(m.nodesets, L.tok[2]);
                    if (!s1) return r.fail("unknown node set name in TIE NODES", &L);
                    if (m.nodesets[s1 - 1].items.size() != (size_t)to_int(L.tok[1]))
                        return r.fail("the node set in a TIE NODES constraint does not have the correct number of nodes", &L);
                    for (int n : m.nodesets[s1 - 1].items) {
                        Constraint c;
                        c.flag = 2;
                        c.node1 = n; c.dof1 = to_int(L.tok[3]); c.node2 = to_int(L.tok[4]); c.dof2 = to_int(L.tok[5]);
                        c.index_parameters = param_index(L);
                        m.constraints.push_back(c);
                    }
                } else if (kw(L.tok[0], "CONSTRAINNODESET")) {
                    return r.fail("CONSTRAIN NODE SET calls user code (user_constraint): outside the supported subset; two-node ties "
                                  "(CONSTRAIN NODE PAIRS, TIE NODES) are supported", &L);
                } else return r.fail("CONSTRAINTS block was not terminated by an END CONSTRAINTS", &L);
            }
            for (const Constraint& c : m.constraints)
                if (c.node1 < 1 || c.node1 > m.n_nodes || c.node2 < 1 || c.node2 > m.n_nodes || c.dof1 < 1 || c.dof1 > 3 || c.dof2 < 1 ||
                    c.dof2 > 3) return r.fail("constraint: node or DOF out of range");
        } else if (kw(L.tok[0], "ENDBOUN")) return true;
        else return r.fail("unknown key in BOUNDARY CONDITIONS block", &L);
    }
    return r.fail("BOUNDARY CONDITIONS block not terminated");
}

// PRINT STATE block (read_input_file.f90:1772-1866 in a static step, :1987-2080 in an explicit dynamic step); L holds the
// PRINT STATE line
bool parse_print_state(Reader& r, Line& L) {
    Model& m = r.m;
    if (L.n() != 2 || L.typ[1] != 2) return r.fail("expecting a file name following PRINT STATE key", &L);
    m.stateprint = true;
    m.state_print_filename = L.tok[1];
    while (true) {
        if (!r.next(L)) return r.fail("PRINT STATE block not terminated");
        if (kw(L.tok[0], "ENDPRINTS")) return true;
        if (L.typ[0] < 2) return r.fail("unrecognized option in PRINT STATE", &L);
        if (kw(L.tok[0], "DEGREE")) m.print_dof = true;
        if (kw(L.tok[0], "DISPLACEDM")) m.print_displacedmesh = true;
        if (kw(L.tok[0], "DISPLACEMENTS")) {
            if (L.n() < 2 || L.typ[1] == 2) return r.fail("expecting value for displacement scale factor", &L);
            m.displacementscalefactor = to_real(L.tok[1]);
        }
        if (kw(L.tok[0], "FIELD")) {
            if (L.n() < 2) return r.fail("expecting names of field variables", &L);
            // names are stored from slot 1 on every FIELD VARIABLES line (field_variable_names(k-1), :1830-1833)
            for (int k = 1; k < L.n(); ++k) {
                if ((int)m.field_variable_names.size() < k) m.field_variable_names.resize(k);
                m.field_variable_names[k - 1] = L.tok[k];
            }
        }
        if (kw(L.tok[0], "ZONES")) {                           // single-zone meshes: the list is read and ignored
            if (!skip_block(r, "ENDZ")) return false;
        }
    }
}

// EXPLICIT DYNAMIC STEP block (read_input_file.f90:1920-2103)
bool parse_explicit_dynamic_step(Reader& r) {
    Model& m = r.m;
    Line L;
    m.explicitdynamicstep = true;
    m.use_lumped_projection_matrix = true;                             // :1922
    while (r.next(L)) {
        if (kw(L.tok[0], "ENDEXP")) {
            if (m.dynamic_timestep == 0.0) return r.fail("a value must be provided for the TIME STEP in an explicit dynamic analysis");
            if (m.no_dynamic_steps == 0) return r.fail("a value must be provided for the NUMBER OF STEPS in an explicit dynamic analysis");
            for (int f : m.element_flag)
                if (f == 10003) return r.fail("explicit dynamics with built-in continuum elements (VUMAT) is outside the supported subset");
            if (!m.abaqusformat)
                for (double d : m.element_density)
                    if (d == 0.0) return r.fail("densities must be specified for an explicit dynamic analysis");
            return true;
        } else if (kw(L.tok[0], "TIMESTEP")) {
            if (L.n() != 2 || L.typ[1] > 1) return r.fail("expecting value for a TIME STEP", &L);
            m.dynamic_timestep = to_real(L.tok[1]);
        } else if (kw(L.tok[0], "NUMB")) {
            if (L.n() != 2 || L.typ[1] != 0) return r.fail("expecting value for NUMBER OF STEPS", &L);
            m.no_dynamic_steps = to_int(L.tok[1]);
        } else if (kw(L.tok[0], "STATEPRINTS")) {
            if (L.n() != 2 || L.typ[1] != 0) return r.fail("expecting value for STATE PRINT STEPS", &L);
            m.state_print_steps = to_int(L.tok[1]);
        } else if (kw(L.tok[0], "USERPRINTST")) {
        } else if (kw(L.tok[0], "PRINTST")) { if (!parse_print_state(r, L)) return false; }
        else if (kw(L.tok[0], "USERPRINTPA") || kw(L.tok[0], "USERPRINTFI")) { if (!skip_block(r, "ENDU")) return false; }
        else return r.fail("unknown key in EXPLICIT DYNAMIC STEP block", &L);
    }
    return r.fail("EXPLICIT DYNAMIC STEP block not terminated");
}

bool parse_static_step(Reader& r) {
    Model& m = r.m;
    Line L;
    m.staticstep = true;
    while (r.next(L)) {
        if (kw(L.tok[0], "ENDSTA")) return true;
        else if (kw(L.tok[0], "INITIALTIMESTEP")) {                    // :1661-1672
            if (L.n() != 2 || L.typ[1] > 1) return r.fail("expecting value for an INITIAL TIME STEP", &L);
            m.timestep_initial = to_real(L.tok[1]);
            if (m.timestep_max == 0.0) m.timestep_max = m.timestep_initial;
            if (m.timestep_min == 0.0) m.timestep_min = m.timestep_initial;
        } else if (kw(L.tok[0], "MAXTIMESTEP")) m.timestep_max = to_real(L.tok[1]);
        else if (kw(L.tok[0], "MINTIMESTEP")) m.timestep_min = to_real(L.tok[1]);
        else if (kw(L.tok[0], "MAXN")) {                               // :1693-1703
            m.max_no_steps = to_int(L.tok[1]);
            if (m.max_total_time == 0.0) m.max_total_time = m.timestep_max * m.max_no_steps;
        } else if (kw(L.tok[0], "STOPT")) {                            // :1704-1714
            m.max_total_time = to_real(L.tok[1]);
            if (m.max_no_steps == 0) m.max_no_steps = (int)(m.max_total_time / m.timestep_min);
        } else if (kw(L.tok[0], "STATEPRINTS")) {                      // :1715-1726
            if (L.n() != 2 || L.typ[1] != 0) return r.fail("expecting value for STATE PRINT STEPS", &L);
            m.state_print_steps = to_int(L.tok[1]);
        } else if (kw(L.tok[0], "USERPRINTST")) {
        } else if (kw(L.tok[0], "SOLVER")) {                           // :1735-1771
            if (L.n() < 3) return r.fail("SOLVER needs a type and LINEAR/NONLINEAR", &L);
            m.unsymmetric_stiffness = false;
            m.solvertype = 1;
            m.nonlinear = false;
            if (kw(L.tok[1], "DIREC")) m.solvertype = 1;
            else if (kw(L.tok[1], "CONJU")) m.solvertype = 2;
            else return r.fail("unknown solver type", &L);
            if (kw(L.tok[2], "LINEAR")) { m.nonlinear = false; m.max_newton_iterations = 1; m.newtonraphson_tolerance = 1.e-08; }
            else if (kw(L.tok[2], "NONLINEAR")) {
                m.nonlinear = true;
                if (L.n() >= 5 && L.typ[3] < 2 && L.typ[4] == 0) {
                    m.newtonraphson_tolerance = to_real(L.tok[3]);
                    m.max_newton_iterations = to_int(L.tok[4]);
                } else return r.fail("expecting convergence tolerance and max # iterations following NONLINEAR", &L);
            } else return r.fail("unknown solver option", &L);
            if (kw(L.tok[L.n() - 1], "UNSYMM")) m.unsymmetric_stiffness = true;
        } else if (kw(L.tok[0], "CGTOL")) {                            // extension: run-time cg_tolerance (Stiffness.f90:35)
            m.cg_tolerance = to_real(L.tok[1]);
        } else if (kw(L.tok[0], "CGMAXIT")) {                          // extension: run-time max_cg_iterations (Stiffness.f90:37)
            m.max_cg_iterations = to_int(L.tok[1]);
        } else if (kw(L.tok[0], "PRINTST")) { if (!parse_print_state(r, L)) return false; }
        else if (kw(L.tok[0], "USERPRINTPA") || kw(L.tok[0], "USERPRINTFI")) { if (!skip_block(r, "ENDU")) return false; }
        else return r.fail("unknown key in STATIC STEP block", &L);
    }
    return r.fail("STATIC STEP block not terminated");
}

bool parse_stream(std::istream& in, Model& m) {
    Reader r(in, m);
    Line L;
    while (r.next(L)) {
        if (kw(L.tok[0], "MESH")) { if (!parse_mesh(r)) return false; }
        else if (kw(L.tok[0], "BOUNDARYCON")) { if (!parse_boundary(r)) return false; }
        else if (kw(L.tok[0], "TIME") || kw(L.tok[0], "PRINTINI") || kw(L.tok[0], "CHECKMATER")) {}
        else if (kw(L.tok[0], "CHECKSTIFF")) {                         // :1613-1648
            if (L.n() > 1) {
                if (L.typ[1] == 0) m.checkstiffness_flag = to_int(L.tok[1]);
                else if (kw(L.tok[1], "U")) m.checkstiffness_flag = to_int(L.tok[1].substr(1)) + 99999;
            }
        } else if (kw(L.tok[0], "STATICSTEP")) { if (!parse_static_step(r)) return false; }
        else if (kw(L.tok[0], "EXPLICITDYNAMICSTEP")) { if (!parse_explicit_dynamic_step(r)) return false; }
        else if (kw(L.tok[0], "STOP")) break;
        else return r.fail("unknown top-level key", &L);
    }
    if (m.n_nodes == 0 || m.n_elements == 0) return r.fail("no mesh was defined");
    m.element_density.resize(m.n_elements, 0.0);
    for (int c : m.connectivity)
        if (c < 1 || c > m.n_nodes) return r.fail("connectivity refers to a node that does not exist");
    // the face tables of the load routines are the hex8 ones (Element_Utilities.f90:1102-1109)
    if (m.nodes_per_element != 8 && !m.distributed_loads.empty())
        return r.fail("DISTRIBUTED LOADS are implemented for 8-node elements only (use FORCES / DEGREES OF FREEDOM on 4-, 10- and 20-node meshes)");
    if (m.nodes_per_element != 8 && (m.stateprint || m.explicitdynamicstep))
        return r.fail("PRINT STATE and EXPLICIT DYNAMIC STEP are implemented for 8-node elements only");
    return true;
}

}  // namespace

bool read_input_file(const std::string& path, Model& m) {
    std::ifstream f(path);
    if (!f) { m.error = "cannot open input file " + path; return false; }
    return parse_stream(f, m);
}
bool read_input_text(const std::string& text, Model& m) {
    std::istringstream f(text);
    return parse_stream(f, m);
}

// modules/Boundaryconditions.f90:161-221
double interpolate_history_table(const History& h, double time) {
    const int n = (int)h.t.size();
    if (n == 1) return h.v[0];
    if (time <= h.t.front()) return h.v.front();
    if (time >= h.t.back()) return h.v.back();
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int k = (hi + lo + 2) / 2 - 1;
        if (h.t[k] > time) hi = k; else lo = k;
    }
    if (h.t[hi] == h.t[lo]) return 0.5 * (h.v[hi] + h.v[lo]);
    return ((h.t[hi] - time) * h.v[lo] + (time - h.t[lo]) * h.v[hi]) / (h.t[hi] - h.t[lo]);
}

}  // namespace en234
