// ORACLE — test infrastructure only (see or_model.h).  SURVEY 8(f) rank-1 rows: C3D8 B-bar continuum
// element + UMAT and Gibbs-Poole-Stockmeyer renumbering (needed only to reproduce the golden log
// Output_Files/Abaqus_umat_holeplate_3d.out).
#include <cstdio>
#include <cstring>

#include "or_model.h"

namespace oracle {

void c3d8_bbar_static(const double*, const double*, const double*, const double*, int, double, double* K,
                      double* R, double*, int) {
    std::fprintf(stderr, "oracle: C3D8 B-bar element not restated yet\n");
    std::memset(K, 0, sizeof(double) * 576);
    std::memset(R, 0, sizeof(double) * 24);
}

void gps_node_order(const Model& m, std::vector<int>& node_order) {
    node_order.resize(m.n_nodes);
    for (int i = 0; i < m.n_nodes; ++i) node_order[i] = i;
}

}  // namespace oracle
