// ORACLE — test infrastructure only (see or_model.h).
// Synthetic structured hex8 block of BASELINE.json (SURVEY.md 8d "cube(n, jitter, seed)"), written from the specification
// of the MESH / USER SUBROUTINE hook (the reference's hook is user_codes/src/user_mesh.f90:1-169: user code fills the Mesh
// module arrays from a parameter list) independently of the product's generator (csrc/host/user_mesh.cpp), so that the
// reference arm of bench.py and the parity gates do not go through the product's parser:
//   node id  = 1 + i + (nx+1) j + (nx+1)(ny+1) k at (i, j, k) h with h = 1/nx; nodes strictly inside the block are moved
//              by jitter * h * U(-1,1) per coordinate, drawn x, y, z in node order from std::mt19937_64(seed) as
//              2 * (r >> 11) / 2^53 - 1
//   element  = 1 + i + nx j + nx ny k, ABAQUS C3D8 node order (Element_Utilities.f90:908-916)
#include <cstdint>
#include <random>

extern "C" int or_synthetic_block(int nx, int ny, int nz, double jitter, unsigned long long seed, double* coords /*3N*/,
                                  int* connectivity /*8E, 1-based*/) {
    if (nx < 1 || ny < 1 || nz < 1) return 1;
    std::mt19937_64 gen(seed);
    const double h = 1.0 / nx;
    const long long sx = 1, sy = nx + 1, sz = (long long)(nx + 1) * (ny + 1);
    long long n = 0;
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i, ++n) {
                double p[3] = {i * h, j * h, k * h};
                const bool inside = i != 0 && i != nx && j != 0 && j != ny && k != 0 && k != nz;
                if (inside && jitter != 0.0)
                    for (int c = 0; c < 3; ++c) {
                        const double u = 2.0 * ((double)(gen() >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
                        p[c] += jitter * h * u;
                    }
                for (int c = 0; c < 3; ++c) coords[3 * n + c] = p[c];
            }
    long long e = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i, ++e) {
                const long long base = 1 + i * sx + j * sy + k * sz;
                const long long corner[8] = {0, sx, sx + sy, sy, sz, sz + sx, sz + sx + sy, sz + sy};
                for (int a = 0; a < 8; ++a) connectivity[8 * e + a] = (int)(base + corner[a]);
            }
    return 0;
}
