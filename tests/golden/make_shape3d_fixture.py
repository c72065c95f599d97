"""Golden vectors for the 3-D shape-function table of the reference, generated FROM THE FORTRAN TEXT ITSELF:
the assignments f(k) = ..., df(k,j) = ... of calculate_shapefunctions (user_codes/modules/Element_Utilities.f90:863-1025)
are read as text, translated token by token to Python (xi(i) -> xi, **2. -> **2.0, Fortran real literals) and evaluated in
IEEE double at fixed points.  Run here (the reference is not present on the GPU box):
    python tests/golden/make_shape3d_fixture.py     ->  tests/golden/shape3d_reference.npz
The table contains entries that are not the derivatives of the reference's own shape functions (tet10 df(10,1); hex20
df(11,1), df(11,2), df(15,1)); they are part of the reference's behaviour and therefore of the fixture."""
import os
import re

import numpy as np

REF = "/root/reference/EN234_FEA/user_codes/modules/Element_Utilities.f90"
HERE = os.path.dirname(os.path.abspath(__file__))


def table_blocks():
    src = open(REF).read().splitlines()
    i0 = next(i for i, l in enumerate(src) if "size(df)==60" in l)
    i1 = next(i for i, l in enumerate(src) if "end subroutine calculate_shapefunctions" in l)
    blocks, cur = {}, None
    for l in src[i0:i1]:
        m = re.search(r"n_nodes == (\d+)", l)
        if m:
            cur = int(m.group(1))
            blocks[cur] = []
            continue
        if cur and "=" in l and not l.strip().startswith("!"):
            blocks[cur].append(l.strip())
    return blocks


def evaluate(blocks, nn, xi):
    env = {"x1": float(xi[0]), "x2": float(xi[1]), "x3": float(xi[2])}
    f, df = np.zeros(nn), np.zeros((nn, 3))
    for l in blocks[nn]:
        lhs, rhs = l.split("=", 1)
        rhs = rhs.split("!")[0]
        rhs = re.sub(r"xi\((\d)\)", r"x\1", rhs).replace("**2.", "**2.0")
        rhs = re.sub(r"(\d)\.(?!\d)", r"\1.0", rhs)
        val = eval(rhs, {}, env)
        lhs = lhs.strip()
        if lhs == "xi4":
            env["xi4"] = val
        elif lhs.startswith("df("):
            a, j = map(int, re.findall(r"\d+", lhs))
            df[a - 1, j - 1] = val
        elif lhs.startswith("f("):
            f[int(re.findall(r"\d+", lhs)[0]) - 1] = val
    return f, df


def points(nn, n=12):
    rng = np.random.default_rng(100 + nn)
    return rng.uniform(-1, 1, (n, 3)) if nn in (8, 20) else rng.uniform(0.02, 0.3, (n, 3))


if __name__ == "__main__":
    blocks = table_blocks()
    out = {}
    for nn in (4, 10, 8, 20):
        P = points(nn)
        out["xi_%d" % nn] = P
        vals = [evaluate(blocks, nn, p) for p in P]
        out["f_%d" % nn] = np.array([v[0] for v in vals])
        out["df_%d" % nn] = np.array([v[1] for v in vals])
    np.savez(os.path.join(HERE, "shape3d_reference.npz"), **out)
    print("wrote shape3d_reference.npz")
