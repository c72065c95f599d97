"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bars (BASELINE.json north_star): displacements and reaction forces within 1e-10 relative, same Newton history;
K_e within 1e-13 relative (Frobenius); index structures exact."""
import gzip
import os

import numpy as np
import pytest

import oracle_binding as ob
from conftest import GOLDEN
from en234_fea_b200 import GpuSystem, Model, synthetic_block_input

pytestmark = pytest.mark.gpu

TOL_U = 1e-10      # north_star: nodal displacements and reaction forces within 1e-10 relative
TOL_KE = 1e-13     # every K_e vs oracle, relative Frobenius


def golden_model(name):
    p = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        return Model.from_text(gzip.open(p, "rb").read().decode())
    return Model.read(p)


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def cube8():
    m = Model.from_text(synthetic_block_input(8, jitter=0.2, cg_tolerance=1e-26, cg_maxit=5000))
    g = GpuSystem(m)
    om = ob.OracleModel(m.arrays())
    yield m, g, om
    g.close()


def test_element_batch_matches_oracle(cube8):
    m, g, om = cube8
    a = m.arrays()
    rng = np.random.default_rng(0)
    ut, du = 1e-2 * rng.standard_normal(g.ndof), 1e-3 * rng.standard_normal(g.ndof)
    els = np.arange(1, m.n_elements + 1, 7)
    K, R, fail = g.element_batch(els, ut, du)
    assert fail == 0
    X = a["coords"].reshape(-1, 3)
    conn = a["connectivity"].reshape(-1, 8) - 1
    worst = 0.0
    for q, e in enumerate(els - 1):
        nodes = conn[e]
        dofs = (3 * nodes[:, None] + np.arange(3)).ravel()
        Ko, Ro, _, _, _ = ob.element(1001, X[nodes].ravel(), du[dofs], ut[dofs], [100.0, 0.3])
        worst = max(worst, np.linalg.norm(K[q] - Ko) / np.linalg.norm(Ko))
        assert np.abs(R[q] - Ro).max() <= 1e-13 * np.abs(Ro).max()
        assert np.abs(K[q] - K[q].T).max() <= 1e-14 * np.abs(K[q]).max()
    assert worst < TOL_KE, worst


def test_assembly_and_bc_match_oracle(cube8):
    m, g, om = cube8
    rng = np.random.default_rng(1)
    ut, du = 1e-2 * rng.standard_normal(g.ndof), 1e-3 * rng.standard_normal(g.ndof)
    s = ob.OracleSystem(om, 2)
    rf_o, nrm_o, fail_o = s.assemble(ut, du, 0.0, 1.0)
    A_o, rhs_o = s.matrix()
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0, ut, du)
    rf, nrm, fail = g.assemble(ut, du, None)
    assert fail == 0 and fail_o == 0
    A = g.csr()
    assert abs(A - A_o).max() <= 1e-13 * abs(A_o).max()
    assert abs(A - A.T).max() <= 1e-14 * abs(A_o).max()
    assert relmax(rf, rf_o) < 1e-13 and abs(nrm - nrm_o) <= 1e-13 * nrm_o
    assert relmax(g.rhs(), rhs_o) < 1e-13
    # boundary conditions: fixdof semantics on matrix and right-hand side
    unb_o = s.apply_bc(ut, du, 0.0, 1.0)
    A_o2, rhs_o2 = s.matrix()
    unb = g.apply_boundaryconditions(fx, fd, fc)
    A2 = g.csr()
    rhs2, diag2 = g.rhs(want_diag=True)
    assert abs(A2 - A_o2).max() <= 1e-13 * abs(A_o).max()
    assert relmax(rhs2, rhs_o2) < 1e-12 and abs(unb - unb_o) <= 1e-12 * unb_o
    assert np.allclose(diag2, A_o2.diagonal(), rtol=1e-13)
    assert np.all(A2.diagonal()[fd] == 1.0) and np.all(rhs2[fd] == fc)
    # NaN in the state -> fail flag, like the NaN scan of solver_direct.f90:402-415
    bad = ut.copy(); bad[10] = np.nan
    _, _, fail = g.assemble(bad, du, None)
    assert fail == 1


def test_assembly_is_bitwise_deterministic(cube8):
    m, g, om = cube8
    rng = np.random.default_rng(2)
    ut, du = rng.standard_normal(g.ndof), rng.standard_normal(g.ndof)
    g.assemble(ut, du, None)
    v1, r1 = g.csr().data.copy(), g.rhs().copy()
    g.assemble(ut, du, None)
    assert np.array_equal(v1, g.csr().data) and np.array_equal(r1, g.rhs())


def test_static_step_matches_oracle_direct_and_cg(cube8):
    m, g, om = cube8
    a = m.arrays()
    ref_direct = ob.OracleModel(a, solvertype=1).run_static()          # skyline LDU: exact linear solve
    ref_cg = ob.OracleModel(a, solvertype=2, cg_check_linear=1).run_static()
    res = g.run_static(method=0, check_linear_cg=True)
    assert relmax(res["dof_total"], ref_direct["dof_total"]) < TOL_U
    assert relmax(res["rforce"], ref_direct["rforce"]) < TOL_U
    # same CG iteration count as the reference recurrences on the CPU (+-2), same Newton record
    assert abs(int(res["cg_iterations"][0]) - int(ref_cg["cg_iterations"][0])) <= 2
    assert res["newton"].shape == ref_cg["newton"].shape
    assert np.array_equal(res["newton"][:, [0, 1, 7]], ref_cg["newton"][:, [0, 1, 7]])
    assert np.allclose(res["newton"][:, 2:4], ref_cg["newton"][:, 2:4], rtol=1e-9)
    # host-buffer entry points give the same answer as the device-resident ones, bit for bit
    res_h = g.run_static(method=0, check_linear_cg=True, use_host_api=True)
    assert np.array_equal(res_h["dof_total"], res["dof_total"])
    # and the solve is run-to-run deterministic
    res2 = g.run_static(method=0, check_linear_cg=True)
    assert np.array_equal(res2["dof_total"], res["dof_total"]) and np.array_equal(res2["cg_iterations"], res["cg_iterations"])


def test_cg_stopping_rule_and_failure(cube8):
    m, g, om = cube8
    n = g.ndof
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
    s = ob.OracleSystem(om, 2)
    for tol in (1e-17, 1e-8):
        g.assemble(np.zeros(n), np.zeros(n), None)
        g.apply_boundaryconditions(fx, fd, fc)
        du, corr, its, fail = g.solve(np.zeros(n), 0, tol, 1000)
        s.assemble(np.zeros(n), np.zeros(n)); s.apply_bc(np.zeros(n), np.zeros(n))
        du_o, corr_o, its_o, fail_o = s.solve(np.zeros(n), tol, 1000)
        assert fail == 0 and fail_o == 0 and abs(its - its_o) <= 2
        assert g.stats().last_cg_err < tol
        assert relmax(du, du_o) < 1e-6 if tol > 1e-10 else relmax(du, du_o) < 1e-7
    # iteration cap -> fail, increment untouched (solver_conjugate_gradient.f90:854-860, :745)
    g.assemble(np.zeros(n), np.zeros(n), None)
    g.apply_boundaryconditions(fx, fd, fc)
    du, corr, its, fail = g.solve(np.ones(n), 0, 1e-17, 5)
    assert fail == 1 and np.all(du == 1.0)
    # zero right-hand side -> zero solution, no iterations (:805-814)
    g.assemble(np.zeros(n), np.zeros(n), None)
    g.apply_boundaryconditions(None, fd, np.zeros(fd.size))
    du, corr, its, fail = g.solve(np.zeros(n), 0, 1e-17, 100)
    assert fail == 0 and its == 0 and np.all(du == 0) and corr == 0


def test_config1_reference_input_analytic_and_oracle():
    m = golden_model("linear_elastic_3d_uel.in")
    g = GpuSystem(m)
    res = g.run_static(method=0)
    a = m.arrays()
    X = a["coords"].reshape(-1, 3)
    exact = np.stack([0.02 * X[:, 0], -0.006 * X[:, 1], -0.006 * X[:, 2]], 1).ravel()
    assert np.abs(res["dof_total"] - exact).max() < 1e-14
    gold = np.load(os.path.join(GOLDEN, "linear_elastic_3d_uel_oracle_output.npz"))
    assert relmax(res["dof_total"], gold["dof_total"]) < TOL_U
    assert relmax(res["rforce"], gold["rforce"]) < TOL_U
    assert res["n_steps"] == 3 and np.array_equal(res["newton"][:, [0, 1, 7]], gold["newton"][:, [0, 1, 7]])
    assert np.allclose(res["newton"][:, 3], gold["newton"][:, 3], rtol=1e-12)        # generalized force norms
    g.close()


def test_holeplate_reference_input_matches_golden_direct_solution():
    # config 1b: 3075 unstructured hex8 (SOLVER, DIRECT, LINEAR; 3 steps with a displacement ramp)
    m = golden_model("holeplate_3d_uel.in.gz")
    g = GpuSystem(m)
    res = g.run_static(method=0)
    gold = np.load(os.path.join(GOLDEN, "holeplate_3d_uel_oracle_output.npz"))
    assert res["n_steps"] == 3
    assert relmax(res["dof_total"], gold["dof_total"]) < TOL_U
    assert relmax(res["rforce"], gold["rforce"]) < TOL_U
    assert np.array_equal(res["newton"][:, [0, 1, 7]], gold["newton"][:, [0, 1, 7]])
    assert np.allclose(res["newton"][:, 3], gold["newton"][:, 3], rtol=1e-10)
    assert np.allclose(res["newton"][:, 2], gold["newton"][:, 2], rtol=1e-6, atol=1e-10)
    # Gauss-point stresses (SVARS) against the oracle's UEL on a few elements
    a = m.arrays()
    sv = g.svars(res["dof_total"], np.zeros(g.ndof), m.n_elements)
    X = a["coords"].reshape(-1, 3); conn = a["connectivity"].reshape(-1, 8) - 1
    for e in (0, 1500, 3074):
        dofs = (3 * conn[e][:, None] + np.arange(3)).ravel()
        _, _, svo, _, _ = ob.element(100000, X[conn[e]].ravel(), np.zeros(24), res["dof_total"][dofs], [100.0, 0.3])
        assert np.abs(sv[e] - svo).max() <= 1e-11 * np.abs(svo).max()
    g.close()


def test_matrix_free_operator_and_solve(cube8):
    m, g, om = cube8
    n = g.ndof
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    g.set_operator_mode(0)
    g.assemble(np.zeros(n), np.zeros(n), None)
    A = g.csr()
    y_bsr = g.apply_operator(x, 0)
    assert relmax(y_bsr, A @ x) < 1e-13
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
    g.apply_boundaryconditions(None, np.zeros(0, np.int32), np.zeros(0))     # clears the prescribed-DOF mask
    y_mf = g.apply_operator(x, 1)
    assert relmax(y_mf, A @ x) < 1e-12
    res_mf = g.run_static(method=1, check_linear_cg=True)
    res = g.run_static(method=0, check_linear_cg=True)
    assert relmax(res_mf["dof_total"], res["dof_total"]) < 1e-9
    assert abs(int(res_mf["cg_iterations"][0]) - int(res["cg_iterations"][0])) <= 2
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    assert relmax(res_mf["dof_total"], ref["dof_total"]) < TOL_U


def test_prescribed_displacement_ramp_multi_step():
    # prescribed u_x ramp on x = max over 3 increments (HISTORY-type DOF, src/solver_direct.f90:690-694)
    m = Model.from_text(synthetic_block_input(5, jitter=0.15, load="stretch", steps=3, cg_tolerance=1e-26, cg_maxit=5000))
    g = GpuSystem(m)
    res = g.run_static(method=0, check_linear_cg=True)
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    assert res["n_steps"] == ref["n_steps"] == 3
    assert relmax(res["dof_total"], ref["dof_total"]) < TOL_U
    assert relmax(res["rforce"], ref["rforce"]) < TOL_U
    g.close()


def test_neohooke_element_and_newton_history():
    # BASELINE configs[3] in small: neo-Hookean user element (mu = 1, K = 200 as input_files/Abaqus_uel_hyperelastic.in:44),
    # prescribed stretch ramp over 4 increments, NONLINEAR 1e-5 / 15 with full re-assembly every Newton iteration.
    # The reference ships no hyperelastic element: parity is against the repo's oracle twin, which is pinned by the FD
    # stiffness check and patch tests in test_oracle_cpu.py (parity unpinned by reference outputs).
    txt = synthetic_block_input(5, jitter=0.15, element="U2", props=[1.0, 200.0], load="stretch", steps=4,
                                solver="DIRECT", nonlinear=(1e-5, 15))
    m = Model.from_text(txt)
    a = m.arrays()
    assert a["nonlinear"][0] == 1 and a["element_flag"][0] == 100001
    g = GpuSystem(m)
    rng = np.random.default_rng(11)
    ut, du = 2e-2 * rng.standard_normal(g.ndof), 5e-3 * rng.standard_normal(g.ndof)
    els = np.arange(1, m.n_elements + 1, 9)
    K, R, fail = g.element_batch(els, ut, du)
    assert fail == 0
    X = a["coords"].reshape(-1, 3); conn = a["connectivity"].reshape(-1, 8) - 1
    for q, e in enumerate(els - 1):
        dofs = (3 * conn[e][:, None] + np.arange(3)).ravel()
        Ko, Ro, _, _, f = ob.element(1002, X[conn[e]].ravel(), du[dofs], ut[dofs], [1.0, 200.0])
        assert f == 0
        assert np.linalg.norm(K[q] - Ko) / np.linalg.norm(Ko) < 1e-12
        assert np.abs(R[q] - Ro).max() <= 1e-12 * np.abs(Ro).max()
    # assembled tangent and residual
    s = ob.OracleSystem(ob.OracleModel(a), 2)
    rf_o, nrm_o, _ = s.assemble(ut, du)
    A_o, _ = s.matrix()
    rf, nrm, fail = g.assemble(ut, du, None)
    assert fail == 0 and abs(g.csr() - A_o).max() <= 1e-12 * abs(A_o).max() and relmax(rf, rf_o) < 1e-12
    # whole analysis: same Newton convergence history, same displacements / reactions
    res = g.run_static(method=0)
    ref = ob.OracleModel(a).run_static()
    assert res["n_steps"] == ref["n_steps"] == 4 and res["n_cutbacks"] == ref["n_cutbacks"] == 0
    assert res["newton"].shape == ref["newton"].shape and res["newton"].shape[0] >= 8
    assert np.array_equal(res["newton"][:, [0, 1, 7]], ref["newton"][:, [0, 1, 7]])
    big = ref["newton"][:, 5] > 1e-7                       # ratios well above the linear-solver noise floor
    assert np.allclose(res["newton"][big][:, 2:6], ref["newton"][big][:, 2:6], rtol=1e-6)
    assert relmax(res["dof_total"], ref["dof_total"]) < TOL_U
    assert relmax(res["rforce"], ref["rforce"]) < 1e-8
    # an inverted element makes the assembly report fail (-> time step cutback in the driver)
    bad = np.zeros(g.ndof); bad[0::3] = -3.0 * X[:, 0]
    _, _, fail = g.assemble(bad, np.zeros(g.ndof), None)
    assert fail == 1
    g.close()


def test_full_size_properties_64cubed():
    # BASELINE configs[1] at full size: properties that need no CPU solve — operator symmetry, residual of the
    # returned solution evaluated with an independent operator (matrix-free vs assembled), determinism
    m = Model.from_text(synthetic_block_input(64, jitter=0.2, cg_tolerance=1e-20, cg_maxit=5000))
    g = GpuSystem(m)
    n = g.ndof
    assert n == 823875 and g.nnz_blocks() * 9 == 64701513              # SURVEY section 8 size table
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
    _, nrm, fail = g.assemble(np.zeros(n), np.zeros(n), None, want_rforce=False)
    assert fail == 0 and nrm == 0.0
    g.apply_boundaryconditions(fx, fd, fc)
    b = g.rhs()
    rng = np.random.default_rng(9)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Ax, Ay = g.apply_operator(x, 0), g.apply_operator(y, 0)
    assert abs(y @ Ax - x @ Ay) <= 1e-12 * abs(y @ Ax)
    du, corr, its, fail = g.solve(np.zeros(n), 0, 1e-20, 5000)
    assert fail == 0 and 300 < its < 1500
    r = b - g.apply_operator(du, 0)
    assert np.linalg.norm(r) / np.linalg.norm(b) < 2e-10
    r_mf = b - g.apply_operator(du, 1)          # matrix-free operator honours the same prescribed-DOF mask
    assert np.linalg.norm(r_mf) / np.linalg.norm(b) < 2e-10
    du2, _, its2, _ = g.solve(np.zeros(n), 0, 1e-20, 5000)
    assert its2 == its and np.array_equal(du, du2)
    assert np.all(du[fd] == 0.0)
    g.close()


def test_fortran_linkage_plugin_twins_match_oracle():
    """user_element_static_ / uel_ (include/en234host.h) called the way a gfortran driver calls them
    (user_codes/src/user_element.f90:5-49, abaqus_uel_linelastic_3d.for:16-28, call site solver_direct.f90:231-240)."""
    from en234_fea_b200 import uel, user_element_static
    rng = np.random.default_rng(5)
    ref = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    for trial in range(3):
        X = (ref * [1.0, 0.7, 1.3] + 0.1 * rng.standard_normal((8, 3))).ravel()
        ut, du = 1e-2 * rng.standard_normal(24), 1e-3 * rng.standard_normal(24)
        for ident, props in ((1001, [100.0, 0.3]), (1002, [1.0, 200.0])):
            K, R, sv, fail = user_element_static(ident, X, du, ut, props, initial_state_variables=[1.0, 2.0])
            Ko, Ro, _, _, _ = ob.element(ident, X, du, ut, props)
            assert fail == 0 and np.array_equal(sv, [1.0, 2.0])           # state copied (user_element.f90:60)
            assert np.linalg.norm(K - Ko) <= TOL_KE * np.linalg.norm(Ko)
            assert np.abs(R - Ro).max() <= 1e-13 * np.abs(Ro).max()
    # UEL on a box-shaped element (flat rectangular faces: SURVEY Q2) with a pressure on face 2 and on face 4
    X = (ref * [1.0, 0.7, 1.3]).ravel()
    u, du = 1e-2 * rng.standard_normal(24), 1e-3 * rng.standard_normal(24)
    A, RHS, SV, EN, pnewdt = uel(1, X, u, du, [100.0, 0.3], jdltyp=[2, 4], adlmag=[3.0, -1.5])
    Ko, Ro, svo, eno, _ = ob.element(100000, X, du, u - du, [100.0, 0.3], jdltyp=[2, 4], adlmag=[3.0, -1.5])
    assert pnewdt == 1.0
    assert np.linalg.norm(A - Ko) <= TOL_KE * np.linalg.norm(Ko)
    assert np.abs(RHS - Ro).max() <= 1e-13 * np.abs(Ro).max()
    assert np.abs(SV.reshape(8, 6) - svo).max() <= 1e-13 * np.abs(svo).max()
    assert abs(EN[1] - eno[1]) <= 1e-13 * abs(eno[1]) and EN[0] == 0.0
    # an element type the device path does not have: fail / PNEWDT < 1, never a process stop
    _, _, _, fail = user_element_static(77, X, du, u, [1.0, 0.3])
    assert fail == 1
    assert uel(1, X[:12], u[:12], du[:12], [100.0, 0.3])[4] < 1.0


def test_two_kernel_assembly_equals_row_gather_assembly(monkeypatch):
    """k_element_blocks + k_gather_rows (default) against the single-kernel row-gather assembly (EN234_ASM=rows): both sum
    every entry in ascending element order, so the matrices are bitwise equal; residuals agree to rounding."""
    for element, props in (("1001", None), ("1002", (1.0, 200.0))):
        kw = dict(jitter=0.2, element=element)
        if props:
            kw.update(props=list(props))
        m = Model.from_text(synthetic_block_input(5, 4, 3, **kw))
        rng = np.random.default_rng(3)
        ut, du = 1e-2 * rng.standard_normal(3 * m.n_nodes), 1e-3 * rng.standard_normal(3 * m.n_nodes)
        out = []
        for mode in ("", "rows"):
            monkeypatch.setenv("EN234_ASM", mode)
            g = GpuSystem(m)
            rf, nrm, fail = g.assemble(ut, du)
            assert fail == 0
            out.append((np.asarray(g.csr().data).copy(), rf.copy(), nrm))
            g.close()
        monkeypatch.delenv("EN234_ASM")
        assert np.array_equal(out[0][0], out[1][0])
        assert np.abs(out[0][1] - out[1][1]).max() <= 1e-14 * np.abs(out[1][1]).max()
        assert abs(out[0][2] - out[1][2]) <= 1e-14 * out[1][2]


def test_matrix_free_congruent_mesh_fast_path():
    """Uniform block with power-of-two spacing: every element is a bitwise translate of element 0, the matrix-free
    operator multiplies by the single K_e (k_mf_congruent).  Same operator as the assembled matrix, same solution."""
    m = Model.from_text(synthetic_block_input(8, 4, 16, jitter=0.0, cg_tolerance=1e-26, cg_maxit=5000))
    g = GpuSystem(m)
    assert g.stats().mf_congruent == 1
    n = g.ndof
    x = np.random.default_rng(7).standard_normal(n)
    g.assemble(np.zeros(n), np.zeros(n), None)
    A = g.csr()
    g.apply_boundaryconditions(None, np.zeros(0, np.int32), np.zeros(0))
    assert relmax(g.apply_operator(x, 1), A @ x) < 1e-12
    res_mf = g.run_static(method=1, check_linear_cg=True)
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    assert relmax(res_mf["dof_total"], ref["dof_total"]) < TOL_U
    assert relmax(res_mf["rforce"], ref["rforce"]) < TOL_U
    g.close()
    # a jittered mesh (and a non-dyadic spacing) must not take the fast path
    for kw in (dict(jitter=0.2), dict(jitter=0.0)):
        g2 = GpuSystem(Model.from_text(synthetic_block_input(3 if kw["jitter"] == 0.0 else 4, **kw)))
        assert g2.stats().mf_congruent == 0
        g2.close()


def test_golden_log_newton_history_on_the_gpu():
    """SURVEY 8(f) row 1 on the device: the reference's only golden output, Output_Files/Abaqus_umat_holeplate_3d.out
    (C3D8 B-bar + elastic UMAT, SOLVER, DIRECT, NONLINEAR 1e-5/15, UNSYMMETRIC, 3 steps).  The device element
    (k_element_c3d8), the ordered assembly, fixdof, Jacobi-BiCGSTAB and the host Newton/time-step loop must reproduce
    every printed Newton record to the 6 significant digits of the log (.out:26-31,42-47,58-63,91-96,124-129).
    One record sits at the linear-solver noise floor: iteration 3 of step 1 (ratio 1.6e-8) is the second-order remainder
    of a 1.2e-4 correction, so a 1e-11 relative difference between two solves of the FIRST system (cond x eps, which is
    what separates any two solvers here — the reference's skyline LDU included) moves its 6th digit; it gets 2 units of
    the last printed digit instead of 0.51."""
    m = golden_model("holeplate_3d_umat.in.gz")
    assert (m.n_elements, m.n_nodes) == (3075, 4104)
    g = GpuSystem(m)
    res = g.run_static(method=0)
    golden = [(1, 1, 0.390615e+00, 0.568430e+01, 0.240174e-01, 0.395353e-02),
              (1, 2, 0.124137e-03, 0.568438e+01, 0.601132e-04, 0.105749e-04),
              (1, 3, 0.273510e-06, 0.568438e+01, 0.914709e-07, 0.160916e-07),
              (2, 1, 0.322128e-03, 0.113544e+02, 0.109670e-03, 0.965856e-05),
              (3, 1, 0.320594e-03, 0.170101e+02, 0.122905e-03, 0.722529e-05)]
    assert res["n_steps"] == 3 and res["n_cutbacks"] == 0 and len(res["newton"]) == len(golden)
    for row, gold in zip(res["newton"], golden):
        assert (int(row[0]), int(row[1])) == gold[:2]
        for got, want in zip(row[2:6], gold[2:]):
            ulp = 10.0 ** (np.floor(np.log10(abs(want))) + 1 - 6)                     # last printed digit of 0.dddddd D+ee
            assert abs(got - want) <= (0.51 if gold[5] > 1e-7 else 2.0) * ulp, (row, gold)
    assert [int(r[7]) for r in res["newton"]] == [0, 0, 1, 1, 1]
    # and against the oracle's run of the same model (skyline LDU): displacements, reactions, state variables
    a = dict(np.load(os.path.join(GOLDEN, "holeplate_3d_umat_model.npz")))
    ref = ob.OracleModel(a, node_order=ob.OracleModel(a).gps_node_order()).run_static()
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-9
    assert relmax(res["rforce"], ref["rforce"]) < 1e-8
    sv = g.state_variables(updated=False)
    assert sv.size == 3075 * 120 and np.abs(sv).max() > 0
    g.close()


def test_edge_cases_isolated_node_zero_load_and_repeated_prescribed_dof():
    """A node no element uses (kept as an identity block row), a step without any load (zero right-hand side: x = 0 without
    iterating, solver_conjugate_gradient.f90:805-814), a prescribed DOF listed twice (the last value wins, as when the
    reference applies fixdof in list order)."""
    m = Model.from_text(synthetic_block_input(3, jitter=0.1))
    a = m.arrays()
    N = m.n_nodes
    coords = np.concatenate([a["coords"], [5.0, 5.0, 5.0]])               # node N+1 is not referenced by any element
    g = GpuSystem(arrays=dict(n_nodes=N + 1, n_elements=m.n_elements, coords=coords, connectivity=a["connectivity"],
                              element_flag=a["element_flag"], element_prop_index=a["element_prop_index"],
                              element_properties=a["element_properties"]))
    n = g.ndof
    rf, nrm, fail = g.assemble(np.zeros(n), np.zeros(n), None)
    assert fail == 0 and nrm == 0.0
    A = g.csr().toarray()
    assert np.all(A[3 * N:, :] == 0.0) and np.all(A[:, 3 * N:] == 0.0)
    fixed = np.array([0, 1, 2, 3 * N, 3 * N + 1, 3 * N + 2, 0], np.int32)   # DOF 0 twice
    corr = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.25])
    g.apply_boundaryconditions(None, fixed, corr)
    rhs = g.rhs()
    assert rhs[0] == 0.25                                                 # the later entry wins
    du, corr_norm, its, sfail = g.solve(np.zeros(n), 0, 1e-24, 5000)
    assert sfail == 0 and its > 0 and abs(du[0] - 0.25) < 1e-12 and np.all(du[3 * N:] == 0.0)
    # no load at all: zero rhs -> zero solution, zero iterations
    g.assemble(np.zeros(n), np.zeros(n), None)
    g.apply_boundaryconditions(None, fixed[:6], np.zeros(6))
    du, corr_norm, its, sfail = g.solve(np.zeros(n), 0, 1e-24, 5000)
    assert sfail == 0 and its == 0 and corr_norm == 0.0 and np.all(du == 0.0)
    g.close()


def test_check_stiffness_through_the_device_twins(tmp_path):
    """The reference's own element self-test (src/checkstiffness.f90, CHECK STIFFNESS key) run by the host driver on the
    device twins of the element routines: coded stiffness vs numerical derivative of the residual, criterion 1e-6."""
    rng = np.random.default_rng(17)
    for element, props in (("1001", None), ("U1", None), ("1002", [1.0, 200.0]), ("U2", [1.0, 200.0])):
        kw = dict(jitter=0.2, element=element)
        if props:
            kw["props"] = props
        m = Model.from_text(synthetic_block_input(2, **kw))
        flag = int(m.arrays()["element_flag"][0])
        ut, du = 2e-2 * rng.standard_normal(3 * m.n_nodes), 5e-3 * rng.standard_normal(3 * m.n_nodes)
        log = str(tmp_path / ("check_%s.log" % element))
        err, nu, nnu = m.check_stiffness(flag, ut, du, log_path=log)
        assert err < 1e-6 and nu == 0, (element, err, nu)
        text = open(log).read()
        assert "STIFFNESS CHECK" in text and "Stiffness matrix is consistent with residual" in text and text.count("Column") == 24
        # same figure as the oracle's restatement of the routine on the same element and state
        a = m.arrays()
        conn = a["connectivity"][:8] - 1
        dofs = (3 * conn[:, None] + np.arange(3)).ravel()
        eo, _ = ob.check_stiffness(flag if flag < 99999 else 100000 + (flag - 100000), a["coords"].reshape(-1, 3)[conn].ravel(),
                                   du[dofs], ut[dofs], props or [100.0, 0.3])
        assert abs(err - eo) <= 0.05 * max(err, eo) + 1e-16, (element, err, eo)


def test_c3d8_umat_two_element_reference_input():
    """input_files/Abaqus_umat_linear_elastic_3d.in (two C3D8 + elastic UMAT elements, uniaxial tension by a NORMAL
    distributed load ramped over 3 steps, NONLINEAR 1e-9 / 15, UNSYMMETRIC): device path against the oracle's skyline-LDU
    run, and against the small-strain analytic values it must approach (sigma = 6, E = 1e4, nu = 0.3)."""
    m = golden_model("linear_elastic_3d_umat.in")
    g = GpuSystem(m)
    res = g.run_static(method=0)
    ref = ob.OracleModel(m.arrays()).run_static()
    assert res["n_steps"] == ref["n_steps"] == 3 and res["n_cutbacks"] == ref["n_cutbacks"] == 0
    assert res["newton"].shape == ref["newton"].shape
    assert np.array_equal(res["newton"][:, [0, 1, 7]], ref["newton"][:, [0, 1, 7]])
    big = ref["newton"][:, 5] > 1e-7
    assert np.allclose(res["newton"][big][:, 2:6], ref["newton"][big][:, 2:6], rtol=1e-6)
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-9
    u = res["dof_total"].reshape(-1, 3)
    assert abs(u[8, 0] - 2 * 6e-4) < 2e-6 and abs(u[6, 1] + 1.8e-4) < 1e-6 and abs(u[6, 2] + 1.8e-4) < 1e-6
    g.close()


def test_c3d8_stretch_then_rigid_rotation_keeps_the_stress_state():
    """The scenario of input_files/Abaqus_umat_stretch_rotate.in (a two-element C3D8 bar stretched 1 % and then rotated
    rigidly by 90 degrees about axis 3 in 10 steps) with the end-face motion given as HISTORY tables instead of the
    reference's SUBROUTINE-type BCs: the Hughes-Winget rotation of the stored stress (continuum_elements.f90:591-640)
    must carry the axial stress from the 1-axis to the 2-axis.  Device path against the oracle, state variables included."""
    X = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1],
                  [2, 0, 0], [2, 1, 0], [2, 0, 1], [2, 1, 1]], float)
    ends = [1, 4, 5, 8, 9, 10, 12, 11]
    steps = 11
    lam = 1.01
    def position(n, t):                                   # t = 1: stretched; t = 2..11: rotated by 9 degrees per step
        x = X[n - 1] * [lam if t >= 1 else 1.0 + (lam - 1.0) * t, 1.0, 1.0]
        th = np.radians(9.0 * max(0.0, t - 1.0))
        c, s = np.cos(th), np.sin(th)
        return np.array([c * x[0] - s * x[1], s * x[0] + c * x[1], x[2]])
    t = ["MESH", " NODES", "  PARAMETERS, 3, 3, 1", "  DISPLACEMENT DOF, 1, 2, 3", "  COORDINATES"]
    t += ["   %d, %.17g, %.17g, %.17g" % (i + 1, *X[i]) for i in range(12)]
    t += ["  END COORDINATES", " END NODES", " MATERIAL, mat", "  STATE VARIABLES, 0", "  PROPERTIES", "   10000.d0, 0.3d0",
          "  END PROPERTIES", " END MATERIAL", " ELEMENTS, INTERNAL", "  TYPE, C3D8", "  PROPERTIES, mat", "  CONNECTIVITY, zone1",
          "   1, 1, 2, 3, 4, 5, 6, 7, 8", "   2, 2, 9, 10, 3, 6, 11, 12, 7", "  END CONNECTIVITY", " END ELEMENTS", "END MESH",
          "BOUNDARY CONDITIONS"]
    for n in ends:
        for d in range(3):
            t += [" HISTORY, h%d_%d" % (n, d + 1)] + ["  %.17g, %.17g" % (k, position(n, k)[d] - X[n - 1][d]) for k in range(steps + 1)] + [" END HISTORY"]
    t += [" DEGREES OF FREEDOM"] + ["  %d, %d, HISTORY, h%d_%d" % (n, d + 1, n, d + 1) for n in ends for d in range(3)] + [" END DEGREES OF FREEDOM",
          "END BOUNDARY CONDITIONS", "STATIC STEP", " INITIAL TIME STEP, 1.d0", " MAX TIME STEP, 1.d0", " MIN TIME STEP, 0.001d0",
          " MAX NUMBER OF STEPS, %d" % steps, " STOP TIME, %d.d0" % steps, " SOLVER, DIRECT, NONLINEAR, 1.d-08, 15, UNSYMMETRIC",
          "END STATIC STEP", "STOP", ""]
    m = Model.from_text("\n".join(t))
    g = GpuSystem(m)
    res = g.run_static(method=0)
    ref = ob.OracleModel(m.arrays()).run_static()
    assert res["n_steps"] == ref["n_steps"] == steps and res["n_cutbacks"] == ref["n_cutbacks"] == 0
    assert res["newton"].shape == ref["newton"].shape
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-9
    sv = g.state_variables(updated=False).reshape(2, 8, 15)
    svo = ref["state_variables"].reshape(2, 8, 15)
    assert np.abs(sv - svo).max() <= 1e-8 * np.abs(svo).max()
    # after the 90 degree rotation the bar axis is the 2-axis: s22 carries the axial stress (~ E x 1 %), larger than s11
    assert np.all(sv[..., 1] > 80.0) and np.all(sv[..., 1] > 2.0 * np.abs(sv[..., 0]))
    g.close()


def test_stretch_rotate_reference_input_with_subroutine_bcs():
    """input_files/Abaqus_umat_stretch_rotate.in as shipped: SUBROUTINE-type prescribed DOFs (user_prescribeddof driven by
    current_step_number) on the two-element C3D8 bar — device path against the oracle, and the objectivity result itself:
    after the rigid 90 degree rotation the uniaxial stress sits on the 2-axis, magnitude unchanged."""
    m = golden_model("stretch_rotate_3d_umat.in")
    g = GpuSystem(m)
    res = g.run_static(method=0)
    ref = ob.OracleModel(m.arrays()).run_static()
    assert res["n_steps"] == ref["n_steps"] == 11 and res["n_cutbacks"] == ref["n_cutbacks"] == 0
    assert res["newton"].shape == ref["newton"].shape
    assert np.array_equal(res["newton"][:, [0, 1, 7]], ref["newton"][:, [0, 1, 7]])
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-9
    sv = g.state_variables(updated=False).reshape(2, 8, 15)
    svo = ref["state_variables"].reshape(2, 8, 15)
    assert np.abs(sv - svo).max() <= 1e-8 * np.abs(svo).max()
    assert np.allclose(sv[..., 1], 49.875, atol=1e-3) and np.abs(np.delete(sv[..., :6], 1, axis=-1)).max() < 1e-4
    g.close()


# ---------------------------------------------------------------------------------------------- state print
def test_field_projection_matches_oracle_linear_elastic_elements(cube8):
    """print_state's nodal projection (src/printing.f90:457-490) on the device against the oracle's restatement (skyline
    LDU of the 27-point matrix): element 1001 on the jittered 8^3 block, consistent and lumped matrices.  The device
    solves the consistent system by Jacobi-PCG to r.r/b.b <= 1e-28, so the bar is 1e-10 of the largest nodal value."""
    m, g, om = cube8
    names = ["S11", "S22", "S33", "S12", "S13", "S23", "SMISES", "nonsense"]
    rng = np.random.default_rng(11)
    ut, du = 0.01 * rng.uniform(-1, 1, g.ndof), 0.002 * rng.uniform(-1, 1, g.ndof)
    for lumped in (False, True):
        got = g.project_fields(names, ut, du, lumped=lumped)
        want = om.project_fields(names, ut, du, lumped=lumped)
        assert got.shape == want.shape == (m.n_nodes, 8)
        assert np.abs(got[:, :7] - want[:, :7]).max() < 1e-10 * np.abs(want).max(), lumped
        assert not got[:, 7].any() and not want[:, 7].any()             # unknown names project to zero
    # the stiffness storage was used for the projection matrix: solving without a new assembly is refused ...
    with pytest.raises(RuntimeError):
        g.solve_dev(method=0)
    # ... and a fresh static step on the same handle still reproduces the oracle
    res = g.run_static(method=0)
    ref = om.run_static()
    assert relmax(res["dof_total"], ref["dof_total"]) < TOL_U


def test_state_print_file_of_the_uel_reference_input(tmp_path):
    """input_files/Abaqus_uel_linear_elastic_3d.in with its PRINT STATE block (DEGREES OF FREEDOM, six stresses, DISPLACED
    MESH): run_static writes one Tecplot zone per step (STATE PRINT STEPS 1, 3 steps); the last zone against the oracle's
    print_state of the same state, number by number, and against the analytic uniaxial stress of 2."""
    m = golden_model("linear_elastic_3d_uel.in")
    names = ["S11", "S22", "S33", "S12", "S13", "S23"]
    out = tmp_path / "contourplots.dat"
    m.set_state_print(str(out), field_variables=names, print_dof=True, displaced_mesh=True, displacement_scale_factor=1.0)
    g = GpuSystem(m)
    res = g.run_static(method=0)
    assert res["n_steps"] == 3 and res["n_state_prints"] == 3
    lines = out.read_text().splitlines()
    per_zone = 2 + 12 + 2
    assert len(lines) == 3 * per_zone
    assert [l for l in lines if l.startswith(" ZONE")] == [' ZONE, T="  0.%d000D+01" F=FEPOINT, I=   12 J=    2 ET=BRICK' % k for k in (1, 2, 3)]
    # the load is constant in time (dload_history), so step 1 carries the whole displacement and the increments of steps
    # 2 and 3 are zero to rounding: the state printed at step 3 is (dof_total, 0)
    om = ob.OracleModel(m.arrays())
    ref = om.run_static()
    u3 = ref["dof_total"]
    want = tmp_path / "oracle.dat"
    om.print_state(want, names, u3, None, svars=ref["state_variables"], print_dof=True, displaced_mesh=True, time_end=3.0)
    wl = want.read_text().splitlines()
    gl = lines[2 * per_zone:]
    assert gl[0] == wl[0] and gl[1] == wl[1] and gl[14:] == wl[14:]
    for a_, b_ in zip(gl[2:14], wl[2:14]):
        va = np.array([float(a_[19 * k:19 * k + 19]) for k in range(15)])
        vb = np.array([float(b_[19 * k:19 * k + 19]) for k in range(15)])
        assert np.allclose(va, vb, rtol=1e-9, atol=1e-11)
        assert abs(va[9] - 2.0) < 1e-7 and np.abs(va[10:]).max() < 1e-11
    # one more zone through the direct entry point, appended
    g.print_state(str(out), u3, None, time_end=4.0, append=True)
    assert len(out.read_text().splitlines()) == 4 * per_zone
    g.close()


def test_field_projection_c3d8_state_variables_golden_model():
    """The golden-log model (3075 C3D8 B-bar elements, elastic UMAT): stresses and strains of the converged state are read
    from the device state variables (continuum_elements.f90:1206-1351) and projected; against the oracle's projection of
    its own converged state.  The two runs differ by their linear solvers (1e-9 on the state), the projection adds 1e-10."""
    m = golden_model("holeplate_3d_umat.in.gz")
    g = GpuSystem(m)
    res = g.run_static(method=0)
    names = ["S11", "S22", "S33", "S12", "S13", "S23", "E11", "E22", "E33", "E12", "E13", "E23", "U1", "SMISES"]
    got = g.project_fields(names)                                        # device-resident state of the last step
    a = dict(np.load(os.path.join(GOLDEN, "holeplate_3d_umat_model.npz")))
    om = ob.OracleModel(a, node_order=ob.OracleModel(a).gps_node_order())
    ref = om.run_static()
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-9
    want = om.project_fields(names, ref["dof_total"], svars=ref["state_variables"])
    for k in range(12):
        assert np.abs(got[:, k] - want[:, k]).max() < 1e-7 * np.abs(want[:, k]).max(), names[k]
    assert not got[:, 12:].any() and not want[:, 12:].any()             # U1: the elastic UMAT has no state; SMISES: not a C3D8 field
    # stress concentration at the hole: the projected S11 peaks at about three times the remote stress (Kirsch), here a
    # finite-width plate, so between 2.5 and 4.5
    s11 = got[:, 0]
    far = np.median(s11)
    assert 2.5 < s11.max() / far < 4.5
    lum = g.project_fields(names[:6], lumped=True)
    wl = om.project_fields(names[:6], ref["dof_total"], svars=ref["state_variables"], lumped=True)
    assert np.abs(lum - wl).max() < 1e-7 * np.abs(wl).max()
    g.close()


# ---------------------------------------------------------------------------------------------- explicit dynamics
def test_explicit_dynamics_matches_oracle_on_a_jittered_block():
    """explicit_dynamic_step (src/explicit_dynamic_step.f90:1-83) on the device against the oracle's restatement: VUEL
    elements (density = third property), a prescribed-displacement ramp on x = max, a nodal force history, clamped x = 0;
    70 steps = four 16-step CUDA graphs + 6 single steps.  The two differ by roundoff only (fused multiply-adds, the scale of a
    load applied after instead of inside the face integral): 1e-10 of the field maximum."""
    text = synthetic_block_input(6, 5, 4, jitter=0.2, element="U1", props=[100.0, 0.3, 2.5], load="stretch", steps=4, dt=0.5,
                                 dynamic=(1e-3, 70))
    text = text.replace(" END DEGREES OF FREEDOM", " END DEGREES OF FREEDOM\n FORCES\n  xmax, 3, HISTORY, ramp\n END FORCES")
    m = Model.from_text(text)
    g = GpuSystem(m)
    res = g.run_dynamic()
    ref = ob.OracleModel(m.arrays()).dynamic().steps(70).get()
    assert res["steps"] == 70 and res["n_state_prints"] == 0
    assert relmax(res["lumped_mass"], ref["lumped_mass"]) < 1e-13
    for k in ("dof_total", "velocity", "acceleration"):
        assert relmax(res[k], ref[k]) < 1e-10, k
    st = g.dynamic_state()
    assert abs(st["time"] - ref["time"]) < 1e-15 and relmax(st["dof_increment"], ref["dof_increment"]) < 1e-10
    # 30 more steps from the device-resident state (single launches: below the graph threshold)
    fail, ms = g.dynamic_steps(1e-3, 30)
    assert fail == 0 and ms > 0
    ref2 = ob.OracleModel(m.arrays()).dynamic().steps(100).get()
    assert relmax(g.dynamic_state()["dof_total"], ref2["dof_total"]) < 1e-10
    g.close()


def test_explicit_dynamics_vuel_reference_input_with_state_prints(tmp_path):
    """input_files/Abaqus_vuel_linear_elastic_3d.in as shipped (two VUEL elements, pressure history on face 4 of the end
    element, u_x = 0 on the left face, 5 steps of 0.01, STATE PRINT STEPS 1 with the lumped projection matrix): final state
    against the oracle, one Tecplot zone per step written where the reference prints (after the boundary conditions of the
    step), header time TIME + DTIME."""
    m = golden_model("linear_elastic_3d_vuel.in")
    out = tmp_path / "contourplots.dat"
    sp = m.state_print()
    m.set_state_print(str(out))
    g = GpuSystem(m)
    res = g.run_dynamic()
    ref = ob.OracleModel(m.arrays()).dynamic().steps(5).get()
    assert res["steps"] == 5 and res["n_state_prints"] == 5
    for k in ("dof_total", "velocity", "acceleration", "lumped_mass"):
        assert relmax(res[k], ref[k]) < 1e-11, k
    assert abs(res["lumped_mass"][0::3].sum() - 20.0) < 1e-6                       # two unit cubes of density 10
    lines = out.read_text().splitlines()
    assert len(lines) == 5 * (2 + 12 + 2)
    zones = [l for l in lines if l.startswith(" ZONE")]
    assert zones == [' ZONE, T="  0.%d000D-01" F=FEPOINT, I=   12 J=    2 ET=BRICK' % k for k in (1, 2, 3, 4, 5)]
    assert lines[0] == " VARIABLES = X,Y,Z,DU1,DU2,DU3,U1,U2,U3," + ",".join(sp["field_variables"])
    # last zone: printed with (dof_total after 4 steps, predicted increment of step 5); U + DU of every node is the final state
    last = lines[4 * 16 + 2:4 * 16 + 14]
    order = list(dict.fromkeys((m.arrays()["connectivity"]).tolist()))
    for line, node in zip(last, order):
        v = np.array([float(line[19 * k:19 * k + 19]) for k in range(15)])
        assert np.allclose(v[3:6] + v[6:9], ref["dof_total"][3 * (node - 1):3 * node], rtol=1e-8, atol=1e-14)
    g.close()


def test_explicit_dynamics_vuel_holeplate_reference_input():
    """input_files/Abaqus_vuel_holeplate_3d.in as shipped (3,075 VUEL elements, 1000 steps of 0.002, state prints every 20 steps
    switched off here): the whole run on the device against the oracle's central-difference loop."""
    m = golden_model("holeplate_3d_vuel.in.gz")
    a = m.arrays()
    assert a["explicitdynamicstep"][0] == 1 and a["no_dynamic_steps"][0] == 1000 and m.n_elements == 3075
    g = GpuSystem(m)
    res = g.run_dynamic()
    ref = ob.OracleModel(a).dynamic().steps(1000).get()
    assert res["steps"] == 1000
    assert relmax(res["lumped_mass"], ref["lumped_mass"]) < 1e-13
    for k in ("dof_total", "velocity", "acceleration"):
        assert relmax(res[k], ref[k]) < 1e-9, k
    assert np.abs(res["dof_total"]).max() > 0
    g.close()


def test_explicit_dynamics_native_element_with_density_and_tractions():
    """Element 1001 with a DENSITY key (user_element_lumped_mass, el_linelast_3dbasic_dynamic) under the three native
    distributed-load types of traction_boundarycondition_dynamic — VALUE, HISTORY (normalised direction x history) and NORMAL
    (as that routine writes it) — against the oracle."""
    base = Model.from_text(synthetic_block_input(4, 3, 3, jitter=0.15)).text()
    base = base.replace("  PARAMETERS, 8, 0, 1001\n", "  PARAMETERS, 8, 0, 1001\n  DENSITY, 3.d0\n")
    head = base[:base.index(" DISTRIBUTED LOADS")]
    loads = (" HISTORY, pulse\n  0.d0, 0.d0\n  0.02d0, 4.d0\n  1.d0, 4.d0\n END HISTORY\n"
             " DISTRIBUTED LOADS\n  xmax_el, 4, VALUE, 0.d0, 0.5d0, -1.d0\n  ymax_el, 5, HISTORY, pulse, 1.d0, 1.d0, 0.d0\n"
             "  zmax_el, 2, NORMAL, pulse\n END DISTRIBUTED LOADS\nEND BOUNDARY CONDITIONS\n")
    text = head + loads + "EXPLICIT DYNAMIC STEP\n TIME STEP, 5.d-4\n NUMBER OF STEPS, 40\nEND EXPLICIT DYNAMIC STEP\nSTOP\n"
    m = Model.from_text(text)
    a = m.arrays()
    assert a["dload_flag"].tolist() == [1, 2, 3] and a["element_density"].min() == 3.0
    g = GpuSystem(m)
    res = g.run_dynamic()
    ref = ob.OracleModel(a).dynamic().steps(40).get()
    for k in ("dof_total", "velocity", "acceleration", "lumped_mass"):
        assert relmax(res[k], ref[k]) < 1e-10, k
    assert np.abs(res["dof_total"]).max() > 0
    g.close()


# ------------------------------------------------------------------------------------------- several parts behind one handle
def _devices(n_parts):
    """One GPU per part.  On a box with fewer GPUs the test is skipped: parts sharing a device would wait for one another
    from separately launched kernels, which nothing guarantees to run at the same time (B200_PROFILING.md) — the library
    supports it as a debugging aid (EN234_ALLOW_SHARED_DEVICE=1, used here only when EN234_TEST_SHARED_DEVICE=1 asks for
    it); the N > 1 host logic is covered on the CPU (tests/test_partition_gloo.py, the partition tests) and the multi-GPU
    numerics by bench.py's parity gate at every N."""
    from en234_fea_b200 import api
    nd = api.gpu_lib().en234gpu_device_count()
    if nd >= n_parts:
        return list(range(n_parts))
    if os.environ.get("EN234_TEST_SHARED_DEVICE") == "1":
        os.environ["EN234_ALLOW_SHARED_DEVICE"] = "1"
        return [0] * n_parts
    pytest.skip("%d GPUs needed, %d present" % (n_parts, nd))


@pytest.mark.parametrize("dims,n_parts,method", [((6, 5, 8), 2, 0), ((5, 4, 9), 3, 0), ((6, 5, 8), 2, 1), ((4, 4, 8), 4, 0)])
def test_multi_handle_from_plain_c_matches_one_device(dims, n_parts, method):
    """tests/c/multi_check.c: a C program that includes include/en234gpu.h only drives en234gpu_create_multi (partition
    inside the library, peer-memory collectives folded into the PCG kernels, one host thread per part) and compares the
    static step with the one-device handle."""
    import subprocess
    exe = os.path.join(os.path.dirname(__file__), "c", "multi_check")
    assert os.path.exists(exe), "tests/c/multi_check is built by __graft_entry__.build()"
    devs = _devices(n_parts)
    env = dict(os.environ, EN234_P2P_TIMEOUT_MS="20000")
    out = subprocess.run([exe] + [str(d) for d in dims] + [str(n_parts), ",".join(str(d) for d in devs), str(method)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    print(out.stdout)
    assert out.returncode == 0, out.stdout


@pytest.mark.parametrize("fused", ["1", "0"])
def test_multi_handle_static_step_matches_oracle(fused, monkeypatch):
    """The whole-mesh C ABI over 3 parts against the oracle (skyline LDU) on the same model: u and rforce within 1e-10,
    PCG iteration count within 2 of the CPU recurrences; collectives folded into the kernels (default) and as separate
    kernels (EN234_P2P_FUSED=0)."""
    from en234_fea_b200 import MultiGpuSystem
    monkeypatch.setenv("EN234_P2P_FUSED", fused)
    monkeypatch.setenv("EN234_P2P_TIMEOUT_MS", "20000")
    m = Model.from_text(synthetic_block_input(6, 5, 9, jitter=0.2, cg_tolerance=1e-26, cg_maxit=5000))
    fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
    g = MultiGpuSystem(m, devices=_devices(3))
    z = np.zeros(g.ndof)
    rf, nrm, fail = g.assemble(z, z)
    assert not fail
    g.apply_boundaryconditions(fx, fd, fc)
    du, corr, its, sfail = g.solve(z, 0, 1e-26, 5000)
    assert not sfail
    rf, nrm2, fail = g.assemble(z, du)
    unb = g.apply_boundaryconditions(fx, fd, fv - du[fd])
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    refcg = ob.OracleModel(m.arrays(), solvertype=2, cg_check_linear=1).run_static()
    assert relmax(du, ref["dof_total"]) < TOL_U
    assert relmax(rf, ref["rforce"]) < TOL_U
    assert abs(its - int(refcg["cg_iterations"][0])) <= 2
    assert unb / nrm2 < 1e-9
    # device-resident variants
    g.upload_state(z, z, None, fx, fd, fv)
    g.zero_increment_dev()
    g.assemble_dev(); g.apply_boundaryconditions_dev()
    _, its2, sfail = g.solve_dev(0, 1e-26, 5000)
    du2, rf2 = g.download_state()
    assert its2 == its and np.array_equal(du2, du)
    g.close()


# ------------------------------------------------------------------------------------------- other 3-D element shapes (1001)
def _assembled(g):
    return g.csr().toarray()


@pytest.mark.parametrize("shape,dims,jitter", [("tet4", (3, 2, 2), 0.15), ("tet10", (2, 2, 1), 0.15), ("hex20", (2, 2, 2), 0.2)])
def test_other_element_shapes_assembly_matches_oracle(shape, dims, jitter):
    """4-node / 10-node tetrahedra and 20-node bricks of element 1001 through the general device path: assembled matrix,
    residual and reaction forces against the oracle (which carries the reference's shape-function table, quirks included)."""
    import meshes
    m = Model.from_text(meshes.input_text(shape, *dims, jitter=jitter))
    a = m.arrays()
    g = GpuSystem(m)
    rng = np.random.default_rng(5)
    ut, du = 1e-2 * rng.standard_normal(g.ndof), 1e-3 * rng.standard_normal(g.ndof)
    rf, nrm, fail = g.assemble(ut, du)
    s = ob.OracleSystem(ob.OracleModel(a, solvertype=2), 2)
    rfo, nrmo, failo = s.assemble(ut, du)
    assert fail == failo == 0
    A, _ = s.matrix()
    K = _assembled(g)
    assert np.abs(K - A.toarray()).max() <= 1e-13 * np.abs(A.toarray()).max()
    assert relmax(rf, rfo) < 1e-12 and abs(nrm - nrmo) <= 1e-12 * nrmo
    g.close()


def test_tet4_static_step_matches_oracle():
    """A tet4 mesh (the one shape besides hex8 whose reference table is right) through the whole device step — parser,
    general assembly, boundary conditions, PCG — against the oracle's skyline LDU: u and rforce within 1e-10."""
    import meshes
    m = Model.from_text(meshes.input_text("tet4", 5, 4, 4))
    g = GpuSystem(m)
    res = g.run_static(method=0)
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    refcg = ob.OracleModel(m.arrays(), solvertype=2, cg_check_linear=1).run_static()
    assert relmax(res["dof_total"], ref["dof_total"]) < TOL_U
    assert relmax(res["rforce"], ref["rforce"]) < TOL_U
    assert abs(int(res["cg_iterations"][0]) - int(refcg["cg_iterations"][0])) <= 2
    with pytest.raises(RuntimeError, match="hex8"):
        g.set_operator_mode(1)
    g.close()


def test_regular_hex20_fails_like_the_reference():
    """Undistorted 20-node bricks: the reference's derivative table makes the Jacobian singular at an integration point, where
    its invert_small stops the program; the device path and the oracle report fail instead."""
    import meshes
    m = Model.from_text(meshes.input_text("hex20", 2, 1, 1, jitter=0.0))
    g = GpuSystem(m)
    z = np.zeros(g.ndof)
    assert g.assemble(z, z)[2] == 1
    s = ob.OracleSystem(ob.OracleModel(m.arrays(), solvertype=2), 2)
    assert s.assemble(z, z)[2] == 1
    g.close()


# ---- two-node tie constraints (CONSTRAINTS block): Lagrange multipliers in the reference, group elimination inside PCG here ----
TOL_TIE = 2e-9     # the reference regularises every multiplier row with a compliance of 1e-12 |diag| / neq, which the device
                   # path drops; on these meshes that moves the oracle's displacements by a few 1e-11 of their maximum


def _tie_case(text, before_last_node=True):
    """The oracle's skyline LDU has no pivoting, so the position of the multiplier equations matters to IT (not to the device
    path): before the last of its two nodes where a body is held by its ties only, after all DOFs where the master is
    prescribed (a multiplier row placed before its only free DOF would start with the bare 1e-12 compliance as its pivot)."""
    import meshes
    m = Model.from_text(text)
    a = m.arrays()
    ref = (ob.OracleModel(a, node_order=meshes.order_with_constraints(a)) if before_last_node else ob.OracleModel(a)).run_static()
    return m, ref


def _same_newton_history(res, ref, rtol=1e-6):
    assert res["newton"].shape == ref["newton"].shape
    k = ref["newton"][:, 7] == 0                               # records before convergence: norms well above round-off
    assert np.allclose(res["newton"][k, 2:5], ref["newton"][k, 2:5], rtol=rtol, atol=0)
    # every record: correction norm (it contains the multiplier corrections) and generalized force norm (multiplier rows)
    assert np.allclose(res["newton"][:, 2:4], ref["newton"][:, 2:4], rtol=rtol, atol=0)
    assert np.array_equal(res["newton"][:, 7], ref["newton"][:, 7])


@pytest.mark.parametrize("use_host_api", [False, True])
def test_tied_cut_block_equals_uncut_block_and_oracle(use_host_api):
    """A block cut along a node plane, the two sides tied node to node in all three DOFs (CONSTRAIN NODE PAIRS over node sets):
    the device path (tied groups reduced onto one unknown inside PCG) gives the displacements of the UNCUT block to solver
    accuracy, and the oracle's Lagrange-multiplier solution (skyline LDU of the saddle-point system) within TOL_TIE; multipliers,
    reaction forces and the Newton log (whose norms run over the multiplier rows too) agree with the oracle's."""
    import meshes
    text, info = meshes.tied_block_text(5, 4, 3, cut=2)
    m, ref = _tie_case(text)
    g = GpuSystem(m)
    res = g.run_static(method=0, use_host_api=use_host_api)
    m0 = Model.from_text(meshes.tied_block_text(5, 4, 3, cut=None)[0])
    g0 = GpuSystem(m0)
    res0 = g0.run_static(method=0)
    n0 = 3 * info["n_orig"]
    assert relmax(res["dof_total"][:n0], res0["dof_total"]) < 1e-12
    for o, c in info["copy"].items():
        assert np.abs(res["dof_total"][3 * (o - 1):3 * o] - res["dof_total"][3 * (c - 1):3 * c]).max() <= 1e-15 * np.abs(res0["dof_total"]).max()
    assert relmax(res["dof_total"], ref["dof_total"]) < TOL_TIE
    assert relmax(res["rforce"], ref["rforce"]) < TOL_TIE
    assert relmax(res["lagrange_multipliers"], ref["lagrange_multipliers"]) < 1e-7
    nc = len(info["copy"])
    assert abs(res["lagrange_multipliers"][2 * nc:].sum() - 0.01 * len(info["xmax"])) < 1e-12     # the ties carry the whole load
    _same_newton_history(res, ref)
    with pytest.raises(RuntimeError, match="assembled operator"):
        g.run_static(method=1)
    again = g.run_static(method=0, use_host_api=use_host_api)     # a second analysis on the handle starts from zero multipliers
    assert np.allclose(again["newton"][:, 2:4], res["newton"][:, 2:4], rtol=1e-9, atol=0) and relmax(again["dof_total"], res["dof_total"]) < 1e-12
    g.close(); g0.close()


def test_tie_nodes_face_follows_master_node_forces_and_prescribed_master():
    """TIE NODES: DOF 1 of every node of the face x = max follows the first node of that face.  (a) nodal forces on the face;
    (b) the master's DOF is PRESCRIBED over two steps (the tied group then follows a prescribed root: nothing is solved for it,
    the multipliers carry the reactions).  Both against the oracle's multiplier solution."""
    import meshes
    for kw in ({"force": (0.02, 0.0, -0.01)}, {"pull": 0.05, "steps": 2}):
        text, info = meshes.tied_block_text(4, 3, 3, tie_face_dof=1, **kw)
        m, ref = _tie_case(text, before_last_node=False)
        g = GpuSystem(m)
        res = g.run_static(method=0)
        u = res["dof_total"]
        face = np.array(info["xmax"]) - 1
        assert np.abs(u[3 * face] - u[3 * face[0]]).max() <= 1e-15 * np.abs(u).max()             # exactly tied
        if "pull" in kw:
            assert abs(u[3 * face[0]] - 0.05) < 1e-15
        assert relmax(u, ref["dof_total"]) < TOL_TIE
        assert relmax(res["rforce"], ref["rforce"]) < TOL_TIE
        assert relmax(res["lagrange_multipliers"], ref["lagrange_multipliers"]) < 1e-7
        assert res["n_steps"] == ref["n_steps"] == kw.get("steps", 1)
        _same_newton_history(res, ref)
        g.close()


def test_tie_errors_loops_and_doubly_prescribed_groups():
    m = Model.from_text(synthetic_block_input(3, jitter=0.1))
    g = GpuSystem(m)
    with pytest.raises(RuntimeError, match="loop"):
        g.set_ties([40, 41, 40], [41, 44, 44])
    with pytest.raises(RuntimeError, match="itself"):
        g.set_ties([40], [40])
    g.set_ties([0], [12])          # nodes 1 and 5 (DOFs 0 and 12) both lie on the clamped face x = 0
    with pytest.raises(RuntimeError, match="two prescribed DOFs"):
        g.run_static(method=0)
    g.set_ties([], [])
    g.run_static(method=0)
    g.close()


def test_tied_face_on_neohookean_block_same_newton_history():
    """Tie constraints inside a NONLINEAR analysis: neo-Hookean user elements (U2), the face x = max tied in DOF 1 to a master
    node that is pulled by 8 % over two steps.  Every Newton iteration re-assembles the tangent, adds -/+ lambda to the tied
    rows and solves the reduced system; the records (whose norms include the multiplier rows and corrections) follow the
    oracle's Lagrange-multiplier solution."""
    import meshes
    text, info = meshes.tied_block_text(4, 3, 3, tie_face_dof=1, pull=0.32, steps=2, element="U2", props=[1.0, 200.0],
                                        solver="DIRECT, NONLINEAR, 1.d-7, 15")
    m, ref = _tie_case(text, before_last_node=False)
    g = GpuSystem(m)
    res = g.run_static(method=0)
    assert res["n_steps"] == ref["n_steps"] == 2 and res["newton"].shape == ref["newton"].shape and res["newton"].shape[0] >= 6
    assert np.array_equal(res["newton"][:, [0, 1, 7]], ref["newton"][:, [0, 1, 7]])
    big = ref["newton"][:, 5] > 1e-6
    assert np.allclose(res["newton"][big][:, 2:6], ref["newton"][big][:, 2:6], rtol=1e-5)
    face = np.array(info["xmax"]) - 1
    assert np.abs(res["dof_total"][3 * face] - 0.32).max() < 1e-15
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-8
    assert relmax(res["lagrange_multipliers"], ref["lagrange_multipliers"]) < 1e-6
    g.close()


def test_tied_cut_block_of_c3d8_elements_bicgstab():
    """Tie constraints with the UNSYMMETRIC solver: the cut-and-tied block made of built-in C3D8 B-bar + UMAT elements
    (BiCGSTAB with the same expand / reduce wrappers around its operator) against the uncut block on the device and against
    the oracle's unsymmetric skyline LDU with multiplier rows."""
    import meshes
    import re

    def c3d8(text):
        head = text[:text.index(" ELEMENTS, USER")]
        conn = text[text.index("  CONNECTIVITY"):]
        t = (head + " MATERIAL, elastic_umat\n  STATE VARIABLES, 0\n  PROPERTIES\n   100.d0\n   0.3d0\n  END PROPERTIES\n END MATERIAL\n"
             " ELEMENTS, INTERNAL\n  TYPE, C3D8\n  PROPERTIES, elastic_umat\n" + conn)
        return re.sub(r" SOLVER, [^\n]*", " SOLVER, DIRECT, NONLINEAR, 1.d-9, 15, UNSYMMETRIC", t)

    text, info = meshes.tied_block_text(4, 3, 3, cut=2, force=(0.0, 0.0, -0.002))
    m, ref = _tie_case(c3d8(text))
    assert m.arrays()["unsymmetric"][0] == 1 and m.arrays()["element_flag"][0] == 10003
    g = GpuSystem(m)
    res = g.run_static(method=0)
    m0 = Model.from_text(c3d8(meshes.tied_block_text(4, 3, 3, cut=None, force=(0.0, 0.0, -0.002))[0]))
    g0 = GpuSystem(m0)
    res0 = g0.run_static(method=0)
    n0 = 3 * info["n_orig"]
    assert relmax(res["dof_total"][:n0], res0["dof_total"]) < 1e-9
    for o, c in info["copy"].items():
        assert np.abs(res["dof_total"][3 * (o - 1):3 * o] - res["dof_total"][3 * (c - 1):3 * c]).max() <= 1e-15 * np.abs(res0["dof_total"]).max()
    assert relmax(res["dof_total"], ref["dof_total"]) < 1e-8
    assert relmax(res["lagrange_multipliers"], ref["lagrange_multipliers"]) < 1e-6
    assert res["newton"].shape == ref["newton"].shape and np.array_equal(res["newton"][:, 7], ref["newton"][:, 7])
    big = ref["newton"][:, 5] > 1e-6
    assert np.allclose(res["newton"][big][:, 2:5], ref["newton"][big][:, 2:5], rtol=1e-5)
    g.close(); g0.close()


def test_symmetric_storage_spmv_variant_gives_the_same_solution(monkeypatch):
    """EN234_SPMV_SYM=1 (A/B, off by default: measured 2x slower at 128^3): rows read only their diagonal and
    upper blocks and take the lower ones as transposes of other rows' blocks.  Same PCG iteration count, solution within
    rounding of the full-storage product."""
    m = Model.from_text(synthetic_block_input(9, 8, 7, jitter=0.2, cg_tolerance=1e-26, cg_maxit=20000))
    g = GpuSystem(m)
    full = g.run_static(method=0)
    g.close()
    monkeypatch.setenv("EN234_SPMV_SYM", "1")
    g = GpuSystem(m)
    sym = g.run_static(method=0)
    g.close()
    assert abs(int(sym["cg_iterations"][0]) - int(full["cg_iterations"][0])) <= 1
    assert relmax(sym["dof_total"], full["dof_total"]) < 1e-12 and relmax(sym["rforce"], full["rforce"]) < 1e-12
    ref = ob.OracleModel(m.arrays(), solvertype=1).run_static()
    assert relmax(sym["dof_total"], ref["dof_total"]) < TOL_U
