"""ctypes binding of the CPU oracle (oracle/liben234oracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_ROOT, "oracle", "liben234oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-j8"], cwd=os.path.join(_ROOT, "oracle"), stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.or_model_new.restype = C.c_void_p
        L.or_model_free.argtypes = [C.c_void_p]
        L.or_model_set_int.argtypes = [C.c_void_p, C.c_char_p, ip, C.c_long]
        L.or_model_set_double.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_long]
        L.or_model_finalize.argtypes = [C.c_void_p]
        L.or_system_new.restype = C.c_void_p
        L.or_system_new.argtypes = [C.c_void_p, C.c_int]
        L.or_system_free.argtypes = [C.c_void_p]
        L.or_system_nnz_upper.restype = C.c_long
        L.or_system_nnz_upper.argtypes = [C.c_void_p]
        L.or_dynamic_new.restype = C.c_void_p
        L.or_dynamic_new.argtypes = [C.c_void_p, dp, dp]
        L.or_dynamic_free.argtypes = [C.c_void_p]
        L.or_dynamic_steps.argtypes = [C.c_void_p, C.c_int]
        L.or_dynamic_get.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


class OracleModel:
    def __init__(self, arrays, **overrides):
        L = lib()
        self.h = C.c_void_p(L.or_model_new())
        self.arrays = dict(arrays)
        self.arrays.update({k: np.atleast_1d(v) for k, v in overrides.items()})
        self._keep = []
        for k, v in self.arrays.items():
            if k == "checkstiffness_flag":
                continue
            if v.dtype.kind in "iu":
                a = np.ascontiguousarray(v, dtype=np.int32)
                rc = L.or_model_set_int(self.h, k.encode(), _i(a), a.size)
            else:
                a = np.ascontiguousarray(v, dtype=np.float64)
                rc = L.or_model_set_double(self.h, k.encode(), _d(a), a.size)
            if rc:
                raise KeyError("oracle does not know field " + k)
            self._keep.append(a)
        rc = L.or_model_finalize(self.h)
        if rc:
            raise ValueError("oracle model invalid: %d" % rc)
        self.n_nodes = int(self.arrays["n_nodes"][0])
        self.n_elements = int(self.arrays["n_elements"][0])
        self.ndof = 3 * self.n_nodes

    def close(self):
        if self.h:
            lib().or_model_free(self.h)
            self.h = None

    def gps_node_order(self):
        """The reference's Gibbs-Poole-Stockmeyer numbering (src/bandwidth_minimizer.f90): position -> 0-based node."""
        order = np.zeros(self.n_nodes, np.int32)
        lib().or_gps_node_order(self.h, _i(order))
        return order

    def projection_mass_lu(self, order):
        """(max reduced diagonal, min reduced diagonal, digits lost) of the print_state projection mass matrix LU."""
        out = np.zeros(3)
        lib().or_projection_mass_lu(self.h, _i(np.ascontiguousarray(order, np.int32)), _d(out))
        return out

    def project_fields(self, names, dof_total, dof_increment=None, svars=None, lumped=False, raw=False):
        """Nodal projection of the named integration-point fields (src/printing.f90:457-490): (n_nodes, n_fields).
        raw=True returns the assembled right-hand sides int f N dV instead."""
        out = np.zeros((self.n_nodes, len(names)))
        du = np.zeros(self.ndof) if dof_increment is None else np.ascontiguousarray(dof_increment, np.float64)
        ut = np.ascontiguousarray(dof_total, np.float64)
        sv = None if svars is None else np.ascontiguousarray(svars, np.float64)
        rc = lib().or_project_fields(self.h, _d(ut), _d(du), _d(sv) if sv is not None else None, ",".join(names).encode(),
                                     C.c_int(int(lumped)), C.c_int(int(raw)), _d(out))
        if rc:
            raise RuntimeError("oracle field projection failed rc=%d" % rc)
        return out

    def lumped_projection_mass(self):
        out = np.zeros(self.n_nodes)
        lib().or_lumped_projection_mass(self.h, _d(out))
        return out

    def print_state(self, path, names, dof_total, dof_increment=None, svars=None, print_dof=False, displaced_mesh=False,
                    lumped=False, scale_factor=1.0, time_end=1.0):
        du = np.zeros(self.ndof) if dof_increment is None else np.ascontiguousarray(dof_increment, np.float64)
        ut = np.ascontiguousarray(dof_total, np.float64)
        sv = None if svars is None else np.ascontiguousarray(svars, np.float64)
        flags = int(print_dof) | 2 * int(displaced_mesh) | 4 * int(lumped)
        rc = lib().or_print_state(self.h, _d(ut), _d(du), _d(sv) if sv is not None else None, ",".join(names).encode(),
                                  C.c_int(flags), C.c_double(scale_factor), C.c_double(time_end), str(path).encode())
        if rc:
            raise RuntimeError("oracle print_state failed rc=%d" % rc)

    def lumped_mass(self):
        out = np.zeros(self.ndof)
        lib().or_lumped_mass(self.h, _d(out))
        return out

    def dynamic(self, dof_total0=None, velocity0=None):
        """Explicit dynamic step state (src/explicit_dynamic_step.f90); advance with .steps(n), read with .get()."""
        return OracleDynamic(self, dof_total0, velocity0)

    def run_static(self):
        L = lib()
        u, rf = np.zeros(self.ndof), np.zeros(self.ndof)
        newton, cg = np.zeros((4096, 8)), np.zeros(4096, np.int32)
        info = (C.c_longlong * 10)()
        lu = np.zeros((256, 3))
        nsv = int(np.sum(self.arrays["element_n_states"])) if "element_n_states" in self.arrays else 0
        sv = np.zeros(max(nsv, 1))                       # element state variables, zero initial state in, final state out
        rc = L.or_run_static(self.h, _d(u), _d(rf), _d(sv) if nsv else None, _d(newton), C.c_int(4096), _i(cg), C.c_int(4096),
                             info, _d(lu), C.c_int(256))
        if rc:
            raise RuntimeError("oracle static step failed rc=%d" % rc)
        ncon = len(self.arrays.get("con_node1", ()))
        lam = np.zeros(ncon)
        if ncon and L.or_last_lagrange(_d(lam), C.c_int(ncon)):
            raise RuntimeError("oracle: Lagrange multipliers unavailable")
        return {"dof_total": u, "rforce": rf, "newton": newton[:info[2]].copy(), "cg_iterations": cg[:info[3]].copy(),
                "lagrange_multipliers": lam,
                "n_steps": int(info[0]), "n_cutbacks": int(info[1]),
                "profile": [int(info[k]) for k in range(4, 9)], "lu": lu[:info[9]].copy(), "state_variables": sv[:nsv].copy()}


class OracleSystem:
    """kind 1 = skyline direct (solver_direct.f90), 2 = COO conjugate gradient (solver_conjugate_gradient.f90)"""

    def __init__(self, model, kind=2):
        self.m = model
        self.kind = kind
        self.h = C.c_void_p(lib().or_system_new(model.h, C.c_int(kind)))

    def close(self):
        if self.h:
            lib().or_system_free(self.h)
            self.h = None

    def assemble(self, dof_total, dof_increment, time=0.0, dtime=1.0):
        ut, du = np.ascontiguousarray(dof_total, float), np.ascontiguousarray(dof_increment, float)
        rf, norms = np.zeros(self.m.ndof), np.zeros(3)
        fail = lib().or_system_assemble(self.h, _d(ut), _d(du), C.c_double(time), C.c_double(dtime), _d(rf), _d(norms), None)
        return rf, norms[0], fail

    def apply_bc(self, dof_total, dof_increment, time=0.0, dtime=1.0):
        ut, du = np.ascontiguousarray(dof_total, float), np.ascontiguousarray(dof_increment, float)
        norms = np.zeros(3)
        lib().or_system_apply_bc(self.h, _d(ut), _d(du), C.c_double(time), C.c_double(dtime), _d(norms))
        return norms[1]

    def matrix(self):
        """assembled matrix as scipy CSR (full symmetric) + rhs"""
        import scipy.sparse as sp
        n = self.m.ndof
        nz = lib().or_system_nnz_upper(self.h)
        rows, cols, a = np.zeros(nz, np.int32), np.zeros(nz, np.int32), np.zeros(nz)
        diag, rhs = np.zeros(n), np.zeros(n)
        lib().or_system_get_coo(self.h, _i(rows), _i(cols), _d(a), _d(diag), _d(rhs))
        A = sp.coo_matrix((a, (rows, cols)), shape=(n, n)).tocsr()
        return (A + A.T + sp.diags(diag)).tocsr(), rhs

    def rhs(self):
        n = self.m.ndof
        diag, rhs = np.zeros(n), np.zeros(n)
        lib().or_system_get_coo(self.h, None, None, None, _d(diag), _d(rhs))
        return rhs, diag

    def solve(self, dof_increment, tol=1e-17, maxit=1000, fixed_iters=-1):
        du = np.ascontiguousarray(dof_increment, float).copy()
        norms, its = np.zeros(3), C.c_int()
        fail = lib().or_system_solve(self.h, C.c_double(tol), C.c_int(maxit), C.c_int(fixed_iters), _d(du), _d(norms), C.byref(its))
        return du, norms[2], its.value, fail


def element(flag, coords24, dof_increment24, dof_total24, props, jdltyp=None, adlmag=None):
    K, R, sv, en = np.zeros(576), np.zeros(24), np.zeros(48), np.zeros(8)
    jd = np.ascontiguousarray(jdltyp if jdltyp is not None else [], np.int32)
    ad = np.ascontiguousarray(adlmag if adlmag is not None else [], float)
    rc = lib().or_element(C.c_int(flag), _d(np.ascontiguousarray(coords24, float)), _d(np.ascontiguousarray(dof_increment24, float)),
                          _d(np.ascontiguousarray(dof_total24, float)), _d(np.ascontiguousarray(props, float)), C.c_int(jd.size),
                          _i(jd), _d(ad), _d(K), _d(R), _d(sv), _d(en))
    if rc < 0:
        raise ValueError("unknown element flag")
    return K.reshape(24, 24).T.copy(), R, sv.reshape(8, 6), en, rc


def check_stiffness(flag, coords24, dof_increment24, dof_total24, props):
    err, asym = C.c_double(), C.c_double()
    rc = lib().or_check_stiffness(C.c_int(flag), _d(np.ascontiguousarray(coords24, float)),
                                  _d(np.ascontiguousarray(dof_increment24, float)), _d(np.ascontiguousarray(dof_total24, float)),
                                  _d(np.ascontiguousarray(props, float)), C.byref(err), C.byref(asym))
    if rc:
        raise RuntimeError("oracle element failed")
    return err.value, asym.value


class OracleDynamic:
    def __init__(self, model, dof_total0=None, velocity0=None):
        self.model = model
        u0 = None if dof_total0 is None else np.ascontiguousarray(dof_total0, np.float64)
        v0 = None if velocity0 is None else np.ascontiguousarray(velocity0, np.float64)
        self.h = C.c_void_p(lib().or_dynamic_new(model.h, _d(u0), _d(v0)))

    def steps(self, n):
        rc = lib().or_dynamic_steps(self.h, C.c_int(n))
        if rc:
            raise RuntimeError("oracle explicit dynamic step failed rc=%d" % rc)
        return self

    def get(self):
        nd = self.model.ndof
        out = {k: np.zeros(nd) for k in ("dof_total", "dof_increment", "velocity", "acceleration", "rforce", "lumped_mass")}
        ts = np.zeros(2)
        lib().or_dynamic_get(self.h, _d(out["dof_total"]), _d(out["dof_increment"]), _d(out["velocity"]), _d(out["acceleration"]),
                             _d(out["rforce"]), _d(out["lumped_mass"]), _d(ts))
        out["time"], out["steps"] = float(ts[0]), int(ts[1])
        return out

    def close(self):
        if self.h:
            lib().or_dynamic_free(self.h)
            self.h = None


def synthetic_block_arrays(nx, ny=None, nz=None, jitter=0.2, seed=1234, E=100.0, nu=0.3, cg_tolerance=1e-17, cg_maxit=1000,
                           solvertype=2):
    """The oracle's own model of the synthetic hex8 block (oracle/or_mesh.cpp + the boundary-condition tables below): face
    x = 0 clamped, traction (0, 0, -1) on face 4 of the elements at x = max, one LINEAR static step — the same arrays the
    product's parser produces for synthetic_block_input(...) (tests/test_oracle_cpu.py checks the equality), built without
    loading any product library."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    N, Ne = (nx + 1) * (ny + 1) * (nz + 1), nx * ny * nz
    coords, conn = np.zeros(3 * N), np.zeros(8 * Ne, np.int32)
    L = lib()
    L.or_synthetic_block.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_ulonglong, dp, ip]
    if L.or_synthetic_block(nx, ny, nz, float(jitter), int(seed), _d(coords), _i(conn)):
        raise ValueError("bad block dimensions")
    i32 = lambda v: np.asarray(v, np.int32)
    node = (1 + np.arange(N)).reshape(nz + 1, ny + 1, nx + 1)
    elem = (1 + np.arange(Ne)).reshape(nz, ny, nx)
    nsets = [node[:, :, 0], node[:, :, nx], node[:, 0, :], node[:, ny, :], node[0], node[nz]]
    esets = [elem[:, :, 0], elem[:, :, nx - 1], elem[:, 0, :], elem[:, ny - 1, :], elem[0], elem[nz - 1]]
    z = i32([])
    return {
        "n_nodes": i32([N]), "n_elements": i32([Ne]), "coords": coords, "connectivity": conn,
        "element_flag": np.full(Ne, 1001, np.int32), "element_prop_index": np.zeros(Ne, np.int32),
        "element_n_props": np.full(Ne, 2, np.int32), "element_n_states": np.zeros(Ne, np.int32),
        "element_material": np.zeros(Ne, np.int32), "material_prop_index": z, "material_n_props": z,
        "element_properties": np.array([E, nu]), "material_properties": np.zeros(0), "element_density": np.zeros(Ne),
        "history_ptr": i32([0]), "history_time": np.zeros(0), "history_value": np.zeros(0),
        "nodeset_ptr": i32(np.concatenate([[0], np.cumsum([s.size for s in nsets])])),
        "nodeset_nodes": i32(np.concatenate([s.ravel() for s in nsets])),
        "elementset_ptr": i32(np.concatenate([[0], np.cumsum([s.size for s in esets])])),
        "elementset_elements": i32(np.concatenate([s.ravel() for s in esets])),
        # DEGREES OF FREEDOM: node set 1 (xmin), DOFs 1..3, VALUE 0
        "pdof_flag": i32([1, 1, 1]), "pdof_dof": i32([1, 2, 3]), "pdof_node": i32([0, 0, 0]), "pdof_nodeset": i32([1, 1, 1]),
        "pdof_history": i32([0, 0, 0]), "pdof_rate": i32([0, 0, 0]), "pdof_subparam": i32([0, 0, 0]), "subparam_ptr": i32([0]),
        "pdof_value": np.zeros(3), "subparam_values": np.zeros(0),
        "pforce_flag": z, "pforce_dof": z, "pforce_node": z, "pforce_nodeset": z, "pforce_history": z, "pforce_value": np.zeros(0),
        # DISTRIBUTED LOADS: element set 2 (xmax_el), face 4, VALUE (0, 0, -1)
        "dload_flag": i32([1]), "dload_elset": i32([2]), "dload_face": i32([4]), "dload_history": i32([0]),
        "dload_val_ptr": i32([0, 3]), "dload_values": np.array([0.0, 0.0, -1.0]),
        "max_no_steps": i32([1]), "solvertype": i32([solvertype]), "nonlinear": i32([0]), "max_newton_iterations": i32([1]),
        "unsymmetric": i32([0]), "abaqusformat": i32([0]), "max_cg_iterations": i32([cg_maxit]), "checkstiffness_flag": i32([0]),
        "explicitdynamicstep": i32([0]), "no_dynamic_steps": i32([0]),
        "timestep_initial": np.array([1.0]), "timestep_min": np.array([1.0 / 1024.0]), "timestep_max": np.array([1.0]),
        "max_total_time": np.array([1.0]), "newtonraphson_tolerance": np.array([1e-8]), "cg_tolerance": np.array([cg_tolerance]),
        "dynamic_timestep": np.array([0.0]),
    }
