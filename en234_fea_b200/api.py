"""ctypes bindings of the two C ABIs (include/en234gpu.h, include/en234host.h).

Python is only glue here (tests, bench, torch.distributed plumbing): every numerical step runs inside
liben234gpu.so on the device.  There is no CPU fallback — creating a GpuSystem without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Stats(C.Structure):
    _fields_ = [("assemble_ms", C.c_double), ("bc_ms", C.c_double), ("solve_ms", C.c_double),
                ("spmv_ms_avg", C.c_double), ("kernel_launches", C.c_longlong), ("last_cg_iterations", C.c_int),
                ("last_cg_err", C.c_double), ("mf_congruent", C.c_int)]


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


_gpu = None
_host = None

GPU_SYMBOLS = [
    "en234gpu_last_error", "en234gpu_device_count", "en234gpu_create", "en234gpu_create_nodes", "en234gpu_destroy", "en234gpu_set_stream",
    "en234gpu_assemble", "en234gpu_apply_boundaryconditions", "en234gpu_solve", "en234gpu_set_operator_mode",
    "en234gpu_upload_state", "en234gpu_assemble_dev", "en234gpu_apply_boundaryconditions_dev", "en234gpu_solve_dev",
    "en234gpu_download_state", "en234gpu_zero_increment_dev", "en234gpu_element_batch", "en234gpu_element_single", "en234gpu_get_svars",
    "en234gpu_nnz_blocks", "en234gpu_get_csr", "en234gpu_get_rhs", "en234gpu_apply_operator",
    "en234gpu_get_unique_id", "en234gpu_comm_init", "en234gpu_set_partition", "en234gpu_get_stats",
    "en234gpu_bench_operator", "en234gpu_comm_arena", "en234gpu_comm_connect",
    "en234gpu_measure_fp64_peak", "en234gpu_n_state_variables", "en234gpu_commit_state", "en234gpu_get_state_variables", "en234gpu_set_state_variables",
    "en234gpu_project_fields", "en234gpu_dynamic_setup", "en234gpu_dynamic_set_state", "en234gpu_dynamic_steps",
    "en234gpu_dynamic_half_step", "en234gpu_dynamic_get_state",
    "en234gpu_create_multi", "en234gpu_multi_destroy", "en234gpu_multi_set_operator_mode", "en234gpu_multi_assemble",
    "en234gpu_multi_apply_boundaryconditions", "en234gpu_multi_solve", "en234gpu_multi_upload_state",
    "en234gpu_multi_assemble_dev", "en234gpu_multi_apply_boundaryconditions_dev", "en234gpu_multi_solve_dev",
    "en234gpu_multi_zero_increment_dev", "en234gpu_multi_download_state", "en234gpu_multi_n_parts",
    "en234gpu_multi_part_handle", "en234gpu_multi_part_info", "en234gpu_multi_get_stats",
    "en234gpu_partition_sizes", "en234gpu_partition_arrays",
    "en234gpu_set_ties", "en234gpu_get_lagrange_multipliers", "en234gpu_reset_lagrange_multipliers", "en234gpu_comm_allreduce_host", "en234gpu_tie_plan", "en234gpu_bench_collective",
]
HOST_SYMBOLS = [
    "en234host_model_read", "en234host_model_from_text", "en234host_model_free", "en234host_last_error",
    "en234host_model_int", "en234host_model_double", "en234host_model_find_set", "en234host_model_set",
    "en234host_model_write_text", "en234host_evaluate_loads", "en234host_gpu_create", "en234host_run_static",
    "en234host_partition", "en234host_part_free", "en234host_part_int", "en234host_part_owned",
    "en234host_part_gpu_create", "en234host_run_static_part", "en234host_check_stiffness",
    "en234host_model_get_state_print", "en234host_model_set_state_print", "en234host_project_fields", "en234host_print_state",
    "en234host_run_dynamic", "en234host_run_dynamic_part", "en234host_dynamic_tables",
    "en234host_model_device_support",
]
# Fortran-linkage twins of the reference plugin entry points (include/en234host.h)
PLUGIN_SYMBOLS = ["user_element_static_", "uel_"]


def gpu_lib():
    """liben234gpu.so (raises if it was not built: the product never falls back to a CPU path)."""
    global _gpu
    if _gpu is None:
        path = os.environ.get("EN234_GPU_LIB") or os.path.join(_CSRC, "liben234gpu.so")   # override: A/B builds of kernel variants
        if not os.path.exists(path):
            raise RuntimeError("liben234gpu.so is not built (run python -c 'import __graft_entry__ as g; g.build()')")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.en234gpu_last_error.restype = C.c_char_p
        L.en234gpu_nnz_blocks.restype = C.c_longlong
        L.en234gpu_nnz_blocks.argtypes = [C.c_void_p]
        _gpu = L
    return _gpu


def host_lib():
    global _host
    if _host is None:
        gpu_lib()
        path = os.path.join(_CSRC, "liben234host.so")
        if not os.path.exists(path):
            raise RuntimeError("liben234host.so is not built")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.en234host_last_error.restype = C.c_char_p
        L.en234host_model_read.restype = C.c_void_p
        L.en234host_model_read.argtypes = [C.c_char_p]
        L.en234host_model_from_text.restype = C.c_void_p
        L.en234host_model_from_text.argtypes = [C.c_char_p]
        L.en234host_model_free.argtypes = [C.c_void_p]
        L.en234host_model_write_text.restype = C.c_long
        L.en234host_model_write_text.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
        L.en234host_partition.restype = C.c_void_p
        L.en234host_partition.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.en234host_part_free.argtypes = [C.c_void_p]
        L.en234host_part_owned.restype = C.POINTER(C.c_ubyte)
        L.en234host_part_owned.argtypes = [C.c_void_p]
        _host = L
    return _host


INT_ARRAYS = ["connectivity", "element_flag", "element_prop_index", "element_n_props", "element_n_states",
              "element_material", "material_prop_index", "material_n_props",
              "history_ptr", "nodeset_ptr", "nodeset_nodes", "elementset_ptr", "elementset_elements",
              "pdof_flag", "pdof_dof", "pdof_node", "pdof_nodeset", "pdof_history", "pdof_rate", "pdof_subparam", "subparam_ptr",
              "pforce_flag", "pforce_dof", "pforce_node", "pforce_nodeset", "pforce_history",
              "dload_flag", "dload_elset", "dload_face", "dload_history", "dload_val_ptr",
              "con_node1", "con_dof1", "con_node2", "con_dof2"]
INT_SCALARS = ["n_nodes", "n_elements", "nodes_per_element", "max_no_steps", "solvertype", "nonlinear", "max_newton_iterations",
               "unsymmetric", "abaqusformat", "max_cg_iterations", "checkstiffness_flag", "explicitdynamicstep", "no_dynamic_steps"]
DBL_ARRAYS = ["coords", "element_properties", "material_properties", "history_time", "history_value", "pdof_value", "pforce_value",
              "dload_values", "subparam_values", "element_density"]
DBL_SCALARS = ["timestep_initial", "timestep_min", "timestep_max", "max_total_time", "newtonraphson_tolerance",
               "cg_tolerance", "dynamic_timestep"]


class Model:
    """A parsed EN234_FEA input file (reference: src/read_input_file.f90)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("input error: " + host_lib().en234host_last_error().decode())
        self.h = C.c_void_p(handle)

    @classmethod
    def read(cls, path):
        return cls(host_lib().en234host_model_read(os.fsencode(path)))

    @classmethod
    def from_text(cls, text):
        return cls(host_lib().en234host_model_from_text(text.encode()))

    def close(self):
        if self.h:
            host_lib().en234host_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ints(self, name):
        p, n = c_int_p(), C.c_long()
        if host_lib().en234host_model_int(self.h, name.encode(), C.byref(p), C.byref(n)):
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int32)

    def doubles(self, name):
        p, n = c_double_p(), C.c_long()
        if host_lib().en234host_model_double(self.h, name.encode(), C.byref(p), C.byref(n)):
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float64)

    def set(self, name, value):
        if host_lib().en234host_model_set(self.h, name.encode(), C.c_double(value)):
            raise KeyError(name)

    def arrays(self):
        """All flat arrays/scalars as a dict of numpy arrays (the schema the oracle's or_model_set_* accepts)."""
        d = {k: self.ints(k) for k in INT_ARRAYS + INT_SCALARS}
        d.update({k: self.doubles(k) for k in DBL_ARRAYS + DBL_SCALARS})
        return d

    @property
    def n_nodes(self):
        return int(self.ints("n_nodes")[0])

    @property
    def n_elements(self):
        return int(self.ints("n_elements")[0])

    def text(self):
        n = host_lib().en234host_model_write_text(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        host_lib().en234host_model_write_text(self.h, buf, n + 1)
        return buf.value.decode()

    def find_set(self, name, element_set=False):
        return host_lib().en234host_model_find_set(self.h, 1 if element_set else 0, name.encode())

    def check_stiffness(self, element_flag=None, dof_total=None, dof_increment=None, log_path=None):
        """check_stiffness (src/checkstiffness.f90) on the device twins: returns (err, n_unsymmetric, n_numerical_unsymmetric)."""
        flag = int(self.ints("checkstiffness_flag")[0]) if element_flag is None else int(element_flag)
        err, nu, nnu = C.c_double(), C.c_int(), C.c_int()
        rc = host_lib().en234host_check_stiffness(self.h, C.c_int(flag), _dp(_f64(dof_total)), _dp(_f64(dof_increment)),
                                                  os.fsencode(log_path) if log_path else None, C.byref(err), C.byref(nu), C.byref(nnu))
        if rc:
            raise RuntimeError("check_stiffness failed: " + host_lib().en234host_last_error().decode())
        return err.value, nu.value, nnu.value

    def state_print(self):
        """The PRINT STATE block as parsed (read_input_file.f90:1772-1866)."""
        flags, steps, scale = C.c_int(), C.c_int(), C.c_double()
        names, fname = C.create_string_buffer(4096), C.create_string_buffer(1024)
        host_lib().en234host_model_get_state_print(self.h, C.byref(flags), C.byref(steps), C.byref(scale), names, C.c_int(4096),
                                                   fname, C.c_int(1024))
        f = flags.value
        return {"stateprint": bool(f & 1), "print_dof": bool(f & 2), "displaced_mesh": bool(f & 4), "lumped": bool(f & 8),
                "state_print_steps": steps.value, "displacement_scale_factor": scale.value,
                "field_variables": [n for n in names.value.decode().split(",") if n], "file_name": fname.value.decode()}

    def set_state_print(self, path, field_variables=None, print_dof=None, displaced_mesh=None, lumped=None,
                        state_print_steps=0, displacement_scale_factor=None, enabled=True):
        """Turn the state prints of run_static on (path) or off (path=None); unspecified options keep their parsed values."""
        cur = self.state_print()
        pick = lambda v, k: cur[k] if v is None else v
        flags = (1 if enabled else 0) | (2 if pick(print_dof, "print_dof") else 0) | (4 if pick(displaced_mesh, "displaced_mesh") else 0) \
            | (8 if pick(lumped, "lumped") else 0)
        names = None if field_variables is None else ",".join(field_variables).encode()
        rc = host_lib().en234host_model_set_state_print(self.h, C.c_int(flags), C.c_int(state_print_steps),
                                                        C.c_double(pick(displacement_scale_factor, "displacement_scale_factor")),
                                                        names, os.fsencode(path) if path else None)
        if rc:
            raise RuntimeError(host_lib().en234host_last_error().decode())

    def device_support(self):
        """None when the device implements every element of the model, else the reason (en234host_model_device_support)."""
        if host_lib().en234host_model_device_support(self.h) == 0:
            return None
        return host_lib().en234host_last_error().decode()

    def dynamic_tables(self, part=None):
        """The explicit-dynamics load tables (apply_dynamic_boundaryconditions, time-independent part) as handed to the device:
        list of (dense unit nodal forces, history id, constant), prescribed DOFs (dof, history, value, mode), element densities."""
        L = host_lib()
        nd = 3 * (part.ints("n_nodes")[0] if part is not None else self.n_nodes)
        ne = part.ints("n_elements")[0] if part is not None else self.n_elements
        ph = part.h if part is not None else None
        cap = int(nd)
        n_pd = C.c_int()
        pd, phs, pm, pv, dens = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap), np.zeros(ne)
        n_loads = L.en234host_dynamic_tables(self.h, ph, C.c_int(-1), None, None, None, C.byref(n_pd), _ip(pd), _ip(phs), _dp(pv),
                                             _ip(pm), C.c_int(cap), _dp(dens))
        if n_loads < 0:
            raise RuntimeError(L.en234host_last_error().decode())
        loads = []
        for l in range(n_loads):
            dense, h, c = np.zeros(nd), C.c_int(), C.c_double()
            L.en234host_dynamic_tables(self.h, ph, C.c_int(l), _dp(dense), C.byref(h), C.byref(c), None, None, None, None, None,
                                       C.c_int(0), None)
            loads.append((dense, h.value, c.value))
        k = n_pd.value
        return {"loads": loads, "pdof": (pd[:k].copy(), phs[:k].copy(), pv[:k].copy(), pm[:k].copy()), "element_density": dens}

    def evaluate_loads(self, time, dtime, dof_total=None, dof_increment=None):
        nd = 3 * self.n_nodes
        ut = _f64(dof_total) if dof_total is not None else np.zeros(nd)
        du = _f64(dof_increment) if dof_increment is not None else np.zeros(nd)
        fe, fx = np.zeros(nd), np.zeros(nd)
        cap = int(self.ints("nodeset_nodes").size + self.ints("pdof_flag").size + 16)
        fd, fv, fc = np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap)
        n = C.c_int()
        rc = host_lib().en234host_evaluate_loads(self.h, C.c_double(time), C.c_double(dtime), _dp(ut), _dp(du), _dp(fe),
                                                 _dp(fx), C.byref(n), _ip(fd), _dp(fv), _dp(fc), C.c_int(cap))
        if rc:
            raise RuntimeError(host_lib().en234host_last_error().decode())
        return fe, fx, fd[:n.value].copy(), fv[:n.value].copy(), fc[:n.value].copy()


class GpuSystem:
    """One device handle (en234gpu_create ... en234gpu_destroy)."""

    def __init__(self, model=None, device=0, arrays=None, part=None):
        L = gpu_lib()
        self.h = C.c_void_p()
        self.model = model
        self.part = part
        if part is not None:
            rc = host_lib().en234host_part_gpu_create(model.h, part.h, C.c_int(device), C.byref(self.h))
            err = host_lib().en234host_last_error
            self.n_nodes = part.n_nodes
        elif arrays is not None:
            co, cn = _f64(arrays["coords"]), _i32(arrays["connectivity"])
            fl, pi, pr = _i32(arrays["element_flag"]), _i32(arrays["element_prop_index"]), _f64(arrays["element_properties"])
            self.n_nodes = co.size // 3
            rc = L.en234gpu_create(C.byref(self.h), C.c_int(device), C.c_int(self.n_nodes), C.c_int(fl.size), _dp(co), _ip(cn),
                                   _ip(fl), _ip(pi), _dp(pr), C.c_int(pr.size))
            err = L.en234gpu_last_error
        else:
            rc = host_lib().en234host_gpu_create(model.h, C.c_int(device), C.byref(self.h))
            err = host_lib().en234host_last_error
            self.n_nodes = model.n_nodes
        if rc:
            self.h = None
            raise RuntimeError("en234gpu_create failed: " + err().decode())
        self.ndof = 3 * self.n_nodes

    def _ck(self, rc):
        if rc:
            raise RuntimeError(gpu_lib().en234gpu_last_error().decode())

    def close(self):
        if self.h:
            gpu_lib().en234gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_operator_mode(self, method):
        self._ck(gpu_lib().en234gpu_set_operator_mode(self.h, C.c_int(method)))

    def set_stream(self, cuda_stream):
        self._ck(gpu_lib().en234gpu_set_stream(self.h, C.c_void_p(cuda_stream)))

    # ---- host-buffer entry points
    def assemble(self, dof_total, dof_increment, f_element_ext=None, want_rforce=True, out_rforce=None):
        ut, du, fe = _f64(dof_total), _f64(dof_increment), _f64(f_element_ext)
        rf = out_rforce if out_rforce is not None else (np.zeros(self.ndof) if want_rforce else None)
        nrm, fail = C.c_double(), C.c_int()
        self._ck(gpu_lib().en234gpu_assemble(self.h, _dp(ut), _dp(du), _dp(fe), _dp(rf), C.byref(nrm), C.byref(fail)))
        return rf, nrm.value, fail.value

    def apply_boundaryconditions(self, f_ext, fixed_dof, fixed_correction):
        fx, fd, fc = _f64(f_ext), _i32(fixed_dof), _f64(fixed_correction)
        nrm = C.c_double()
        self._ck(gpu_lib().en234gpu_apply_boundaryconditions(self.h, _dp(fx), C.c_int(fd.size), _ip(fd), _dp(fc), C.byref(nrm)))
        return nrm.value

    def solve(self, dof_increment, method=0, cg_tolerance=1e-17, max_cg_iterations=1000, inplace=False):
        du = _f64(dof_increment) if inplace else _f64(dof_increment).copy()
        nrm, its, fail = C.c_double(), C.c_int(), C.c_int()
        self._ck(gpu_lib().en234gpu_solve(self.h, C.c_int(method), C.c_double(cg_tolerance), C.c_int(max_cg_iterations),
                                          _dp(du), C.byref(nrm), C.byref(its), C.byref(fail)))
        return du, nrm.value, its.value, fail.value

    # ---- device-resident variants
    def upload_state(self, dof_total, dof_increment, f_element_ext, f_ext, fixed_dof, fixed_value):
        ut, du, fe, fx = _f64(dof_total), _f64(dof_increment), _f64(f_element_ext), _f64(f_ext)
        fd, fv = _i32(fixed_dof), _f64(fixed_value)
        self._ck(gpu_lib().en234gpu_upload_state(self.h, _dp(ut), _dp(du), _dp(fe), _dp(fx), C.c_int(fd.size), _ip(fd), _dp(fv)))

    def assemble_dev(self):
        nrm, fail = C.c_double(), C.c_int()
        self._ck(gpu_lib().en234gpu_assemble_dev(self.h, C.byref(nrm), C.byref(fail)))
        return nrm.value, fail.value

    def apply_boundaryconditions_dev(self):
        nrm = C.c_double()
        self._ck(gpu_lib().en234gpu_apply_boundaryconditions_dev(self.h, C.byref(nrm)))
        return nrm.value

    def solve_dev(self, method=0, cg_tolerance=1e-17, max_cg_iterations=1000):
        nrm, its, fail = C.c_double(), C.c_int(), C.c_int()
        self._ck(gpu_lib().en234gpu_solve_dev(self.h, C.c_int(method), C.c_double(cg_tolerance), C.c_int(max_cg_iterations),
                                              C.byref(nrm), C.byref(its), C.byref(fail)))
        return nrm.value, its.value, fail.value

    def download_state(self, out_du=None, out_rf=None):
        du = out_du if out_du is not None else np.zeros(self.ndof)
        rf = out_rf if out_rf is not None else np.zeros(self.ndof)
        self._ck(gpu_lib().en234gpu_download_state(self.h, _dp(du), _dp(rf)))
        return du, rf

    def zero_increment_dev(self):
        self._ck(gpu_lib().en234gpu_zero_increment_dev(self.h))

    # ---- plugin surface / inspection
    def element_batch(self, element_numbers, dof_total, dof_increment):
        el = _i32(element_numbers)
        K, R = np.zeros((el.size, 576)), np.zeros((el.size, 24))
        fail = C.c_int()
        self._ck(gpu_lib().en234gpu_element_batch(self.h, C.c_int(el.size), _ip(el), _dp(_f64(dof_total)), _dp(_f64(dof_increment)),
                                                  _dp(K), _dp(R), C.byref(fail)))
        # column-major (ROW,COL) -> K[e][row, col]
        return K.reshape(el.size, 24, 24).transpose(0, 2, 1).copy(), R, fail.value

    def svars(self, dof_total, dof_increment, n_elements):
        sv = np.zeros(48 * n_elements)
        self._ck(gpu_lib().en234gpu_get_svars(self.h, _dp(_f64(dof_total)), _dp(_f64(dof_increment)), _dp(sv)))
        return sv.reshape(n_elements, 8, 6)

    def nnz_blocks(self):
        return gpu_lib().en234gpu_nnz_blocks(self.h)

    def state_variables(self, updated=True):
        L = gpu_lib()
        L.en234gpu_n_state_variables.restype = C.c_longlong
        sv = np.zeros(L.en234gpu_n_state_variables(self.h))
        self._ck(L.en234gpu_get_state_variables(self.h, C.c_int(1 if updated else 0), _dp(sv)))
        return sv

    def commit_state(self):
        self._ck(gpu_lib().en234gpu_commit_state(self.h))

    def project_fields(self, names, dof_total=None, dof_increment=None, lumped=False):
        """Nodal projection of integration-point fields (print_state, src/printing.f90:457-490) on the device:
        (n_nodes, n_fields).  dof arrays None: the device-resident state of the last step."""
        out = np.zeros((len(names), self.model.n_nodes))
        ut = None if dof_total is None else _f64(dof_total)
        du = None if dof_increment is None else _f64(dof_increment)
        if ut is not None and du is None:
            du = np.zeros_like(ut)
        rc = host_lib().en234host_project_fields(self.model.h, self.h, ",".join(names).encode(), C.c_int(1 if lumped else 0),
                                                 _dp(ut), _dp(du), _dp(out))
        if rc:
            raise RuntimeError("project_fields failed: " + host_lib().en234host_last_error().decode())
        return out.T.copy()

    def run_dynamic(self, n_steps=0, initial_velocity=None):
        """explicit_dynamic_step (src/explicit_dynamic_step.f90) for an input with an EXPLICIT DYNAMIC STEP block; n_steps=0
        runs the input's NUMBER OF STEPS.  State prints go to the path set with Model.set_state_print."""
        nd = self.ndof
        out = {k: np.zeros(nd) for k in ("dof_total", "velocity", "acceleration", "rforce", "lumped_mass")}
        info = (C.c_longlong * 8)()
        ms = C.c_double()
        v0 = None if initial_velocity is None else _f64(initial_velocity)
        if self.part is not None:      # rank-local arrays; collective over all ranks
            rc = host_lib().en234host_run_dynamic_part(self.model.h, self.part.h, self.h, C.c_int(n_steps), _dp(v0),
                                                       _dp(out["dof_total"]), _dp(out["velocity"]), _dp(out["acceleration"]),
                                                       _dp(out["rforce"]), _dp(out["lumped_mass"]), info, C.byref(ms))
        else:
            rc = host_lib().en234host_run_dynamic(self.model.h, self.h, C.c_int(n_steps), _dp(v0), _dp(out["dof_total"]),
                                                  _dp(out["velocity"]), _dp(out["acceleration"]), _dp(out["rforce"]),
                                                  _dp(out["lumped_mass"]), info, C.byref(ms))
        if rc:
            raise RuntimeError("explicit dynamic step failed: " + host_lib().en234host_last_error().decode())
        out.update(steps=int(info[0]), n_state_prints=int(info[1]), kernel_launches=int(info[2]), device_ms=ms.value)
        return out

    def dynamic_steps(self, dt, n_steps):
        """n_steps more central-difference steps from the device-resident state (after run_dynamic / dynamic set-up)."""
        fail, ms = C.c_int(), C.c_double()
        self._ck(gpu_lib().en234gpu_dynamic_steps(self.h, C.c_double(dt), C.c_int(n_steps), C.byref(fail), C.byref(ms)))
        return fail.value, ms.value

    def dynamic_state(self):
        nd = self.ndof
        out = {k: np.zeros(nd) for k in ("dof_total", "dof_increment", "velocity", "acceleration", "rforce", "lumped_mass")}
        t = C.c_double()
        self._ck(gpu_lib().en234gpu_dynamic_get_state(self.h, _dp(out["dof_total"]), _dp(out["dof_increment"]), _dp(out["velocity"]),
                                                      _dp(out["acceleration"]), _dp(out["rforce"]), _dp(out["lumped_mass"]),
                                                      C.byref(t)))
        out["time"] = t.value
        return out

    def project_field_codes(self, codes, dof_total=None, dof_increment=None, lumped=False):
        """en234gpu_project_fields with numeric field codes (include/en234gpu.h) on this handle's own (rank-local when
        partitioned) numbering: (n_local_nodes, n_fields).  Collective on partitioned handles."""
        n = self.ndof // 3
        out = np.zeros((len(codes), n))
        ut = None if dof_total is None else _f64(dof_total)
        du = None if dof_increment is None else _f64(dof_increment)
        if ut is not None and du is None:
            du = np.zeros_like(ut)
        its = C.c_int()
        self._ck(gpu_lib().en234gpu_project_fields(self.h, C.c_int(len(codes)), _ip(_i32(codes)), C.c_int(1 if lumped else 0),
                                                   C.c_double(0.0), C.c_int(0), _dp(ut), _dp(du), _dp(out), C.byref(its)))
        return out.T.copy()

    def print_state(self, path, dof_total, dof_increment=None, time_end=1.0, append=False):
        """One Tecplot zone of the current PRINT STATE options for the given state."""
        ut = _f64(dof_total)
        du = None if dof_increment is None else _f64(dof_increment)
        rc = host_lib().en234host_print_state(self.model.h, self.h, C.c_double(time_end), _dp(ut), _dp(du), os.fsencode(path),
                                              C.c_int(1 if append else 0))
        if rc:
            raise RuntimeError("print_state failed: " + host_lib().en234host_last_error().decode())

    def csr(self):
        import scipy.sparse as sp
        nnz = 9 * self.nnz_blocks()
        rowptr, col, val = np.zeros(self.ndof + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz)
        self._ck(gpu_lib().en234gpu_get_csr(self.h, rowptr.ctypes.data_as(C.POINTER(C.c_longlong)), _ip(col), _dp(val)))
        return sp.csr_matrix((val, col, rowptr), shape=(self.ndof, self.ndof))

    def rhs(self, want_diag=False):
        r = np.zeros(self.ndof)
        d = np.zeros(self.ndof) if want_diag else None
        self._ck(gpu_lib().en234gpu_get_rhs(self.h, _dp(r), _dp(d)))
        return (r, d) if want_diag else r

    def apply_operator(self, x, method=0):
        y = np.zeros(self.ndof)
        self._ck(gpu_lib().en234gpu_apply_operator(self.h, C.c_int(method), _dp(_f64(x)), _dp(y)))
        return y

    def stats(self):
        s = Stats()
        self._ck(gpu_lib().en234gpu_get_stats(self.h, C.byref(s)))
        return s

    def fp64_peak(self):
        t = C.c_double()
        self._ck(gpu_lib().en234gpu_measure_fp64_peak(self.h, C.byref(t)))
        return t.value

    def bench_operator(self, method=0, warm=3, reps=20):
        ms = C.c_double()
        self._ck(gpu_lib().en234gpu_bench_operator(self.h, C.c_int(method), C.c_int(warm), C.c_int(reps), C.byref(ms)))
        return ms.value

    # ---- multi-GPU
    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        if gpu_lib().en234gpu_get_unique_id(buf):
            raise RuntimeError(gpu_lib().en234gpu_last_error().decode())
        return buf.raw

    def comm_init(self, n_ranks, rank, uid):
        self._ck(gpu_lib().en234gpu_comm_init(self.h, C.c_int(n_ranks), C.c_int(rank), C.c_char_p(uid)))

    def comm_arena(self, max_shared_nodes):
        buf = C.create_string_buffer(64)
        self._ck(gpu_lib().en234gpu_comm_arena(self.h, C.c_int(max_shared_nodes), buf))
        return buf.raw

    def comm_connect(self, handles):
        self._ck(gpu_lib().en234gpu_comm_connect(self.h, C.c_char_p(b"".join(handles))))

    def setup_distributed(self, use_peer_memory=True):
        """torch.distributed plumbing for a partitioned handle: NCCL communicator from a broadcast unique id, then
        (optionally) the peer-memory arena handles exchanged with an all-gather."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        uid = [GpuSystem.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        self.comm_init(world, rank, uid[0])
        if not use_peer_memory or os.environ.get("EN234_COMM", "") == "nccl" or world > 16:
            return "nccl"
        sp = self.part.ints("shared_ptr")
        mine = int(np.diff(sp).max()) if sp.size > 1 else 0
        sizes = [None] * world
        dist.all_gather_object(sizes, mine)
        handle = self.comm_arena(max(sizes))
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        self.comm_connect(handles)
        dist.barrier()
        return "peer-memory"

    # ---- whole static analysis through the host driver (src/staticstep.f90 loop)
    def run_static(self, method=0, cg_tolerance=0.0, max_cg_iterations=0, check_linear_cg=True, use_host_api=False,
                   log_path=None):
        nd = self.ndof
        u, rf = np.zeros(nd), np.zeros(nd)
        newton, cg = np.zeros((4096, 8)), np.zeros(4096, np.int32)
        info = (C.c_longlong * 8)()
        self.set_operator_mode(method)
        if self.part is not None:      # rank-local arrays; collective over all ranks
            rc = host_lib().en234host_run_static_part(self.model.h, self.part.h, self.h, C.c_int(method),
                                                      C.c_double(cg_tolerance), C.c_int(max_cg_iterations),
                                                      C.c_int(1 if check_linear_cg else 0), C.c_int(1 if use_host_api else 0),
                                                      _dp(u), _dp(rf), _dp(newton), C.c_int(4096), _ip(cg), C.c_int(4096), info,
                                                      os.fsencode(log_path) if log_path else None)
        else:
            rc = host_lib().en234host_run_static(self.model.h, self.h, C.c_int(method), C.c_double(cg_tolerance),
                                                 C.c_int(max_cg_iterations), C.c_int(1 if check_linear_cg else 0),
                                                 C.c_int(1 if use_host_api else 0), _dp(u), _dp(rf), _dp(newton), C.c_int(4096),
                                                 _ip(cg), C.c_int(4096), info, os.fsencode(log_path) if log_path else None)
        if rc:
            raise RuntimeError("static step failed: " + host_lib().en234host_last_error().decode())
        out = {"dof_total": u, "rforce": rf, "newton": newton[:info[2]].copy(), "cg_iterations": cg[:info[3]].copy(),
               "n_steps": int(info[0]), "n_cutbacks": int(info[1]), "n_state_prints": int(info[4])}
        ncon = self.model.ints("con_node1").size if self.model is not None and self.part is None else 0
        if ncon:                       # two-node tie constraints: the Lagrange multipliers the reference keeps in module Boundaryconditions
            out["lagrange_multipliers"] = self.lagrange_multipliers(ncon)
        return out

    def set_ties(self, dof1, dof2):
        """u[dof1[k]] = u[dof2[k]] (0-based DOFs): en234gpu_set_ties."""
        a, b = _i32(dof1), _i32(dof2)
        self._ck(gpu_lib().en234gpu_set_ties(self.h, C.c_int(a.size), _ip(a), _ip(b)))

    def lagrange_multipliers(self, n):
        lam = np.zeros(n)
        self._ck(gpu_lib().en234gpu_get_lagrange_multipliers(self.h, C.c_int(n), _dp(lam)))
        return lam



def tie_plan(dof1, dof2, fixed_dof=()):
    """Elimination plan of en234gpu_set_ties for a prescribed-DOF list (host code of liben234gpu.so; no device needed)."""
    L = gpu_lib()
    a, b, f = _i32(dof1), _i32(dof2), _i32(np.asarray(fixed_dof, np.int32))
    n = a.size
    sizes = np.zeros(4, np.int32)
    groot, gptr, gfix, mem = np.zeros(2 * n, np.int32), np.zeros(2 * n + 1, np.int32), np.zeros(2 * n, np.int32), np.zeros(2 * n, np.int32)
    ptie, pdof = np.zeros(n, np.int32), np.zeros(n, np.int32)
    if L.en234gpu_tie_plan(C.c_int(n), _ip(a), _ip(b), C.c_int(f.size), _ip(f) if f.size else None, _ip(sizes), _ip(groot), _ip(gptr),
                           _ip(gfix), _ip(mem), _ip(ptie), _ip(pdof)):
        raise RuntimeError(L.en234gpu_last_error().decode())
    ng, nm = int(sizes[0]), int(sizes[1])
    return {"group_root": groot[:ng].copy(), "group_ptr": gptr[:ng + 1].copy(), "group_fixed": gfix[:ng].copy(),
            "member_dof": mem[:nm].copy(), "peel_tie": ptie.copy(), "peel_dof": pdof.copy(), "n_touched": int(sizes[2])}


def library_partition(connectivity, n_nodes, n_parts, part):
    """The element partition en234gpu_create_multi builds (host code of liben234gpu.so; no device needed)."""
    L = gpu_lib()
    cn = _i32(connectivity)
    ne = cn.size // 8
    sizes = np.zeros(6, np.int32)
    if L.en234gpu_partition_sizes(C.c_int(n_parts), C.c_int(part), C.c_int(n_nodes), C.c_int(ne), _ip(cn), _ip(sizes)):
        raise RuntimeError(L.en234gpu_last_error().decode())
    nl, el, e0, nif, nn, ns = (int(x) for x in sizes)
    out = {"local_to_global_node": np.zeros(nl, np.int32), "connectivity": np.zeros(8 * el, np.int32),
           "neighbour_rank": np.zeros(nn, np.int32), "shared_ptr": np.zeros(nn + 1, np.int32),
           "shared_nodes": np.zeros(ns, np.int32), "owned": np.zeros(nl, np.uint8)}
    if L.en234gpu_partition_arrays(C.c_int(n_parts), C.c_int(part), C.c_int(n_nodes), C.c_int(ne), _ip(cn),
                                   _ip(out["local_to_global_node"]), _ip(out["connectivity"]), _ip(out["neighbour_rank"]),
                                   _ip(out["shared_ptr"]), _ip(out["shared_nodes"]),
                                   out["owned"].ctypes.data_as(C.POINTER(C.c_ubyte))):
        raise RuntimeError(L.en234gpu_last_error().decode())
    out.update(first_element=e0, n_interface=nif, local_elements=np.arange(e0, e0 + el, dtype=np.int32))
    return out


class MultiGpuSystem:
    """Several GPUs behind one handle in one process (en234gpu_create_multi): whole-mesh arrays in, whole-mesh arrays out.
    devices may repeat a device (several parts co-scheduled on one GPU: tests on a one-GPU box)."""

    def __init__(self, model=None, devices=(0, 1), arrays=None):
        L = gpu_lib()
        self.h = C.c_void_p()
        a = arrays if arrays is not None else model.arrays()
        co, cn = _f64(a["coords"]), _i32(a["connectivity"])
        fl, pi, pr = _i32(a["element_flag"]), _i32(a["element_prop_index"]), _f64(a["element_properties"])
        self.n_nodes = co.size // 3
        self.ndof = 3 * self.n_nodes
        dv = _i32(list(devices))
        rc = L.en234gpu_create_multi(C.byref(self.h), C.c_int(dv.size), _ip(dv), C.c_int(self.n_nodes), C.c_int(fl.size), _dp(co),
                                     _ip(cn), _ip(fl), _ip(pi), _dp(pr), C.c_int(pr.size))
        if rc:
            self.h = None
            raise RuntimeError("en234gpu_create_multi failed: " + L.en234gpu_last_error().decode())
        self.n_parts = dv.size

    def _ck(self, rc):
        if rc:
            raise RuntimeError(gpu_lib().en234gpu_last_error().decode())

    def close(self):
        if self.h:
            gpu_lib().en234gpu_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_operator_mode(self, method):
        self._ck(gpu_lib().en234gpu_multi_set_operator_mode(self.h, C.c_int(method)))

    def assemble(self, dof_total, dof_increment, f_element_ext=None):
        rf, nrm, fail = np.zeros(self.ndof), C.c_double(), C.c_int()
        self._ck(gpu_lib().en234gpu_multi_assemble(self.h, _dp(_f64(dof_total)), _dp(_f64(dof_increment)), _dp(_f64(f_element_ext)),
                                                   _dp(rf), C.byref(nrm), C.byref(fail)))
        return rf, nrm.value, fail.value

    def apply_boundaryconditions(self, f_ext, fixed_dof, fixed_correction):
        fd, nrm = _i32(fixed_dof), C.c_double()
        self._ck(gpu_lib().en234gpu_multi_apply_boundaryconditions(self.h, _dp(_f64(f_ext)), C.c_int(fd.size), _ip(fd),
                                                                   _dp(_f64(fixed_correction)), C.byref(nrm)))
        return nrm.value

    def solve(self, dof_increment, method=0, cg_tolerance=1e-17, max_cg_iterations=1000):
        du = _f64(dof_increment).copy()
        nrm, its, fail = C.c_double(), C.c_int(), C.c_int()
        self._ck(gpu_lib().en234gpu_multi_solve(self.h, C.c_int(method), C.c_double(cg_tolerance), C.c_int(max_cg_iterations),
                                                _dp(du), C.byref(nrm), C.byref(its), C.byref(fail)))
        return du, nrm.value, its.value, fail.value

    def upload_state(self, dof_total, dof_increment, f_element_ext, f_ext, fixed_dof, fixed_value):
        fd = _i32(fixed_dof)
        self._ck(gpu_lib().en234gpu_multi_upload_state(self.h, _dp(_f64(dof_total)), _dp(_f64(dof_increment)), _dp(_f64(f_element_ext)),
                                                       _dp(_f64(f_ext)), C.c_int(fd.size), _ip(fd), _dp(_f64(fixed_value))))

    def zero_increment_dev(self):
        self._ck(gpu_lib().en234gpu_multi_zero_increment_dev(self.h))

    def assemble_dev(self):
        nrm, fail = C.c_double(), C.c_int()
        self._ck(gpu_lib().en234gpu_multi_assemble_dev(self.h, C.byref(nrm), C.byref(fail)))
        return nrm.value, fail.value

    def apply_boundaryconditions_dev(self):
        nrm = C.c_double()
        self._ck(gpu_lib().en234gpu_multi_apply_boundaryconditions_dev(self.h, C.byref(nrm)))
        return nrm.value

    def solve_dev(self, method=0, cg_tolerance=1e-17, max_cg_iterations=1000):
        nrm, its, fail = C.c_double(), C.c_int(), C.c_int()
        self._ck(gpu_lib().en234gpu_multi_solve_dev(self.h, C.c_int(method), C.c_double(cg_tolerance), C.c_int(max_cg_iterations),
                                                    C.byref(nrm), C.byref(its), C.byref(fail)))
        return nrm.value, its.value, fail.value

    def download_state(self):
        du, rf = np.zeros(self.ndof), np.zeros(self.ndof)
        self._ck(gpu_lib().en234gpu_multi_download_state(self.h, _dp(du), _dp(rf)))
        return du, rf

    def stats(self):
        s = Stats()
        self._ck(gpu_lib().en234gpu_multi_get_stats(self.h, C.byref(s)))
        return s

    def part_info(self, part):
        v = [C.c_int() for _ in range(5)]
        self._ck(gpu_lib().en234gpu_multi_part_info(self.h, C.c_int(part), *[C.byref(x) for x in v]))
        return dict(zip(("device", "n_nodes", "n_elements", "n_interface", "n_neighbours"), (x.value for x in v)))


# ---- the reference's plugin entry points, called the way a gfortran-compiled driver would (everything by reference)
def user_element_static(element_identifier, element_coords, dof_increment, dof_total, element_properties,
                        initial_state_variables=None):
    """user_element_static (user_codes/src/user_element.f90:5-49) through its Fortran-linkage twin.
    Returns element_stiffness[row, col], element_residual, updated_state_variables, fail."""
    L = host_lib()
    ci = lambda v: C.byref(C.c_int(int(v)))
    xc, du, ut, pr = _f64(element_coords), _f64(dof_increment), _f64(dof_total), _f64(element_properties)
    n_nodes = xc.size // 3
    sv0 = _f64(initial_state_variables if initial_state_variables is not None else [])
    sv1 = np.zeros_like(sv0)
    node = np.zeros((n_nodes, 7), np.int32)                      # type(node), modules/Mesh.f90:5-14
    for k in range(n_nodes):
        node[k] = (0, 3 * k + 1, 3, 3 * k + 1, 3, 0, 0)
    iprops = np.zeros(1, np.int32)
    K, R = np.zeros(du.size * du.size), np.zeros(du.size)
    fail = C.c_int(-1)
    L.user_element_static_(ci(1), ci(element_identifier), ci(n_nodes), _ip(node), ci(pr.size), _dp(pr), ci(0), _ip(iprops),
                           _dp(xc), ci(xc.size), _dp(du), _dp(ut), ci(du.size), ci(sv0.size), _dp(sv0), _dp(sv1), _dp(K),
                           _dp(R), C.byref(fail))
    return K.reshape(du.size, du.size).T.copy(), R, sv1, fail.value


def uel(jtype, coords, u, du, props, jdltyp=(), adlmag=(), nsvars=48):
    """UEL (user_codes/src/abaqus_uel_linelastic_3d.for:16-28) through its Fortran-linkage twin, with the argument
    values the reference's assembly passes (src/solver_direct.f90:231-240).  coords is (NNODE, 3) row-major, i.e. the
    memory of Fortran's COORDS(MCRD, NNODE).  Returns AMATRX[row, col], RHS, SVARS, ENERGY, PNEWDT."""
    L = host_lib()
    ci = lambda v: C.byref(C.c_int(int(v)))
    cd = lambda v: C.byref(C.c_double(float(v)))
    xc, U, DU, pr = _f64(coords), _f64(u), _f64(du), _f64(props)
    nnode, nd = xc.size // 3, U.size
    jd, ad = _i32(list(jdltyp) or [0]), _f64(list(adlmag) or [0.0])
    ndload = len(jdltyp)
    mdload = max(ndload, 1)
    A, RHS, SV, EN = np.zeros(nd * nd), np.zeros(nd), np.zeros(nsvars), np.zeros(8)
    zeros, time2, lflags = np.zeros(nd), np.zeros(2), np.array([1, 0, 1, 0, 0], np.int32)
    params, predef, ddl, jprops = np.zeros(3), np.zeros(2 * nnode), np.zeros(mdload), np.zeros(1, np.int32)
    pnewdt = C.c_double(-1.0)
    L.uel_(_dp(RHS), _dp(A), _dp(SV), _dp(EN), ci(nd), ci(1), ci(nsvars), _dp(pr), ci(pr.size), _dp(xc), ci(3), ci(nnode),
           _dp(U), _dp(DU), _dp(zeros), _dp(zeros), ci(jtype), _dp(time2), cd(1.0), ci(1), ci(1), ci(1), _dp(params),
           ci(ndload), _ip(jd), _dp(ad), _dp(predef), ci(1), _ip(lflags), ci(nd), _dp(ddl), ci(mdload),
           C.byref(pnewdt), _ip(jprops), ci(0), cd(1.0))
    return A.reshape(nd, nd).T.copy(), RHS, SV, EN, pnewdt.value


class Partition:
    """Element partition of a model for one rank (en234host_partition)."""

    def __init__(self, model, n_ranks, rank):
        self.h = C.c_void_p(host_lib().en234host_partition(model.h, C.c_int(n_ranks), C.c_int(rank)))
        if not self.h:
            raise RuntimeError(host_lib().en234host_last_error().decode())
        self.n_ranks, self.rank = n_ranks, rank

    def ints(self, name):
        p, n = c_int_p(), C.c_long()
        if host_lib().en234host_part_int(self.h, name.encode(), C.byref(p), C.byref(n)):
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int32)

    @property
    def n_nodes(self):
        return int(self.ints("n_nodes")[0])

    def owned(self):
        p = host_lib().en234host_part_owned(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_nodes,)).copy()

    def close(self):
        if self.h:
            host_lib().en234host_part_free(self.h)
            self.h = None


def synthetic_block_input(nx, ny=None, nz=None, jitter=0.2, seed=1234, element="1001", E=100.0, nu=0.3,
                          solver="CONJUGATE GRADIENT", cg_tolerance=1e-17, cg_maxit=1000, load="traction",
                          nonlinear=None, steps=1, props=None, dt=1.0, dynamic=None):
    """`.in` text of the synthetic structured hex8 blocks of BASELINE.json: face x=0 clamped,
    face x=max loaded by the traction (0,0,-1) ("traction") or a prescribed u_x ramp ("stretch")."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    flag = int(element[1:]) + 99999 if element.upper().startswith("U") else int(element)
    plist = props if props is not None else [E, nu]
    t = ["MESH", " USER SUBROUTINE",
         "  %d, %d, %d, %.17g, %d, %d, 0" % (nx, ny, nz, jitter, seed, flag),
         "  " + ", ".join("%.17g" % p for p in plist), " END USER SUBROUTINE", "END MESH", "BOUNDARY CONDITIONS"]
    if load == "stretch":
        t += [" HISTORY, ramp", "  0.d0, 0.d0", "  %.17g, 0.2d0" % (steps * dt), " END HISTORY"]
    t += [" DEGREES OF FREEDOM", "  xmin, 1, VALUE, 0.d0", "  xmin, 2, VALUE, 0.d0", "  xmin, 3, VALUE, 0.d0"]
    if load == "stretch":
        t += ["  xmax, 1, HISTORY, ramp"]
    t += [" END DEGREES OF FREEDOM"]
    if load == "traction":
        t += [" DISTRIBUTED LOADS", "  xmax_el, 4, VALUE, 0.d0, 0.d0, -1.d0", " END DISTRIBUTED LOADS"]
    if dynamic:          # (time step, number of steps): EXPLICIT DYNAMIC STEP instead of the static step
        t += ["END BOUNDARY CONDITIONS", "EXPLICIT DYNAMIC STEP", " TIME STEP, %.17g" % dynamic[0], " NUMBER OF STEPS, %d" % dynamic[1],
              "END EXPLICIT DYNAMIC STEP", "STOP", ""]
        return "\n".join(t)
    t += ["END BOUNDARY CONDITIONS", "STATIC STEP", " INITIAL TIME STEP, %.17g" % dt, " MAX TIME STEP, %.17g" % dt,
          " MIN TIME STEP, %.17g" % (dt / 1024.0), " MAX NUMBER OF STEPS, %d" % steps, " STOP TIME, %.17g" % (steps * dt)]
    if nonlinear:
        t += [" SOLVER, %s, NONLINEAR, %.17g, %d" % (solver, nonlinear[0], nonlinear[1])]
    else:
        t += [" SOLVER, %s, LINEAR" % solver]
    t += [" CG TOLERANCE, %.17g" % cg_tolerance, " CG MAX ITERATIONS, %d" % cg_maxit, "END STATIC STEP", "STOP", ""]
    return "\n".join(t)
