"""The single-process multi-GPU entry (en234gpu_create_multi — what a Fortran driver binds) on the bench workload:
  python tools/multi_bench.py N [nx ny nz] [steps]
One static linear step per 'step' (assemble -> BC -> PCG -> re-assemble -> BC) on the 128^3 block through the
device-resident en234gpu_multi_* calls and through the host-buffer ones; wall-clock around the (blocking) calls."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from en234_fea_b200 import Model, MultiGpuSystem, synthetic_block_input

n = int(sys.argv[1])
dims = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (128, 128, 128)
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
TOL, MAXIT = 1e-20, 20000
m = Model.from_text(synthetic_block_input(*dims, jitter=0.2, cg_tolerance=TOL, cg_maxit=MAXIT))
fe, fx, fd, fv, fc = m.evaluate_loads(0.0, 1.0)
t0 = time.perf_counter()
g = MultiGpuSystem(m, devices=list(range(n)))
t_create = time.perf_counter() - t0
nd = g.ndof
z = np.zeros(nd)


def step_dev():
    g.zero_increment_dev()
    g.assemble_dev()
    g.apply_boundaryconditions_dev()
    _, its, fail = g.solve_dev(0, TOL, MAXIT)
    nrm, _ = g.assemble_dev()
    unb = g.apply_boundaryconditions_dev()
    assert not fail
    return its, unb / nrm


def step_host():
    rf, _, _ = g.assemble(z, z)
    g.apply_boundaryconditions(fx, fd, fc)
    du, _, its, fail = g.solve(z, 0, TOL, MAXIT)
    rf, nrm, _ = g.assemble(z, du)
    g.apply_boundaryconditions(fx, fd, fv - du[fd])
    assert not fail
    return its


g.upload_state(z, z, None, fx, fd, fv)
for _ in range(2):
    its, ratio = step_dev()
t0 = time.perf_counter()
for _ in range(steps):
    its, ratio = step_dev()
ms_dev = (time.perf_counter() - t0) * 1e3 / steps
st = g.stats()
step_host()
t0 = time.perf_counter()
for _ in range(steps):
    its_h = step_host()
ms_host = (time.perf_counter() - t0) * 1e3 / steps
print(json.dumps({"entry": "en234gpu_create_multi (one process, %d GPUs, one host thread per GPU, peer memory only)" % n,
                  "workload": "%dx%dx%d hex8 block, jitter 0.2" % dims, "n_dof": nd, "cg_iterations": int(its), "final_newton_ratio": ratio,
                  "value_MDOF_it_s": nd * its / (ms_dev * 1e-3) / 1e6, "ms_per_step": ms_dev, "e2e_MDOF_it_s": nd * its_h / (ms_host * 1e-3) / 1e6,
                  "ms_per_step_host_buffers": ms_host, "solve_ms_slowest_part": st.solve_ms, "ms_per_iteration": st.solve_ms / max(its, 1),
                  "create_s": t_create, "timing": "host wall clock around blocking calls"}))
g.close()
