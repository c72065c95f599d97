#!/usr/bin/env python
"""bench.py — static-step benchmark of the EN234_FEA hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithms (the C++ oracle port — the
                                                           Fortran reference cannot be compiled in this image) on host cores

A step = one static (linear) step of the synthetic hex8 block: assemble -> apply BCs -> Jacobi-PCG solve ->
re-assemble -> apply BCs -> convergence check (src/staticstep.f90:44-70 loop body).
metric  = MDOF-CG-iters/s = n_dof x CG iterations / step seconds / 1e6  (whole job, all ranks)
value   : inputs resident in HBM (device-resident entry points);  e2e : host-buffer C-ABI entry points with pinned host
          arrays, host<->device copies inside the timed region.

Default workload at EVERY N: the 128^3 block of BASELINE configs[2] / the north_star target, split into z-slabs over the N
GPUs (strong scaling).  Every line carries
  parity     a 12x12x32 block of the same family solved with the SAME N ranks before the timed region, against the oracle
             (skyline LDU displacements / reaction forces, CPU PCG iteration count) on rank 0
  secondary  (default run only; --no-secondary skips) c2 = configs[1] 64^3 on one GPU, c4 = configs[3] first increment of
             the 96^3 neo-Hookean Newton analysis on one GPU (N = 1 lines), c5 = configs[4] matrix-free PCG with 128^3
             elements per GPU (256^3 at N = 8)
Other workloads as the primary one: --workload c2 | c4 | c5 | weak (64^3 elements per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MDOF-CG-iters/s"
CG_TOL = 1e-20          # squared ratio (r.r)/(b.b): ||r||/||b|| <= 1e-10, the north_star residual
CG_MAXIT = 20000
PARITY_DIMS = (12, 12, 32)


def workload(n_gpus, name):
    if os.environ.get("EN234_BENCH_DIMS"):                 # test hook: a small block instead of the BASELINE sizes
        d = tuple(int(x) for x in os.environ["EN234_BENCH_DIMS"].split(","))
        return d, "test: %dx%dx%d hex8 block (EN234_BENCH_DIMS)" % d
    if name == "c4":
        return (96, 96, 96), "c4: synthetic 96^3 hex8 neo-Hookean block, Newton-Raphson with re-assembly (BASELINE configs[3])"
    if name == "c5":
        dims = {1: (128, 128, 128), 2: (128, 128, 256), 4: (128, 256, 256), 8: (256, 256, 256)}.get(n_gpus, (128, 128, 128 * n_gpus))
        return dims, "c5: synthetic %dx%dx%d hex8 block, matrix-free PCG, 128^3 elements per GPU%s" % (
            dims + (" (= BASELINE configs[4])" if n_gpus == 8 else "",))
    if name == "c2":
        return (64, 64, 64), "c2: synthetic 64^3 hex8 linear-elastic cube, Jacobi-PCG (BASELINE configs[1])"
    if name == "weak":
        dims = {1: (64, 64, 64), 2: (64, 64, 128), 4: (64, 128, 128), 8: (128, 128, 128)}.get(n_gpus, (64, 64, 64 * n_gpus))
        return dims, "weak: %dx%dx%d hex8 block, 64^3 elements per GPU, z-slabs" % dims
    return (128, 128, 128), ("c3: synthetic 128^3 hex8 linear-elastic block, assembled and solved to 1e-10 residual "
                             "(BASELINE configs[2], north_star target), z-slabs over the GPUs")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 5]
        os.unlink(self.f.name)
        if not rows:
            return out
        try:
            sm = sorted(float(r[0]) for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for k, n in enumerate(names) if any("Active" in r[2 + k] and "Not" not in r[2 + k] for r in rows)]
            out["samples"] = len(rows)
        except ValueError:
            pass
        return out


def measured_traffic(kernel, elements_per_gpu):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture) of the
    dominant kernel at this per-GPU size, from the committed profiles/traffic.json; None when that size was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    first = lambda name: name.replace("void ", "").split("<")[0].split("(")[0].split(" ")[0]
    for row in json.load(open(p)):
        if first(kernel) == first(row["kernel"]) and row["elements_per_gpu"] == elements_per_gpu:
            return row["dram_bytes_per_launch"]
    return None


class Dist:
    """torch.distributed plumbing (rendezvous, barrier, max over ranks); world 1 needs none of it."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class Block:
    """The synthetic block on this job's ranks: model, (partitioned) device handle, rank-local load tables."""

    def __init__(self, D, dims, jitter, method=0, element=None):
        from en234_fea_b200 import GpuSystem, Model, Partition, synthetic_block_input
        self.D, self.dims, self.method = D, dims, method
        kw = dict(element=element[0], props=element[1]) if element else {}
        self.model = Model.from_text(synthetic_block_input(*dims, jitter=jitter, cg_tolerance=CG_TOL, cg_maxit=CG_MAXIT, **kw))
        self.ndof_global = 3 * self.model.n_nodes
        fe, fx, fd, fv, fc = self.model.evaluate_loads(0.0, 1.0)
        if D.world > 1:
            self.part = Partition(self.model, D.world, D.rank)
            self.gpu = GpuSystem(self.model, device=D.local, part=self.part)
            self.comm_kind = self.gpu.setup_distributed()
            if self.comm_kind == "peer-memory":
                self.comm_kind = ("peer-memory, folded into the PCG kernels (EN234_P2P_FUSED=1)" if os.environ.get("EN234_P2P_FUSED") == "1"
                                  else "peer-memory (single-CTA collective kernels; interface rows and their exchange on a side stream)")
            self.l2g = self.part.ints("local_to_global_node").astype(np.int64)
            self.ldof = (3 * self.l2g[:, None] + np.arange(3)).ravel()
            g2l = -np.ones(self.ndof_global, np.int64)
            g2l[self.ldof] = np.arange(self.ldof.size)
            fx = fx[self.ldof]
            keep = g2l[fd] >= 0
            fd, fv, fc = g2l[fd[keep]].astype(np.int32), fv[keep], fc[keep]
            self.owned = np.repeat(self.part.owned().astype(bool), 3)
        else:
            self.part = None
            self.gpu = GpuSystem(self.model, device=D.local)
            self.comm_kind = "none"
            self.ldof = np.arange(self.ndof_global)
            self.owned = np.ones(self.ndof_global, bool)
        self.fx, self.fd, self.fv, self.fc = fx, fd.astype(np.int32), fv, fc
        self.gpu.set_operator_mode(method)
        self.stream = D.torch.cuda.current_stream()
        self.gpu.set_stream(self.stream.cuda_stream)
        self.n = self.gpu.ndof
        self.gpu.upload_state(np.zeros(self.n), np.zeros(self.n), None, self.fx, self.fd, self.fv)

    def step_device(self):
        g = self.gpu
        g.zero_increment_dev()
        _, fail = g.assemble_dev()
        g.apply_boundaryconditions_dev()
        _, its, sfail = g.solve_dev(self.method, CG_TOL, CG_MAXIT)
        nrm, fail2 = g.assemble_dev()
        unb = g.apply_boundaryconditions_dev()
        assert not (fail or sfail or fail2), "device step failed"
        return its, unb / nrm if nrm > 0 else 0.0

    def make_host_step(self):
        t = self.D.torch

        def pinned(a):
            x = t.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return x, x.numpy()

        self._keep = [pinned(np.zeros(self.n)), pinned(np.zeros(self.n)), pinned(self.fx), pinned(self.fd), pinned(self.fc),
                      pinned(self.fv), pinned(np.zeros(self.n))]
        ut, du, fxp, fdp, fcp, fvp, rfp = (k[1] for k in self._keep)
        g, method = self.gpu, self.method

        def step_host():
            du[:] = 0.0
            fcp[:] = fvp
            _, _, fail = g.assemble(ut, du, None, out_rforce=rfp)
            g.apply_boundaryconditions(fxp, fdp, fcp)
            _, corr, its, sfail = g.solve(du, method, CG_TOL, CG_MAXIT, inplace=True)
            _, nrm, fail2 = g.assemble(ut, du, None, out_rforce=rfp)
            fcp[:] = fvp - du[fdp]
            g.apply_boundaryconditions(fxp, fdp, fcp)
            assert not (fail or sfail or fail2), "host-API step failed"
            return its

        nfix = int(self.fd.size)
        self.h2d = 2 * (2 * self.n * 8) + 2 * (self.n * 8 + nfix * 12) + self.n * 8
        self.d2h = 2 * self.n * 8 + self.n * 8
        return step_host

    def timed(self, fn, steps, warmup):
        D, t = self.D, self.D.torch
        out = None
        for _ in range(warmup):
            out = fn()
        D.barrier()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        l0 = self.gpu.stats().kernel_launches
        e0.record(self.stream)
        for _ in range(steps):
            out = fn()
        e1.record(self.stream)
        D.barrier()
        ms = D.max(e0.elapsed_time(e1))
        return ms / steps, out, self.gpu.stats().kernel_launches - l0

    def gather(self, local):
        """whole-mesh image of a rank-local DOF array (owned entries), on every rank"""
        t = self.D.torch
        acc = t.zeros(self.ndof_global, dtype=t.float64, device="cuda")
        own = self.owned
        acc[t.from_numpy(self.ldof[own]).cuda()] = t.from_numpy(np.ascontiguousarray(local[own])).cuda()
        if self.D.world > 1:
            self.D.dist.all_reduce(acc)
        return acc.cpu().numpy()

    def close(self):
        self.gpu.close()


def parity_gate(D):
    """SURVEY 8(d) 'parity gates printed with every perf number': a block of the workload's family, small enough for the
    oracle's skyline LDU, solved by the SAME ranks (same partitioned code path, same collectives) right before the timed
    region.  max_rel_u / max_rel_rforce: against the oracle's direct solution; cg_iters: device vs the CPU recurrences."""
    import oracle_binding as ob
    dims = PARITY_DIMS if not os.environ.get("EN234_BENCH_DIMS") else (4, 4, max(4, 2 * D.world))
    B = Block(D, dims, 0.2)
    g = B.gpu
    tol = 1e-26
    g.zero_increment_dev()
    _, fail = g.assemble_dev()
    g.apply_boundaryconditions_dev()
    _, its, sfail = g.solve_dev(0, tol, CG_MAXIT)
    nrm, fail2 = g.assemble_dev()
    unb = g.apply_boundaryconditions_dev()
    du, rf = g.download_state()
    err = g.stats().last_cg_err
    u, r = B.gather(du), B.gather(rf)
    out = None
    if D.rank == 0:
        a = ob.synthetic_block_arrays(*dims, jitter=0.2, cg_tolerance=tol, cg_maxit=CG_MAXIT)
        ref = ob.OracleModel(a, solvertype=1).run_static()
        refcg = ob.OracleModel(a, solvertype=2, cg_check_linear=1).run_static()
        out = {"block": "%dx%dx%d, jitter 0.2, PCG to (r.r)/(b.b) < %g on %d rank%s" % (dims + (tol, D.world, "" if D.world == 1 else "s")),
               "oracle": "skyline LDU (u, rforce), COO Jacobi-PCG (iteration count); mesh from oracle/or_mesh.cpp",
               "max_rel_u": float(np.abs(u - ref["dof_total"]).max() / np.abs(ref["dof_total"]).max()),
               "max_rel_rforce": float(np.abs(r - ref["rforce"]).max() / np.abs(ref["rforce"]).max()),
               "cg_iters_cpu": int(refcg["cg_iterations"][0]), "cg_iters_gpu": int(its),
               "exit_rr_over_bb": float(err), "final_newton_ratio": float(unb / nrm) if nrm > 0 else 0.0,
               "failed": bool(fail or sfail or fail2)}
        out["ok"] = bool(out["max_rel_u"] <= 1e-10 and out["max_rel_rforce"] <= 1e-10 and not out["failed"]
                         and abs(out["cg_iters_cpu"] - out["cg_iters_gpu"]) <= 2)
    B.close()
    D.barrier()
    return out


def run_newton(D, args, increments=4, steps=None, warmup=None, with_e2e=True):
    """configs[3]: neo-Hookean block, prescribed stretch ramp, NONLINEAR (tolerance 1e-5, <= 15 iterations), full
    re-assembly in every Newton iteration; the whole analysis runs through the host driver (src/staticstep.f90 loop restated
    in csrc/host/staticstep.cpp).  value: device-resident state; e2e: host-buffer API.  1 GPU."""
    from en234_fea_b200 import GpuSystem, Model, synthetic_block_input
    torch = D.torch
    dims, wname = workload(1, "c4")
    steps = args.steps if steps is None else steps
    warmup = max(args.warmup, 1) if warmup is None else warmup
    # the ramp always spans 4 increments; `increments` of them are run
    text = synthetic_block_input(*dims, jitter=args.jitter, element="U2", props=[1.0, 200.0], load="stretch",
                                 steps=4, nonlinear=(1e-5, 15), cg_tolerance=CG_TOL, cg_maxit=10 * CG_MAXIT)
    if increments < 4:
        text = text.replace(" MAX NUMBER OF STEPS, 4", " MAX NUMBER OF STEPS, %d" % increments).replace(
            " STOP TIME, 4", " STOP TIME, %d" % increments)
    model = Model.from_text(text)
    gpu = GpuSystem(model, device=D.local)
    n = gpu.ndof
    sampler = ClockSampler(D.local)
    res = None
    for _ in range(warmup):
        res = gpu.run_static(method=0)
    torch.cuda.synchronize()
    l0 = gpu.stats().kernel_launches
    t0 = time.perf_counter()
    for _ in range(steps):
        res = gpu.run_static(method=0)
    torch.cuda.synchronize()
    ms_dev = (time.perf_counter() - t0) * 1e3 / steps
    launches = gpu.stats().kernel_launches - l0
    clocks = sampler.stop()
    ms_e2e, res_h = None, None
    if with_e2e:
        t0 = time.perf_counter()
        res_h = gpu.run_static(method=0, use_host_api=True)
        ms_e2e = (time.perf_counter() - t0) * 1e3
    its = int(res["cg_iterations"].sum())
    st = gpu.stats()
    nnzb = gpu.nnz_blocks()
    alg_it = 8 * 9 * nnzb + 4 * nnzb + 4 * (gpu.n_nodes + 1) + 16 * n + 104 * n
    peak, peak_src = peaks()
    n_solves = len(res["cg_iterations"])
    line = {"metric": METRIC, "value": n * its / (ms_dev * 1e-3) / 1e6, "unit": "MDOF-it/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname + (", first %d of 4 increments" % increments if increments < 4 else ", 4 increments"),
                       "n_dof": n, "elements": dims[0] * dims[1] * dims[2], "jitter": args.jitter,
                       "increments": res["n_steps"], "cutbacks": res["n_cutbacks"], "newton_iterations": int(res["newton"].shape[0]),
                       "linear_solves": n_solves, "cg_iterations": its, "cg_rule": "(r.r)/(b.b) < 1e-20",
                       "cg_iterations_per_solve": [int(x) for x in res["cg_iterations"]],
                       "newton_ratios": [float("%.4e" % r) for r in res["newton"][:, 5]],
                       "timing": "host wall clock around the whole analysis (device-synchronised)",
                       "l2": "matrix stream per PCG iteration (%.0f MB) exceeds the 126 MB L2" % ((alg_it - 104 * n) / 1e6)},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "PCG iteration (k_spmv + vector kernels) inside the Newton loop",
                         "achieved": alg_it * its / (ms_dev * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_it * its / (ms_dev * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                         "phases_ms_last_call": {"assemble": st.assemble_ms, "bc": st.bc_ms, "solve": st.solve_ms}},
            "parity": "neo-Hookean element: parity unpinned by the reference (it ships no hyperelastic hex8 user element); "
                      "device vs oracle Newton history at 5^3: tests/test_gpu_parity.py::test_neohooke_element_and_newton_history"}
    if with_e2e:
        line["e2e"] = {"value": n * int(res_h["cg_iterations"].sum()) / (ms_e2e * 1e-3) / 1e6, "unit": "MDOF-it/s",
                       "ms_per_step": ms_e2e, "h2d_bytes_per_step": int((3 * res["newton"].shape[0] + n_solves) * n * 8),
                       "d2h_bytes_per_step": int((2 * res["newton"].shape[0] + n_solves) * n * 8)}
    gpu.close()
    return line


def run_reference(args):
    """Reference arm: the reference's own CPU algorithms (COO Jacobi-PCG of src/solver_conjugate_gradient.f90:776-912 on
    the system assembled by its element loop), restated in C++ (oracle/) because no Fortran compiler exists here.
    Serial, like the reference (no OpenMP/MPI; module-level scratch makes its element routines non-re-entrant).
    The mesh comes from the oracle's own generator (oracle/or_mesh.cpp): no product library is loaded on this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_binding as ob
    dims, wname = workload(args.gpus, args.workload)
    its_per_step = args.ref_iters
    t0 = time.time()
    om = ob.OracleModel(ob.synthetic_block_arrays(*dims, jitter=0.2, cg_tolerance=CG_TOL, cg_maxit=CG_MAXIT))
    s = ob.OracleSystem(om, 2)
    n = om.ndof
    z = np.zeros(n)
    t1 = time.time()
    s.assemble(z, z, 0.0, 1.0)
    t_asm = time.time() - t1
    t1 = time.time()
    s.apply_bc(z, z, 0.0, 1.0)
    t_bc = time.time() - t1
    times = []
    for k in range(args.warmup + args.steps):
        t1 = time.time()
        s.solve(z, CG_TOL, CG_MAXIT, fixed_iters=its_per_step)
        if k >= args.warmup:
            times.append(time.time() - t1)
    t_step = float(np.mean(times))
    value = n * its_per_step / t_step / 1e6
    sample = ("%d PCG iterations per step on the assembled %dx%dx%d system (COO upper-triangle SpMV, 1 thread); assembly %.1f s "
              "and BC %.2f s done once outside the timed steps" % (its_per_step, dims[0], dims[1], dims[2], t_asm, t_bc))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "MDOF-it/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "n_dof": n, "elements": dims[0] * dims[1] * dims[2]},
            "cpu_baseline": {"value": value, "unit": "MDOF-it/s", "cores": 1, "kind": "port", "sample": sample,
                             "cpu": cpu_model(), "host_cores_available": os.cpu_count(), "assemble_s": t_asm, "bc_s": t_bc,
                             "cg_iterations_per_step": its_per_step, "setup_s": t1 - t0},
            "e2e": {"value": value, "unit": "MDOF-it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(dims, gpu_iters):
    """Bounded CPU sample for the N=1 line (about 20 s): the oracle on a 64^3 corner of the workload's mesh family — one
    assembly + BC + a fixed window of PCG iterations; the CPU's per-DOF iteration rate does not depend on the block size
    (memory-bound COO SpMV), so the step time is put together for the same step definition as the GPU line."""
    import oracle_binding as ob
    sd = tuple(min(d, 64) for d in dims)
    om = ob.OracleModel(ob.synthetic_block_arrays(*sd, jitter=0.2, cg_tolerance=CG_TOL, cg_maxit=CG_MAXIT))
    s = ob.OracleSystem(om, 2)
    n = om.ndof
    z = np.zeros(n)
    t1 = time.time(); s.assemble(z, z, 0.0, 1.0); t_asm = time.time() - t1
    t1 = time.time(); s.apply_bc(z, z, 0.0, 1.0); t_bc = time.time() - t1
    win = 100
    s.solve(z, CG_TOL, CG_MAXIT, fixed_iters=2)
    t1 = time.time(); s.solve(z, CG_TOL, CG_MAXIT, fixed_iters=win); t_it = (time.time() - t1) / win
    # same step definition as the GPU line: 2 assemblies + 2 BC passes + gpu_iters PCG iterations, per DOF of the sample
    t_step = 2 * (t_asm + t_bc) + gpu_iters * t_it
    return {"value": n * gpu_iters / t_step / 1e6, "unit": "MDOF-it/s", "cores": 1, "kind": "port", "cpu": cpu_model(),
            "sample": "oracle (C++ port of the reference, 1 thread) on a %dx%dx%d block of the same family: 1 assembly (%.1f s) + BC "
                      "(%.2f s) + %d PCG iterations (%.3f s each); step time put together as 2 assemblies + 2 BC passes + %d "
                      "iterations (the GPU line's step), rate per DOF" % (sd + (t_asm, t_bc, win, t_it, gpu_iters)),
            "pcg_iteration_only": n / t_it / 1e6, "host_cores_available": os.cpu_count()}


def run_linear(D, args, wl, method, steps, warmup, with_e2e=True, jitter=None):
    """One linear static step workload on the job's ranks -> the JSON line (without parity / secondary / cpu_baseline)."""
    jitter = args.jitter if jitter is None else jitter
    dims, wname = workload(D.world, wl)
    B = Block(D, dims, jitter, method)
    gpu, n, world = B.gpu, B.n, D.world
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    ms_dev, (its, ratio), launches = B.timed(B.step_device, steps, warmup)
    clocks = sampler.stop() if sampler else {}
    value = B.ndof_global * its / (ms_dev * 1e-3) / 1e6
    st = gpu.stats()
    it_ms = D.max(st.solve_ms) / max(st.last_cg_iterations, 1)
    phases = {"assemble": D.max(st.assemble_ms), "bc": D.max(st.bc_ms), "solve": D.max(st.solve_ms)}
    e2e = None
    if with_e2e:
        step_host = B.make_host_step()
        ms_e2e, its_h, _ = B.timed(step_host, steps, 1)
        e2e = {"value": B.ndof_global * its_h / (ms_e2e * 1e-3) / 1e6, "unit": "MDOF-it/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": B.h2d, "d2h_bytes_per_step": B.d2h}
        gpu.upload_state(np.zeros(n), np.zeros(n), None, B.fx, B.fd, B.fv)
        B.step_device()                                   # leave a valid assembled matrix for bench_operator

    # roofline of the dominant kernel (operator application inside PCG), measured live with CUDA events on the launch stream
    peak, peak_src = peaks()
    N_loc = gpu.n_nodes
    e_loc = dims[0] * dims[1] * dims[2] // world
    if method == 0:
        ms_op = gpu.bench_operator(0, 5, 50)
        nnzb = gpu.nnz_blocks()
        alg = 8 * 9 * nnzb + 4 * nnzb + 4 * (N_loc + 1) + 16 * n
        kern = "k_spmv (3x3-block CSR, FP64)"
        extra = {"algorithmic_bytes_scalar_csr": 12 * 9 * nnzb + 20 * n}
    elif gpu.stats().mf_congruent:
        ms_op = gpu.bench_operator(1, 5, 50)
        nnzb = gpu.nnz_blocks()
        # column indices (ELL, one per block) + sorted-order maps + operand read / masked copy written and read / result
        alg = 4 * nnzb + 8 * N_loc + (8 + 1 + 8 + 8 + 8) * n
        kern = "k_mask_soa + k_stencil_nodes (matrix-free, congruent mesh: distinct block rows from a table)"
        extra = {"note": "uniform mesh: every element a bitwise translate of element 0; no matrix values are read"}
    else:
        ms_op = gpu.bench_operator(1, 5, 50)
        geom = 640 * e_loc if os.environ.get("EN234_MF_GEOM", "1") != "0" else 0
        alg = 72 * N_loc + 32 * e_loc + 2 * 192 * e_loc + geom
        kern = "k_element_forces + k_mf_gather (matrix-free, element by element)"
        # FP64 roof: flops of one element application as the kernel is written
        fp64_peak = gpu.fp64_peak()
        per_gp = (3 * 72 * 2 + 100) if geom else (4 * 72 * 2 + 134)
        flops = e_loc * 8 * per_gp
        extra = {"fp64": {"peak_tflops_measured": fp64_peak, "achieved_tflops": flops / (ms_op * 1e-3) / 1e12,
                          "frac": flops / (ms_op * 1e-3) / 1e12 / fp64_peak, "flops_per_launch": flops},
                 "note": "general mesh: FP64 pipe and HBM are both in play (cached J^-1 and w|J| per Gauss point: %d B per "
                         "element); fp64.frac is the fraction of the measured FMA peak" % (640 if geom else 0)}
    ms_op = D.max(ms_op)
    ach = alg / (ms_op * 1e-3) / 1e9
    b_iter = alg + 104 * n
    roof = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "frac_of_nominal_8000": ach / 8000.0, "per": "GPU (slowest rank)",
            "peak_source": peak_src, "traffic": measured_traffic(kern, e_loc),
            "ms_per_launch": ms_op, "algorithmic_bytes_per_launch": alg,
            "cg_iteration": {"ms": it_ms, "algorithmic_bytes": b_iter, "achieved": b_iter / (it_ms * 1e-3) / 1e9,
                             "frac": b_iter / (it_ms * 1e-3) / 1e9 / peak},
            "phases_ms": phases}
    roof.update(extra)
    line = {"metric": METRIC, "value": value, "unit": "MDOF-it/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong" if wl in ("auto", "c3", "c2") else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "n_dof": B.ndof_global, "elements": dims[0] * dims[1] * dims[2], "jitter": jitter,
                       "cg_iterations": int(its), "cg_rule": "(r.r)/(b.b) < 1e-20", "final_newton_ratio": ratio,
                       "method": "assembled 3x3-block CSR" if method == 0 else
                                 ("matrix-free (congruent-mesh row table)" if gpu.stats().mf_congruent else "matrix-free (element by element)"),
                       "l2": "matrix stream per PCG iteration (%.0f MB per GPU) exceeds the 126 MB L2" % (alg / 1e6) if alg > 2.5e8 else
                             "operator traffic per iteration %.0f MB per GPU; vectors (%.0f MB) are rewritten every iteration" % (alg / 1e6, 40 * n / 1e6),
                       "step": "assemble+BC+PCG solve+re-assemble+BC (src/staticstep.f90:44-70)",
                       "collectives": B.comm_kind},
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof}
    if e2e:
        line["e2e"] = e2e
    B.close()
    D.barrier()
    return line, dims, int(its)


def brief(line):
    """what a secondary entry keeps of a full line"""
    keep = {k: line[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "gpu_launches") if k in line}
    if "e2e" in line:
        keep["e2e"] = line["e2e"]
    r = line.get("roofline", {})
    keep["roofline"] = {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "ms_per_launch", "cg_iteration", "fp64", "phases_ms",
                                          "phases_ms_last_call") if k in r}
    return keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c3", "c4", "c5", "weak"])
    ap.add_argument("--jitter", type=float, default=0.2)
    ap.add_argument("--method", type=int, default=0, help="0 assembled block-CSR PCG, 1 matrix-free PCG")
    ap.add_argument("--ref-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the single JSON line
    D = Dist()
    warmup = max(args.warmup, 3)
    parity = None if args.no_parity else parity_gate(D)
    if args.workload == "c4":
        if D.world > 1:
            raise SystemExit("bench.py --workload c4 is a 1-GPU measurement")
        line = run_newton(D, args)
        dims, its = None, None
    else:
        method = 1 if args.workload == "c5" else args.method
        line, dims, its = run_linear(D, args, args.workload, method, args.steps, warmup)
    if parity is not None:
        line["parity"] = parity
    test_hook = bool(os.environ.get("EN234_BENCH_DIMS"))
    if args.workload == "auto" and not args.no_secondary and not test_hook:
        sec = {}
        if D.world == 1:
            l2, _, _ = run_linear(D, args, "c2", 0, 3, 3)
            sec["c2"] = brief(l2)
            sec["c4"] = brief(run_newton(D, args, increments=1, steps=1, warmup=1, with_e2e=False))
        l5, _, _ = run_linear(D, args, "c5", 1, 2, 3, with_e2e=False)
        sec["c5_jittered"] = brief(l5)
        l5u, _, _ = run_linear(D, args, "c5", 1, 2, 3, with_e2e=False, jitter=0.0)
        sec["c5_uniform"] = brief(l5u)
        line["secondary"] = sec
    if D.rank == 0:
        if D.world == 1 and not args.no_cpu_baseline and dims is not None:
            line["cpu_baseline"] = cpu_baseline_sample(dims, its)
        print(json.dumps(line), flush=True)
    D.close()


if __name__ == "__main__":
    main()
